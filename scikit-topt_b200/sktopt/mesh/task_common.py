"""Task container shared by the elasticity and heat problems.

Keeps the 22 fields and the constructors of the reference's ``FEMDomain``
(``mesh/task_common.py:24-347``); index bookkeeping is vectorised NumPy
instead of per-element Python loops.  This is setup-time host code: it produces
the arrays the GPU path consumes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal

import numpy as np

from sktopt.mesh import utils
from sktopt.fea import composer

_lit_bc = Literal['u^1', 'u^2', 'u^3', 'all']
_lit_force = Literal['u^1', 'u^2', 'u^3']


def setdiff1d(a, b):
    """Order-preserving difference (``mesh/task_common.py:18-21``)."""
    a = np.asarray(a)
    return np.ascontiguousarray(a[~np.isin(a, b)])


def _dofs_of(basis, nodes, direction):
    view = basis.get_dofs(nodes=nodes)
    return view.all() if direction in (None, 'all') else view.nodal[direction]


def _nodes_of_facets(facets, ids):
    if ids is None:
        return None
    if isinstance(ids, list):
        return [np.unique(facets[:, np.asarray(i, dtype=np.int64)].ravel()) for i in ids]
    if isinstance(ids, np.ndarray):
        return np.unique(facets[:, ids.astype(np.int64)].ravel())
    raise ValueError("facet ids should be list[np.ndarray] or np.ndarray")


@dataclass
class FEMDomain():
    """Mesh, boundary-condition sets and design/fixed element sets of a task."""

    basis: object
    dirichlet_nodes: np.ndarray | list[np.ndarray] | None
    dirichlet_dofs: np.ndarray | list[np.ndarray] | None
    dirichlet_elements: np.ndarray | None
    dirichlet_values: float | list[float] | None

    neumann_nodes: np.ndarray | list[np.ndarray] | None
    neumann_elements: np.ndarray | None
    neumann_dir_type: str | list[str] | None
    neumann_values: float | list[float] | None

    robin_facets_ids: np.ndarray | list[np.ndarray] | None
    robin_nodes: np.ndarray | list[np.ndarray] | None
    robin_elements: np.ndarray | None
    robin_coefficient: float | list[float] | None
    robin_bc_value: float | list[float] | None
    design_robin_boundary: bool | None

    design_elements: np.ndarray
    free_dofs: np.ndarray
    free_elements: np.ndarray
    all_elements: np.ndarray
    fixed_elements: np.ndarray
    dirichlet_neumann_elements: np.ndarray
    elements_volume: np.ndarray

    @property
    def n_tasks(self) -> int:
        raise NotImplementedError("")

    @property
    def design_mask(self):
        return np.isin(self.all_elements, self.design_elements)

    @property
    def mesh(self):
        return self.basis.mesh

    @classmethod
    def from_nodes(cls, basis, dirichlet_nodes, dirichlet_dir, dirichlet_values,
                   neumann_nodes, neumann_dir_type, neumann_values,
                   robin_facets_ids, robin_nodes, robin_coefficient,
                   robin_bc_value, design_robin_boundary,
                   design_elements) -> 'FEMDomain':
        mesh = basis.mesh
        # ---- Dirichlet
        if dirichlet_nodes is None:
            dirichlet_dofs = None
            dirichlet_elements = None
        elif isinstance(dirichlet_nodes, list):
            if dirichlet_dir is None:
                assert isinstance(dirichlet_values, (list, float))
                if isinstance(dirichlet_values, list):
                    assert len(dirichlet_nodes) == len(dirichlet_values)
                dirs = [None] * len(dirichlet_nodes)
            else:
                assert isinstance(dirichlet_dir, list)
                assert len(dirichlet_nodes) == len(dirichlet_dir)
                dirs = dirichlet_dir
            dirichlet_dofs = [_dofs_of(basis, n, d) for n, d in zip(dirichlet_nodes, dirs)]
            dirichlet_elements = utils.get_elements_by_nodes(mesh, [np.concatenate(dirichlet_nodes)])
        elif isinstance(dirichlet_nodes, np.ndarray):
            assert isinstance(dirichlet_dir, str)
            dirichlet_dofs = _dofs_of(basis, dirichlet_nodes, dirichlet_dir)
            dirichlet_elements = utils.get_elements_by_nodes(mesh, [dirichlet_nodes])
        else:
            raise ValueError("dirichlet_nodes should be list or np.ndarray")

        # ---- Neumann / Robin element sets
        def touching(nodes, label):
            if nodes is None:
                return None
            group = [nodes] if isinstance(nodes, np.ndarray) else nodes
            elems = utils.get_elements_by_nodes(mesh, group)
            if elems.shape[0] == 0:
                raise ValueError(f"{label}_elements has not been set.")
            return elems

        neumann_elements = touching(neumann_nodes, "neumann")
        robin_elements = touching(robin_nodes, "robin")

        # ---- design set: drop loaded (and, on request, Robin) elements
        excluded = np.array([])
        if neumann_elements is not None:
            excluded = np.concatenate([excluded, neumann_elements])
        if design_robin_boundary is False:
            excluded = np.concatenate([excluded, robin_elements])
        design_elements = setdiff1d(design_elements, excluded)
        if len(design_elements) == 0:
            raise ValueError("⚠️Warning: `design_elements` is empty")

        all_elements = np.arange(mesh.nelements)
        fixed_elements = setdiff1d(all_elements, design_elements)
        parts = [s for s in (dirichlet_elements, neumann_elements)
                 if s is not None and len(s) > 0]
        dirichlet_neumann_elements = (
            np.concatenate(parts) if parts else np.array([], dtype=int)
        )
        flat_dir = (np.concatenate(dirichlet_dofs)
                    if isinstance(dirichlet_dofs, list) else dirichlet_dofs)
        free_dofs = setdiff1d(np.arange(basis.N), flat_dir)
        # the reference passes DOF ids as node ids here;
        # metadata only, so out-of-range ids are ignored
        free_elements = utils.get_elements_by_nodes(mesh, [free_dofs])
        elements_volume = composer.get_elements_volume(mesh)
        return cls(
            basis, dirichlet_nodes, dirichlet_dofs, dirichlet_elements,
            dirichlet_values, neumann_nodes, neumann_elements, neumann_dir_type,
            neumann_values, robin_facets_ids, robin_nodes, robin_elements,
            robin_coefficient, robin_bc_value, design_robin_boundary,
            design_elements, free_dofs, free_elements, all_elements,
            fixed_elements, dirichlet_neumann_elements, elements_volume,
        )

    @classmethod
    def from_facets(cls, basis, dirichlet_facets_ids, dirichlet_dir,
                    dirichlet_values, neumann_facets_ids, neumann_dir_type,
                    neumann_values, robin_facets_ids, robin_coefficient,
                    robin_bc_value, design_robin_boundary,
                    design_elements) -> 'FEMDomain':
        facets = basis.mesh.facets
        dirichlet_nodes = _nodes_of_facets(facets, dirichlet_facets_ids)

        def merged(ids):
            return np.concatenate(ids) if isinstance(ids, list) else ids

        if neumann_facets_ids is not None:
            neumann_nodes = np.unique(
                facets[:, np.asarray(merged(neumann_facets_ids), dtype=np.int64)].ravel())
        else:
            neumann_nodes = neumann_dir_type = neumann_values = None
        if robin_facets_ids is not None:
            robin_nodes = np.unique(
                facets[:, np.asarray(merged(robin_facets_ids), dtype=np.int64)].ravel())
        else:
            robin_nodes = robin_coefficient = robin_bc_value = None
            design_robin_boundary = None
        return cls.from_nodes(
            basis, dirichlet_nodes, dirichlet_dir, dirichlet_values,
            neumann_nodes, neumann_dir_type, neumann_values,
            robin_facets_ids, robin_nodes, robin_coefficient, robin_bc_value,
            design_robin_boundary, design_elements,
        )

    @classmethod
    def from_json(self, path: str):
        raise NotImplementedError("not implmented yet")

    @property
    def neumann_nodes_all(self) -> np.ndarray:
        if isinstance(self.neumann_nodes, list):
            return np.unique(np.concatenate(self.neumann_nodes))
        return self.neumann_nodes

    def export_analysis_condition_on_mesh(self, dst_path: str):
        """``condition.vtu`` with the node / element colouring of the reference
        (``mesh/task_common.py:360-396``), written by the native VTU writer."""
        from sktopt.core.visualization import export_mesh_with_info
        mesh = self.basis.mesh
        node_color = np.zeros(mesh.nvertices, dtype=int)
        if self.neumann_nodes is not None:
            node_color[self.neumann_nodes_all] = 1
        if self.dirichlet_nodes is not None:
            dn = self.dirichlet_nodes
            node_color[np.concatenate(dn) if isinstance(dn, list) else dn] = 2
        if self.robin_nodes is not None:
            rn = self.robin_nodes
            node_color[np.concatenate(rn) if isinstance(rn, list) else rn] = 3
        elem_color = np.zeros(mesh.nelements, dtype=int)
        elem_color[self.free_elements] = 1
        elem_color[self.fixed_elements] = 2
        elem_color[self.design_elements] = 3
        try:
            export_mesh_with_info(mesh, point_data_values=[node_color],
                                  point_data_names=["node_color"],
                                  cell_data_values=[elem_color], cell_data_names=["condition"],
                                  filepath=f"{dst_path}/condition.vtu")
        except OSError:
            pass

    def nodes_and_elements_stats(self, dst_path: str | None = None) -> dict:
        """Nearest-neighbour distance statistics of nodes and element centres
        (reference ``mesh/task_common.py:424-449``; the histogram figure it also
        saves needs matplotlib and is not produced here).  Returns the numbers it
        prints."""
        from scipy.spatial import cKDTree
        mesh = self.basis.mesh
        out = {}
        for key, title, pts in (
                ("nodes", "=== Distance between nodes ===", mesh.p.T),
                ("elements", "\n=== Distance between elements ===",
                 np.mean(mesh.p[:, mesh.t], axis=1).T)):
            d = cKDTree(pts).query(pts, k=2)[0][:, 1]
            st = dict(min=float(np.min(d)), max=float(np.max(d)), mean=float(np.mean(d)),
                      median=float(np.median(d)), std=float(np.std(d)))
            print(title)
            for name in ("min", "max", "mean", "median", "std"):
                print(f"{name + ':':8s}{st[name]:.4f}")
            out[key] = st
        return out

    def exlude_dirichlet_from_design(self):
        self.design_elements = setdiff1d(self.design_elements, self.dirichlet_elements)

    def scale(self, L_scale: float, F_scale: float):
        from sktopt._fem import Basis
        mesh = self.basis.mesh
        scaled = type(mesh)(mesh.p * L_scale, mesh.t, mesh.boundaries, mesh.subdomains)
        self.basis = Basis(scaled, self.basis.elem, intorder=self.basis.intorder)
        # Scale the load arrays in place, never through the ``force`` property:
        # for a single-load task its getter returns neumann_linear[0] and the
        # setter would store that ndarray back as the list itself, so n_tasks
        # would become n_dof (the reference, mesh/task_common.py:466-476, has
        # that defect; it is not reproduced).
        loads = getattr(self, "neumann_linear", None)
        if loads is None:
            loads = self.force
        if isinstance(loads, np.ndarray):
            loads *= F_scale
        elif isinstance(loads, list):
            for f in loads:
                f *= F_scale
        else:
            raise ValueError("should be ndarray or list of ndarray")
