"""Linear-elasticity task (reference ``mesh/task_elastic.py``)."""
from __future__ import annotations

import re
from dataclasses import dataclass, fields
from typing import List, Literal, Union

import numpy as np

from sktopt._fem import facet_area, facet_load
from sktopt.mesh.task_common import FEMDomain

_lit_bc = Literal['u^1', 'u^2', 'u^3', 'all']
_lit_force = Literal['u^1', 'u^2', 'u^3']


def _component(s: str) -> int:
    if not (isinstance(s, str) and s.startswith('u^') and s[2:].isdigit()):
        raise ValueError(f"force_dir_type must be like 'u^1','u^2','u^3', got: {s}")
    c = int(s[2:]) - 1
    if c < 0:
        raise ValueError(f"Invalid component index parsed from {s}")
    return c


def assemble_surface_forces(
    basis,
    force_facets_ids: Union[np.ndarray, List[np.ndarray]],
    force_dir_type: Union[str, List[str]],
    force_value: Union[float, List[float]],
):
    """Consistent surface load vectors: the total ``force_value`` is spread as
    a uniform traction value/area over the listed facets
    (reference ``mesh/task_elastic.py:15-81``)."""
    as_list = lambda x: x if isinstance(x, list) else [x]
    facets_l, dirs_l, vals_l = as_list(force_facets_ids), as_list(force_dir_type), as_list(force_value)
    if not (len(facets_l) == len(dirs_l) == len(vals_l)):
        raise ValueError(
            "Lengths of force_facets_ids, force_dir_type, and force_value must match when lists."
        )
    out = []
    for facets, dir_s, val in zip(facets_l, dirs_l, vals_l):
        comp = _component(dir_s)
        ids = np.asarray(facets, dtype=int)
        pressure = float(val) / facet_area(basis.mesh, ids)
        out.append(facet_load(basis.mesh, ids, pressure, dpn=basis.dpn, comp=comp))
    return out[0] if len(out) == 1 else out


@dataclass
class LinearElasticity(FEMDomain):
    """FEMDomain + material constants and the assembled load vector(s)."""

    E: float
    nu: float
    neumann_linear: list
    body_force: np.ndarray | None = None

    @property
    def material_coef(self) -> float:
        return self.E

    @property
    def n_tasks(self) -> int:
        return len(self.neumann_linear)

    @property
    def force(self):
        return self.neumann_linear[0] if len(self.neumann_linear) == 1 else self.neumann_linear

    @force.setter
    def force(self, value):
        self.neumann_linear = value

    @property
    def force_elements(self):
        return self.neumann_elements

    @property
    def force_elements_all(self) -> np.ndarray:
        """Union of the loaded elements of every load case.  (The reference
        property, ``mesh/task_elastic.py:261-263``, forwards to an attribute
        ``neumann_elements_all`` that ``FEMDomain`` never defines.)"""
        ne = self.neumann_elements
        if isinstance(ne, list):
            return np.unique(np.concatenate([np.asarray(a).ravel() for a in ne]))
        return np.asarray(ne)

    @force_elements.setter
    def force_elements(self, value):
        self.neumann_elements = value

    @property
    def force_nodes(self):
        return self.neumann_elements

    @force_nodes.setter
    def force_nodes(self, value):
        self.neumann_nodes = value

    @property
    def dirichlet_force_elements(self):
        return self.dirichlet_neumann_elements

    @dirichlet_force_elements.setter
    def dirichlet_force_elements(self, value):
        self.dirichlet_neumann_elements = value

    @classmethod
    def from_facets(cls, basis, dirichlet_facets_ids, dirichlet_dir,
                    force_facets_ids, force_dir_type, force_value,
                    design_elements, E: float, nu: float) -> 'LinearElasticity':
        base = FEMDomain.from_facets(
            basis, dirichlet_facets_ids, dirichlet_dir, None,
            force_facets_ids, force_dir_type, force_value,
            None, None, None, None, design_elements,
        )
        loads = assemble_surface_forces(
            base.basis, force_facets_ids=force_facets_ids,
            force_dir_type=base.neumann_dir_type, force_value=base.neumann_values,
        )
        if isinstance(loads, np.ndarray):
            loads = [loads]
        shared = {f.name: getattr(base, f.name) for f in fields(FEMDomain)}
        return cls(**shared, E=E, nu=nu, neumann_linear=loads)

    @classmethod
    def from_mesh_tags(cls, basis, dirichlet_dir, neumann_dir_type,
                       neumann_values, E: float, nu: float) -> 'FEMDomain':
        mesh = basis.mesh
        design_elements = mesh.subdomains["design"]
        keys = mesh.boundaries.keys()

        def numbered(pattern):
            found = [k for k in keys if re.match(pattern, k)]
            return sorted(found, key=lambda x: int(re.search(r"\d+$", x).group()))

        # (sic) the reference matches "dirichlete_<n>" for numbered Dirichlet tags
        dk = numbered(r"dirichlete_\d+$")
        if dk:
            dirichlet_facets_ids = [mesh.boundaries[k] for k in dk]
        elif "dirichlet" in keys:
            dirichlet_facets_ids = mesh.boundaries["dirichlet"]
        else:
            dirichlet_facets_ids = np.array([])
        nk = numbered(r"neumann_\d+$")
        if nk:
            neumann_facets_ids = [mesh.boundaries[k] for k in nk]
        elif "neumann" in keys:
            neumann_facets_ids = [mesh.boundaries["neumann"]]
        else:
            neumann_facets_ids = np.array([])
        return cls.from_facets(
            basis, dirichlet_facets_ids, dirichlet_dir, neumann_facets_ids,
            neumann_dir_type, neumann_values, design_elements, E, nu,
        )
