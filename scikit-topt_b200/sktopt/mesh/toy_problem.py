"""Box meshes and the toy cantilever / multi-load tasks
(reference ``mesh/toy_problem.py:10-144``)."""
from __future__ import annotations

import numpy as np

from sktopt._fem import Basis, ElementHex1, ElementTetP1, ElementVector, MeshHex, MeshTet
from sktopt.mesh import task_elastic, utils


def _box_axes(x_len, y_len, z_len, mesh_size):
    counts = [int(np.ceil(L / mesh_size)) for L in (x_len, y_len, z_len)]
    return [np.linspace(0, L, n + 1) for L, n in zip((x_len, y_len, z_len), counts)]


def create_box_hex(x_len, y_len, z_len, mesh_size):
    """Tensor hexahedral box with ceil(L/h) cells per axis (``:10-37``)."""
    x, y, z = _box_axes(x_len, y_len, z_len, mesh_size)
    mesh = MeshHex.init_tensor(x, y, z)
    return MeshHex(mesh.p, utils.fix_hexahedron_orientation(mesh.t, mesh.p))


def create_box_tet(x_len, y_len, z_len, mesh_size):
    """Tetrahedral box.  The reference refines skfem's unit ``MeshTet()``
    (``:40-48``), which cannot be reproduced without skfem; here each cell of the
    tensor grid is split into 6 Kuhn tetrahedra instead."""
    x, y, z = _box_axes(x_len, y_len, z_len, mesh_size)
    mesh = MeshTet.init_tensor(x, y, z)
    return MeshTet(mesh.p, utils.fix_tetrahedron_orientation(mesh.t, mesh.p))


def toy_base(mesh_size: float, intorder: int = 2):
    """8 x 6 x 4 cantilever: clamped x=0 face, -100 in u^3 over the facets whose
    midpoint lies in [6.8, 8.1] x [2.4, 3.6] x [2.8, 4] (``:51-91``).  As in the
    reference the quadrature order is fixed to 2 regardless of ``intorder``."""
    x_len, y_len, z_len, eps = 8.0, 6.0, 4.0, 1.2
    mesh = create_box_hex(x_len, y_len, z_len, mesh_size)
    clamp = utils.get_points_in_range((0.0, 0.03), (0.0, y_len), (0.0, z_len))
    load = utils.get_points_in_range(
        (x_len - eps, x_len + 0.1), (y_len * 2 / 5, y_len * 3 / 5), (z_len - eps, z_len))
    everywhere = utils.get_points_in_range((0.0, x_len), (0.0, y_len), (0.0, z_len))
    basis = Basis(mesh, ElementVector(ElementHex1()), intorder=2)
    return task_elastic.LinearElasticity.from_facets(
        basis,
        mesh.facets_satisfying(clamp),
        "all",
        mesh.facets_satisfying(load),
        "u^3",
        -100.0,
        mesh.elements_satisfying(everywhere),
        210e3,
        0.30,
    )


def toy_test():
    return toy_base(1.0)


def toy1():
    return toy_base(0.3)


def toy1_fine():
    return toy_base(0.2)


def toy2(mesh_size: float = 0.3):
    """8 x 8 x 1 plate, two load cases (-1 / +1 in u^2 on two end patches)
    defined through mesh tags (``:106-144``)."""
    x_len, y_len, z_len = 8.0, 8.0, 1.0
    mesh = create_box_hex(x_len, y_len, z_len, mesh_size)
    eps = mesh_size
    mesh = mesh.with_boundaries({
        "dirichlet": utils.get_points_in_range((0.0, 0.05), (0.0, y_len), (0.0, z_len)),
        "neumann_0": utils.get_points_in_range((x_len, x_len), (y_len - eps, y_len), (0, z_len)),
        "neumann_1": utils.get_points_in_range((x_len, x_len), (0, eps), (0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.array(range(mesh.nelements))})
    basis = Basis(mesh, ElementVector(ElementHex1()), intorder=2)
    return task_elastic.LinearElasticity.from_mesh_tags(
        basis, "all", ["u^2", "u^2"], [-1.0, 1.0], 210e3, 0.30)


def toy_msh(task_elastic_name: str = "down", msh_path: str = 'plate.msh'):
    raise NotImplementedError(
        "reading .msh files needs meshio, which is outside the B200 hot path")
