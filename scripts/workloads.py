"""BASELINE.json configs 3 and 4 at their named sizes,
built from the package's own task API.  Used by bench.py (``--workload c3|c4``),
scripts/ and the GPU tests; sizes can be scaled down for parity tests.

C3: 55 x 41 x 37 cells of the 8 x 6 x 4 box, Kuhn 6-tet split -> 500,610 tets /
    89,376 nodes / 268,128 DOF; interior nodes jittered by U(-0.2 h, 0.2 h) with
    ``np.random.default_rng(0)``; two load cases as ``toy2`` (u^2 = -1 / +1 on
    two end patches); objective = mean of the two compliances.
C4: 8 x 8 x 1 plate, h = 0.0317 -> 253 x 253 x 32 = 2,048,288 hex / 2,129,028
    scalar DOF; the reference's heat smoke task scaled up
    (scikit-topt/tests/test_global_flow.py:53-103): Dirichlet patch T = 600,
    Robin on the x = 0 and y = 8 faces (h = 4e-5, T_env = 300),
    ``design_robin_boundary=True``, k = 10, objective "compliance".
"""
from __future__ import annotations

import numpy as np

C3_CELLS = (55, 41, 37)
C4_MESH_SIZE = 0.0317


def c3_task(sktopt, cells=C3_CELLS, jitter=0.2, seed=0):
    from sktopt._fem import Basis, ElementTetP1, ElementVector, MeshTet
    x_len, y_len, z_len = 8.0, 6.0, 4.0
    axes = [np.linspace(0, L, n + 1) for L, n in zip((x_len, y_len, z_len), cells)]
    mesh = MeshTet.init_tensor(*axes)
    h = x_len / cells[0]
    p = mesh.p.copy()
    hi = np.array([[x_len], [y_len], [z_len]])
    interior = np.all((p > 1e-9) & (p < hi - 1e-9), axis=0)
    p[:, interior] += np.random.default_rng(seed).uniform(-jitter * h, jitter * h,
                                                          (3, int(interior.sum())))
    mesh = MeshTet(p, sktopt.mesh.utils.fix_tetrahedron_orientation(mesh.t, p))
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({
        "dirichlet": rng((0.0, 0.0), (0.0, y_len), (0.0, z_len)),
        "neumann_0": rng((x_len, x_len), (y_len - 1.4, y_len), (0.0, z_len)),
        "neumann_1": rng((x_len, x_len), (0.0, 1.4), (0.0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.arange(mesh.nelements)})
    basis = Basis(mesh, ElementVector(ElementTetP1()), intorder=2)
    return sktopt.mesh.LinearElasticity.from_mesh_tags(
        basis, "all", ["u^2", "u^2"], [-1.0, 1.0], 210e3, 0.30)


def c4_task(sktopt, mesh_size=C4_MESH_SIZE, intorder=2, design_robin_boundary=True):
    from sktopt._fem import Basis, ElementHex1
    x_len, y_len, z_len = 8.0, 8.0, 1.0
    mesh = sktopt.mesh.toy_problem.create_box_hex(x_len, y_len, z_len, mesh_size)
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({
        "robin_0": rng((0.0, 0.0), (0.0, y_len), (0.0, z_len)),
        "robin_1": rng((0.0, x_len), (y_len, y_len), (0.0, z_len)),
        "dirichlet_0": rng((x_len - 1.0 * x_len / 20, x_len), (0.0, 1.0 * y_len / 20),
                           (0.0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.array(range(mesh.nelements))})
    basis = Basis(mesh, ElementHex1(), intorder=intorder)
    return sktopt.mesh.LinearHeatConduction.from_mesh_tags(
        basis, 600.0, 4.0e-5, 300.0, design_robin_boundary, 10.0, "compliance")
