"""CPU test of the heat task construction (mesh/task_heat.py:60-139,252-322 of the
reference): Robin facet matrices / loads and Dirichlet nodes of the reference's
heat smoke task against the oracle's independent construction."""
import numpy as np
import pytest


@pytest.mark.parametrize("h", [1.0, 0.5])
def test_heat_smoke_task_matches_oracle(h):
    import sktopt
    from oracle import heat as oheat
    from sktopt._fem import Basis, ElementHex1
    x_len, y_len, z_len = 8.0, 8.0, 1.0
    mesh = sktopt.mesh.toy_problem.create_box_hex(x_len, y_len, z_len, h)
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({
        "robin_0": rng((0.0, 0.0), (0.0, y_len), (0.0, z_len)),
        "robin_1": rng((0.0, x_len), (y_len, y_len), (0.0, z_len)),
        "dirichlet_0": rng((x_len - x_len / 20, x_len), (0.0, y_len / 20), (0.0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.arange(mesh.nelements)})
    basis = Basis(mesh, ElementHex1(), intorder=2)
    tsk = sktopt.mesh.LinearHeatConduction.from_mesh_tags(
        basis, 600.0, 4.0e-5, 300.0, True, 10.0, "heat_exchange")
    p, t, Bs, fs, D = oheat.smoke_task_inputs(h)
    assert np.array_equal(mesh.p, p) and np.array_equal(mesh.t, t)
    assert tsk.objective == "heat_exchange" and tsk.k == 10.0
    assert tsk.robin_coefficient == 4.0e-5 and tsk.robin_bc_value == 300.0
    assert len(tsk.robin_bilinear) == 2 and len(tsk.robin_linear) == 2
    for B, Bg, f, fg in zip(Bs, tsk.robin_bilinear, fs, tsk.robin_linear):
        assert abs(B - Bg).max() <= 1e-18 + 1e-13 * abs(B).max()
        assert np.max(np.abs(f - fg)) <= 1e-13 * np.abs(f).max()
        # h * area of the face: 8 x 1
        np.testing.assert_allclose(B.sum(), 4.0e-5 * 8.0, rtol=1e-12)
        np.testing.assert_allclose(np.sum(f), 4.0e-5 * 300.0 * 8.0, rtol=1e-12)
    dn = tsk.dirichlet_nodes
    dn = np.unique(np.concatenate(dn)) if isinstance(dn, list) else np.unique(dn)
    if D.size:
        assert np.array_equal(dn, D)
    else:
        # at h = 1.0 no boundary facet midpoint falls inside the 0.4 x 0.4 patch
        assert dn.size == 0
