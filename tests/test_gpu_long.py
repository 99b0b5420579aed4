"""Longer GPU parity runs: the BASELINE tolerance 'densities after 50
iterations <= 1e-4 L-inf, per-iteration compliance <= 1e-6 relative' on a
mid-size cantilever, and the two-load tetrahedral case (BASELINE config 3
scaled down: Kuhn tets with jittered interior nodes, toy2-style +/-1 loads)."""
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sktopt
    from sktopt._b200 import device as dev
    return sktopt, dev


def test_oc_50_iterations_match_oracle(gpu):
    sktopt, dev = gpu
    from oracle import mesh as omesh, optim
    h = 0.5                                   # 16 x 12 x 8 = 1536 hex
    o = omesh.toy_base(h)
    pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                       o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    ref = optim.run(pr, "oc", max_iters=50, filter_radius=0.6, vol_frac=0.4)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(
            dst_path=tmp, max_iters=50, record_times=50, solver_option="cg_pyamg",
            vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.4),
            filter_radius=sktopt.tools.SchedulerConfig.constant(target_value=0.6))
        opt = sktopt.core.OC_Optimizer(cfg, sktopt.mesh.toy_problem.toy_base(h))
        opt.parameterize()
        opt.optimize()
        comp = np.asarray(opt.recorder.as_object().compliance)
        rho = opt._state.rho.cpu().numpy()
    assert comp.size == 50
    assert np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])) <= 1e-6
    assert np.max(np.abs(rho - ref["rho_final"])) <= 1e-4
    assert opt.bisection_steps == ref["bisection_steps"]


def _tet_task(sktopt, cells=(12, 9, 8), jitter=0.2, seed=0):
    """8 x 6 x 4 box of Kuhn tets, interior nodes jittered by U(-j h, j h),
    clamped at x = 0, two load cases -/+1 in u^2 on two end patches."""
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


def test_two_load_tets_match_oracle(gpu):
    sktopt, dev = gpu
    from oracle import optim
    tsk = _tet_task(sktopt)
    tsk_ref = _tet_task(sktopt)
    tsk_ref.exlude_dirichlet_from_design()
    assert tsk.n_tasks == 2 and tsk.mesh.t.shape[0] == 4
    assert np.all(tsk.elements_volume > 0)
    pr = optim.Problem(tsk_ref.mesh.p, tsk_ref.mesh.t, tsk_ref.dirichlet_dofs,
                       list(tsk_ref.neumann_linear), tsk_ref.design_elements,
                       tsk_ref.dirichlet_neumann_elements, tsk_ref.elements_volume,
                       tsk_ref.E, tsk_ref.nu, fixed=tsk_ref.fixed_elements)
    ref = optim.run(pr, "logmoc", max_iters=6, vol_frac=0.5, filter_radius=0.5)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.LogMOC_Config(
            dst_path=tmp, max_iters=6, record_times=6,
            vol_frac=sktopt.tools.SchedulerConfig.constant(target_value=0.5),
            filter_radius=sktopt.tools.SchedulerConfig.constant(target_value=0.5))
        opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
        opt.parameterize()
        opt.optimize()
        comp = np.asarray(opt.recorder.as_object().compliance)
        rho = opt._state.rho.cpu().numpy()
        u_max = np.asarray(opt.recorder.as_object().u_max)
    # objective = mean of the two compliances (common_density.py:1060)
    assert np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])) <= 1e-6
    assert np.max(np.abs(rho - ref["rho_final"])) <= 1e-4
    assert u_max.shape == (6, 4)              # min / mean / max / std over the two loads
