"""GPU parity of the heat-conduction path (BASELINE config 4
scaled down): conduction + real Robin + virtual Robin assembly, enforce with
non-zero Dirichlet values, J = T^T K T, element energy and the compliance
sensitivity including the explicit Robin term, at intorder 1 and 2."""
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


def heat_task(sktopt, mesh_size=0.5, intorder=2, design_robin_boundary=True):
    """The reference's heat smoke task (tests/test_global_flow.py:53-103)."""
    from sktopt._fem import Basis, ElementHex1
    x_len, y_len, z_len = 8.0, 8.0, 1.0
    mesh = sktopt.mesh.toy_problem.create_box_hex(x_len, y_len, z_len, mesh_size)
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({
        "robin_0": rng((0.0, 0.0), (0.0, y_len), (0.0, z_len)),
        "robin_1": rng((0.0, x_len), (y_len, y_len), (0.0, z_len)),
        "dirichlet_0": rng((x_len - 1.0 * x_len / 20, x_len), (0.0, 1.0 * y_len / 20), (0.0, z_len)),
    })
    mesh = mesh.with_subdomains({"design": np.array(range(mesh.nelements))})
    basis = Basis(mesh, ElementHex1(), intorder=intorder)
    return sktopt.mesh.LinearHeatConduction.from_mesh_tags(
        basis, 600.0, 4.0e-5, 300.0, design_robin_boundary, 10.0, "compliance")


def oracle_inputs(tsk):
    """Robin facet terms for the oracle, built independently from the mesh."""
    from oracle import heat as oheat, mesh as omesh
    p, t = tsk.mesh.p, tsk.mesh.t
    srt, cyc = omesh.hex_facets(t)
    # boundary facets = facets owned by exactly one element
    allf = np.sort(np.hstack([t[list(f)] for f in omesh._HEX_FACES]).astype(np.int64), axis=0)
    _, inv, cnt = np.unique(allf, axis=1, return_inverse=True, return_counts=True)
    mid = p[:, srt].mean(axis=1)
    on_bnd = cnt == 1
    Bs, fs = [], []
    for sel in (omesh.in_box(mid, (0.0, 0.0), (0.0, 8.0), (0.0, 1.0)),
                omesh.in_box(mid, (0.0, 8.0), (8.0, 8.0), (0.0, 1.0))):
        ids = np.nonzero(sel & on_bnd)[0]
        Bs.append(oheat.quad_facet_mass(p, cyc[:, ids], 4.0e-5))
        f, _ = omesh.quad_facet_load(p, cyc[:, ids], 4.0e-5 * 300.0)
        fs.append(f)
    dsel = np.nonzero(omesh.in_box(mid, (7.6, 8.0), (0.0, 0.4), (0.0, 1.0)) & on_bnd)[0]
    D = np.unique(srt[:, dsel])
    return Bs, fs, D


@pytest.mark.parametrize("intorder", [2, 1])
def test_heat_objective_and_sensitivity(gpu, intorder):
    sktopt, dev = gpu
    from oracle import heat as oheat
    tsk = heat_task(sktopt, 0.5, intorder)
    p, t = tsk.mesh.p, tsk.mesh.t
    Bs, fs, D = oracle_inputs(tsk)
    assert np.array_equal(D, np.unique(np.concatenate(tsk.dirichlet_nodes)))
    for B, Bg in zip(Bs, tsk.robin_bilinear):
        assert abs(B - Bg).max() <= 1e-18 + 1e-13 * abs(B).max()
    rho = np.random.default_rng(0).uniform(0.1, 0.95, t.shape[1])
    fem_gpu = sktopt.fea.FEM_SimpLinearHeatConduction(tsk, 1e-3)
    T = np.zeros((tsk.basis.N, 1))
    J = fem_gpu.objectives_multi_load(rho, 3.0, T)
    J_ref, T_ref, K_ref = oheat.solve_compliance(
        p, t, rho, 10.0, 1e-2, 3.0, 4, 4.0e-5, 300.0, Bs, fs, D, 600.0, intorder)
    # assembled total matrix (conduction + Robin + virtual Robin)
    eng = fem_gpu.engine
    K_gpu = sktopt.fea.composer._csr_to_scipy(eng.n_dof, eng.row_ptr, eng.col_idx, eng.vals)
    assert abs(K_gpu - K_ref).max() <= 1e-10 * abs(K_ref).max()
    if intorder == 2:
        # (one-point quadrature gives an hourglass-singular conduction matrix:
        # only the assembled operators are compared at intorder 1)
        assert abs(J[0] - J_ref) <= 1e-6 * abs(J_ref)
        assert np.max(np.abs(T[:, 0] - T_ref)) <= 1e-6 * np.abs(T_ref).max()
        assert np.all(T[D, 0] == 600.0)
        assert np.allclose(fem_gpu.λ_all, -2.0 * T)
    Tin = T_ref[:, None].copy()
    U = fem_gpu.energy_multi_load(rho, 3.0, Tin)
    g = fem_gpu.compliance_sensitivity_multi_load(rho, 3.0, Tin)
    g_ref, U_ref = oheat.sensitivity(p, t, rho, T_ref, 10.0, 1e-2, 3.0, 4, 4.0e-5, 300.0, intorder)
    assert np.max(np.abs(U[:, 0] - U_ref)) <= 1e-10 * np.abs(U_ref).max()
    assert np.max(np.abs(g[:, 0] - g_ref)) <= 1e-10 * np.abs(g_ref).max()


def test_heat_without_robin_is_uniform(gpu):
    """no Robin facets -> T = T_D everywhere, J = 0."""
    sktopt, dev = gpu
    from sktopt._fem import Basis, ElementHex1
    mesh = sktopt.mesh.toy_problem.create_box_hex(2.0, 1.0, 1.0, 0.25)
    rng = sktopt.mesh.utils.get_points_in_range
    mesh = mesh.with_boundaries({"dirichlet_0": rng((0.0, 0.0), (0.0, 1.0), (0.0, 1.0))})
    mesh = mesh.with_subdomains({"design": np.arange(mesh.nelements)})
    tsk = sktopt.mesh.LinearHeatConduction.from_mesh_tags(
        Basis(mesh, ElementHex1(), intorder=2), 350.0, None, None, None, 10.0, "compliance")
    fem_gpu = sktopt.fea.FEM_SimpLinearHeatConduction(tsk, 1e-3)
    T = np.zeros((tsk.basis.N, 1))
    J = fem_gpu.objectives_multi_load(np.full(mesh.nelements, 0.5), 3.0, T)
    assert np.max(np.abs(T - 350.0)) <= 1e-6
    assert abs(J[0]) <= 1e-6


def test_heat_oc_smoke_like_reference(gpu):
    """Reference tests/test_global_flow.py:154-157: one OC iteration on the heat
    task (intorder=1, design-dependent Robin) gives a finite objective."""
    sktopt, dev = gpu
    tsk = heat_task(sktopt, 0.5, 1)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=1, record_times=1)
        opt = sktopt.core.OC_Optimizer(cfg, tsk)
        opt.parameterize()
        opt.optimize()
        res = opt.recorder.as_object_latest()
    assert np.isfinite(res.compliance)


def test_heat_oc_loop_runs_intorder2(gpu):
    sktopt, dev = gpu
    tsk = heat_task(sktopt, 0.5, 2)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=3, record_times=3)
        opt = sktopt.core.OC_Optimizer(cfg, tsk)
        opt.parameterize()
        opt.optimize()
        comp = np.asarray(opt.recorder.as_object().compliance)
        vol_err = np.asarray(opt.recorder.as_object().vol_error)
    assert np.all(np.isfinite(comp)) and comp.size == 3
    assert np.all(np.abs(vol_err) < 1e-3)


def test_unbuilt_objectives_raise(gpu):
    sktopt, dev = gpu
    tsk = heat_task(sktopt, 1.0, 2)
    tsk.objective = "heat_exchange"
    fem_gpu = sktopt.fea.FEM_SimpLinearHeatConduction(tsk, 1e-3)
    with pytest.raises(NotImplementedError):
        fem_gpu.objectives_multi_load(np.full(tsk.mesh.nelements, 0.5), 3.0,
                                      np.zeros((tsk.basis.N, 1)))
