"""GPU parity tests: every kernel of the hot path, called through the C ABI
(ctypes), against the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: CSR pattern / numbering bit-exact; assembled K
and sensitivities <= 1e-10 relative; compliance <= 1e-6 relative; densities
<= 1e-4 L-inf.
"""
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sktopt
    from sktopt._b200 import device as dev
    return sktopt, dev


def rel_err(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _toy(sktopt, h=1.0):
    tsk = sktopt.mesh.toy_problem.toy_base(h)
    tsk.exlude_dirichlet_from_design()
    return tsk


def _rand_rho(n, seed=0):
    return np.random.default_rng(seed).uniform(0.05, 1.0, n)


# ---------------------------------------------------------------- assembly --
@pytest.mark.parametrize("h", [1.0, 0.45])
def test_pattern_bit_exact_and_K_values(gpu, h):
    sktopt, dev = gpu
    from oracle import fem
    tsk = _toy(sktopt, h)
    p, t = tsk.mesh.p, tsk.mesh.t
    rho = _rand_rho(t.shape[1])
    K = sktopt.fea.composer.assemble_stiffness_matrix(tsk.basis, rho, 210e3, 210.0, 3.0, 0.3)
    Kref = fem.assemble_stiffness(p, t, rho, 210e3, 210.0, 3.0, 0.3, intorder=2)
    assert K.shape == Kref.shape
    assert np.array_equal(K.indptr, Kref.indptr)          # bit-exact pattern
    assert np.array_equal(K.indices, Kref.indices)
    scale = np.abs(Kref.data).max()
    assert np.max(np.abs(K.data - Kref.data)) <= 1e-10 * scale
    # row-wise relative check on the significant entries
    big = np.abs(Kref.data) > 1e-6 * scale
    assert np.max(np.abs(K.data[big] - Kref.data[big]) / np.abs(Kref.data[big])) <= 1e-10


def test_conduction_matrix(gpu):
    sktopt, dev = gpu
    from oracle import fem
    from sktopt._fem import Basis, ElementHex1
    mesh = sktopt.mesh.toy_problem.create_box_hex(2.0, 1.0, 1.0, 0.25)
    for intorder in (1, 2):
        basis = Basis(mesh, ElementHex1(), intorder=intorder)
        rho = _rand_rho(mesh.nelements, 1)
        K = sktopt.fea.composer.assemble_conduction_matrix(basis, rho, 10.0, 1e-2, 3.0)
        k = fem.simp(rho, 10.0, 1e-2, 3.0)
        Kref = fem.assemble_scalar(mesh.p, mesh.t, k, intorder, "laplace")
        assert np.array_equal(K.indptr, Kref.indptr) and np.array_equal(K.indices, Kref.indices)
        assert np.max(np.abs(K.data - Kref.data)) <= 1e-10 * np.abs(Kref.data).max()


def test_tet_assembly_unstructured(gpu):
    sktopt, dev = gpu
    from oracle import fem
    from sktopt._fem import Basis, ElementTetP1, ElementVector, MeshTet
    mesh = sktopt.mesh.toy_problem.create_box_tet(2.0, 1.0, 1.0, 0.34)
    rng = np.random.default_rng(0)
    p = mesh.p.copy()
    interior = np.all((p > 1e-9) & (p < np.array([[2.0], [1.0], [1.0]]) - 1e-9), axis=0)
    p[:, interior] += rng.uniform(-0.05, 0.05, (3, int(interior.sum())))
    mesh = MeshTet(p, mesh.t)
    basis = Basis(mesh, ElementVector(ElementTetP1()), intorder=2)
    rho = _rand_rho(mesh.nelements, 2)
    K = sktopt.fea.composer.assemble_stiffness_matrix(basis, rho, 1.0, 1e-3, 3.0, 0.3)
    Kref = fem.assemble_stiffness(mesh.p, mesh.t, rho, 1.0, 1e-3, 3.0, 0.3, intorder=2)
    assert np.array_equal(K.indptr, Kref.indptr) and np.array_equal(K.indices, Kref.indices)
    assert np.max(np.abs(K.data - Kref.data)) <= 1e-10 * np.abs(Kref.data).max()
    assert dev.device_mesh(mesh).elem_class is None       # per-element classes


# ------------------------------------------------------- solve / compliance --
def test_spmv_matches_scipy(gpu):
    sktopt, dev = gpu
    tsk = _toy(sktopt, 0.45)
    rho = _rand_rho(tsk.mesh.nelements)
    K = sktopt.fea.composer.assemble_stiffness_matrix(tsk.basis, rho, 210e3, 210.0, 3.0, 0.3)
    x = np.random.default_rng(3).standard_normal(K.shape[0])
    y = dev.spmv(dev.to_dev(K.indptr, dev.I32), dev.to_dev(K.indices, dev.I32),
                 dev.to_dev(K.data), dev.to_dev(x), 3).cpu().numpy()
    assert rel_err(y, K @ x) <= 1e-13
    # node-block column indices (the PCG's format for 3 dofs per node)
    dm = dev.device_mesh(tsk.mesh)
    rp, ci = dm.node_graph()
    yb = dev.spmv_bsr3(dev.to_dev(rp, dev.I32), dev.to_dev(ci, dev.I32), dev.to_dev(K.data),
                       dev.to_dev(x)).cpu().numpy()
    assert rel_err(yb, K @ x) <= 1e-13
    yt = dev.spmv_bsr3_tma(dev.to_dev(rp, dev.I32), dev.to_dev(ci, dev.I32), dev.to_dev(K.data),
                           dev.to_dev(x), int(np.diff(rp).max())).cpu().numpy()
    assert rel_err(yt, K @ x) <= 1e-13
    # scalar rows (8 lanes per row pair)
    A = sp.random(3001, 3001, density=0.01, random_state=0, format="csr") + sp.eye(3001)
    A = A.tocsr()
    A.sort_indices()
    x = np.random.default_rng(4).standard_normal(3001)
    y = dev.spmv(dev.to_dev(A.indptr, dev.I32), dev.to_dev(A.indices, dev.I32),
                 dev.to_dev(A.data), dev.to_dev(x), 1).cpu().numpy()
    assert rel_err(y, A @ x) <= 1e-13


@pytest.mark.parametrize("h", [1.0, 0.45])
def test_compliance_and_displacement(gpu, h):
    sktopt, dev = gpu
    from oracle import fem
    tsk = _toy(sktopt, h)
    p, t = tsk.mesh.p, tsk.mesh.t
    rho = _rand_rho(t.shape[1], 5)
    fem_gpu = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
    u = np.zeros((tsk.basis.N, 1))
    c = fem_gpu.objectives_multi_load(rho, 3.0, u)
    c_ref, u_ref = fem.compliance_single(p, t, rho, 210e3, 210.0, 3.0, 0.3,
                                         tsk.neumann_linear[0], tsk.dirichlet_dofs)
    assert abs(c[0] - c_ref) <= 1e-6 * abs(c_ref)
    assert rel_err(u[:, 0], u_ref) <= 1e-6
    assert np.all(u[tsk.dirichlet_dofs, 0] == 0.0)
    # element energies and the identity sum U_e = 1/2 f.u
    U = fem_gpu.energy_multi_load(rho, 3.0, u)
    U_ref = fem.strain_energy(p, t, rho, u, 210e3, 210.0, 3.0, 0.3)
    assert rel_err(U, U_ref) <= 1e-10
    assert abs(U.sum() - 0.5 * c[0]) <= 1e-6 * abs(c[0])


def test_multi_load(gpu):
    sktopt, dev = gpu
    from oracle import fem
    tsk = sktopt.mesh.toy_problem.toy2(0.5)
    tsk.exlude_dirichlet_from_design()
    p, t = tsk.mesh.p, tsk.mesh.t
    rho = _rand_rho(t.shape[1], 6)
    fem_gpu = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3)
    u = np.zeros((tsk.basis.N, 2))
    c = fem_gpu.objectives_multi_load(rho, 2.0, u)
    c_ref, U_ref = fem.compliance_multi(p, t, rho, 210e3, 210.0, 2.0, 0.3,
                                        tsk.neumann_linear, tsk.dirichlet_dofs)
    assert np.max(np.abs(c - c_ref) / np.abs(c_ref)) <= 1e-6
    assert rel_err(u, U_ref) <= 1e-6
    E = fem_gpu.energy_multi_load(rho, 2.0, u)
    assert E.shape == (t.shape[1], 2)
    assert rel_err(E, fem.strain_energy(p, t, rho, u, 210e3, 210.0, 2.0, 0.3)) <= 1e-9


# ------------------------------------------------------------------ filters --
@pytest.mark.parametrize("radius", [0.01, 0.6])
def test_helmholtz_filter(gpu, radius):
    sktopt, dev = gpu
    from oracle.filters import HelmholtzOracle
    tsk = _toy(sktopt, 0.6)
    mesh = tsk.mesh
    f = sktopt.filters.HelmholtzFilterNodal.from_defaults(
        mesh, tsk.elements_volume, radius, design_mask=tsk.design_mask)
    ref = HelmholtzOracle(mesh.p, mesh.t, tsk.elements_volume, tsk.design_mask)
    ref.set_radius(radius)
    rho = _rand_rho(mesh.nelements, 7)
    assert np.max(np.abs(f.forward(rho) - ref.forward(rho))) <= 1e-9
    v = -np.random.default_rng(8).uniform(0.0, 1.0, mesh.nelements)
    g, g_ref = f.gradient(v), ref.gradient(v)
    assert np.max(np.abs(g - g_ref)) <= 1e-9 * max(1.0, np.abs(g_ref).max())
    assert np.all(g <= 0.0)
    # no design mask: pure Neumann problem
    f2 = sktopt.filters.HelmholtzFilterNodal.from_defaults(mesh, tsk.elements_volume, radius)
    ref2 = HelmholtzOracle(mesh.p, mesh.t, tsk.elements_volume, None)
    ref2.set_radius(radius)
    assert np.max(np.abs(f2.forward(rho) - ref2.forward(rho))) <= 1e-9


def test_spatial_filter(gpu):
    sktopt, dev = gpu
    from oracle.filters import SpatialOracle
    tsk = _toy(sktopt, 0.6)
    mesh = tsk.mesh
    f = sktopt.filters.SpacialFilter.from_defaults(
        mesh, tsk.elements_volume, 1.3, design_mask=tsk.design_mask)
    ref = SpatialOracle(mesh.p, mesh.t, tsk.design_mask)
    ref.set_radius(1.3)
    rho = _rand_rho(mesh.nelements, 9)
    assert np.max(np.abs(f.forward(rho) - ref.forward(rho))) <= 1e-12
    v = np.random.default_rng(10).standard_normal(mesh.nelements)
    assert np.max(np.abs(f.gradient(v) - ref.gradient(v))) <= 1e-12


# --------------------------------------------------------- elementwise / K15 --
def test_projection_sensitivity_kernels(gpu):
    sktopt, dev = gpu
    from oracle import optim
    n = 10007
    x = np.random.default_rng(11).uniform(0.0, 1.0, n)
    U = np.random.default_rng(12).uniform(0.0, 5.0, n)
    xd, Ud = dev.to_dev(x), dev.to_dev(U)
    for beta in (1.0, 2.0, 8.0):
        out = torch.empty_like(xd)
        dH = torch.empty_like(xd)
        dev.heaviside(xd, beta, 0.5, out=out, dH=dH)
        assert rel_err(out.cpu().numpy(), optim.heaviside(x, beta, 0.5)) <= 1e-13
        assert rel_err(dH.cpu().numpy(), optim.heaviside_derivative(x, beta, 0.5)) <= 1e-13
    g = dev.dc_drho(xd, Ud, 210e3, 210.0, 3.0).cpu().numpy()
    assert rel_err(g, optim.dC_drho_simp(x, U, 210e3, 210.0, 3.0)) <= 1e-12
    E = dev.interpolate_modulus(xd, 210e3, 210.0, 3.0).cpu().numpy()
    assert rel_err(E, 210.0 + (210e3 - 210.0) * x ** 3.0) <= 1e-14


@pytest.mark.parametrize("n", [1, 2, 17, 4096, 100003])
def test_abs_percentile_matches_numpy(gpu, n):
    sktopt, dev = gpu
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 3, n)
    a[rng.integers(0, n, max(1, n // 7))] = a[0]          # duplicates
    ad = dev.to_dev(a)
    for q in (0.0, 5.0, 50.0, 95.0, 99.9, 100.0):
        assert dev.abs_percentile(ad, q) == np.percentile(np.abs(a), q)
    assert dev.reduce_absmax(ad) == np.max(np.abs(a))
    mn, mean, mx, sd = dev.reduce_stats(ad)
    assert mn == a.min() and mx == a.max()
    assert abs(mean - a.mean()) <= 1e-12 * np.abs(a).max()
    assert abs(sd - a.std()) <= 1e-10 * max(a.std(), 1e-300)


def test_update_kernels(gpu):
    sktopt, dev = gpu
    from oracle import optim
    n = 5000
    rng = np.random.default_rng(13)
    rho = rng.uniform(0.01, 1.0, n)
    dL = rng.standard_normal(n) * 2.0
    rd = dev.to_dev(rho.copy())
    s, lo, hi = (torch.empty(n, dtype=dev.F64, device="cuda") for _ in range(3))
    dev.logmoc_update(rd, dev.to_dev(dL), 0.6, 0.2, 1e-2, 1.0, 1.0, s, lo, hi)
    assert rel_err(rd.cpu().numpy(), optim.logmoc_step(rho, dL, 0.6, 0.2, 1e-2, 1.0, 1.0)) <= 1e-13
    dC = -rng.uniform(0.0, 1.0, n)
    sr, cand = torch.empty_like(s), torch.empty_like(s)
    dev.oc_candidate(dev.to_dev(dC), dev.to_dev(rho), 0.37, 1e-12, 0.5, 0.2, 1e-2, 1.0,
                     0.7, 1.3, None, sr, cand, None)
    sr_ref = np.clip((-dC / (0.37 + 1e-12)) ** 0.5, 0.7, 1.3)
    cand_ref = np.clip(rho * sr_ref, np.maximum(rho - 0.2, 1e-2), np.minimum(rho + 0.2, 1.0))
    assert rel_err(sr.cpu().numpy(), sr_ref) <= 1e-14
    assert rel_err(cand.cpu().numpy(), cand_ref) <= 1e-14


# ------------------------------------------------------------ optimiser loop --
def _run_gpu(sktopt, kind, tsk, max_iters, **cfg_kw):
    with tempfile.TemporaryDirectory() as tmp:
        if kind == "oc":
            cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=max_iters,
                                        record_times=max_iters, **cfg_kw)
            opt = sktopt.core.OC_Optimizer(cfg, tsk)
        else:
            cfg = sktopt.core.LogMOC_Config(dst_path=tmp, max_iters=max_iters,
                                            record_times=max_iters, **cfg_kw)
            opt = sktopt.core.LogMOC_Optimizer(cfg, tsk)
        opt.parameterize()
        opt.optimize()
        hist = opt.recorder.as_object()
        return (np.asarray(hist.compliance), opt._state.rho.cpu().numpy(),
                np.asarray(hist.vol_error), opt)


def test_oc_loop_matches_oracle(gpu, toy_oracle):
    sktopt, dev = gpu
    from oracle import optim
    o, pr = toy_oracle
    comp, rho, vol_err, opt = _run_gpu(sktopt, "oc", sktopt.mesh.toy_problem.toy_test(), 8)
    ref = optim.run(pr, "oc", max_iters=8)
    assert np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])) <= 1e-6
    assert np.max(np.abs(rho - ref["rho_final"])) <= 1e-4
    assert opt.bisection_steps == ref["bisection_steps"]
    assert np.max(np.abs(vol_err - ref["vol_error"])) <= 1e-7


def test_logmoc_loop_matches_oracle(gpu, toy_oracle):
    sktopt, dev = gpu
    from oracle import optim
    o, pr = toy_oracle
    vf = sktopt.tools.SchedulerConfig.constant(target_value=0.6)
    comp, rho, vol_err, _ = _run_gpu(sktopt, "logmoc", sktopt.mesh.toy_problem.toy_test(),
                                     8, vol_frac=vf)
    ref = optim.run(pr, "logmoc", max_iters=8, vol_frac=0.6)
    assert np.max(np.abs(comp - ref["compliance"]) / np.abs(ref["compliance"])) <= 1e-6
    assert np.max(np.abs(rho - ref["rho_final"])) <= 1e-4
    assert np.max(np.abs(vol_err - ref["vol_error"])) <= 1e-7


def test_optimize_full_vs_steps_bit_identical(gpu):
    """Reference tests/test_optimize_steps_equivalence.py: atol=0, rtol=0."""
    sktopt, dev = gpu
    _, rho_full, _, o1 = _run_gpu(sktopt, "oc", sktopt.mesh.toy_problem.toy_test(), 5)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = sktopt.core.OC_Config(dst_path=tmp, max_iters=5, record_times=5)
        opt = sktopt.core.OC_Optimizer(cfg, sktopt.mesh.toy_problem.toy_test())
        opt.parameterize()
        for _ in range(5):
            opt.optimize_steps(1)
        rho_step = opt._state.rho.cpu().numpy()
        c_step = np.asarray(opt.recorder.as_object().compliance)
        data = np.load(os.path.join(tmp, "data", "000005-rho.npz"))["rho_design_elements"]
    np.testing.assert_allclose(rho_full, rho_step, rtol=0.0, atol=0.0)
    np.testing.assert_allclose(np.asarray(o1.recorder.as_object().compliance), c_step)
    assert data.shape == (opt.tsk.design_elements.size,)


def test_reference_smoke_cases(gpu):
    """Reference tests/test_global_flow.py: finite compliance after one
    iteration of OC on toy_test and LogMOC on the two-load toy2."""
    sktopt, dev = gpu
    comp, _, _, _ = _run_gpu(sktopt, "oc", sktopt.mesh.toy_problem.toy_test(), 1)
    assert np.isfinite(comp[-1])
    cfgkw = dict(
        p=sktopt.tools.SchedulerConfig(init_value=1.0, target_value=3.0, num_steps=3,
                                        scheduler_type="Step"),
        vol_frac=sktopt.tools.SchedulerConfig(target_value=0.6, scheduler_type="Step"),
    )
    comp, _, _, _ = _run_gpu(sktopt, "logmoc", sktopt.mesh.toy_problem.toy2(0.5), 1, **cfgkw)
    assert np.isfinite(comp[-1])
