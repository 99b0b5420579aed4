"""Geometric multigrid preconditioner: Galerkin coarse operators equal
P^T A P, the V-cycle is a symmetric operator, and MG-PCG converges to the same
solution as Jacobi-PCG / the direct solver in far fewer iterations."""
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


def _engine(sktopt, dims=(2.8, 2.0, 1.6), h=0.4, monkey=None):
    from sktopt._fem import Basis, ElementHex1, ElementVector
    from sktopt.fea._engine import FeaEngine, KE_ELASTIC
    mesh = sktopt.mesh.toy_problem.create_box_hex(*dims, h)       # 7 x 5 x 4 cells (odd sizes)
    basis = Basis(mesh, ElementVector(ElementHex1()), intorder=2)
    clamp = np.nonzero(mesh.p[0] == 0.0)[0]
    D = np.unique((3 * clamp[:, None] + np.arange(3)[None, :]).ravel())
    return mesh, basis, D, FeaEngine(basis, D, KE_ELASTIC, 0.3)


def _prolongation(mg, level):
    """scipy P (fine nodes x coarse nodes) from the axis tables, per component."""
    from sktopt.fea._multigrid import axis_tables
    fine_cells = tuple(c.size - 1 for c in mg.coords[level])
    mats = []
    for n in fine_cells:
        c0, c1, w0, w1, _, _ = axis_tables(n)
        nc = (n + 1) // 2
        P = sp.lil_matrix((n + 1, nc + 1))
        for i in range(n + 1):
            P[i, c0[i]] += w0[i]
            P[i, c1[i]] += w1[i]
        mats.append(P.tocsr())
    Px, Py, Pz = mats
    # node = iy + npy*ix + npy*npx*iz  ->  kron(Pz, kron(Px, Py))
    Pn = sp.kron(Pz, sp.kron(Px, Py)).tocsr()
    return sp.kron(Pn, sp.eye(3)).tocsr()


def test_galerkin_coarse_operator_equals_PtAP(gpu, monkeypatch):
    sktopt, dev = gpu
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    mesh, basis, D, eng = _engine(sktopt)
    assert eng.precond == "mg" and eng.mg.n_levels >= 2
    rho = np.random.default_rng(0).uniform(0.05, 1.0, mesh.nelements)
    eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
    eng.assemble(enforce=False)
    rp, ci = eng.dm.dof_pattern(3)
    K = sktopt.fea.composer._csr_to_scipy(eng.n_dof, rp, ci, eng.vals)
    mg = eng.mg
    lv = mg.levels[1]
    from sktopt._b200 import lib as _lib
    _lib.check(mg.lib.sktb_elem_restrict(
        lv["n_elem"], dev._ptr(lv["child"]), dev._ptr(lv["ptype"]), dev._ptr(mg.Qtab),
        None, dev._ptr(eng.unit_ke), dev._ptr(eng.dm.elem_class), dev._ptr(eng.scale),
        dev._ptr(lv["ke"]), dev._stream()))
    vals = lv["dm"].assemble(3, lv["ke"], scale=None, dir_mask=None, per_element=True)
    rpc, cic = lv["dm"].dof_pattern(3)
    Kc = sktopt.fea.composer._csr_to_scipy(3 * lv["n_nodes"], rpc, cic, vals)
    P = _prolongation(mg, 0)
    ref = (P.T @ K @ P).tocsr()
    assert abs(Kc - ref).max() <= 1e-10 * abs(ref).max()
    # uniform modulus: the Galerkin element matrix of a full 2x2x2 parent is the
    # geometric element matrix of the coarse element
    eng.set_modulus(dev.to_dev(np.ones(mesh.nelements)), 1.0, 0.0, 1.0)
    _lib.check(mg.lib.sktb_elem_restrict(
        lv["n_elem"], dev._ptr(lv["child"]), dev._ptr(lv["ptype"]), dev._ptr(mg.Qtab),
        None, dev._ptr(eng.unit_ke), dev._ptr(eng.dm.elem_class), dev._ptr(eng.scale),
        dev._ptr(lv["ke"]), dev._stream()))
    geo = lv["dm"].unit_ke(0, basis.X, basis.W, nu=0.3)
    cls = lv["dm"].elem_class.cpu().numpy()
    full = np.nonzero(lv["ptype"].cpu().numpy() == 0)[0]
    got = lv["ke"].cpu().numpy()[full].reshape(-1, 24, 24)
    want = geo.cpu().numpy()[cls[full]]
    assert np.max(np.abs(got - want)) <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("fp32", ["0", "1"])
def test_vcycle_is_symmetric_and_mg_pcg_converges(gpu, monkeypatch, fp32):
    sktopt, dev = gpu
    from oracle import fem
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    # level-0 products of the V-cycle in fp64 / fp32 (the operator stays
    # symmetric to rounding: 1e-10 vs 1e-5)
    monkeypatch.setenv("SKTOPT_B200_MG_FP32", fp32)
    sym_tol = 1e-10 if fp32 == "0" else 2e-5
    mesh, basis, D, eng = _engine(sktopt, dims=(4.0, 3.0, 2.0), h=0.25)   # 16 x 12 x 8
    rho = np.random.default_rng(1).uniform(0.01, 1.0, mesh.nelements)
    eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
    eng.assemble(enforce=True)
    eng.update_preconditioner()
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(eng.n_dof), rng.standard_normal(eng.n_dof)
    a[D] = 0.0
    b[D] = 0.0
    Ma = eng.mg.vcycle(dev.to_dev(a)).cpu().numpy()
    Mb = eng.mg.vcycle(dev.to_dev(b)).cpu().numpy()
    assert abs(Ma @ b - a @ Mb) <= sym_tol * abs(Ma @ b)
    assert a @ Ma > 0 and b @ Mb > 0
    assert np.all(Ma[D] == 0.0)
    # solve with MG-PCG and with Jacobi-PCG
    f = np.zeros(eng.n_dof)
    tip = np.nonzero(mesh.p[0] == mesh.p[0].max())[0]
    f[3 * tip + 2] = -1.0
    f[D] = 0.0
    fd = dev.to_dev(f)
    eng.warm_start = False
    u_mg = eng.solve(fd, 0, 1e-9, None).cpu().numpy().copy()
    it_mg = eng.pcg_log[-1][0]
    eng.mg_enabled = False
    u_j = eng.solve(fd, 1, 1e-9, None).cpu().numpy().copy()
    it_j = eng.pcg_log[-1][0]
    K = fem.assemble_stiffness(mesh.p, mesh.t, rho, 210e3, 210.0, 3.0, 0.3)
    K_e, f_e = fem.enforce(K, f, D)
    u_ref, _, _ = fem.solve(K_e, f_e, "spsolve")
    assert eng.pcg_log[-2][1] and eng.pcg_log[-1][1]
    assert np.max(np.abs(u_mg - u_ref)) <= 1e-6 * np.abs(u_ref).max()
    assert np.max(np.abs(u_j - u_ref)) <= 1e-6 * np.abs(u_ref).max()
    print("iterations: multigrid", it_mg, "jacobi", it_j)
    assert it_mg * 4 < it_j


def test_selector_cg_jacobi_disables_multigrid(gpu):
    sktopt, dev = gpu
    tsk = sktopt.mesh.toy_problem.toy_base(0.45)
    rho = np.full(tsk.mesh.nelements, 0.5)
    u = np.zeros((tsk.basis.N, 1))
    f_mg = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
    c_mg = f_mg.objectives_multi_load(rho, 3.0, u)
    assert f_mg.engine.precond == "mg"
    it_mg = f_mg.engine.pcg_log[-1][0]
    f_j = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_jacobi")
    f_j.engine.warm_start = False
    c_j = f_j.objectives_multi_load(rho, 3.0, u)
    it_j = f_j.engine.pcg_log[-1][0]
    assert abs(c_mg[0] - c_j[0]) <= 1e-7 * abs(c_j[0])
    assert it_mg < it_j


def test_fused_tail_matches_per_level_launches(gpu, monkeypatch):
    """The cooperative coarse-tail kernel applies the same V-cycle as the
    launch-per-operation path (only the summation order inside a row differs)."""
    sktopt, dev = gpu
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    monkeypatch.setenv("SKTOPT_B200_MG_FP32", "0")
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SKTOPT_B200_MG_FUSED_TAIL", flag)
        mesh, basis, D, eng = _engine(sktopt, dims=(4.0, 3.0, 2.0), h=0.25)
        rho = np.random.default_rng(1).uniform(0.01, 1.0, mesh.nelements)
        eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
        eng.prepare()
        assert eng.mg.n_levels >= 3
        a = np.random.default_rng(2).standard_normal(eng.n_dof)
        a[D] = 0.0
        out[flag] = eng.mg.vcycle(dev.to_dev(a)).cpu().numpy()
    assert np.max(np.abs(out["1"] - out["0"])) <= 1e-11 * np.max(np.abs(out["0"]))


def test_projected_start_vector(gpu, monkeypatch):
    """Galerkin projection of the new system onto the last solutions: the same
    converged displacement as the plain warm start, in no more PCG iterations,
    over a sequence of slowly changing density fields."""
    sktopt, dev = gpu
    from oracle import fem
    rng = np.random.default_rng(3)
    out = {}
    for hist in ("1", "3"):
        monkeypatch.setenv("SKTOPT_B200_START_HIST", hist)
        tsk = sktopt.mesh.toy_problem.toy_base(0.45)
        f = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
        rho = np.full(tsk.mesh.nelements, 0.5)
        noise = np.random.default_rng(4).uniform(-1.0, 1.0, rho.size)
        u = np.zeros((tsk.basis.N, 1))
        its, comp = [], []
        for k in range(5):
            c = f.objectives_multi_load(np.clip(rho + 0.03 * k * noise, 0.05, 1.0), 3.0, u)
            its.append(f.engine.pcg_log[-1][0])
            comp.append(c[0])
            assert f.engine.pcg_log[-1][1]
        assert f.engine.start_hist == int(hist)
        out[hist] = (its, comp, u[:, 0].copy(), tsk, np.clip(rho + 0.03 * 4 * noise, 0.05, 1.0))
    print("PCG iterations: warm start", out["1"][0], "projected start", out["3"][0])
    assert np.max(np.abs(np.array(out["1"][1]) / np.array(out["3"][1]) - 1.0)) <= 1e-7
    assert sum(out["3"][0][1:]) <= sum(out["1"][0][1:])
    tsk, rho4 = out["3"][3], out["3"][4]
    K = fem.assemble_stiffness(tsk.mesh.p, tsk.mesh.t, rho4, tsk.E, tsk.E * 1e-3, 3.0, tsk.nu)
    fl = tsk.neumann_linear if isinstance(tsk.neumann_linear, list) else [tsk.neumann_linear]
    F = np.asarray(fl[0], dtype=float).copy()
    K_e, f_e = fem.enforce(K, F, tsk.dirichlet_dofs)
    u_ref, _, _ = fem.solve(K_e, f_e, "spsolve")
    assert np.max(np.abs(out["3"][2] - u_ref)) <= 1e-6 * np.abs(u_ref).max()


def test_fused_jacobi_sweeps_match_separate_kernels(gpu, monkeypatch):
    """Coarse-level sweeps fused into the SpMV epilogue (one kernel, ping-pong
    iterates) apply the same V-cycle as product + update launches."""
    sktopt, dev = gpu
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    monkeypatch.setenv("SKTOPT_B200_MG_FP32", "0")
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SKTB_MG_FUSED_SWEEPS", flag)
        mesh, basis, D, eng = _engine(sktopt, dims=(8.0, 4.0, 2.0), h=0.1)   # level 1: 9471 nodes (bulk-async kernel)
        rho = np.random.default_rng(1).uniform(0.01, 1.0, mesh.nelements)
        eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
        eng.prepare()
        assert eng.mg.n_levels >= 4
        a = np.random.default_rng(2).standard_normal(eng.n_dof)
        a[D] = 0.0
        ad = dev.to_dev(a)
        z1 = eng.mg.vcycle(ad).cpu().numpy()
        z2 = eng.mg.vcycle(ad).cpu().numpy()       # buffers swapped an odd/even number of times
        assert np.array_equal(z1, z2)
        out[flag] = z1
    assert np.max(np.abs(out["1"] - out["0"])) <= 1e-12 * np.max(np.abs(out["0"]))


def test_multigrid_pcg_at_late_stage_contrast(gpu):
    """Harder than anything a filtered run produces: p = 3 on a SHARP 0/1 truss with
    three-element webs (modulus contrast 1e3, no filter, cold start).  The
    damped-Jacobi-smoothed geometric multigrid degrades here (82 iterations measured,
    against 18-20 warm iterations in the late stage of the real C2 run, bench.py
    `late_stage`) but still converges, 18x faster than Jacobi-PCG (1490), to the
    same solution."""
    sktopt, dev = gpu
    tsk = sktopt.mesh.toy_problem.toy_base(0.2)
    tsk.exlude_dirichlet_from_design()
    ne = tsk.mesh.nelements
    cen = np.mean(tsk.mesh.p[:, tsk.mesh.t], axis=1)
    # a crude "truss": solid skins and diagonal webs, void elsewhere
    solid = ((cen[2] < 0.6) | (cen[2] > 3.4) | (np.abs((cen[0] % 2.0) - cen[2] / 2.0) < 0.3))
    rho = np.where(solid, 1.0, 0.01)
    rho[np.random.default_rng(0).uniform(size=ne) < 0.02] = 0.5
    u_mg = np.zeros((tsk.basis.N, 1))
    f_mg = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_pyamg")
    c_mg = f_mg.objectives_multi_load(rho, 3.0, u_mg)
    it_mg = f_mg.engine.pcg_log[-1][0]
    assert f_mg.engine.precond == "mg" and f_mg.engine.pcg_log[-1][1]
    f_j = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3, solver_option="cg_jacobi")
    f_j.engine.warm_start = False
    u_j = np.zeros_like(u_mg)
    c_j = f_j.objectives_multi_load(rho, 3.0, u_j)
    it_j = f_j.engine.pcg_log[-1][0]
    print("late-stage contrast: MG-PCG", it_mg, "Jacobi-PCG", it_j)
    assert it_mg <= 120 and 10 * it_mg < it_j
    assert abs(c_mg[0] - c_j[0]) <= 1e-6 * abs(c_j[0])
    assert np.abs(u_mg - u_j).max() <= 1e-5 * np.abs(u_j).max()


def test_single_precision_level_values(gpu, monkeypatch):
    """Large multigrid levels keep an fp32 copy of their values for the V-cycle's
    products (csrc/spmv_bsr_tma.cu, V = float): the product must agree with the fp64
    one to single-precision rounding on ragged rows (boundary nodes: 8 / 12 / 18 /
    27 blocks), and a hierarchy that uses it must converge like the fp64 one."""
    sktopt, dev = gpu
    monkeypatch.setenv("SKTOPT_B200_MATFREE", "0")
    mesh, basis, D, eng = _engine(sktopt, dims=(4.0, 3.0, 2.0), h=0.125)   # 32 x 24 x 16
    rho = np.random.default_rng(4).uniform(0.01, 1.0, mesh.nelements)
    eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
    eng.assemble(enforce=True)
    x = dev.to_dev(np.random.default_rng(5).standard_normal(eng.n_dof))
    v32 = dev.to_f32(eng.vals)
    assert torch.equal(v32, eng.vals.to(torch.float32))
    y64 = dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, eng.vals, x, eng.max_deg)
    y32 = dev.spmv_bsr3_tma_f32(eng.node_ptr_loc, eng.node_col_loc, v32, x, eng.max_deg)
    # against the same rounded values in fp64 arithmetic: exact up to summation order
    yr = dev.spmv_bsr3_tma(eng.node_ptr_loc, eng.node_col_loc, v32.to(torch.float64), x,
                           eng.max_deg)
    scale = float(y64.abs().max())
    assert float((y32 - yr).abs().max()) <= 1e-13 * scale
    assert float((y32 - y64).abs().max()) <= 1e-6 * scale
    # odd length / conversion tail
    w = dev.to_dev(np.random.default_rng(6).standard_normal(1001))
    assert torch.equal(dev.to_f32(w), w.to(torch.float32))
    # a hierarchy with fp32 level values: 64 x 48 x 32 cells, level 1 = 14,025 nodes
    # (above the bulk-async kernel's threshold; the copy's own threshold is lowered)
    from sktopt.fea._multigrid import Multigrid
    monkeypatch.delenv("SKTOPT_B200_MATFREE")
    its = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("SKTOPT_B200_MG_FP32_LEVELS", flag)
        monkeypatch.setenv("SKTOPT_B200_MG_FP32_MIN_NODES", "10000")
        mesh2, _, D2, e2 = _engine(sktopt, dims=(4.0, 3.0, 2.0), h=0.0625)
        assert e2.precond == "mg" and e2.matrix_free
        assert ("vals32" in e2.mg.levels[1]) == (flag == "1")
        assert "vals32" not in e2.mg.levels[2]
        rho2 = np.random.default_rng(4).uniform(0.01, 1.0, mesh2.nelements)
        e2.set_modulus(dev.to_dev(rho2), 210e3, 210.0, 3.0)
        e2.prepare()
        f = np.zeros(e2.n_dof)
        tip = np.nonzero(mesh2.p[0] == mesh2.p[0].max())[0]
        f[3 * tip + 2] = -1.0
        f[D2] = 0.0
        e2.warm_start = False
        u = e2.solve(dev.to_dev(f), 0, 1e-9, None).cpu().numpy().copy()
        assert e2.pcg_log[-1][1]
        its[flag] = (e2.pcg_log[-1][0], u)
    assert abs(its["1"][0] - its["0"][0]) <= 2, (its["1"][0], its["0"][0])
    assert np.max(np.abs(its["1"][1] - its["0"][1])) <= 1e-6 * np.abs(its["0"][1]).max()
