"""CPU tests: the oracle against analytic known answers (the reference pins no
numbers on this path), the scheduler scalars the reference's own
tests pin, and the product's host-side task construction against the oracle's
independent restatement (bit-exact numbering)."""
import numpy as np
import pytest

from oracle import fem, filters as ofilters, mesh as omesh, optim


# ------------------------------------------------------------ known answers --
def test_patch_test_constant_strain():
    """K u_lin has zero interior residual and U_e = 1/2 eps:C:eps V."""
    p, t = omesh.box_hex(2.0, 1.0, 1.0, 0.5)
    rng = np.random.default_rng(0)
    interior = np.all((p > 1e-9) & (p < np.array([[2.0], [1.0], [1.0]]) - 1e-9), axis=0)
    p = p.copy()
    p[:, interior] += rng.uniform(-0.08, 0.08, (3, int(interior.sum())))   # distorted hexes
    ne = t.shape[1]
    E, nu = 3.0, 0.3
    K = fem.assemble_stiffness(p, t, np.ones(ne), E, 0.0, 1.0, nu, intorder=3)
    A = rng.standard_normal((3, 3))
    u = (A @ p).T.ravel()                       # u_i = A_ij x_j, dof = 3*node+i
    r = (K @ u).reshape(-1, 3)
    assert np.max(np.abs(r[interior])) <= 1e-11 * np.abs(K.data).max()
    eps = 0.5 * (A + A.T)
    lam, mu = nu * E / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu))
    dens = 0.5 * (2 * mu * np.sum(eps * eps) + lam * np.trace(eps) ** 2)
    U = fem.strain_energy(p, t, np.ones(ne), u, E, 0.0, 1.0, nu, intorder=3)[:, 0]
    np.testing.assert_allclose(U.sum(), dens * 2.0, rtol=1e-12)
    np.testing.assert_allclose(U.sum(), 0.5 * u @ (K @ u), rtol=1e-12)


def test_rigid_body_modes_and_symmetry():
    p, t = omesh.box_hex(1.0, 1.0, 1.0, 0.5)
    K = fem.assemble_stiffness(p, t, np.full(t.shape[1], 0.7), 10.0, 1e-3, 3.0, 0.3)
    assert abs(K - K.T).max() <= 1e-12 * np.abs(K.data).max()
    n = p.shape[1]
    for d in range(3):                       # translations
        u = np.zeros(3 * n)
        u[d::3] = 1.0
        assert np.max(np.abs(K @ u)) <= 1e-11 * np.abs(K.data).max()
    rot = np.zeros(3 * n)                    # rotation about z
    rot[0::3], rot[1::3] = -p[1], p[0]
    assert np.max(np.abs(K @ rot)) <= 1e-11 * np.abs(K.data).max()


def test_toy_mesh_sizes_match_survey():
    """toy_test = 192 hex, 315 nodes, 945 DOF, nnz 55,575."""
    o = omesh.toy_base(1.0)
    assert o["t"].shape == (8, 192) and o["p"].shape == (3, 315)
    K = fem.assemble_stiffness(o["p"], o["t"], np.ones(192), 210e3, 210.0, 3.0, 0.3)
    assert K.shape == (945, 945) and K.nnz == 55575
    np.testing.assert_allclose(o["force"].sum(), -100.0, rtol=1e-13)
    # the reference's hex volume formula gives 5/6 of the true volume under
    # skfem's local vertex order
    np.testing.assert_allclose(o["volumes"], 5.0 / 6.0, rtol=1e-13)


def test_energy_identity_and_enforce(toy_oracle):
    o, pr = toy_oracle
    rho = np.random.default_rng(1).uniform(0.1, 1.0, o["t"].shape[1])
    c, u = fem.compliance_single(o["p"], o["t"], rho, 210e3, 210.0, 3.0, 0.3,
                                 o["force"], o["dirichlet_dofs"])
    assert np.all(u[o["dirichlet_dofs"]] == 0.0)
    U = fem.strain_energy(o["p"], o["t"], rho, u, 210e3, 210.0, 3.0, 0.3)
    np.testing.assert_allclose(2.0 * U.sum(), c, rtol=1e-10)
    # cg + Jacobi reaches the same solution as the direct solver
    K = fem.assemble_stiffness(o["p"], o["t"], rho, 210e3, 210.0, 3.0, 0.3)
    K_e, F_e = fem.enforce(K, o["force"], o["dirichlet_dofs"])
    assert K_e.nnz == K.nnz                  # pattern kept (explicit zeros)
    u2, info, _ = fem.solve(K_e, F_e, "cg_jacobi", rtol=1e-10, maxiter=5000)
    assert info == 0
    np.testing.assert_allclose(u2, u, atol=1e-7 * np.abs(u).max())


def test_enforce_nonzero_dirichlet():
    p, t = omesh.box_hex(1.0, 1.0, 1.0, 0.34)
    K = fem.assemble_scalar(p, t, np.ones(t.shape[1]), 2, "laplace")
    D = np.nonzero(p[0] == 0.0)[0]
    D2 = np.nonzero(p[0] == 1.0)[0]
    both = np.concatenate([D, D2])
    xD = np.concatenate([np.full(D.size, 2.0), np.full(D2.size, 5.0)])
    K_e, b = fem.enforce(K, np.zeros(p.shape[1]), both, xD)
    T, _, _ = fem.solve(K_e, b, "spsolve")
    np.testing.assert_allclose(T, 2.0 + 3.0 * p[0], atol=1e-12)     # linear profile


def test_helmholtz_oracle_properties(toy_oracle):
    o, pr = toy_oracle
    f = ofilters.HelmholtzOracle(o["p"], o["t"], o["volumes"], None)
    f.set_radius(0.5)
    ne = o["t"].shape[1]
    np.testing.assert_allclose(f.forward(np.full(ne, 0.37)), 0.37, atol=1e-12)  # constants preserved
    # with a design mask the nodes of non-design elements are pinned to 1
    g = ofilters.HelmholtzOracle(o["p"], o["t"], o["volumes"], pr.design_mask)
    g.set_radius(0.5)
    out = g.forward(np.full(ne, 0.2))
    assert np.all(np.abs(out[~pr.design_mask] - 1.0) <= 1e-12)
    assert out[pr.design_mask].min() >= 0.2 - 1e-12
    assert np.all(f.gradient(-np.ones(ne)) <= 0.0)
    assert np.all(f.gradient(np.ones(ne)) == 0.0)               # clamp


def test_spatial_oracle_adjoint(toy_oracle):
    o, pr = toy_oracle
    f = ofilters.SpatialOracle(o["p"], o["t"], pr.design_mask)
    f.set_radius(1.5)
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal(192), rng.standard_normal(192)
    m = pr.design_mask
    np.testing.assert_allclose((f.forward(a)[m]) @ b[m], a[m] @ f.gradient(b)[m], rtol=1e-12)
    np.testing.assert_allclose(f.forward(np.ones(192)), 1.0, atol=1e-14)   # rows normalised


def test_heaviside_and_bisection(toy_oracle):
    assert optim.heaviside(np.array([0.0]), 4.0, 0.5)[0] == pytest.approx(0.0, abs=1e-15)
    assert optim.heaviside(np.array([1.0]), 4.0, 0.5)[0] == pytest.approx(1.0, abs=1e-11)
    x = np.linspace(0.01, 0.99, 7)
    fd = (optim.heaviside(x + 1e-6, 3.0, 0.5) - optim.heaviside(x - 1e-6, 3.0, 0.5)) / 2e-6
    np.testing.assert_allclose(optim.heaviside_derivative(x, 3.0, 0.5), fd, rtol=1e-8)
    o, pr = toy_oracle
    h = optim.run(pr, "oc", max_iters=3)
    assert np.all(np.abs(h["vol_error"]) < 1e-4)               # bisection hits V* within vol_tol
    assert np.all(np.isfinite(h["compliance"]))


def test_heat_exchange_forms_known_answers():
    """rho_n linear in x => |grad rho| = b everywhere: J_den = b V, J_num =
    -T_env h (T0 - T_env) b V for a uniform T0, the adjoint load sums to
    -T_env b int h_eff; grad T . grad lambda of two linear fields = (gT.gL) V."""
    from oracle import heat as oheat
    p, t = omesh.box_hex(2.0, 1.0, 1.0, 0.25)
    a, b, h, T_env, T0, pw, q = 0.2, 0.3, 4e-5, 300.0, 450.0, 3.0, 4
    rho_n = a + b * p[0]
    vol = 2.0
    J_num, J_den, rhs = oheat.heat_exchange_forms(p, t, rho_n, np.full(p.shape[1], T0),
                                                  h, T_env, pw, q)
    np.testing.assert_allclose(J_den, b * vol, rtol=1e-12)
    np.testing.assert_allclose(J_num, -T_env * h * (T0 - T_env) * b * vol, rtol=1e-12)
    # int of the nodal interpolant of h_eff: trapezoid in x on the uniform grid
    xs = np.unique(p[0])
    he = h * (a + b * xs) ** pw * (1.0 - a - b * xs) ** q
    integral = np.sum(0.5 * (he[1:] + he[:-1]) * np.diff(xs)) * 1.0 * 1.0
    np.testing.assert_allclose(rhs.sum(), -T_env * b * integral, rtol=1e-12)
    gT, gL = np.array([1.0, -2.0, 0.5]), np.array([0.3, 0.1, -0.7])
    U = oheat.grad_dot_energy(p, t, gT @ p, gL @ p, 2)
    np.testing.assert_allclose(U.sum(), float(gT @ gL) * vol, rtol=1e-12)
    np.testing.assert_allclose(U, float(gT @ gL) * 0.25 ** 3, rtol=1e-11)


def test_fast_diagonalisation_matches_sparse_lu():
    """Host-side check of the product's direct Helmholtz solve (tensor grids, no
    fixed nodes): V D^-1 V^T equals the LU solve of M + r^2 K assembled by the
    oracle with skfem's default quadrature, on uniform and graded grids."""
    import scipy.sparse.linalg as spla
    import torch
    from sktopt._fem import MeshHex
    from sktopt.filters._fastdiag import FastDiagHelmholtz
    rng = np.random.default_rng(5)
    io = fem.default_intorder(8)
    for axes in ((np.linspace(0, 2.0, 9), np.linspace(0, 1.5, 7), np.linspace(0, 1.0, 5)),
                 (np.array([0, 0.1, 0.4, 0.5, 1.0]), np.array([0, 0.3, 0.35, 1.0]),
                  np.array([0, 0.5, 0.6]))):
        m = MeshHex.init_tensor(*axes)
        one = np.ones(m.t.shape[1])
        M = fem.assemble_scalar(m.p, m.t, one, io, "mass")
        K = fem.assemble_scalar(m.p, m.t, one, io, "laplace")
        fd = FastDiagHelmholtz(axes, device="cpu")
        for r in (0.01, 0.3, 1.0):
            fd.set_radius(r)
            b = rng.standard_normal(m.p.shape[1])
            ref = spla.splu((M + r * r * K).tocsc()).solve(b)
            x = fd.solve(torch.as_tensor(b)).numpy()
            assert np.max(np.abs(x - ref)) <= 1e-13 * np.abs(ref).max()


def test_oracle_is_deterministic(toy_oracle):
    o, pr = toy_oracle
    a = optim.run(pr, "logmoc", max_iters=3, vol_frac=0.6)
    b = optim.run(pr, "logmoc", max_iters=3, vol_frac=0.6)
    assert np.array_equal(a["rho_final"], b["rho_final"])


# ------------------------------------------- scheduler scalars (reference) --
def test_scheduler_scalars_pinned_by_reference_tests():
    """Numbers from the reference's tests/test_scheduler.py:34-35,77-80,99-113
    hold for the product's scheduler and for the oracle's restatement."""
    from sktopt.tools.scheduler import (
        Scheduler, SchedulerConfig, schedule_step, schedule_step_accelerating,
        schedule_step_decelerating, SchedulerStepAcceleratingToOne,
        SchedulerStepDeceleratingToOne, schedule_sawtooth_decay)
    kw = dict(total=10, initial_value=1.0, target_value=5.0, num_steps=5, curvature=2.0)
    vals = [schedule_step_decelerating(it=i, **kw) for i in (1, 3, 5, 7, 9)]
    assert np.isclose(vals[0], 1.0) and np.isclose(vals[-1], 5.0)
    assert all(a <= b for a, b in zip(vals, vals[1:]))
    assert schedule_step_accelerating(it=3, **kw) < schedule_step(it=3, **kw) < schedule_step_decelerating(it=3, **kw)
    assert schedule_step_accelerating(it=7, **kw) < schedule_step(it=7, **kw) < schedule_step_decelerating(it=7, **kw)
    k1 = dict(total=10, initial_value=0.5, target_value=1.5, num_steps=4, curvature=1.0)
    lin = [schedule_step(it=i, **k1) for i in (1, 4, 7, 9)]
    np.testing.assert_allclose(lin, [schedule_step_decelerating(it=i, **k1) for i in (1, 4, 7, 9)])
    np.testing.assert_allclose(lin, [schedule_step_accelerating(it=i, **k1) for i in (1, 4, 7, 9)])
    cfg = SchedulerConfig.step_to_one(name="x", num_steps=4, iters_max=8)
    s = Scheduler.from_config(cfg)
    assert np.isclose(cfg.target_value, 1.0) and np.isclose(cfg.init_value, 0.25)
    assert np.isclose(s.value(1), 0.25) and np.isclose(s.value(8), 1.0)
    with pytest.raises(ValueError):
        SchedulerConfig.from_defaults(name="x", num_steps=3, target_value=0.5,
                                      scheduler_type="StepToOne")
    for factory in (SchedulerConfig.step_accelerating_to_one,
                    SchedulerConfig.step_decelerating_to_one):
        c = factory(name="x", num_steps=5, iters_max=10, curvature=2.0)
        s = Scheduler.from_config(c)
        assert np.isclose(c.target_value, 1.0) and np.isclose(c.init_value, 0.2)
        assert np.isclose(s.value(1), 0.2) and s.value(10) <= 1.0
    assert np.isclose(SchedulerStepAcceleratingToOne("x", 5, 10, 2.0).value(1), 0.2)
    assert np.isclose(SchedulerStepDeceleratingToOne("x", 5, 10, 2.0).value(1), 0.2)
    # docstring examples of the reference (tools/scheduler.py:703-707, 1118-1133)
    st = Scheduler.from_config(SchedulerConfig.step(
        name="p", init_value=1.0, target_value=3.0, num_steps=5, iters_max=100))
    assert st.value(1) == 1.0 and st.value(50) == 2.0 and st.value(100) == 3.0
    assert np.isclose(schedule_sawtooth_decay(1, 120, 0.3, 0.1, 4), 0.3)
    assert np.isclose(schedule_sawtooth_decay(61, 120, 0.3, 0.1, 4), 0.3)
    # the oracle's restatement agrees with the product's on every iteration
    for it in range(1, 61):
        assert optim.sched_step(it, 60, 1.0, 3.0, 3) == schedule_step(it, 60, 1.0, 3.0, 3)
        assert optim.sched_step(it, 60, 1.0, 2.0, 3, 2.0, "accelerating") == \
            schedule_step_accelerating(it, 60, 1.0, 2.0, 3, 2.0)
        assert optim.sched_sawtooth(it, 60, 0.3, 0.1, 6) == \
            (schedule_sawtooth_decay(it, 60, 0.3, 0.1, 6) if it < 60 else 0.1)


# ------------------------------- product host logic vs oracle (bit-exact) --
@pytest.mark.parametrize("h", [1.0, 0.45, 0.31])
def test_task_arrays_bit_exact_against_oracle(h):
    import sktopt
    tsk = sktopt.mesh.toy_problem.toy_base(h)
    tsk.exlude_dirichlet_from_design()
    o = omesh.toy_base(h)
    assert np.array_equal(tsk.mesh.p, o["p"])                     # node numbering
    assert np.array_equal(tsk.mesh.t, o["t"])                     # connectivity
    assert tsk.mesh.t.dtype == np.int32
    assert np.array_equal(tsk.dirichlet_dofs, o["dirichlet_dofs"])  # DOF numbering
    assert np.array_equal(tsk.design_elements, o["design"])
    assert np.array_equal(tsk.fixed_elements, o["fixed"])
    assert np.array_equal(np.sort(tsk.dirichlet_neumann_elements), np.sort(o["pinned"]))
    np.testing.assert_allclose(tsk.elements_volume, o["volumes"], rtol=1e-13)
    np.testing.assert_allclose(tsk.neumann_linear[0], o["force"], rtol=0, atol=1e-13)
    assert tsk.basis.N == 3 * o["p"].shape[1]
    ed = tsk.basis.element_dofs
    assert np.array_equal(ed[3 * 5 + 2], 3 * o["t"][5].astype(np.int64) + 2)


def test_toy2_and_config_defaults():
    import sktopt
    t2 = sktopt.mesh.toy_problem.toy2()
    assert t2.n_tasks == 2 and t2.mesh.nelements == 27 * 27 * 4
    np.testing.assert_allclose([f.sum() for f in t2.neumann_linear], [-1.0, 1.0], rtol=1e-12)
    cfg = sktopt.core.OC_Config()
    assert cfg.solver_option == "spsolve" and cfg.filter_type == "helmholtz"
    assert cfg.move_limit.scheduler_type == "SawtoothDecay" and cfg.move_limit.num_steps == 6
    assert cfg.eta.target_value == 0.5 and cfg.lambda_lower == 1e-7
    lm = sktopt.core.LogMOC_Config()
    assert lm.eta == 0.6 and lm.lambda_lower == -1e7 and lm.mu_p == 5.0
    with pytest.raises(RuntimeError):
        sktopt.core.OC_Config(solver_option="petsc")


def test_c_port_matches_the_numpy_oracle():
    """oracle/cport (C / OpenMP: the CPU baseline at full size) against the
    NumPy / SciPy restatement: same pattern, K to 1e-14, the same PCG iterates,
    element energies to 1e-12, and a 3-iteration LogMOC loop to 1e-10."""
    from oracle import cport, fem, mesh as omesh, optim
    o = omesh.toy_base(0.8)
    be = cport.CBackend(o["p"], o["t"], 3, cport.unit_elasticity_ke(o["p"], o["t"], o["nu"]),
                        o["dirichlet_dofs"])
    rho = np.random.default_rng(0).uniform(0.1, 1.0, o["t"].shape[1])
    E = fem.simp(rho, 210e3, 210.0, 3.0)
    K1 = be.assemble(E)
    K = fem.assemble_stiffness(o["p"], o["t"], rho, 210e3, 210.0, 3.0, 0.3)
    K2, _ = fem.enforce(K, o["force"], o["dirichlet_dofs"])
    assert np.array_equal(K1.indptr, K2.indptr) and np.array_equal(K1.indices, K2.indices)
    assert np.abs(K1.data - K2.data).max() <= 1e-14 * np.abs(K2.data).max()
    F = o["force"].copy()
    F[o["dirichlet_dofs"]] = 0.0
    u1, it1, rel = be.pcg(F, 1e-8)
    u2, _, it2 = fem.solve(K2, F, "cg_jacobi", 1e-8)
    assert abs(it1 - it2) <= 1 and rel <= 1e-8
    assert np.abs(u1 - u2).max() <= 1e-7 * np.abs(u2).max()
    e1 = be.energy(E, u2)
    e2 = fem.strain_energy(o["p"], o["t"], rho, u2, 210e3, 210.0, 3.0, 0.3)[:, 0]
    assert np.abs(e1 - e2).max() <= 1e-12 * np.abs(e2).max()
    M, Ks = cport.scalar_matrices(o["p"], o["t"])
    M2 = fem.assemble_scalar(o["p"], o["t"], None, 6, "mass")
    assert abs(M - M2).max() <= 1e-14 * abs(M2).max()
    pr = optim.Problem(o["p"], o["t"], o["dirichlet_dofs"], o["force"], o["design"],
                       o["pinned"], o["volumes"], o["E"], o["nu"], fixed=o["fixed"])
    a = optim.run(pr, "logmoc", max_iters=200, iters=3, vol_frac=0.3, solver="cg_jacobi")
    b = optim.run(pr, "logmoc", max_iters=200, iters=3, vol_frac=0.3, solver="cg_jacobi",
                  backend=be, filter_solver="cg", filter_matrices=(M, Ks))
    ca, cb = np.array(a["compliance"]), np.array(b["compliance"])
    assert np.abs(ca - cb).max() <= 1e-10 * np.abs(ca).max()
    assert np.abs(a["rho_final"] - b["rho_final"]).max() <= 1e-9
