"""Function-level entry points of the reference that wrap the same kernels as the
classes: Helmholtz steps (filters/helmholtz_filter_nodal.py:26-56,121-157), heat
energies and adjoint gradient densities (fea/solver_heat.py:256-303,327-383,518-549),
composer re-exports of the strain energy."""
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


def test_helmholtz_function_api(gpu):
    sktopt, dev = gpu
    import scipy.sparse.linalg as spla
    from oracle import fem, filters as ofilters
    from sktopt.filters import helmholtz_filter_nodal as hf
    tsk = sktopt.mesh.toy_problem.toy_base(0.6)
    tsk.exlude_dirichlet_from_design()
    mesh, vol, mask = tsk.mesh, tsk.elements_volume, tsk.design_mask
    p, t = mesh.p, mesh.t
    rng = np.random.default_rng(0)
    rho = rng.uniform(0.05, 1.0, mesh.nelements)
    # element -> node averaging (weighted / plain, with and without a mask)
    got = hf.element_to_node_density_averaging(mesh, vol, rho, mask)
    assert np.max(np.abs(got - ofilters.element_to_node(t, vol, rho, mask, 1.0))) <= 1e-14
    got = hf.element_to_node_density_averaging(mesh, vol, rho, mask, weighted=False,
                                               fixed_value_for_design=0.0)
    ref = ofilters.element_to_node(t, np.ones_like(vol), rho, mask, 0.0)
    assert np.max(np.abs(got - ref)) <= 1e-14
    got = hf.element_to_node_density_averaging(mesh, vol, rho)
    assert np.max(np.abs(got - ofilters.element_to_node(t, vol, rho, None, 1.0))) <= 1e-14
    # node -> element mean
    xn = rng.standard_normal(mesh.nvertices)
    assert np.max(np.abs(hf.node_to_element_density(mesh, xn) - ofilters.node_to_element(t, xn))) <= 1e-14
    # the solve
    io = fem.default_intorder(8)
    M = fem.assemble_scalar(p, t, None, io, "mass")
    K = fem.assemble_scalar(p, t, None, io, "laplace")
    r = 0.4
    A = (M + r * r * K).tocsc()
    b = M @ xn
    x_ref = spla.splu(A).solve(b)
    x = hf.solve_helmholtz("gradient", mesh, xn, r)
    assert np.max(np.abs(x - x_ref)) <= 1e-9 * np.abs(x_ref).max()
    fixed = np.unique(t[:, ~mask].ravel())
    free = np.setdiff1d(np.arange(mesh.nvertices), fixed)
    xf = np.ones(mesh.nvertices)
    rhs = (b - A @ np.where(np.isin(np.arange(mesh.nvertices), fixed), 1.0, 0.0))[free]
    xf[free] = spla.splu(A[free][:, free].tocsc()).solve(rhs)
    x = hf.solve_helmholtz("forward", mesh, xn, r, mask)
    assert np.max(np.abs(x - xf)) <= 1e-9 * np.abs(xf).max()
    assert np.max(np.abs(x[fixed] - 1.0)) <= 1e-10
    with pytest.raises(NotImplementedError):
        hf.solve_helmholtz("gradient", mesh, xn, r, mask)
    with pytest.raises(ValueError):
        hf.solve_helmholtz("sideways", mesh, xn, r)
    # composition of the three functions = the filter class
    filt = hf.HelmholtzFilterNodal.from_defaults(mesh, vol, r, design_mask=mask)
    composed = hf.node_to_element_density(mesh, hf.solve_helmholtz(
        "forward", mesh, hf.element_to_node_density_averaging(mesh, vol, rho, mask), r, mask))
    assert np.max(np.abs(composed - filt.forward(rho))) <= 1e-9


def test_heat_and_energy_function_api(gpu):
    sktopt, dev = gpu
    from oracle import fem, heat as oheat
    from sktopt._fem import Basis, ElementHex1
    from sktopt.fea import composer, solver_heat
    mesh = sktopt.mesh.toy_problem.create_box_hex(2.0, 1.0, 1.0, 0.25)
    basis = Basis(mesh, ElementHex1(), intorder=2)
    rng = np.random.default_rng(1)
    rho = rng.uniform(0.05, 1.0, mesh.nelements)
    T = 300.0 + rng.standard_normal((mesh.nvertices, 2))
    lam = rng.standard_normal((mesh.nvertices, 2))
    U = solver_heat.heat_energy_skfem_multi(basis, rho, T, 10.0, 1e-2, 3.0)
    U_ref = fem.heat_energy(mesh.p, mesh.t, rho, T, 10.0, 1e-2, 3.0, 2)
    assert U.shape == (mesh.nelements, 2)
    assert np.max(np.abs(U - U_ref)) <= 1e-10 * np.abs(U_ref).max()
    u1 = solver_heat.heat_energy_skfem(basis, rho, T[:, 1], 10.0, 1e-2, 3.0)
    assert np.array_equal(u1, U[:, 1])
    G = solver_heat.heat_exchange_grad_density_multi(basis, T, lam)
    for i in range(2):
        ref = oheat.grad_dot_energy(mesh.p, mesh.t, T[:, i], lam[:, i], 2)
        assert np.max(np.abs(G[:, i] - ref)) <= 1e-10 * np.abs(ref).max()
    assert np.array_equal(solver_heat.avg_temp_grad_density_multi(basis, T, lam), G)
    # composer re-exports of the strain energy
    tsk = sktopt.mesh.toy_problem.toy_test()
    u = rng.standard_normal((tsk.basis.N, 2))
    rho_t = rng.uniform(0.05, 1.0, tsk.mesh.nelements)
    E = composer.strain_energy_skfem_multi(tsk.basis, rho_t, u, 210e3, 210.0, 3.0, 0.3)
    E_ref = fem.strain_energy(tsk.mesh.p, tsk.mesh.t, rho_t, u, 210e3, 210.0, 3.0, 0.3)
    assert np.max(np.abs(E - E_ref)) <= 1e-10 * np.abs(E_ref).max()
    e0 = composer.strain_energy_skfem(tsk.basis, rho_t, u[:, 0], 210e3, 210.0, 3.0, 0.3)
    assert np.array_equal(e0, E[:, 0])


def test_fast_diagonalisation_hand_written_products_match_cublas(gpu, monkeypatch):
    """The six dense products of the direct Helmholtz solve run on csrc/dgemm.cu;
    torch.matmul (cuBLAS) is kept as the cross-check: same result to 1e-13, on a
    grid whose sizes are not multiples of the 64 x 64 x 16 tile."""
    sktopt, dev = gpu
    import torch
    from sktopt.filters._fastdiag import FastDiagHelmholtz
    axes = (np.linspace(0, 8, 71), np.linspace(0, 6, 54), np.linspace(0, 4, 37))
    fd = FastDiagHelmholtz(axes)
    fd.set_radius(0.35)
    n = 71 * 54 * 37
    b = torch.as_tensor(np.random.default_rng(0).standard_normal(n), device="cuda")
    x1 = fd.solve(b).clone()
    monkeypatch.setenv("SKTOPT_B200_FD_TORCH", "1")
    x2 = fd.solve(b)
    assert float((x1 - x2).abs().max()) <= 1e-13 * float(x2.abs().max())
    out = torch.empty_like(b)
    monkeypatch.delenv("SKTOPT_B200_FD_TORCH")
    assert fd.solve(b, out=out) is out and torch.equal(out, x1)
    # plain product, batched with strides, against numpy
    A = np.random.default_rng(1).standard_normal((3, 70, 33))
    B = np.random.default_rng(2).standard_normal((3, 33, 45))
    C = torch.empty((3, 70, 45), dtype=torch.float64, device="cuda")
    dev.dgemm(dev.to_dev(A), dev.to_dev(B), C, 70, 45, 33, 33, 45, 45, batch=3,
              stride_a=70 * 33, stride_b=33 * 45, stride_c=70 * 45)
    assert np.abs(C.cpu().numpy() - A @ B).max() <= 1e-13 * np.abs(A @ B).max()


def test_stress_tensor_and_von_mises(gpu):
    """fea/composer.py:444-519: sigma = 2 mu eps + lam tr(eps) I at the quadrature
    points.  A linear displacement field has a constant, known strain; a general
    field is checked against the oracle's gradient tables, hex and tet."""
    sktopt, dev = gpu
    from oracle import fem
    from sktopt._fem import Basis, ElementHex1, ElementTetP1, ElementVector
    comp = sktopt.fea.composer
    E0, Emin, p_pow, nu = 210e3, 210.0, 3.0, 0.3
    for mesh, elem in ((sktopt.mesh.toy_problem.create_box_hex(2.0, 1.0, 1.0, 0.25), ElementHex1()),
                       (sktopt.mesh.toy_problem.create_box_tet(2.0, 1.0, 1.0, 0.34), ElementTetP1())):
        basis = Basis(mesh, ElementVector(elem), intorder=2)
        P, t = mesh.p, mesh.t
        rho = np.random.default_rng(0).uniform(0.2, 1.0, t.shape[1])
        H = np.array([[1e-3, 2e-3, 0.0], [-1e-3, 5e-4, 3e-3], [2e-3, 0.0, -1e-3]])
        u = (H @ P).T.ravel()                          # u_i = H_ij x_j, dof = 3 node + i
        s = comp.stress_tensor_skfem(basis, rho, u, E0, Emin, p_pow, nu)
        nq = basis.X.shape[1]
        assert s.shape == (3, 3, t.shape[1], nq)
        Ee = fem.simp(rho, E0, Emin, p_pow)
        lam, mu = nu * Ee / ((1 + nu) * (1 - 2 * nu)), Ee / (2 * (1 + nu))
        eps = 0.5 * (H + H.T)
        ref = (2 * mu[None, None, :] * eps[:, :, None]
               + lam[None, None, :] * np.trace(eps) * np.eye(3)[:, :, None])
        assert np.abs(s - ref[:, :, :, None]).max() <= 1e-10 * np.abs(ref).max()
        # general field against the oracle's physical gradients
        u2 = np.random.default_rng(1).standard_normal(3 * P.shape[1]) * 1e-3
        s2 = comp.stress_tensor_skfem(basis, rho, u2, E0, Emin, p_pow, nu)
        _, G, _ = fem.physical_gradients(P, t, basis.X)            # (ne, nq, nen, 3)
        ue = np.stack([u2[3 * t.astype(np.int64) + c] for c in range(3)], axis=-1)  # (nen, ne, 3)
        grad = np.einsum("aei,eqaj->ijeq", ue, G)
        w = {"uh": grad, "mu_elem": np.tile(mu, (nq, 1)), "lam_elem": np.tile(lam, (nq, 1))}
        ref2 = comp.compute_element_stress_tensor(w)
        assert np.abs(s2 - ref2).max() <= 1e-10 * np.abs(ref2).max()
        vm = comp.von_mises_from_stress_tensor(s2)
        dev_s = s2 - np.trace(s2, axis1=0, axis2=1)[None, None] * np.eye(3)[:, :, None, None] / 3
        assert np.abs(vm - np.sqrt(1.5 * np.einsum("ijeq,ijeq->eq", dev_s, dev_s))).max() \
            <= 1e-10 * vm.max()
        vm_d = comp.von_mises_from_stress_tensor(dev.to_dev(s2))
        assert np.abs(vm_d.cpu().numpy() - vm).max() <= 1e-12 * vm.max()


def test_helmholtz_filter_element(gpu):
    """Element-graph Helmholtz filter (filters/helmholtz_filter_element.py:195-317,
    448-536) against the oracle's sparse LU, hex and tet, NumPy and CUDA inputs."""
    sktopt, dev = gpu
    from oracle.filters import HelmholtzElementOracle
    from sktopt.filters.helmholtz_filter_element import (HelmholtzFilterElement,
                                                         prepare_helmholtz_filter)
    for mesh in (sktopt.mesh.toy_problem.create_box_hex(2.0, 1.0, 1.0, 0.25),
                 sktopt.mesh.toy_problem.create_box_tet(2.0, 1.0, 1.0, 0.34)):
        vol = sktopt.fea.composer.get_elements_volume(mesh)
        rho = np.random.default_rng(0).uniform(0.0, 1.0, mesh.nelements)
        ref = HelmholtzElementOracle(mesh.p, mesh.t, 0.4)
        A, V = prepare_helmholtz_filter(mesh, 0.4)
        assert abs(A - ref.A).max() <= 1e-13 * abs(ref.A).max()
        f = HelmholtzFilterElement.from_defaults(mesh, vol, 0.4, solver_option="spsolve")
        y = f.forward(rho)
        assert np.abs(y - ref.forward(rho)).max() <= 1e-9
        g = f.gradient(dev.to_dev(rho))
        assert g.is_cuda and np.abs(g.cpu().numpy() - ref.gradient(rho)).max() <= 1e-9
        # the filter preserves the volume-weighted mean (V-weighted column sums of A^-1 V)
        w = vol / vol.mean()
        assert abs(np.sum(w * y) - np.sum(w * rho)) <= 1e-8 * np.sum(w)
        f.update_radius(0.2)
        assert np.abs(f.forward(rho) - HelmholtzElementOracle(mesh.p, mesh.t, 0.2).forward(rho)).max() <= 1e-9
