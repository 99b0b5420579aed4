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
