"""Scalar stencil multigrid (csrc/mg_scalar.cu): the CSR -> stencil conversion
reproduces the operator, the coarse stencils equal the algebraic Galerkin product
P^T A P (Dirichlet nodes as identity rows), the V-cycle is symmetric, and the
MG-PCG reaches the direct solution of the heat system in few iterations --
including a ~98k-element plate against the oracle's sparse LU (BASELINE config 4
scaled down)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sktopt
    from sktopt._b200 import device as dev
    return sktopt, dev


def _scalar_engine(sktopt, dims=(5.0, 3.8, 2.2), h=0.2):
    from sktopt._fem import Basis, ElementHex1
    from sktopt.fea._engine import FeaEngine, KE_LAPLACE
    mesh = sktopt.mesh.toy_problem.create_box_hex(*dims, h)      # 25 x 19 x 11 cells (odd sizes)
    basis = Basis(mesh, ElementHex1(), intorder=2)
    D = np.nonzero((mesh.p[0] <= 0.2 + 1e-9) & (mesh.p[1] <= 0.4 + 1e-9))[0]
    return mesh, basis, D, FeaEngine(basis, D, KE_LAPLACE)


def _prolongation(smg, level, mask_f, mask_c):
    from sktopt.fea._multigrid import axis_tables
    mats = []
    for c in smg.coords[level]:
        n = c.size - 1
        c0, c1, w0, w1, _, _ = axis_tables(n)
        P = sp.lil_matrix((n + 1, (n + 1) // 2 + 1))
        for i in range(n + 1):
            P[i, c0[i]] += w0[i]
            P[i, c1[i]] += w1[i]
        mats.append(P.tocsr())
    Px, Py, Pz = mats
    P = sp.kron(Pz, sp.kron(Px, Py)).tocsr()
    return sp.diags(1.0 - mask_f) @ P @ sp.diags(1.0 - mask_c)


def _dia_to_scipy(vals, np_axes):
    npx, npy, npz = (int(v) for v in np_axes)
    n = npx * npy * npz
    idx = np.arange(n)
    rows, cols, data = [], [], []
    for k in range(27):
        dz, dx, dy = k // 9 - 1, (k // 3) % 3 - 1, k % 3 - 1
        j = idx + dz * npx * npy + dx * npy + dy
        ok = (vals[k] != 0.0) & (j >= 0) & (j < n)
        rows.append(idx[ok]); cols.append(j[ok]); data.append(vals[k][ok])
    return sp.csr_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(n, n))


def test_stencil_levels_equal_the_galerkin_products(gpu, monkeypatch):
    sktopt, dev = gpu
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    mesh, basis, D, eng = _scalar_engine(sktopt)
    smg = eng.smg
    assert smg is not None and smg.n_levels >= 3
    rho = np.random.default_rng(1).uniform(0.05, 1.0, mesh.nelements)
    eng.set_modulus(dev.to_dev(rho), 10.0, 0.01, 3.0)
    eng.assemble(enforce=True)
    eng.update_preconditioner(eng.vals)
    A = sktopt.fea.composer._csr_to_scipy(eng.n_dof, eng.row_ptr, eng.col_idx, eng.vals)
    A0 = _dia_to_scipy(smg.level_values(0).cpu().numpy(), smg.np_h[0])
    assert abs(A0 - A).max() <= 1e-14 * abs(A).max()
    x = np.random.default_rng(2).standard_normal(eng.n_dof)
    y = smg.apply(0, dev.to_dev(x)).cpu().numpy()
    assert np.abs(y - A @ x).max() <= 1e-12 * np.abs(A @ x).max()
    masks = [np.zeros(int(np.prod(smg.np_h[l]))) for l in range(smg.n_levels)]
    masks[0][D] = 1.0
    Al = A
    from sktopt.fea._multigrid import coarse_index_map
    for l in range(smg.n_levels - 1):
        fine_cells = [c.size - 1 for c in smg.coords[l]]
        fm = [coarse_index_map(n) for n in fine_cells]
        cnp = smg.np_h[l + 1]
        Iz, Ix, Iy = np.meshgrid(np.arange(cnp[2]), np.arange(cnp[0]), np.arange(cnp[1]),
                                 indexing="ij")
        fnode = (fm[1][Iy] + (fine_cells[1] + 1) * fm[0][Ix]
                 + (fine_cells[1] + 1) * (fine_cells[0] + 1) * fm[2][Iz]).ravel()
        masks[l + 1] = masks[l][fnode]
        P = _prolongation(smg, l, masks[l], masks[l + 1])
        Ac = (P.T @ Al @ P + sp.diags(masks[l + 1])).tocsr()
        Ad = _dia_to_scipy(smg.level_values(l + 1).cpu().numpy(), cnp)
        assert abs(Ad - Ac).max() <= 1e-12 * abs(Ac).max(), l
        Al = Ac


def test_vcycle_symmetric_and_pcg_matches_direct_solve(gpu):
    sktopt, dev = gpu
    mesh, basis, D, eng = _scalar_engine(sktopt, dims=(4.0, 4.0, 1.0), h=0.1)   # 40x40x10
    rho = np.random.default_rng(3).uniform(0.01, 1.0, mesh.nelements)
    rho[rho > 0.6] = 1.0
    rho[rho < 0.4] = 0.01                                   # contrast 1e6 in k
    eng.set_modulus(dev.to_dev(rho), 10.0, 1e-2, 3.0)
    eng.assemble(enforce=True)
    eng.update_preconditioner(eng.vals)
    A = sktopt.fea.composer._csr_to_scipy(eng.n_dof, eng.row_ptr, eng.col_idx, eng.vals)
    rng = np.random.default_rng(4)
    r1, r2 = rng.standard_normal(eng.n_dof), rng.standard_normal(eng.n_dof)
    z1 = eng.smg.vcycle(dev.to_dev(r1)).cpu().numpy()
    z2 = eng.smg.vcycle(dev.to_dev(r2)).cpu().numpy()
    assert abs(z1 @ r2 - r1 @ z2) <= 1e-10 * abs(z1 @ r2)
    assert z1 @ r1 > 0 and z2 @ r2 > 0
    b = rng.standard_normal(eng.n_dof)
    b[D] = 0.0
    x = eng.solve(dev.to_dev(b), 0, 1e-10, None, vals=eng.vals).cpu().numpy()
    it_mg = eng.pcg_log[-1][0]
    ref = spla.spsolve(A.tocsc(), b)
    assert np.abs(x - ref).max() <= 1e-7 * np.abs(ref).max()
    eng.mg_enabled = False
    eng.u.clear()
    eng.solve(dev.to_dev(b), 0, 1e-10, None, vals=eng.vals)
    it_jac = eng.pcg_log[-1][0]
    print("scalar MG-PCG iterations", it_mg, "Jacobi-PCG", it_jac)
    assert it_mg <= 40 and it_mg * 5 < it_jac


def test_heat_plate_98k_elements_against_oracle_lu(gpu):
    """BASELINE config 4 scaled to 8 x 8 x 1 @ h = 0.0725 (111 x 111 x 14 = 172k hex
    would take the oracle's LU minutes; 88 x 88 x 11 = 85k hex takes seconds)."""
    sktopt, dev = gpu
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from scripts import workloads
    from oracle import heat as oheat
    from test_gpu_heat import oracle_inputs
    tsk = workloads.c4_task(sktopt, mesh_size=8.0 / 88)
    p, t = tsk.mesh.p, tsk.mesh.t
    assert 80000 <= t.shape[1] <= 100000
    Bs, fs, D = oracle_inputs(tsk)
    rho = np.random.default_rng(0).uniform(0.1, 0.95, t.shape[1])
    fem_gpu = sktopt.fea.FEM_SimpLinearHeatConduction(tsk, 1e-3)
    T = np.zeros((tsk.basis.N, 1))
    J = fem_gpu.objectives_multi_load(rho, 3.0, T)
    eng = fem_gpu.engine
    assert eng.smg is not None
    its = eng.pcg_log[-1][0]
    J_ref, T_ref, _ = oheat.solve_compliance(p, t, rho, 10.0, 1e-2, 3.0, 4, 4.0e-5, 300.0, Bs, fs,
                                             D, 600.0, 2)
    print("heat 85k: MG-PCG iterations", its, "J rel", abs(J[0] - J_ref) / abs(J_ref))
    assert abs(J[0] - J_ref) <= 1e-6 * abs(J_ref)
    assert np.abs(T[:, 0] - T_ref).max() <= 1e-6 * np.abs(T_ref).max()
    assert its <= 40
