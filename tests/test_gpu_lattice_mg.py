"""Lattice multigrid with algebraic Galerkin coarse operators (csrc/galerkin_bsr.cu):
meshes numbered like ``init_tensor`` with ANY geometry or element type (jittered
Kuhn tetrahedra = BASELINE config 3's mesh family, jittered hexahedra).  The coarse
operator must equal P^T A P, and MG-PCG must give the oracle's solution in far
fewer iterations than Jacobi-PCG (the reference's ``cg_pyamg`` vs plain cg)."""
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


def _jittered(sktopt, kind, cells, jitter=0.2, seed=0):
    from sktopt._fem import (Basis, ElementHex1, ElementTetP1, ElementVector, MeshHex, MeshTet)
    L = (4.0, 3.0, 2.0)
    axes = [np.linspace(0, l, n + 1) for l, n in zip(L, cells)]
    M = MeshTet if kind == "tet" else MeshHex
    mesh = M.init_tensor(*axes)
    h = L[0] / cells[0]
    p = mesh.p.copy()
    hi = np.array(L)[:, None]
    interior = np.all((p > 1e-9) & (p < hi - 1e-9), axis=0)
    p[:, interior] += np.random.default_rng(seed).uniform(-jitter * h, jitter * h,
                                                          (3, int(interior.sum())))
    t = mesh.t
    if kind == "tet":
        t = sktopt.mesh.utils.fix_tetrahedron_orientation(t, p)
    mesh = M(p, t)
    elem = ElementTetP1() if kind == "tet" else ElementHex1()
    basis = Basis(mesh, ElementVector(elem), intorder=2)
    clamp = np.nonzero(mesh.p[0] < 1e-9)[0]
    D = np.unique((3 * clamp[:, None] + np.arange(3)[None, :]).ravel())
    return mesh, basis, D


def _prolongation(mg, level, mask_f=None):
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
    Pn = sp.kron(Pz, sp.kron(Px, Py)).tocsr()
    return sp.kron(Pn, sp.eye(3)).tocsr()


def _level_matrix(lv):
    """scipy CSR of a node-block level (values in the sktb_spmv_bsr3 layout)."""
    rp = lv["node_ptr"].cpu().numpy().astype(np.int64)
    ci = lv["node_col"].cpu().numpy().astype(np.int64)
    v = lv["vals"].cpu().numpy()
    n = rp.size - 1
    deg = np.diff(rp)
    indptr = np.concatenate([[0], np.cumsum(np.repeat(3 * deg, 3))])
    indices = np.empty(9 * ci.size, dtype=np.int64)
    pos = 0
    for node in range(n):
        cols = (3 * ci[rp[node]:rp[node + 1]][:, None] + np.arange(3)[None, :]).ravel()
        for a in range(3):
            indices[pos:pos + cols.size] = cols
            pos += cols.size
    return sp.csr_matrix((v, indices, indptr), shape=(3 * n, 3 * n))


@pytest.mark.parametrize("kind,cells", [("tet", (9, 7, 6)), ("hex", (8, 6, 5))])
def test_algebraic_galerkin_equals_PtAP(gpu, monkeypatch, kind, cells):
    sktopt, dev = gpu
    from sktopt.fea._engine import FeaEngine, KE_ELASTIC
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    mesh, basis, D = _jittered(sktopt, kind, cells)
    eng = FeaEngine(basis, D, KE_ELASTIC, 0.3)
    assert eng.lattice == tuple(c + 1 for c in cells)
    assert eng.precond == "mg" and eng.mg.algebraic and not eng.matrix_free
    rho = np.random.default_rng(0).uniform(0.05, 1.0, mesh.nelements)
    eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
    eng.prepare()
    torch.cuda.synchronize()
    mg = eng.mg
    rp, ci = eng.dm.dof_pattern(3)
    A = sktopt.fea.composer._csr_to_scipy(eng.n_dof, rp, ci, eng.vals).tocsr()
    mask = eng.dir_mask.cpu().numpy().astype(bool)
    for l in range(1, mg.n_levels):
        lv = mg.levels[l]
        free_f = sp.diags((~mask).astype(float))
        mask_c = lv["mask"].cpu().numpy().astype(bool)
        free_c = sp.diags((~mask_c).astype(float))
        P = _prolongation(mg, l - 1)
        ref = (free_c @ P.T @ free_f @ A @ free_f @ P @ free_c + sp.diags(mask_c.astype(float))).tocsr()
        got = _level_matrix(lv)
        assert abs(got - ref).max() <= 1e-11 * abs(ref).max(), (kind, l)
        # symmetric, positive diagonal
        assert abs(got - got.T).max() <= 1e-11 * abs(ref).max()
        assert got.diagonal().min() > 0.0
        A, mask = got, mask_c


@pytest.mark.parametrize("kind,cells", [("tet", (20, 15, 12)), ("hex", (16, 12, 10))])
def test_lattice_mg_pcg_matches_the_oracle(gpu, kind, cells):
    sktopt, dev = gpu
    from oracle import fem
    from sktopt.fea._engine import FeaEngine, KE_ELASTIC
    mesh, basis, D = _jittered(sktopt, kind, cells)
    eng = FeaEngine(basis, D, KE_ELASTIC, 0.3)
    assert eng.precond == "mg" and eng.mg.algebraic
    cen = np.mean(mesh.p[:, mesh.t], axis=1)
    rho = np.where(np.sin(cen[0] * 3.0) * np.sin(cen[1] * 2.5 + cen[2]) > 0.0, 1.0, 0.01)
    eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
    eng.prepare()
    f = np.zeros(eng.n_dof)
    tip = np.nonzero(mesh.p[0] > mesh.p[0].max() - 1e-9)[0]
    f[3 * tip + 2] = -1.0
    f[D] = 0.0
    fd = dev.to_dev(f)
    eng.warm_start = False
    u_mg = eng.solve(fd, 0, 1e-9, None).cpu().numpy().copy()
    it_mg, ok_mg = eng.pcg_log[-1][0], eng.pcg_log[-1][1]
    eng.mg_enabled = False
    u_j = eng.solve(fd, 1, 1e-9, None).cpu().numpy().copy()
    it_j, ok_j = eng.pcg_log[-1][0], eng.pcg_log[-1][1]
    K = fem.assemble_stiffness(mesh.p, mesh.t, rho, 210e3, 210.0, 3.0, 0.3)
    K_e, f_e = fem.enforce(K, f, D)
    u_ref, _, _ = fem.solve(K_e, f_e, "spsolve")
    print(f"lattice multigrid ({kind}): MG-PCG {it_mg}, Jacobi-PCG {it_j}")
    assert ok_mg and ok_j
    assert np.max(np.abs(u_mg - u_ref)) <= 1e-6 * np.abs(u_ref).max()
    assert np.max(np.abs(u_j - u_ref)) <= 1e-6 * np.abs(u_ref).max()
    assert it_mg * 5 < it_j and it_mg <= 120


def test_two_load_tet_task_through_the_public_api(gpu):
    """BASELINE config 3's shape at test size: jittered Kuhn tetrahedra, two load
    cases; the facade picks the lattice multigrid and returns the oracle's
    compliances and displacements."""
    sktopt, dev = gpu
    import sys, pathlib
    sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1] / "scripts"))
    import workloads
    from oracle import fem
    tsk = workloads.c3_task(sktopt, cells=(18, 13, 12))
    rho = np.random.default_rng(5).uniform(0.2, 1.0, tsk.mesh.nelements)
    fe = sktopt.fea.FEM_SimpLinearElasticity(tsk, 1e-3)
    u = np.zeros((tsk.basis.N, 2))
    c = fe.objectives_multi_load(rho, 3.0, u)
    assert fe.engine.precond == "mg" and fe.engine.mg.algebraic
    its = [e[0] for e in fe.engine.pcg_log[-2:]]
    assert all(e[1] for e in fe.engine.pcg_log[-2:]) and max(its) <= 100, its
    c_ref, U_ref = fem.compliance_multi(tsk.mesh.p, tsk.mesh.t, rho, tsk.E, tsk.E * 1e-3, 3.0,
                                        tsk.nu, tsk.neumann_linear, tsk.dirichlet_dofs)
    assert np.max(np.abs(c - c_ref) / np.abs(c_ref)) <= 1e-6
    assert np.max(np.abs(u - U_ref)) <= 1e-6 * np.abs(U_ref).max()
