"""Matrix-free grid operator (csrc/gridop.cu) against the assembled K(rho):
same product (with and without Dirichlet dofs, on node sub-ranges), same
diagonal, and the same PCG solution as the oracle's direct solve."""
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


def _engines(sktopt, monkeypatch, dims=(2.8, 2.0, 1.6), h=0.4, partial_bc=False):
    from sktopt._fem import Basis, ElementHex1, ElementVector
    from sktopt.fea._engine import FeaEngine, KE_ELASTIC
    mesh = sktopt.mesh.toy_problem.create_box_hex(*dims, h)
    basis = Basis(mesh, ElementVector(ElementHex1()), intorder=2)
    clamp = np.nonzero(mesh.p[0] == 0.0)[0]
    D = (3 * clamp[:, None] + np.arange(3)[None, :]).ravel()
    if partial_bc:
        # a roller: only u_z fixed on the top face, only u_y on one edge line
        top = np.nonzero(mesh.p[2] == mesh.p[2].max())[0]
        edge = np.nonzero((mesh.p[1] == 0.0) & (mesh.p[2] == 0.0))[0]
        D = np.concatenate([D, 3 * top + 2, 3 * edge + 1])
    D = np.unique(D)
    monkeypatch.setenv("SKTOPT_B200_MATFREE", "1")
    mf = FeaEngine(basis, D, KE_ELASTIC, 0.3)
    monkeypatch.setenv("SKTOPT_B200_MATFREE", "0")
    asm = FeaEngine(basis, D, KE_ELASTIC, 0.3)
    assert mf.matrix_free and not asm.matrix_free
    return mesh, basis, D, mf, asm


@pytest.mark.parametrize("partial_bc", [False, True])
def test_apply_and_diagonal_match_assembled(gpu, monkeypatch, partial_bc):
    sktopt, dev = gpu
    mesh, basis, D, mf, asm = _engines(sktopt, monkeypatch, partial_bc=partial_bc)
    rho = np.random.default_rng(0).uniform(0.01, 1.0, mesh.nelements)
    for eng in (mf, asm):
        eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
        eng.prepare()
    assert mf._pattern is None          # no CSR was built for the matrix-free engine
    rng = np.random.default_rng(1)
    x = rng.standard_normal(mf.n_dof)   # NOT zero at the Dirichlet dofs
    xd = dev.to_dev(x)
    y_ref = dev.spmv(asm.row_ptr, asm.col_idx, asm.vals, xd, 3).cpu().numpy()
    y = mf.gridop.apply(xd).cpu().numpy()
    assert np.max(np.abs(y - y_ref)) <= 1e-12 * np.max(np.abs(y_ref))
    assert np.array_equal(y[D], x[D])
    # node sub-ranges (the rows a rank owns when the operator is sharded)
    n = mesh.p.shape[1]
    for lo, hi in ((0, 17), (17, n - 5), (n - 5, n)):
        part = mf.gridop.apply(xd, lo, hi - lo).cpu().numpy()
        assert np.array_equal(part, y[3 * lo:3 * hi])
    d_ref = asm.inv_diag.cpu().numpy()
    d = mf.inv_diag.cpu().numpy()
    assert np.max(np.abs(d - d_ref)) <= 1e-13 * np.max(np.abs(d_ref))


def test_matrix_free_solve_matches_oracle(gpu, monkeypatch):
    sktopt, dev = gpu
    from oracle import fem
    monkeypatch.setenv("SKTOPT_B200_PRECOND", "mg")
    mesh, basis, D, mf, asm = _engines(sktopt, monkeypatch, dims=(4.0, 3.0, 2.0), h=0.25)
    rho = np.random.default_rng(3).uniform(0.01, 1.0, mesh.nelements)
    f = np.zeros(mf.n_dof)
    tip = np.nonzero(mesh.p[0] == mesh.p[0].max())[0]
    f[3 * tip + 2] = -1.0
    f[D] = 0.0
    K = fem.assemble_stiffness(mesh.p, mesh.t, rho, 210e3, 210.0, 3.0, 0.3)
    K_e, f_e = fem.enforce(K, f, D)
    u_ref, _, _ = fem.solve(K_e, f_e, "spsolve")
    its = {}
    for name, eng in (("mf", mf), ("asm", asm)):
        assert eng.precond == "mg"
        eng.set_modulus(dev.to_dev(rho), 210e3, 210.0, 3.0)
        eng.prepare()
        eng.warm_start = False
        u = eng.solve(dev.to_dev(f), 0, 1e-10, None).cpu().numpy()
        assert eng.pcg_log[-1][1]
        its[name] = eng.pcg_log[-1][0]
        assert np.max(np.abs(u - u_ref)) <= 1e-7 * np.max(np.abs(u_ref))
        # Jacobi PCG on the same operator
        eng.mg_enabled = False
        uj = eng.solve(dev.to_dev(f), 1, 1e-10, None).cpu().numpy()
        assert np.max(np.abs(uj - u_ref)) <= 1e-7 * np.max(np.abs(u_ref))
    assert abs(its["mf"] - its["asm"]) <= 2


def test_helmholtz_filter_matrix_free_matches_assembled(gpu, monkeypatch):
    """Scalar grid operator inside the Helmholtz filter (forward with fixed
    non-design nodes, adjoint without) against the assembled CSR path."""
    sktopt, dev = gpu
    from sktopt.filters.helmholtz_filter_nodal import HelmholtzFilterNodal
    mesh = sktopt.mesh.toy_problem.create_box_hex(2.8, 2.0, 1.6, 0.2)
    from sktopt._fem import Basis, ElementHex1
    basis = Basis(mesh, ElementHex1())
    vol = np.full(mesh.nelements, 0.2 ** 3)
    rng = np.random.default_rng(5)
    design = np.ones(mesh.nelements, dtype=bool)
    design[rng.choice(mesh.nelements, 40, replace=False)] = False
    rho = rng.uniform(0.05, 1.0, mesh.nelements)
    v = -rng.uniform(0.0, 1.0, mesh.nelements)
    out = {}
    # (matrix-free?, direct fast-diagonalisation solve of the adjoint system?)
    for flag, fd in (("1", "1"), ("1", "0"), ("0", "1")):
        monkeypatch.setenv("SKTOPT_B200_MATFREE", flag)
        monkeypatch.setenv("SKTOPT_B200_HELMHOLTZ_FD", fd)
        f = HelmholtzFilterNodal.from_defaults(mesh, vol, radius=0.35, design_mask=design)
        st = f._device()
        assert (st.grid is not None) == (flag == "1")
        assert (st.fd is not None) == (flag == "1" and fd == "1")
        out[flag + fd] = (f.forward(rho), f.gradient(v), list(st.solve_iters))
    ref = out["01"]                            # assembled CSR + PCG
    for key in ("11", "10"):
        for k in (0, 1):
            a, b = out[key][k], ref[k]
            assert np.max(np.abs(a - b)) <= 1e-9 * max(1.0, np.max(np.abs(b)))
    assert out["10"][2] == ref[2]              # same PCG iteration counts
    assert out["11"][2] == [ref[2][0], 0]      # forward: PCG (fixed nodes); adjoint: direct
