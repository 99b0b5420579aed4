"""CPU ORACLE (test infrastructure, not product code): NumPy / SciPy restatement
of the reference's assembly, enforce, solve and element-energy path.

PARITY UNPINNED: the reference (scikit-topt 0.3.9) cannot be imported here
(scikit-fem / pyamg absent) and its own tests pin no numbers on this path
(SURVEY.md 8c), so this restatement is pinned only by analytic known-answer
checks (tests/test_oracle.py) and by the reference's scheduler scalars.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this package.  Every function cites the reference lines it follows (paths
relative to /root/reference/scikit-topt/sktopt/); scikit-fem semantics are
those of SURVEY.md Appendix A.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# reference-cube coordinates of skfem's ElementHex1 local vertices (App. A.1)
HEX_VERTS = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0],
                      [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1]], dtype=float)


# ------------------------------------------------------------- quadrature --
def gauss_unit(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def quadrature(nen: int, intorder: int):
    """skfem Basis quadrature (App. A.1): hex = tensor Gauss with
    ceil((intorder+1)/2) points per axis on [0,1]^3; tet = 1- or 4-point rule."""
    if nen == 8:
        n = int(np.ceil((intorder + 1) / 2.0))
        g, w = gauss_unit(n)
        pts = np.array([[a, b, c] for a in g for b in g for c in g]).T
        wts = np.array([wa * wb * wc for wa in w for wb in w for wc in w])
        return pts, wts
    if intorder <= 1:
        return np.full((3, 1), 0.25), np.array([1.0 / 6.0])
    a, b = 0.5854101966249685, 0.1381966011250105
    return np.array([[a, b, b, b], [b, a, b, b], [b, b, a, b]]), np.full(4, 1.0 / 24.0)


def default_intorder(nen: int) -> int:
    """skfem default: 2 * element maxdeg (Hex1 -> 6, TetP1 -> 2)."""
    return 6 if nen == 8 else 2


def shape(nen: int, X: np.ndarray):
    """N (nen, nq), dN/dX (nen, 3, nq) on the reference element."""
    nq = X.shape[1]
    if nen == 8:
        N = np.ones((8, nq))
        dN = np.ones((8, 3, nq))
        for a in range(8):
            for d in range(3):
                f = X[d] if HEX_VERTS[a, d] else 1.0 - X[d]
                s = 1.0 if HEX_VERTS[a, d] else -1.0
                N[a] *= f
                for k in range(3):
                    dN[a, k] *= s if k == d else f
        return N, dN
    N = np.vstack([1 - X[0] - X[1] - X[2], X[0], X[1], X[2]])
    dN = np.zeros((4, 3, nq))
    dN[0] = -1.0
    dN[1, 0] = dN[2, 1] = dN[3, 2] = 1.0
    return N, dN


def physical_gradients(p, t, X):
    """Per element / quadrature point: N (nen,nq), G (ne,nq,nen,3) physical
    gradients and |det J| (ne,nq) of the isoparametric map."""
    nen = t.shape[0]
    N, dN = shape(nen, X)
    xe = p[:, t]                                    # (3, nen, ne)
    J = np.einsum("dae,akq->eqdk", xe, dN)          # (ne,nq,3,3)  dx_d/dX_k
    detJ = np.linalg.det(J)
    invJ = np.linalg.inv(J)                         # dX_k/dx_d at [k,d]
    G = np.einsum("akq,eqkd->eqad", dN, invJ)       # (ne,nq,nen,3)
    return N, G, np.abs(detJ)


# ------------------------------------------------------------- materials ---
def simp(rho, E0, Emin, p):
    """fea/composer.py:19-22"""
    return Emin + (E0 - Emin) * rho ** p


def ramp(rho, E0, Emin, p):
    """fea/composer.py:25-39"""
    return Emin + (E0 - Emin) * (rho / (1.0 + p * (1.0 - rho)))


# ------------------------------------------------------------- element Ke --
def elasticity_ke(p, t, lam, mu, intorder=2):
    """Element matrices of  lam tr e(u) tr e(v) + 2 mu e(u):e(v)
    (fea/composer.py:82-98); rows = test dof 3a+i, cols = trial dof 3b+j.
    Returns (ne, 3nen, 3nen)."""
    nen, ne = t.shape
    X, W = quadrature(nen, intorder)
    _, G, dJ = physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    K = np.zeros((ne, nen, 3, nen, 3))
    gg = np.einsum("eqad,eqbd,eq->eab", G, G, dx)
    K += np.einsum("eqai,eqbj,eq->eaibj", G, G, dx) * lam[:, None, None, None, None]
    K += np.einsum("eqaj,eqbi,eq->eaibj", G, G, dx) * mu[:, None, None, None, None]
    for i in range(3):
        K[:, :, i, :, i] += gg * mu[:, None, None]
    return K.reshape(ne, 3 * nen, 3 * nen)


def scalar_ke(p, t, intorder, kind):
    """Scalar element matrices: 'laplace' grad u . grad v (fea/composer.py:139-141)
    or 'mass' u v (filters/helmholtz_filter_nodal.py:136-138). (ne, nen, nen)."""
    nen = t.shape[0]
    X, W = quadrature(nen, intorder)
    N, G, dJ = physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    if kind == "laplace":
        return np.einsum("eqad,eqbd,eq->eab", G, G, dx)
    return np.einsum("aq,bq,eq->eab", N, N, dx)


def structural_pattern(t, n_nodes, dpn):
    """Union of element couplings as an all-ones CSR matrix with sorted
    indices (the pattern skfem's COO -> CSR produces, App. A.1)."""
    nen, ne = t.shape
    B = sp.coo_matrix((np.ones(nen * ne), (t.ravel().astype(np.int64),
                                            np.tile(np.arange(ne), nen))),
                      shape=(n_nodes, ne)).tocsr()
    N = (B @ B.T).tocsr()
    N.data[:] = 1.0
    P = sp.kron(N, np.ones((dpn, dpn)), format="csr") if dpn > 1 else N
    P.sort_indices()
    return P


def assemble(t, Ke, dpn, pattern=None, chunk=20000):
    """skfem asm (App. A.1): element contributions rows = element_dofs[i],
    cols = element_dofs[j] summed into CSR with sorted indices.  Entries whose
    contributions cancel to 0.0 stay in the pattern (explicit zeros)."""
    nen, ne = t.shape
    n_nodes = int(t.max()) + 1 if pattern is None else pattern.shape[0] // dpn
    P = structural_pattern(t, n_nodes, dpn) if pattern is None else pattern
    n = P.shape[0]
    keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(P.indptr)) * n + P.indices
    data = np.zeros(P.nnz)
    nde = nen * dpn
    for s in range(0, ne, chunk):
        tc = t[:, s:s + chunk].astype(np.int64)
        nc = tc.shape[1]
        ed = (dpn * tc[:, None, :] + np.arange(dpn)[None, :, None]).reshape(nde, nc)
        rows = np.repeat(ed.T[:, :, None], nde, axis=2).ravel()
        cols = np.repeat(ed.T[:, None, :], nde, axis=1).ravel()
        Kc = sp.coo_matrix((Ke[s:s + chunk].ravel(), (rows, cols)), shape=(n, n)).tocsr()
        Kc.sort_indices()
        ck = np.repeat(np.arange(n, dtype=np.int64), np.diff(Kc.indptr)) * n + Kc.indices
        data[np.searchsorted(keys, ck)] += Kc.data
    return sp.csr_matrix((data, P.indices.copy(), P.indptr.copy()), shape=(n, n))


def assemble_stiffness(p, t, rho, E0, Emin, pw, nu, intorder=2, interp=simp,
                       chunk=20000, pattern=None):
    """composer.assemble_stiffness_matrix (fea/composer.py:53-101)."""
    E = interp(rho, E0, Emin, pw)
    lam = (nu * E) / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 * (1.0 + nu))
    P = structural_pattern(t, p.shape[1], 3) if pattern is None else pattern
    n = P.shape[0]
    keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(P.indptr)) * n + P.indices
    data = np.zeros(P.nnz)
    for s in range(0, t.shape[1], chunk):
        sl = slice(s, min(s + chunk, t.shape[1]))
        Ke = elasticity_ke(p, t[:, sl], lam[sl], mu[sl], intorder)
        Kc = assemble(t[:, sl], Ke, 3, pattern=P, chunk=chunk)
        data += Kc.data
    return sp.csr_matrix((data, P.indices.copy(), P.indptr.copy()), shape=(n, n))


def assemble_scalar(p, t, coeff, intorder, kind, chunk=100000, pattern=None):
    """composer.assemble_conduction_matrix (fea/composer.py:104-145) for
    kind='laplace' with coeff = k_e; mass matrix for kind='mass'."""
    P = structural_pattern(t, p.shape[1], 1) if pattern is None else pattern
    data = np.zeros(P.nnz)
    for s in range(0, t.shape[1], chunk):
        sl = slice(s, min(s + chunk, t.shape[1]))
        Ke = scalar_ke(p, t[:, sl], intorder, kind)
        if coeff is not None:
            Ke = Ke * np.asarray(coeff)[sl, None, None]
        data += assemble(t[:, sl], Ke, 1, pattern=P, chunk=chunk).data
    n = P.shape[0]
    return sp.csr_matrix((data, P.indices.copy(), P.indptr.copy()), shape=(n, n))


# ---------------------------------------------------------------- enforce --
def enforce(K, f, D, xD=None):
    """skfem.enforce (App. A.1): rows and columns of D zeroed, unit diagonal,
    b <- b - K[:, D] x_D, b[D] = x_D.  Pattern (explicit zeros) kept."""
    K = K.tocsr().copy()
    n = K.shape[0]
    x = np.zeros(n)
    if xD is not None:
        x[D] = xD
    b = np.asarray(f, dtype=float) - K @ x
    b[D] = x[D]
    mask = np.zeros(n, dtype=bool)
    mask[D] = True
    rows = np.repeat(np.arange(n), np.diff(K.indptr))
    hit = mask[rows] | mask[K.indices]
    K.data[hit] = 0.0
    K.data[hit & (rows == K.indices)] = 1.0
    return K, b


# ------------------------------------------------------------------ solve --
def solve(K_e, F_e, solver="cg_jacobi", rtol=1e-8, maxiter=None):
    """solve_u (fea/solver_elastic.py:61-143): 'cg_jacobi' = scipy cg with
    M = 1/diag (:84-92); 'spsolve' (:106-109).  Returns (u, info, iterations)."""
    if solver == "spsolve":
        return spla.spsolve(K_e.tocsc(), F_e), 0, 0
    Minv = 1.0 / K_e.diagonal()
    M = spla.LinearOperator(K_e.shape, matvec=lambda x: Minv * x)
    count = [0]

    def cb(_):
        count[0] += 1
    u, info = spla.cg(K_e, F_e, M=M, rtol=rtol, maxiter=maxiter, callback=cb)
    return u, info, count[0]


def compliance_single(p, t, rho, E0, Emin, pw, nu, force, D, intorder=2,
                      solver="spsolve", rtol=1e-8, maxiter=None):
    """compute_compliance_basis (fea/solver_elastic.py:146-237)."""
    K = assemble_stiffness(p, t, rho, E0, Emin, pw, nu, intorder)
    K_e, F_e = enforce(K, force, D)
    u, _, _ = solve(K_e, F_e, solver, rtol, maxiter)
    free = np.setdiff1d(np.arange(K.shape[0]), D, assume_unique=True)
    return float(F_e[free] @ u[free]), u


def compliance_multi(p, t, rho, E0, Emin, pw, nu, forces, D, intorder=2):
    """solve_multi_load + compute_compliance_basis_multi_load
    (fea/solver_elastic.py:240-467): one LU, k right-hand sides."""
    K = assemble_stiffness(p, t, rho, E0, Emin, pw, nu, intorder)
    K_e, _ = enforce(K, forces[0], D)
    F = np.column_stack([enforce(K, f, D)[1] for f in forces])
    U = spla.splu(K_e.tocsc()).solve(F)
    return np.einsum("ij,ij->j", F, U), U


# ----------------------------------------------------------------- energy --
def strain_energy(p, t, rho, U, E0, Emin, pw, nu, intorder=2, interp=simp):
    """strain_energy_skfem_multi (fea/solver_elastic.py:470-537):
    U_e = int_e 1/2 (2 mu eps + lam tr eps I):eps with the basis quadrature.
    U (n_dof, n_loads) -> (n_elem, n_loads)."""
    nen, ne = t.shape
    E = interp(rho, E0, Emin, pw)
    lam = (nu * E) / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = E / (2.0 * (1.0 + nu))
    X, W = quadrature(nen, intorder)
    _, G, dJ = physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    U = U if U.ndim == 2 else U[:, None]
    out = np.zeros((ne, U.shape[1]))
    tt = t.astype(np.int64)
    for l in range(U.shape[1]):
        ue = np.stack([U[3 * tt + c, l] for c in range(3)], axis=-1)   # (nen,ne,3)
        grad = np.einsum("aei,eqaj->eqij", ue, G)                       # du_i/dx_j
        eps = 0.5 * (grad + np.swapaxes(grad, 2, 3))
        tr = np.trace(eps, axis1=2, axis2=3)
        dens = 0.5 * (2.0 * mu[:, None] * np.einsum("eqij,eqij->eq", eps, eps)
                      + lam[:, None] * tr * tr)
        out[:, l] = np.sum(dens * dx, axis=1)
    return out


def heat_energy(p, t, rho, T, k0, kmin, pw, intorder=2, interp=simp):
    """heat_energy_skfem_multi (fea/solver_heat.py:256-303): 1/2 k_e |grad T|^2."""
    nen, ne = t.shape
    k = interp(rho, k0, kmin, pw)
    X, W = quadrature(nen, intorder)
    _, G, dJ = physical_gradients(p, t, X)
    dx = dJ * W[None, :]
    T = T if T.ndim == 2 else T[:, None]
    out = np.zeros((ne, T.shape[1]))
    for l in range(T.shape[1]):
        Te = T[t.astype(np.int64), l]                                  # (nen, ne)
        g = np.einsum("ae,eqad->eqd", Te, G)
        out[:, l] = 0.5 * k * np.sum(np.einsum("eqd,eqd->eq", g, g) * dx, axis=1)
    return out


# ---------------------------------------------------------------- volumes --
_HEX_TETS = ((0, 1, 3, 4), (1, 2, 3, 6), (1, 5, 6, 4), (3, 6, 7, 4), (1, 3, 6, 4), (1, 6, 5, 4))


def element_volumes(p, t):
    """get_elements_volume (fea/composer.py:164-248,314-322), literal formula:
    hex = sum of six |det|/6 on fixed local quadruples; tet = signed det/6."""
    def det6(q):
        v1 = p[:, t[q[1]]] - p[:, t[q[0]]]
        v2 = p[:, t[q[2]]] - p[:, t[q[0]]]
        v3 = p[:, t[q[3]]] - p[:, t[q[0]]]
        return np.einsum("de,de->e", np.cross(v1, v2, axis=0), v3) / 6.0
    if t.shape[0] == 4:
        return det6((0, 1, 2, 3))
    vol = np.zeros(t.shape[1])
    for q in _HEX_TETS:
        vol += np.abs(det6(q))
    return vol
