"""CPU ORACLE (test infrastructure): density filters.  PARITY UNPINNED (see
fem.py).  Restates filters/helmholtz_filter_nodal.py:26-56,121-233 and
filters/spacial.py:19-166 with NumPy / SciPy."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial import cKDTree

from . import fem


def element_to_node(t, volumes, rho_elem, design_mask, fixed_value):
    """element_to_node_density_averaging (helmholtz_filter_nodal.py:30-56):
    volume-weighted nodal average; non-design elements contribute fixed_value."""
    n_nodes = int(t.max()) + 1
    val = rho_elem if design_mask is None else np.where(design_mask, rho_elem, fixed_value)
    num = np.zeros(n_nodes)
    den = np.zeros(n_nodes)
    for a in range(t.shape[0]):
        np.add.at(num, t[a], volumes * val)
        np.add.at(den, t[a], volumes)
    den[den == 0.0] = 1.0
    return num / den


def node_to_element(t, x_node):
    """node_to_element_density (:26-27)."""
    return np.mean(x_node[t], axis=0)


class HelmholtzOracle:
    """A = M + r^2 K on the scalar P1/Q1 basis with skfem's default quadrature
    (:128,:136-145); forward pins the nodes of non-design elements to 1 (:132,
    :147-151); the adjoint has no Dirichlet nodes and clamps to <= 0 (:232).
    The reference re-assembles and calls spsolve per application; the oracle
    factorises once per radius (same linear systems)."""

    def __init__(self, p, t, volumes, design_mask=None, solver="splu", cg_rtol=1e-12,
                 matrices=None):
        """``solver='cg'``: the same linear systems solved by scipy cg + Jacobi at
        ``cg_rtol`` instead of a sparse LU (for meshes where the 3-D factorisation
        does not fit: the CPU baseline at BASELINE's full sizes)."""
        self.solver, self.cg_rtol = solver, cg_rtol
        self.p, self.t, self.vol = p, t, volumes
        self.mask = None if design_mask is None else np.asarray(design_mask, bool)
        io = fem.default_intorder(t.shape[0])
        if matrices is not None:      # (M, K) assembled elsewhere (oracle.cport at full size)
            self.M, self.K = matrices
        else:
            self.M = fem.assemble_scalar(p, t, None, io, "mass")
            self.K = fem.assemble_scalar(p, t, None, io, "laplace")
        n = p.shape[1]
        if self.mask is None:
            self.fixed = np.array([], dtype=np.int64)
        else:
            self.fixed = np.unique(t[:, ~self.mask].ravel())
        self.free = np.setdiff1d(np.arange(n), self.fixed)
        self.radius = None

    def set_radius(self, r):
        if r == self.radius:
            return
        if self.solver == "cg":
            A = (self.M + (r ** 2) * self.K).tocsr()
            self.A = A
            self.lu_full = _CgSolve(A, self.cg_rtol)
            if self.fixed.size:
                Af = A[self.free]
                self.A_free_fixed = Af[:, self.fixed].tocsr()
                self.lu_free = _CgSolve(Af[:, self.free].tocsr(), self.cg_rtol)
            self.radius = r
            return
        self.A = (self.M + (r ** 2) * self.K).tocsc()
        self.lu_full = spla.splu(self.A)
        if self.fixed.size:
            self.lu_free = spla.splu(self.A[self.free][:, self.free].tocsc())
        self.radius = r

    def forward(self, rho_elem):
        rho_n = element_to_node(self.t, self.vol, rho_elem, self.mask, 1.0)
        b = self.M @ rho_n
        if self.fixed.size:
            x = np.zeros(self.p.shape[1])
            x[self.fixed] = 1.0
            Aff = (self.A_free_fixed if self.solver == "cg"
                   else self.A[self.free][:, self.fixed])
            rhs = b[self.free] - (Aff @ x[self.fixed])
            x[self.free] = self.lu_free.solve(rhs)
        else:
            x = self.lu_full.solve(b)
        return node_to_element(self.t, x)

    def gradient(self, v_elem):
        v_n = element_to_node(self.t, self.vol, v_elem, self.mask, 0.0)
        x = self.lu_full.solve(self.M @ v_n)
        return np.minimum(node_to_element(self.t, x), 0.0)


class _CgSolve:
    """``.solve(b)`` by scipy cg + Jacobi (stands in for ``splu(A).solve``)."""

    def __init__(self, A, rtol):
        self.A, self.rtol = A, rtol
        d = 1.0 / A.diagonal()
        self.M = spla.LinearOperator(A.shape, matvec=lambda x: d * x)
        self.iters = []

    def solve(self, b):
        n = [0]

        def cb(_):
            n[0] += 1
        x, info = spla.cg(self.A, b, M=self.M, rtol=self.rtol, atol=0.0, maxiter=5000,
                          callback=cb)
        if info != 0:
            raise RuntimeError("Helmholtz cg did not converge")
        self.iters.append(n[0])
        return x


class SpatialOracle:
    """SpacialFilter (filters/spacial.py:106-166): Gaussian weights with
    s = r/3, support r, design elements only, no volume weighting, rows
    normalised; literal neighbour loop via query_ball_point (:69-85)."""

    def __init__(self, p, t, design_mask=None):
        self.c = np.mean(p[:, t], axis=1)
        self.mask = None if design_mask is None else np.asarray(design_mask, bool)
        self.radius = None

    def set_radius(self, r):
        if r == self.radius:
            return
        c = self.c
        n_all = c.shape[1]
        mask = np.ones(n_all, bool) if self.mask is None else self.mask
        ids = np.nonzero(mask)[0]
        tree = cKDTree(c.T)
        s = r / 3.0
        rows, cols, data = [], [], []
        nbrs = tree.query_ball_point(c[:, ids].T, 3.0 * s)
        for row, (i, js) in enumerate(zip(ids, nbrs)):
            js = np.asarray(js, dtype=int)
            js = js[mask[js]]
            d = np.linalg.norm(c[:, js] - c[:, [i]], axis=0)
            rows.append(np.full(js.size, row))
            cols.append(np.searchsorted(ids, js))
            data.append(np.exp(-0.5 * (d / s) ** 2))
        W = sp.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(ids.size, ids.size)).tocsr()
        rs = np.asarray(W.sum(axis=1)).ravel()
        rs[rs == 0.0] = 1.0
        self.W = (sp.diags(1.0 / rs) @ W).tocsr()
        self.ids = ids
        self.radius = r

    def forward(self, rho):
        out = rho.copy()
        if self.mask is None:
            out[:] = self.W @ rho
        else:
            out[self.mask] = self.W @ rho[self.mask]
        return out

    def gradient(self, v):
        if self.mask is None:
            return self.W.T @ v
        out = np.zeros(self.c.shape[1])
        out[self.mask] = self.W.T @ v[self.mask]
        return out


class HelmholtzElementOracle:
    """HelmholtzFilterElement (filters/helmholtz_filter_element.py:195-317,448-536)
    with the sparse LU option: A = V + r^2 L on the face-adjacency graph, L the
    Gaussian-weighted graph Laplacian (already scaled by r^2 at :216-218), V =
    diag(volume / mean volume); forward and gradient both solve A x = V (.)."""

    def __init__(self, p, t, radius):
        import itertools
        nen, ne = t.shape
        vol = np.abs(fem.element_volumes(p, t))
        # two elements are neighbours iff they share a whole face, i.e. exactly 4
        # (hex) / 3 (tet) vertices.  (The reference's hexahedral face table, :139-146,
        # presumes VTK vertex order; under scikit-fem's order its 4-tuples are not
        # faces and match nothing, which leaves L = 0 -- the module is not exported
        # upstream.  The geometric faces are used here and in the product.)
        k = 4 if nen == 8 else 3
        faces = [list(c) for c in itertools.combinations(range(nen), k)]
        owner = {}
        pairs = set()
        for e in range(ne):
            for f in faces:
                key = tuple(sorted(int(t[a, e]) for a in f))
                for o in owner.setdefault(key, []):
                    pairs.add((o, e))
                owner[key].append(e)
        cen = np.mean(p[:, t], axis=1)
        rows, cols, data = [], [], []
        diag = np.zeros(ne)
        for i, j in sorted(pairs):
            if len(set(t[:, i].tolist()) & set(t[:, j].tolist())) != k:
                continue
            d = np.linalg.norm(cen[:, i] - cen[:, j])
            if d < 1e-12:
                continue
            w = radius ** 2 * np.exp(-d ** 2 / (2 * radius ** 2))
            rows += [i, j]
            cols += [j, i]
            data += [-w, -w]
            diag[i] += w
            diag[j] += w
        L = sp.coo_matrix((data + diag.tolist(), (rows + list(range(ne)), cols + list(range(ne)))),
                          shape=(ne, ne)).tocsc()
        self.V = sp.diags(vol / vol.mean(), format="csc")
        self.A = (self.V + radius ** 2 * L).tocsc()
        self.lu = spla.splu(self.A)

    def forward(self, rho):
        return self.lu.solve(self.V @ rho)

    gradient = forward
