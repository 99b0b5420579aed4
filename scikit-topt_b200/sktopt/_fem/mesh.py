"""Minimal mesh substrate (host, NumPy) standing in for scikit-fem's ``Mesh``.

scikit-fem is a third-party dependency of the reference that is not installed
here, so the attributes the reference reads from ``skfem``
meshes are provided by these duck-typed classes:

* ``p`` (3, n_nodes) float64, ``t`` (nen, n_elem) int32, ``facets``, ``f2t``,
  ``nelements``, ``nvertices``, ``boundaries``, ``subdomains``
* ``facets_satisfying`` / ``elements_satisfying`` / ``with_boundaries`` /
  ``with_subdomains`` (used at reference ``mesh/toy_problem.py:65-76,132-134``)
* ``MeshHex.init_tensor`` (used at ``mesh/toy_problem.py:34``).

Numbering conventions follow SURVEY.md Appendix A.1: node id of a tensor grid
is ``iy + (ny+1)*ix + (ny+1)*(nx+1)*iz``; the hexahedron's local vertices sit
at reference coordinates v0=(0,0,0) v1=(0,0,1) v2=(0,1,0) v3=(1,0,0)
v4=(0,1,1) v5=(1,0,1) v6=(1,1,0) v7=(1,1,1).
"""
from __future__ import annotations

import numpy as np

# reference-cube coordinates of the 8 local vertices (see module docstring)
HEX_REF_VERTS = np.array(
    [
        [0, 0, 0],
        [0, 0, 1],
        [0, 1, 0],
        [1, 0, 0],
        [0, 1, 1],
        [1, 0, 1],
        [1, 1, 0],
        [1, 1, 1],
    ],
    dtype=np.float64,
)

# local faces of the hexahedron, vertices in cyclic order.
# (fixed axis, fixed value) of the reference cube for each face
HEX_FACES = np.array(
    [
        [0, 1, 4, 2],  # X = 0
        [3, 5, 7, 6],  # X = 1
        [0, 1, 5, 3],  # Y = 0
        [2, 4, 7, 6],  # Y = 1
        [0, 2, 6, 3],  # Z = 0
        [1, 4, 7, 5],  # Z = 1
    ],
    dtype=np.int64,
)
HEX_FACE_AXIS = np.array([0, 0, 1, 1, 2, 2])
HEX_FACE_SIDE = np.array([0.0, 1.0, 0.0, 1.0, 0.0, 1.0])

TET_FACES = np.array(
    [
        [0, 1, 2],
        [0, 1, 3],
        [0, 2, 3],
        [1, 2, 3],
    ],
    dtype=np.int64,
)


def _unique_columns(sorted_cols: np.ndarray):
    """Unique columns of an integer array whose columns are sorted tuples.

    Returns (unique (k, n_unique), first_index (n_unique,), inverse (n,)).
    Lexicographic order with row 0 as the primary key, as ``np.unique(axis=1)``.
    """
    k, n = sorted_cols.shape
    if k == 4 and sorted_cols.dtype == np.int64 and n and 0 <= sorted_cols.min() \
            and sorted_cols.max() < 2 ** 31:
        # two packed keys sort like the four rows (non-negative ids below 2^31)
        hi = (sorted_cols[0] << 32) | sorted_cols[1]
        lo = (sorted_cols[2] << 32) | sorted_cols[3]
        order = np.lexsort((lo, hi))
    else:
        order = np.lexsort(sorted_cols[::-1])
    s = sorted_cols[:, order]
    new = np.ones(n, dtype=bool)
    if n > 1:
        new[1:] = np.any(s[:, 1:] != s[:, :-1], axis=0)
    group = np.cumsum(new) - 1
    inverse = np.empty(n, dtype=np.int64)
    inverse[order] = group
    first_sorted = np.nonzero(new)[0]
    # representative = smallest original index in each group (stable lexsort)
    first_index = order[first_sorted]
    return s[:, first_sorted], first_index, inverse


def _mean_of_vertices(p: np.ndarray, conn: np.ndarray) -> np.ndarray:
    """``p[:, conn].mean(axis=1)`` without the (3, k, n) temporary: the reduction over
    the k vertices of a row adds them in order and divides by k, which is what NumPy's
    mean does along a non-contiguous axis of that length (bit-identical; checked in
    tests/test_host_helpers.py)."""
    acc = p[:, conn[0]]                       # fancy indexing: a fresh array
    for k in range(1, conn.shape[0]):
        acc += p[:, conn[k]]
    acc /= float(conn.shape[0])
    return acc


class Mesh:
    """Base class; see module docstring."""

    nen = 0
    local_faces: np.ndarray = None
    cell_type = ""

    def __init__(self, p, t, boundaries=None, subdomains=None):
        self.p = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
        self.t = np.ascontiguousarray(np.asarray(t, dtype=np.int32))
        if self.p.shape[0] != 3:
            raise ValueError("p must have shape (3, n_nodes)")
        if self.t.shape[0] != self.nen:
            raise ValueError(f"t must have shape ({self.nen}, n_elem)")
        self.boundaries = boundaries
        self.subdomains = subdomains
        self._facets = None
        self._f2t = None
        self._t2f = None
        self._f2lf = None

    # -- sizes -------------------------------------------------------------
    @property
    def nelements(self) -> int:
        return self.t.shape[1]

    @property
    def nvertices(self) -> int:
        return self.p.shape[1]

    def dim(self) -> int:
        return 3

    # -- facets --------------------------------------------------------------
    def _build_facets(self):
        lf = self.local_faces
        nf_loc, nfv = lf.shape
        ne = self.nelements
        # (nfv, nf_loc*ne): face-major blocks like a horizontal stack per local face
        allf = np.hstack([self.t[lf[i]].astype(np.int64) for i in range(nf_loc)])
        allf_sorted = np.sort(allf, axis=0)
        facets, first_idx, inverse = _unique_columns(allf_sorted)
        self._facets = facets.astype(np.int32)
        self._t2f = inverse.reshape(nf_loc, ne).astype(np.int32)
        nfac = facets.shape[1]
        f2t = np.full((2, nfac), -1, dtype=np.int32)
        f2lf = np.full((2, nfac), -1, dtype=np.int8)
        elem_of = np.tile(np.arange(ne, dtype=np.int64), nf_loc)
        lface_of = np.repeat(np.arange(nf_loc, dtype=np.int64), ne)
        # first occurrence (smallest stacked index) -> slot 0, other -> slot 1
        order = np.argsort(inverse, kind="stable")
        inv_sorted = inverse[order]
        is_first = np.ones(order.size, dtype=bool)
        is_first[1:] = inv_sorted[1:] != inv_sorted[:-1]
        f2t[0, inv_sorted[is_first]] = elem_of[order[is_first]]
        f2lf[0, inv_sorted[is_first]] = lface_of[order[is_first]]
        f2t[1, inv_sorted[~is_first]] = elem_of[order[~is_first]]
        f2lf[1, inv_sorted[~is_first]] = lface_of[order[~is_first]]
        self._f2t = f2t
        self._f2lf = f2lf

    @property
    def facets(self) -> np.ndarray:
        if self._facets is None:
            self._build_facets()
        return self._facets

    @property
    def f2t(self) -> np.ndarray:
        if self._f2t is None:
            self._build_facets()
        return self._f2t

    @property
    def t2f(self) -> np.ndarray:
        if self._t2f is None:
            self._build_facets()
        return self._t2f

    @property
    def f2lf(self) -> np.ndarray:
        """Local face index of each facet inside f2t[0] / f2t[1]."""
        if self._f2lf is None:
            self._build_facets()
        return self._f2lf

    def boundary_facets(self) -> np.ndarray:
        return np.nonzero(self.f2t[1] == -1)[0].astype(np.int32)

    def facets_satisfying(self, test, boundaries_only: bool = False):
        # the midpoints are shared by every query on this mesh (clamp, load, ...)
        if getattr(self, "_facet_midp", None) is None:
            self._facet_midp = _mean_of_vertices(self.p, self.facets)
        midp = self._facet_midp
        facets = np.nonzero(test(midp))[0]
        if boundaries_only:
            facets = np.intersect1d(facets, self.boundary_facets())
        return facets.astype(np.int32)

    def elements_satisfying(self, test):
        midp = _mean_of_vertices(self.p, self.t)
        return np.nonzero(test(midp))[0].astype(np.int32)

    def _clone(self, boundaries, subdomains):
        m = type(self)(self.p, self.t, boundaries=boundaries, subdomains=subdomains)
        m._facets, m._f2t, m._t2f, m._f2lf = (
            self._facets, self._f2t, self._t2f, self._f2lf
        )
        m._facet_midp = getattr(self, "_facet_midp", None)
        return m

    def with_boundaries(self, boundaries: dict, boundaries_only: bool = True):
        out = dict(self.boundaries) if self.boundaries else {}
        for name, test in boundaries.items():
            if callable(test):
                out[name] = self.facets_satisfying(test, boundaries_only)
            else:
                out[name] = np.asarray(test, dtype=np.int32)
        return self._clone(out, self.subdomains)

    def with_subdomains(self, subdomains: dict):
        out = dict(self.subdomains) if self.subdomains else {}
        for name, test in subdomains.items():
            if callable(test):
                out[name] = self.elements_satisfying(test)
            else:
                out[name] = np.asarray(test, dtype=np.int32)
        return self._clone(self.boundaries, out)

    def scaled(self, factors):
        f = np.asarray(factors, dtype=np.float64).reshape(-1, 1)
        return type(self)(self.p * f, self.t, self.boundaries, self.subdomains)


class MeshHex(Mesh):
    nen = 8
    local_faces = HEX_FACES
    cell_type = "hexahedron"

    # -- facet table without sorting, for lattice-numbered grids -------------------
    def _build_facets(self):
        if not (self._build_facets_native() or self._build_facets_lattice()):
            super()._build_facets()

    def _build_facets_native(self) -> bool:
        """``_build_facets_lattice`` in multi-threaded C++ (``sktb_host_lattice_facets``,
        csrc/host_setup.cu; host pointers, no device work): 9.6 s -> ~1 s at 8M
        elements.  False when the library is not built or the mesh does not qualify."""
        import ctypes as C
        try:
            from sktopt._b200 import lib as _lib
            lib = _lib.load()
        except Exception:
            return False
        ne, nn = self.nelements, self.nvertices
        if ne == 0:
            return False
        s0 = np.sort(self.t[:, 0].astype(np.int64))
        ny, P = int(s0[2] - s0[0]), int(s0[4] - s0[0])
        if ny < 3 or P < 3 * ny:
            return False
        t = np.ascontiguousarray(self.t, dtype=np.int32)
        lf = np.ascontiguousarray(self.local_faces, dtype=np.int32)
        key = np.empty(6 * ne, dtype=np.int64)
        nfac = C.c_int64(0)
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        if lib.sktb_host_lattice_facets(ne, nn, ptr(t), ptr(lf), ny, P, ptr(key),
                                        C.cast(C.byref(nfac), C.c_void_p),
                                        None, None, None, None) != 0 or nfac.value <= 0:
            return False
        n = int(nfac.value)
        facets = np.empty((4, n), dtype=np.int32)
        t2f = np.empty((6, ne), dtype=np.int32)
        f2t = np.empty((2, n), dtype=np.int32)
        f2lf = np.empty((2, n), dtype=np.int8)
        if lib.sktb_host_lattice_facets(ne, nn, ptr(t), ptr(lf), ny, P, ptr(key),
                                        C.cast(C.byref(nfac), C.c_void_p), ptr(facets),
                                        ptr(t2f), ptr(f2t), ptr(f2lf)) != 0:
            return False
        self._facets, self._t2f, self._f2t, self._f2lf = facets, t2f, f2t, f2lf
        return True

    def _build_facets_lattice(self) -> bool:
        """The same ``facets`` / ``t2f`` / ``f2t`` / ``f2lf`` as the generic sort-based
        construction, in O(n) streaming passes, when every element is a cell of a
        lattice numbered like ``init_tensor`` (node = iy + npy*ix + npy*npx*iz, any
        geometry, any valid local vertex order).  A face of such a cell is
        {n, n+a, n+b, n+a+b} with (a, b) one of (1, npy), (1, P), (npy, P), P = npy*npx:
        facets are the distinct (n, family) pairs, and their lexicographic order by
        sorted node tuple is the order of 3 n + family.  Returns False (nothing
        touched) when the mesh does not have that structure."""
        t = self.t.astype(np.int64)
        ne = t.shape[1]
        if ne == 0:
            return False
        ts = np.sort(t, axis=0)
        m = ts[0]
        ny = int(ts[2, 0] - m[0])
        P = int(ts[4, 0] - m[0])
        if ny < 3 or P < 3 * ny:
            return False
        pattern = np.array([0, 1, ny, ny + 1, P, P + 1, P + ny, P + ny + 1], dtype=np.int64)
        if not np.array_equal(ts - m, np.broadcast_to(pattern[:, None], ts.shape)):
            return False
        del ts
        lf = self.local_faces
        nf_loc = lf.shape[0]
        span_of = np.array([ny + 1, P + 1, P + ny], dtype=np.int64)
        base = np.empty((nf_loc, ne), dtype=np.int64)
        fam = np.empty((nf_loc, ne), dtype=np.int64)
        for i in range(nf_loc):
            q = t[lf[i]]
            b = q.min(axis=0)
            span = q.max(axis=0) - b
            f = np.where(span == span_of[0], 0, np.where(span == span_of[1], 1, 2))
            if not (np.array_equal(span, span_of[f])
                    and np.array_equal(q.sum(axis=0), 4 * b + 2 * span)):
                return False
            base[i], fam[i] = b, f
        n_nodes = self.nvertices
        key = (3 * base + fam).ravel()                    # stacked index = lface * ne + elem
        present = np.zeros(3 * n_nodes, dtype=bool)
        present[key] = True
        rank = np.cumsum(present, dtype=np.int64) - 1
        fid = rank[key]
        ids = np.nonzero(present)[0]
        nfac = ids.size
        fb, ff = ids // 3, ids % 3
        a = np.array([1, 1, ny], dtype=np.int64)[ff]
        b = np.array([ny, P, P], dtype=np.int64)[ff]
        facets = np.empty((4, nfac), dtype=np.int32)
        facets[0], facets[1], facets[2], facets[3] = fb, fb + a, fb + b, fb + a + b
        # the (at most two) elements of a facet: slot 0 = the smaller stacked index
        stacked = np.arange(key.size, dtype=np.int64)
        occ1 = np.full(nfac, -1, dtype=np.int64)
        occ1[fid] = stacked                               # some occurrence of every facet
        other = occ1[fid] != stacked
        if np.bincount(fid[other], minlength=nfac).max(initial=0) > 1:
            return False                                  # a facet with > 2 elements
        occ2 = np.full(nfac, -1, dtype=np.int64)
        occ2[fid[other]] = stacked[other]
        two = occ2 >= 0
        first = np.where(two, np.minimum(occ1, occ2), occ1)
        second = np.where(two, np.maximum(occ1, occ2), -1)
        f2t = np.full((2, nfac), -1, dtype=np.int32)
        f2lf = np.full((2, nfac), -1, dtype=np.int8)
        f2t[0], f2lf[0] = first % ne, first // ne
        f2t[1, two], f2lf[1, two] = second[two] % ne, second[two] // ne
        self._facets = facets
        self._t2f = fid.reshape(nf_loc, ne).astype(np.int32)
        self._f2t, self._f2lf = f2t, f2lf
        return True

    @classmethod
    def init_tensor(cls, x, y, z):
        """Tensor-product hexahedral grid (SURVEY.md Appendix A.1 numbering)."""
        x = np.sort(np.asarray(x, dtype=np.float64))
        y = np.sort(np.asarray(y, dtype=np.float64))
        z = np.sort(np.asarray(z, dtype=np.float64))
        npx, npy, npz = len(x), len(y), len(z)
        # node id = iy + npy*ix + npy*npx*iz
        iz, ix, iy = np.meshgrid(
            np.arange(npz), np.arange(npx), np.arange(npy), indexing="ij"
        )
        p = np.vstack((x[ix.ravel()], y[iy.ravel()], z[iz.ravel()]))

        def nid(jy, jx, jz):
            return jy + npy * jx + npy * npx * jz

        # element id = ey + (npy-1)*ex + (npy-1)*(npx-1)*ez
        ez, ex, ey = np.meshgrid(
            np.arange(npz - 1), np.arange(npx - 1), np.arange(npy - 1),
            indexing="ij",
        )
        ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
        t = np.empty((8, ex.size), dtype=np.int64)
        t[0] = nid(ey, ex, ez)
        t[1] = nid(ey + 1, ex, ez)
        t[2] = nid(ey, ex + 1, ez)
        t[3] = nid(ey, ex, ez + 1)
        t[4] = nid(ey + 1, ex + 1, ez)
        t[5] = nid(ey + 1, ex, ez + 1)
        t[6] = nid(ey, ex + 1, ez + 1)
        t[7] = nid(ey + 1, ex + 1, ez + 1)
        return cls(p, t.astype(np.int32))


class MeshTet(Mesh):
    nen = 4
    local_faces = TET_FACES
    cell_type = "tetra"

    @classmethod
    def init_tensor(cls, x, y, z):
        """Tensor grid split into 6 Kuhn tetrahedra per cell."""
        hexm = MeshHex.init_tensor(x, y, z)
        th = hexm.t.astype(np.int64)
        # corner c(dx,dy,dz) of each cell in terms of MeshHex local vertices:
        # local k sits at (X,Y,Z) ref -> phys (x,y,z) = (Y, Z, X)
        # c000=0 c010=1 c100=2 c001=3 c110=4 c011=5 c101=6 c111=7
        c = {
            (0, 0, 0): th[0], (0, 1, 0): th[1], (1, 0, 0): th[2],
            (0, 0, 1): th[3], (1, 1, 0): th[4], (0, 1, 1): th[5],
            (1, 0, 1): th[6], (1, 1, 1): th[7],
        }
        import itertools
        tets = []
        for perm in itertools.permutations(range(3)):
            cur = [0, 0, 0]
            path = [tuple(cur)]
            for ax in perm:
                cur[ax] = 1
                path.append(tuple(cur))
            tets.append(np.vstack([c[v] for v in path]))
        # interleave so that the 6 tets of a cell are contiguous
        t = np.stack(tets, axis=2).reshape(4, -1)
        return cls(hexm.p, t.astype(np.int32))
