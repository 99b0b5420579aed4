"""CPU ORACLE (test infrastructure, NOT product code): ctypes front of
``cport.c``, the C / OpenMP restatement of assembly, Jacobi-PCG and element
energy that lets the CPU baseline run BASELINE's full-size meshes on all host
cores (bench.py ``--impl reference`` and the ``cpu_baseline`` leg).

The NumPy/SciPy oracle (``oracle/fem.py``) stays the parity checker; this port
is validated against it in ``tests/test_oracle.py`` (same K to 1e-12 relative,
same PCG iterates) and is only ever the thing *timed*, never the thing shipped.

The library is compiled on first use with the host's gcc into
``oracle/cport/_build/`` (git-ignored); nothing of /root/reference is needed.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as sp

from .. import fem

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "cport.c")
_SO = os.path.join(_HERE, "_build", "libcport.so")
_LIB = None


def build(force: bool = False) -> str:
    if (force or not os.path.exists(_SO)
            or os.path.getmtime(_SO) < os.path.getmtime(_SRC)):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        tmp = _SO + ".tmp%d" % os.getpid()
        subprocess.run(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", "-o", tmp, _SRC, "-lm"],
                       check=True)
        os.replace(tmp, _SO)
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.cport_num_threads.restype = C.c_int
        L.cport_pcg_jacobi.restype = C.c_int64
        _LIB = L
    return _LIB


def num_threads() -> int:
    return int(lib().cport_num_threads())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CBackend:
    """Assembly / solve / energy of one mesh + physics through the C port.

    ``p`` (3, n_nodes), ``t`` (nen, n_elem) as in the oracle; ``dpn`` dofs per
    node (3 elasticity, 1 scalar); ``Ke`` unit element matrices: (1, nde, nde)
    when every element has the same geometry, else (n_elem, nde, nde);
    ``dirichlet`` dof indices enforced as identity rows / columns."""

    def __init__(self, p, t, dpn, Ke, dirichlet=None):
        self.L = lib()
        self.p = p
        self.t = np.ascontiguousarray(t, dtype=np.int32)
        self.nen, self.ne = self.t.shape
        self.nn = int(p.shape[1])
        self.dpn = int(dpn)
        self.n = self.dpn * self.nn
        self.Ke = np.ascontiguousarray(Ke, dtype=np.float64)
        self.per_elem = int(self.Ke.shape[0] == self.ne and self.ne > 1)
        # node graph (union of element couplings, sorted) and node -> element lists
        ne, nen = self.ne, self.nen
        B = sp.coo_matrix((np.ones(nen * ne, dtype=np.int8),
                           (self.t.ravel().astype(np.int64), np.tile(np.arange(ne), nen))),
                          shape=(self.nn, ne)).tocsr()
        B.sort_indices()
        self.n2e_ptr = B.indptr.astype(np.int64)
        self.n2e_elem = B.indices.astype(np.int32)
        # local vertex of node n in element e
        loc = np.empty(self.n2e_elem.size, dtype=np.uint8)
        rows = np.repeat(np.arange(self.nn), np.diff(self.n2e_ptr))
        for a in range(nen):
            hit = self.t[a, self.n2e_elem] == rows
            loc[hit] = a
        self.n2e_loc = loc
        N = (B.astype(np.int32) @ B.astype(np.int32).T).tocsr()
        N.sort_indices()
        self.nptr = N.indptr.astype(np.int64)
        self.ncol = N.indices.astype(np.int32)
        nnz = self.dpn * self.dpn * int(self.nptr[-1])
        self.indptr = np.empty(self.n + 1, dtype=np.int64)
        self.indices = np.empty(nnz, dtype=np.int32)
        self.L.cport_expand_pattern(C.c_int64(self.nn), C.c_int(self.dpn), _p(self.nptr),
                                    _p(self.ncol), _p(self.indptr), _p(self.indices))
        self.data = np.empty(nnz, dtype=np.float64)
        self.mask = None
        if dirichlet is not None and len(dirichlet):
            self.mask = np.zeros(self.n, dtype=np.uint8)
            self.mask[np.asarray(dirichlet, dtype=np.int64)] = 1

    def assemble(self, scale, enforce=True):
        """K = sum_e scale_e Ke (enforced when the backend has a Dirichlet set):
        a SciPy CSR matrix that shares the backend's value buffer."""
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        self.L.cport_assemble(
            C.c_int64(self.nn), C.c_int64(self.ne), C.c_int(self.nen), C.c_int(self.dpn),
            _p(self.t), _p(self.n2e_ptr), _p(self.n2e_elem), _p(self.n2e_loc), _p(self.nptr),
            _p(self.ncol), _p(self.Ke), None, C.c_int(self.per_elem), _p(sc),
            _p(self.mask if enforce else None), _p(self.data))
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=(self.n, self.n))

    def spmv(self, x, out=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n) if out is None else out
        self.L.cport_spmv(C.c_int64(self.n), _p(self.indptr), _p(self.indices), _p(self.data),
                          _p(x), _p(y))
        return y

    def pcg(self, b, rtol=1e-8, maxiter=None):
        """scipy cg + Jacobi semantics on the last assembled matrix.
        Returns (x, iterations, relres)."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty(self.n)
        rel = C.c_double()
        mi = 10 * self.n if maxiter is None else int(maxiter)
        it = self.L.cport_pcg_jacobi(C.c_int64(self.n), _p(self.indptr), _p(self.indices),
                                     _p(self.data), _p(b), _p(x), C.c_double(rtol),
                                     C.c_int64(mi), C.byref(rel))
        return x, int(it), float(rel.value)

    def energy(self, scale, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        out = np.empty(self.ne)
        Ke = self.Ke
        cls = None
        if self.per_elem:
            cls = np.arange(self.ne, dtype=np.int32)
        self.L.cport_energy(C.c_int64(self.ne), C.c_int(self.nen), C.c_int(self.dpn), _p(self.t),
                            _p(Ke), _p(cls), _p(sc), _p(u), _p(out))
        return out


def unit_elasticity_ke(p, t, nu, intorder=2):
    """Unit (E = 1) element matrices: one if all elements are translates of the
    first (create_box_hex meshes), else one per element."""
    x = p[:, t]                                   # (3, nen, ne)
    d = x - x[:, :1, :]
    uniform = bool(np.all(np.abs(d - d[:, :, :1]) <= 1e-12 * np.abs(d).max()))
    sel = t[:, :1] if uniform else t
    one = np.ones(sel.shape[1])
    lam = nu * one / ((1.0 + nu) * (1.0 - 2.0 * nu))
    mu = one / (2.0 * (1.0 + nu))
    out = np.empty((sel.shape[1], 3 * t.shape[0], 3 * t.shape[0]))
    for s in range(0, sel.shape[1], 20000):
        out[s:s + 20000] = fem.elasticity_ke(p, sel[:, s:s + 20000], lam[s:s + 20000],
                                             mu[s:s + 20000], intorder)
    return out


def scalar_matrices(p, t):
    """(M, K) of the Helmholtz filter (mass and Laplace matrices of the scalar
    P1/Q1 basis with skfem's default quadrature, filters/helmholtz_filter_nodal.py
    :128,136-145) assembled by the C port; copies, so both stay valid."""
    io = fem.default_intorder(t.shape[0])
    x = p[:, t]
    d = x - x[:, :1, :]
    uniform = bool(np.all(np.abs(d - d[:, :, :1]) <= 1e-12 * np.abs(d).max()))
    sel = t[:, :1] if uniform else t
    out = []
    for kind in ("mass", "laplace"):
        Ke = np.concatenate([fem.scalar_ke(p, sel[:, s:s + 100000], io, kind)
                             for s in range(0, sel.shape[1], 100000)])
        be = CBackend(p, t, 1, Ke)
        out.append(be.assemble(None, enforce=False).copy())
    return tuple(out)
