"""Minimal element / basis substrate standing in for scikit-fem's ``Basis``.

Only the surface the reference touches is provided:
``Basis(mesh, elem, intorder=k)``, ``.N``, ``.nodal_dofs``, ``.element_dofs``,
``.get_dofs(nodes=...)`` -> ``.all()`` / ``.nodal['u^k']``, ``.X`` / ``.W``
quadrature, ``.mesh``, ``.elem``.  DOF numbering is ``dpn*node + comp``
(reference comment at ``mesh/task_elastic.py:72``).
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh, MeshHex, MeshTet, HEX_REF_VERTS


class ElementHex1:
    maxdeg = 3
    nen = 8
    dim = 3
    dpn = 1


class ElementTetP1:
    maxdeg = 1
    nen = 4
    dim = 3
    dpn = 1


class ElementVector:
    def __init__(self, elem, dim: int = 3):
        self.elem = elem
        self.maxdeg = elem.maxdeg
        self.nen = elem.nen
        self.dim = 3
        self.dpn = dim


def gauss_legendre_unit(n: int):
    """n-point Gauss-Legendre rule on [0, 1]."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def hex_quadrature(intorder: int):
    """Tensor Gauss rule on [0,1]^3 with ceil((intorder+1)/2) points per axis.

    Returns X (3, nqp) and W (nqp,).
    """
    n = int(np.ceil((intorder + 1) / 2.0))
    x, w = gauss_legendre_unit(n)
    X0, X1, X2 = np.meshgrid(x, x, x, indexing="ij")
    W0, W1, W2 = np.meshgrid(w, w, w, indexing="ij")
    X = np.vstack([X0.ravel(), X1.ravel(), X2.ravel()])
    W = (W0 * W1 * W2).ravel()
    return X, W


def tet_quadrature(intorder: int):
    """Symmetric rules on the reference tetrahedron (volume 1/6)."""
    if intorder <= 1:
        X = np.full((3, 1), 0.25)
        W = np.array([1.0 / 6.0])
    else:
        a = 0.5854101966249685
        b = 0.1381966011250105
        X = np.array([[a, b, b, b], [b, a, b, b], [b, b, a, b]])
        W = np.full(4, 1.0 / 24.0)
    return X, W


def hex_shape(X: np.ndarray):
    """Trilinear shape functions at reference points X (3, nqp).

    Returns N (8, nqp) and dN (8, 3, nqp) w.r.t. reference coordinates.
    """
    nq = X.shape[1]
    N = np.ones((8, nq))
    dN = np.ones((8, 3, nq))
    for a in range(8):
        f = []
        df = []
        for d in range(3):
            if HEX_REF_VERTS[a, d] == 1.0:
                f.append(X[d])
                df.append(np.ones(nq))
            else:
                f.append(1.0 - X[d])
                df.append(-np.ones(nq))
        N[a] = f[0] * f[1] * f[2]
        dN[a, 0] = df[0] * f[1] * f[2]
        dN[a, 1] = f[0] * df[1] * f[2]
        dN[a, 2] = f[0] * f[1] * df[2]
    return N, dN


def tet_shape(X: np.ndarray):
    nq = X.shape[1]
    N = np.vstack([1.0 - X[0] - X[1] - X[2], X[0], X[1], X[2]])
    dN = np.zeros((4, 3, nq))
    dN[0] = -1.0
    dN[1, 0] = 1.0
    dN[2, 1] = 1.0
    dN[3, 2] = 1.0
    return N, dN


def default_intorder(elem) -> int:
    return 2 * elem.maxdeg


class DofsView:
    def __init__(self, nodes: np.ndarray, dpn: int):
        self._nodes = np.asarray(nodes, dtype=np.int64)
        self._dpn = dpn
        if dpn == 1:
            self.nodal = {"u": self._nodes.copy()}
        else:
            self.nodal = {
                f"u^{c + 1}": dpn * self._nodes + c for c in range(dpn)
            }

    def all(self) -> np.ndarray:
        n = self._nodes
        d = self._dpn
        out = (d * n[:, None] + np.arange(d)[None, :]).ravel()
        return np.unique(out)


class Basis:
    def __init__(self, mesh: Mesh, elem, intorder: int | None = None):
        self.mesh = mesh
        self.elem = elem
        self.dpn = getattr(elem, "dpn", 1)
        self.intorder = default_intorder(elem) if intorder is None else int(intorder)
        if isinstance(mesh, MeshHex):
            self.X, self.W = hex_quadrature(self.intorder)
        elif isinstance(mesh, MeshTet):
            self.X, self.W = tet_quadrature(self.intorder)
        else:
            raise NotImplementedError("MeshHex or MeshTet")
        self.N = self.dpn * mesh.nvertices

    @property
    def nodal_dofs(self) -> np.ndarray:
        n = np.arange(self.mesh.nvertices, dtype=np.int64)
        return np.vstack([self.dpn * n + c for c in range(self.dpn)])

    @property
    def element_dofs(self) -> np.ndarray:
        t = self.mesh.t.astype(np.int64)
        nen = t.shape[0]
        out = np.empty((nen * self.dpn, t.shape[1]), dtype=np.int64)
        for a in range(nen):
            for c in range(self.dpn):
                out[self.dpn * a + c] = self.dpn * t[a] + c
        return out

    def get_dofs(self, nodes=None) -> DofsView:
        if nodes is None:
            nodes = np.arange(self.mesh.nvertices)
        return DofsView(np.asarray(nodes), self.dpn)
