"""Facet (surface) quadrature helpers replacing skfem's ``FacetBasis`` usage.

The reference integrates over listed facets with
``asm(LinearForm / BilinearForm / Functional, FacetBasis(mesh, elem, facets=ids))``
(``mesh/task_elastic.py:47-69``, ``mesh/task_heat.py:45-55,106-134``).  Here the
same integrals are evaluated with a Gauss rule on each facet, taking the trace
from the first neighbouring element ``f2t[0]`` like skfem does.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .mesh import MeshHex, MeshTet, HEX_FACES, HEX_FACE_AXIS, HEX_FACE_SIDE, TET_FACES
from .basis import gauss_legendre_unit, hex_shape


def _facet_quadrature_hex(mesh: MeshHex, facets: np.ndarray, npts: int = 2):
    """Returns (nodes (4, nf), N (4, nf, nq), dS (nf, nq)) for hex facets.

    ``nodes`` are the facet's element-local vertices in cyclic order and ``N``
    the trace of their trilinear shape functions at the surface Gauss points;
    ``dS`` = quadrature weight x surface Jacobian.
    """
    facets = np.asarray(facets, dtype=np.int64)
    elems = mesh.f2t[0, facets].astype(np.int64)
    lfs = mesh.f2lf[0, facets].astype(np.int64)
    g, w = gauss_legendre_unit(npts)
    R, S = np.meshgrid(g, g, indexing="ij")
    WR, WS = np.meshgrid(w, w, indexing="ij")
    r, s, wq = R.ravel(), S.ravel(), (WR * WS).ravel()
    nq = r.size
    nf = facets.size
    nodes = np.empty((4, nf), dtype=np.int64)
    Nout = np.empty((4, nf, nq))
    dS = np.empty((nf, nq))
    for lf in range(6):
        sel = np.nonzero(lfs == lf)[0]
        if sel.size == 0:
            continue
        ax = HEX_FACE_AXIS[lf]
        other = [d for d in range(3) if d != ax]
        X = np.empty((3, nq))
        X[ax] = HEX_FACE_SIDE[lf]
        X[other[0]] = r
        X[other[1]] = s
        N, dN = hex_shape(X)  # (8,nq), (8,3,nq)
        te = mesh.t[:, elems[sel]].astype(np.int64)  # (8, ns)
        xe = mesh.p[:, te]  # (3, 8, ns)
        # tangent vectors dx/dr, dx/ds : (3, ns, nq)
        tr = np.einsum("das,aq->dsq", xe, dN[:, other[0], :])
        ts = np.einsum("das,aq->dsq", xe, dN[:, other[1], :])
        nrm = np.cross(tr, ts, axis=0)
        jac = np.sqrt(np.sum(nrm * nrm, axis=0))  # (ns, nq)
        dS[sel] = jac * wq[None, :]
        loc = HEX_FACES[lf]
        nodes[:, sel] = te[loc]
        Nout[:, sel, :] = N[loc][:, None, :]
    return nodes, Nout, dS


def _facet_quadrature_tet(mesh: MeshTet, facets: np.ndarray):
    facets = np.asarray(facets, dtype=np.int64)
    elems = mesh.f2t[0, facets].astype(np.int64)
    lfs = mesh.f2lf[0, facets].astype(np.int64)
    nf = facets.size
    nodes = np.empty((3, nf), dtype=np.int64)
    for lf in range(4):
        sel = np.nonzero(lfs == lf)[0]
        if sel.size:
            nodes[:, sel] = mesh.t[:, elems[sel]][TET_FACES[lf]]
    x = mesh.p[:, nodes]  # (3, 3, nf)
    e1 = x[:, 1] - x[:, 0]
    e2 = x[:, 2] - x[:, 0]
    area = 0.5 * np.linalg.norm(np.cross(e1, e2, axis=0), axis=0)
    # 3-point (edge midpoint) rule: exact for quadratics
    Nq = np.array([[0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]).T  # (3 nodes, 3 qp)
    N = np.repeat(Nq[:, None, :], nf, axis=1)
    dS = np.repeat((area / 3.0)[:, None], 3, axis=1)
    return nodes, N, dS


def facet_quadrature(mesh, facets):
    if isinstance(mesh, MeshHex):
        return _facet_quadrature_hex(mesh, facets)
    if isinstance(mesh, MeshTet):
        return _facet_quadrature_tet(mesh, facets)
    raise NotImplementedError("MeshHex or MeshTet")


def facet_area(mesh, facets) -> float:
    """``asm(Functional(1), FacetBasis)``: total area of the listed facets."""
    _, _, dS = facet_quadrature(mesh, facets)
    return float(dS.sum())


def facet_load(mesh, facets, value: float, dpn: int = 1, comp: int = 0) -> np.ndarray:
    """``asm(LinearForm(value * v[comp]), FacetBasis)`` -> (dpn*n_nodes,)."""
    nodes, N, dS = facet_quadrature(mesh, facets)
    contrib = value * np.einsum("afq,fq->af", N, dS)
    F = np.zeros(dpn * mesh.nvertices)
    np.add.at(F, dpn * nodes.ravel() + comp, contrib.ravel())
    return F


def facet_mass(mesh, facets, coeff: float) -> sp.csr_matrix:
    """``asm(BilinearForm(coeff * u * v), FacetBasis)`` (scalar basis) -> CSR."""
    nodes, N, dS = facet_quadrature(mesh, facets)
    nfv = nodes.shape[0]
    M = coeff * np.einsum("afq,bfq,fq->abf", N, N, dS)
    rows = np.repeat(nodes[:, None, :], nfv, axis=1).ravel()
    cols = np.repeat(nodes[None, :, :], nfv, axis=0).ravel()
    n = mesh.nvertices
    return sp.coo_matrix((M.ravel(), (rows, cols)), shape=(n, n)).tocsr()
