"""Mesh helpers (vectorised restatements of reference ``mesh/utils.py``)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from sktopt._fem import MeshHex, MeshTet


def get_points_in_range(x_rng, y_rng, z_rng):
    """Closed-box predicate on points x of shape (3, n) (``mesh/utils.py:8-15``)."""
    lo = np.array([x_rng[0], y_rng[0], z_rng[0]], dtype=float)
    hi = np.array([x_rng[1], y_rng[1], z_rng[1]], dtype=float)

    def in_range(x):
        ok = np.ones(x.shape[1:], dtype=bool)
        for d in range(3):
            ok &= (x[d] >= lo[d]) & (x[d] <= hi[d])
        return ok
    return in_range


def _triple(p, t, i0, i1, i2, i3):
    v1 = p[:, t[i1]] - p[:, t[i0]]
    v2 = p[:, t[i2]] - p[:, t[i0]]
    v3 = p[:, t[i3]] - p[:, t[i0]]
    return np.einsum("de,de->e", np.cross(v1, v2, axis=0), v3)


def fix_hexahedron_orientation(t, p):
    """Swap local vertices 1 and 3 where (p1-p0)x(p3-p0).(p4-p0) < 0
    (``mesh/utils.py:18-54``)."""
    t_fixed = np.array(t, copy=True)
    flip = _triple(p, t_fixed, 0, 1, 3, 4) < 0
    t_fixed[1, flip], t_fixed[3, flip] = t[3, flip], t[1, flip]
    return t_fixed


def fix_tetrahedron_orientation(t, p):
    """Swap local vertices 1 and 2 of negatively oriented tets
    (``mesh/utils.py:57-91``)."""
    t_fixed = np.array(t, copy=True)
    flip = _triple(p, t_fixed, 0, 1, 2, 3) / 6.0 < 0
    t_fixed[1, flip], t_fixed[2, flip] = t[2, flip], t[1, flip]
    return t_fixed


fix_tetrahedron_orientation_numba = fix_tetrahedron_orientation


def fix_elements_orientation(mesh):
    if isinstance(mesh, MeshTet):
        return fix_tetrahedron_orientation(mesh.t, mesh.p)
    if isinstance(mesh, MeshHex):
        return fix_hexahedron_orientation(mesh.t, mesh.p)
    raise ValueError("MeshTet or MeshHex expected")


def get_elements_by_nodes(mesh, target_nodes) -> np.ndarray:
    """Sorted unique int32 ids of elements touching any of the given nodes
    (``mesh/utils.py:180-228``)."""
    # (duplicates and order are irrelevant for the membership mask: no np.unique)
    if isinstance(target_nodes, np.ndarray):
        nodes = target_nodes.ravel()
    else:
        nodes = np.concatenate([np.asarray(a).ravel() for a in target_nodes])
    nodes = nodes.astype(np.int64)
    hit = np.zeros(mesh.nvertices, dtype=bool)
    hit[nodes[(nodes >= 0) & (nodes < mesh.nvertices)]] = True
    elems = np.nonzero(hit[mesh.t].any(axis=0))[0]
    return np.ascontiguousarray(elems.astype(np.int32))


def get_adjacent_elements(mesh, element_indices):
    """Elements sharing a node with the given ones, excluding them."""
    element_indices = np.asarray(element_indices, dtype=np.int64)
    nodes = np.unique(mesh.t[:, element_indices])
    nb = get_elements_by_nodes(mesh, nodes)
    return sorted(set(nb.tolist()) - set(element_indices.tolist()))


def _element_incidence(mesh) -> sp.csr_matrix:
    """(n_nodes x n_elem) node-element incidence, one per (node, element) pair."""
    t = np.asarray(mesh.t, dtype=np.int64)
    nen, ne = t.shape
    B = sp.coo_matrix((np.ones(nen * ne, dtype=np.int32),
                       (t.ravel(), np.tile(np.arange(ne), nen))),
                      shape=(mesh.nvertices, ne)).tocsr()
    B.data[:] = 1                      # a degenerate element may list a node twice
    return B


def build_element_adjacency_matrix(mesh) -> sp.csr_matrix:
    """uint8 CSR A with A[i, j] = 1 iff elements i and j share at least one node,
    the diagonal included (``mesh/utils.py:139-159``, there a double Python loop;
    here one sparse product)."""
    B = _element_incidence(mesh)
    A = (B.T @ B).tocsr()
    A.sort_indices()
    return sp.csr_matrix((np.ones(A.nnz, dtype=np.uint8), A.indices, A.indptr), shape=A.shape)


def build_element_adjacency_matrix_fast(mesh) -> sp.csr_matrix:
    """Same without the diagonal (``mesh/utils.py:231-254``)."""
    A = build_element_adjacency_matrix(mesh).tolil()
    A.setdiag(0)
    A = A.tocsr()
    A.eliminate_zeros()
    A.sort_indices()
    return A


def get_adjacent_elements_fast(adjacency, element_indices) -> np.ndarray:
    """Sorted int32 ids of the elements adjacent to any of ``element_indices``,
    excluding those (``mesh/utils.py:257-263``)."""
    idx = np.asarray(element_indices, dtype=np.int64).ravel()
    nb = np.unique(adjacency[idx].indices) if idx.size else np.array([], dtype=np.int64)
    return np.setdiff1d(nb, idx).astype(np.int32)
