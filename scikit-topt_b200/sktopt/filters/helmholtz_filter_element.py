"""Element-wise Helmholtz filter on the face-adjacency graph of the mesh.

Same operator as reference ``filters/helmholtz_filter_element.py`` (a module the
reference ships but does not export: ``filters/__init__.py`` comments it out):

* element graph Laplacian (``:195-234``): for every pair (i, j) of elements that
  share a face, w = exp(-d^2 / (2 r^2)) with d the distance of the element
  centres; L_ij = -r^2 w, L_ii = sum_j r^2 w;
* A = V + r^2 L with V = diag(volume / mean volume) (``:264-278``; hex volumes by
  the same six-tetrahedra split as ``get_elements_volume``);
* ``forward(rho)`` solves A x = V rho, ``gradient(v)`` solves A x = V v (A is
  symmetric, ``:281-317``).

Neighbours are the elements sharing a geometric face (``mesh.f2t``).  The
reference's hexahedral face table (``:139-146``) presumes VTK vertex order; under
scikit-fem's local order its 4-tuples are not faces and match nothing (L = 0: the
filter degenerates to the identity on hex meshes), so that accident is not
reproduced.  For tetrahedra both definitions coincide.

The reference offers a sparse LU, scipy cg + Jacobi and pyamg; here the matrix
is built once per radius on the host (vectorised) and every solve is the device
Jacobi-PCG on its CSR form: ``solver_option`` only selects the tolerance
(``spsolve``: 1e-11, the iterative options: ``rtol``, 1e-5 by default as in the
reference).  NumPy arrays are copied in and out; CUDA tensors stay on the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional

import numpy as np
import scipy.sparse as sp
import torch

from sktopt._b200 import device as dev
from sktopt.filters.base import BaseFilter


def _face_pairs(mesh) -> np.ndarray:
    """(2, n_pairs) element pairs sharing a face."""
    f2t = np.asarray(mesh.f2t)
    return f2t[:, (f2t >= 0).all(axis=0)].astype(np.int64)


def _volumes(mesh) -> np.ndarray:
    from sktopt.fea.composer import get_elements_volume
    return np.abs(np.asarray(get_elements_volume(mesh), dtype=np.float64))


def adjacency_matrix_volume_hex_fast(mesh):
    """(face-neighbour lists, element volumes) (``:105-164``)."""
    pairs = _face_pairs(mesh)
    adjacency = [[] for _ in range(mesh.nelements)]
    for i, j in pairs.T.tolist():
        adjacency[i].append(j)
        adjacency[j].append(i)
    return adjacency, _volumes(mesh)


adjacency_matrix_volume_tet_fast = adjacency_matrix_volume_hex_fast
adjacency_matrix_volume_hex = adjacency_matrix_volume_hex_fast
adjacency_matrix_volume_tet = adjacency_matrix_volume_hex_fast


def element_to_element_laplacian(mesh, radius: float):
    """(L csc, volumes): the weighted graph Laplacian of ``:195-234``."""
    ne = mesh.nelements
    pairs = _face_pairs(mesh)
    cen = np.mean(mesh.p[:, mesh.t], axis=1)                       # (3, n_elem)
    d = np.linalg.norm(cen[:, pairs[0]] - cen[:, pairs[1]], axis=0)
    keep = d >= 1e-12
    i, j, d = pairs[0, keep], pairs[1, keep], d[keep]
    w = float(radius) ** 2 * np.exp(-d ** 2 / (2.0 * float(radius) ** 2))
    rows = np.concatenate([i, j, i, j])
    cols = np.concatenate([j, i, i, j])
    data = np.concatenate([-w, -w, w, w])
    L = sp.coo_matrix((data, (rows, cols)), shape=(ne, ne)).tocsc()
    return L, _volumes(mesh)


def prepare_helmholtz_filter(mesh, radius: float,
                             design_elements_mask: Optional[np.ndarray] = None,
                             exclude_nonadjacent: bool = False):
    """(A, V) with A = V + r^2 L, V = diag(volume / mean volume) (``:264-278``;
    the two optional arguments are accepted and unused there as well)."""
    L, vol = element_to_element_laplacian(mesh, radius)
    V = sp.diags(vol / np.mean(vol), format="csc")
    return (V + float(radius) ** 2 * L).tocsc(), V


def _update_radius(mesh, radius: float, design_mask: Optional[np.ndarray] = None):
    return prepare_helmholtz_filter(mesh, radius, design_elements_mask=design_mask,
                                    exclude_nonadjacent=design_mask is not None)


class _DeviceSystem:
    """CSR copy of A on the device + PCG workspace."""

    def __init__(self, A, V):
        dev.require_cuda()
        A = A.tocsr()
        A.sort_indices()
        self.n = A.shape[0]
        self.rp = dev.to_dev(A.indptr, dev.I32)
        self.ci = dev.to_dev(A.indices, dev.I32)
        self.va = dev.to_dev(A.data)
        self.minv = dev.csr_inv_diag(self.rp, self.ci, self.va)
        self.v = dev.to_dev(V.diagonal())
        self.pcg = dev.PcgSolver(self.n)
        self.rhs = torch.empty(self.n, dtype=dev.F64, device="cuda")
        self.x = torch.zeros(self.n, dtype=dev.F64, device="cuda")

    def solve(self, vec, rtol: float, maxiter: int):
        on_dev = isinstance(vec, torch.Tensor) and vec.is_cuda
        b = vec if on_dev else dev.to_dev(np.ascontiguousarray(vec, dtype=np.float64))
        dev.hadamard(1.0, self.v, b, self.rhs)                    # V @ vec
        self.pcg.solve(self.rp, self.ci, self.va, self.minv, self.rhs, self.x, dpn_hint=1,
                       rtol=rtol, maxiter=maxiter, use_x0=True, check_every=8)
        if not self.pcg.last_converged:
            raise RuntimeError("helmholtz_filter_cg does not converge")
        out = self.x.clone()
        return out if on_dev else out.cpu().numpy()


@dataclass
class HelmholtzFilterElement(BaseFilter):
    A: Optional[sp.csc_matrix] = None
    V: Optional[sp.csc_matrix] = None
    solver_option: Literal["spsolve", "cg_jacobi", "cg_pyamg"] = "cg_jacobi"
    dst_path: Optional[str] = None
    rtol: float = 1e-5
    maxiter: int = 1000

    def update_radius(self, radius: float, **args):
        self.radius = radius
        self.A, self.V = _update_radius(self.mesh, radius, self.design_mask)
        self.preprocess(self.solver_option)

    @classmethod
    def from_defaults(cls, mesh, elements_volume: np.ndarray, radius: float = 0.3,
                      design_mask: Optional[np.ndarray] = None,
                      solver_option: Literal["spsolve", "cg_jacobi", "cg_pyamg"] = "cg_pyamg"):
        A, V = _update_radius(mesh, radius, design_mask)
        ret = cls(mesh=mesh, elements_volume=elements_volume, A=A, V=V, radius=radius,
                  design_mask=design_mask, solver_option=solver_option)
        ret.preprocess(solver_option)
        return ret

    def preprocess(self, solver_option: Optional[str] = None):
        if isinstance(solver_option, str):
            if solver_option not in ("cg_jacobi", "cg_pyamg", "spsolve"):
                raise ValueError("should be cg/pyamg/spsolve")
            self.solver_option = solver_option
        if self.maxiter is None or self.maxiter <= 0:
            self.maxiter = max(self.A.shape[0] // 4, 1000)
        self.__dict__["_sys"] = _DeviceSystem(self.A, self.V)

    def _tol(self) -> float:
        return 1e-11 if self.solver_option == "spsolve" else float(self.rtol)

    def forward(self, rho_element):
        return self.__dict__["_sys"].solve(rho_element, self._tol(), int(self.maxiter))

    def gradient(self, v):
        if self.solver_option not in ("spsolve", "cg_jacobi", "cg_pyamg"):
            raise ValueError("solver_option is not set")
        return self.__dict__["_sys"].solve(v, self._tol(), int(self.maxiter))


def apply_helmholtz_filter_cg(rho_element, A, V, M=None, rtol: float = 1e-6,
                              maxiter: Optional[int] = None):
    """Module-level form of ``forward`` (``:320-338``): one device PCG solve of
    A x = V rho (``M`` is accepted for signature parity; Jacobi is built in)."""
    n = A.shape[0]
    mi = min(1000, max(300, n // 5)) if maxiter is None else int(maxiter)
    return _DeviceSystem(A, V).solve(rho_element, float(rtol), mi)


apply_filter_gradient_cg = apply_helmholtz_filter_cg
