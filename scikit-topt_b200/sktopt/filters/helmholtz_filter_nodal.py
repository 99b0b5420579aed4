"""Nodal Helmholtz (PDE) density filter on the GPU.

Same operator as reference ``filters/helmholtz_filter_nodal.py``:

* forward (:180-198): volume-weighted element->node average with non-design
  elements contributing 1.0 (:30-56), solve (M + r^2 K) x = M rho_n with x = 1
  on the nodes of non-design elements (:121-157), node->element mean (:26-27);
* gradient (:200-233): same with the fill value 0.0, no Dirichlet nodes, and
  the result clamped to <= 0 (:232).

The reference re-assembles A and b with scikit-fem and calls a sparse direct
solver on every application; here M and K are assembled once on the device
(scalar basis with skfem's *default* quadrature order, as at :128), A is
rebuilt only when the radius changes, and each application is one gather
kernel, one SpMV, one Jacobi-PCG solve (tight tolerance, warm-started from the
previous application) and one gather kernel.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from sktopt._b200 import device as dev
from sktopt._fem import Basis, ElementHex1, ElementTetP1, MeshHex, MeshTet
from sktopt.filters.base import BaseFilter

KE_LAPLACE, KE_MASS = 1, 2


def infer_element_from_mesh(mesh):
    if isinstance(mesh, MeshTet):
        return ElementTetP1()
    if isinstance(mesh, MeshHex):
        return ElementHex1()
    raise ValueError(f"Unknown mesh type: {type(mesh)}")


class _HelmholtzDevice:
    """Device state shared by forward / gradient of one filter object."""

    RTOL = 1e-11
    MAXITER = 5000

    def __init__(self, mesh, elements_volume, design_mask):
        dev.require_cuda()
        self.dm = dev.device_mesh(mesh)
        basis = Basis(mesh, infer_element_from_mesh(mesh))  # default intorder
        self.row_ptr, self.col_idx = self.dm.dof_pattern(1)
        n = self.dm.n_nodes
        self.n_nodes = n
        self.M = self.dm.assemble(1, self.dm.unit_ke(KE_MASS, basis.X, basis.W))
        self.K = self.dm.assemble(1, self.dm.unit_ke(KE_LAPLACE, basis.X, basis.W))
        self.w = dev.to_dev(elements_volume)
        self.wsum = self.dm.e2n_wsum(self.w)
        if design_mask is None:
            self.design_u8 = None
            fixed_nodes = np.array([], dtype=np.int64)
        else:
            dmask = np.asarray(design_mask, dtype=bool)
            self.design_u8 = dev.to_dev(dmask.astype(np.uint8), dev.U8)
            fixed_nodes = np.unique(mesh.t[:, ~dmask].ravel())
        self.has_fixed = fixed_nodes.size > 0
        fm = np.zeros(n, dtype=np.uint8)
        fm[fixed_nodes] = 1
        self.fixed_u8 = dev.to_dev(fm, dev.U8)
        self.x_fixed = dev.to_dev(fm.astype(np.float64))  # x0: 1 on fixed nodes
        self.A = torch.empty_like(self.M)
        self.A_fwd = torch.empty_like(self.M) if self.has_fixed else None
        self.minv = torch.empty(n, dtype=dev.F64, device="cuda")
        self.minv_fwd = torch.empty(n, dtype=dev.F64, device="cuda") if self.has_fixed else None
        self.c = torch.empty(n, dtype=dev.F64, device="cuda")
        self.node = torch.empty(n, dtype=dev.F64, device="cuda")
        self.b = torch.empty(n, dtype=dev.F64, device="cuda")
        self.rhs = torch.empty(n, dtype=dev.F64, device="cuda")
        self.x_fwd = torch.zeros(n, dtype=dev.F64, device="cuda")
        self.x_adj = torch.zeros(n, dtype=dev.F64, device="cuda")
        self.pcg = dev.PcgSolver(n)
        self.radius = None
        self.solve_iters = []

    def set_radius(self, r: float):
        if self.radius == r:
            return
        self.A.copy_(self.M)
        dev.axpby(float(r) ** 2, self.K, 1.0, self.A)  # A = M + r^2 K
        dev.csr_inv_diag(self.row_ptr, self.col_idx, self.A, out=self.minv)
        if self.has_fixed:
            dev.spmv(self.row_ptr, self.col_idx, self.A, self.x_fixed, 1, out=self.c)
            self.A_fwd.copy_(self.A)
            dev.csr_enforce(self.row_ptr, self.col_idx, self.A_fwd, self.fixed_u8)
            dev.csr_inv_diag(self.row_ptr, self.col_idx, self.A_fwd, out=self.minv_fwd)
        self.radius = r

    def _solve(self, A, minv, rhs, x):
        self.pcg.solve(self.row_ptr, self.col_idx, A, minv, rhs, x, dpn_hint=1,
                       rtol=self.RTOL, maxiter=self.MAXITER, use_x0=True,
                       check_every=8)
        self.solve_iters.append(self.pcg.last_iters)
        if not self.pcg.last_converged:
            raise RuntimeError(
                f"Helmholtz filter PCG did not converge "
                f"(relres={self.pcg.last_relres:.3e})")
        return x

    def forward(self, rho, out=None):
        self.dm.e2n(self.w, rho, self.design_u8, 1.0, self.wsum, out=self.node)
        dev.spmv(self.row_ptr, self.col_idx, self.M, self.node, 1, out=self.b)
        if self.has_fixed:
            dev.enforce_rhs(self.b, self.c, self.fixed_u8, self.x_fixed, out=self.rhs)
            x = self._solve(self.A_fwd, self.minv_fwd, self.rhs, self.x_fwd)
        else:
            x = self._solve(self.A, self.minv, self.b, self.x_fwd)
        return self.dm.n2e_mean(x, clamp_max0=False, out=out)

    def gradient(self, v, out=None):
        self.dm.e2n(self.w, v, self.design_u8, 0.0, self.wsum, out=self.node)
        dev.spmv(self.row_ptr, self.col_idx, self.M, self.node, 1, out=self.b)
        x = self._solve(self.A, self.minv, self.b, self.x_adj)
        return self.dm.n2e_mean(x, clamp_max0=True, out=out)


@dataclass
class HelmholtzFilterNodal(BaseFilter):

    def update_radius(self, radius: float, **args):
        self.radius = radius

    @classmethod
    def from_defaults(cls, mesh, elements_volume: np.ndarray, radius: float = 0.3,
                      design_mask: Optional[np.ndarray] = None) -> 'HelmholtzFilterNodal':
        return cls(mesh, elements_volume, radius, design_mask)

    def _device(self) -> _HelmholtzDevice:
        st = self.__dict__.get("_dev_state")
        if st is None:
            st = _HelmholtzDevice(self.mesh, self.elements_volume, self.design_mask)
            self.__dict__["_dev_state"] = st
        st.set_radius(float(self.radius))
        return st

    def forward(self, rho_element, out=None):
        st = self._device()
        if isinstance(rho_element, torch.Tensor) and rho_element.is_cuda:
            return st.forward(rho_element, out=out)
        return st.forward(dev.to_dev(rho_element)).cpu().numpy()

    def gradient(self, v_ele, out=None):
        st = self._device()
        if isinstance(v_ele, torch.Tensor) and v_ele.is_cuda:
            return st.gradient(v_ele, out=out)
        return st.gradient(dev.to_dev(v_ele)).cpu().numpy()
