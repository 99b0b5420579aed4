"""Direct solve of the Helmholtz filter system on tensor grids (fast
diagonalisation).

``solve_helmholtz`` (reference ``filters/helmholtz_filter_nodal.py:121-157``)
factorises  A = M + r^2 K  with a sparse LU on every call.  On a tensor-product
hexahedral grid with the natural (Neumann) boundary condition -- the adjoint
solves of ``HelmholtzFilterNodal.gradient`` and forward solves without fixed
nodes -- the trilinear mass and stiffness matrices are Kronecker products of the
1-D linear-element matrices (node id = iy + npy*(ix + npx*iz)):

    M = Mz (x) Mx (x) My,   K = Kz (x) Mx (x) My + Mz (x) Kx (x) My + Mz (x) Mx (x) Ky

With the generalised eigen-pairs  K_d V_d = M_d V_d diag(lam_d),  V_d^T M_d V_d = I
of each axis and V = Vz (x) Vx (x) Vy,

    V^T A V = diag(1 + r^2 (lam_z + lam_x + lam_y)) =: D,   A^-1 = V D^-1 V^T,

so one solve is six dense products with the small V_d and one diagonal scaling
(fused into the third product): a direct solve like the reference's LU, exact to
rounding, ~10x cheaper than the PCG it replaces (20 iterations at rtol 1e-11).
The products run on the hand-written batched fp64 kernel ``csrc/dgemm.cu``
(``sktb_dgemm_batched``); ``SKTOPT_B200_FD_TORCH=1`` switches to ``torch.matmul``
(cuBLAS), kept only as the cross-check of ``tests/test_gpu_parity.py``.  Systems
with fixed nodes keep the PCG (``_HelmholtzDevice._solve``).
"""
from __future__ import annotations

import os

import numpy as np
import scipy.linalg
import torch

from sktopt._b200 import device as dev


def axis_matrices(c: np.ndarray):
    """Assembled 1-D linear-element mass and stiffness matrices on nodes ``c``."""
    h = np.diff(np.asarray(c, dtype=np.float64))
    n = h.size + 1
    M = np.zeros((n, n))
    K = np.zeros((n, n))
    for e, he in enumerate(h):
        M[e:e + 2, e:e + 2] += he / 6.0 * np.array([[2.0, 1.0], [1.0, 2.0]])
        K[e:e + 2, e:e + 2] += 1.0 / he * np.array([[1.0, -1.0], [-1.0, 1.0]])
    return M, K


class FastDiagHelmholtz:
    """x = (M + r^2 K)^-1 b for nodal vectors of ``MeshHex.init_tensor(xs, ys, zs)``."""

    def __init__(self, axes, device="cuda"):
        xs, ys, zs = axes
        self.shape = (len(zs), len(xs), len(ys))          # (npz, npx, npy): iy fastest
        self.V, self.lam = [], []
        for c in (zs, xs, ys):
            M, K = axis_matrices(c)
            lam, V = scipy.linalg.eigh(K, M)              # V^T M V = I, V^T K V = diag(lam)
            lam[0] = max(lam[0], 0.0)                     # the constant mode: exactly 0
            self.V.append(torch.as_tensor(np.ascontiguousarray(V), dtype=torch.float64,
                                          device=device))
            self.lam.append(torch.as_tensor(lam, dtype=torch.float64, device=device))
        self.Vt = [v.t().contiguous() for v in self.V]
        self.radius = None
        self.Dinv = None

    def set_radius(self, r: float):
        if self.radius == r:
            return
        lz, lx, ly = self.lam
        D = 1.0 + float(r) ** 2 * (lz[:, None, None] + lx[None, :, None] + ly[None, None, :])
        self.Dinv = (1.0 / D).contiguous()
        self.radius = r

    def solve(self, b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if b.device.type != "cuda" or os.environ.get("SKTOPT_B200_FD_TORCH", "0") == "1":
            # (host tensors: the CPU test of the factorisation itself, tests/test_oracle.py)
            return self._solve_torch(b, out)
        npz, npx, npy = self.shape
        n, plane = npz * npx * npy, npx * npy
        if getattr(self, "_w", None) is None:
            self._w = [torch.empty(n, dtype=torch.float64, device=b.device) for _ in range(2)]
        w0, w1 = self._w
        Vz, Vx, Vy = self.V
        Vzt, Vxt, Vyt = self.Vt
        if out is None:
            out = torch.empty(n, dtype=torch.float64, device=b.device)
        # V^T b: along y (rows x Vy), along x (Vx^T per z-plane), along z (Vz^T), then D^-1
        dev.dgemm(b, Vy, w0, npz * npx, npy, npy, npy, npy, npy)
        dev.dgemm(Vxt, w0, w1, npx, npy, npx, npx, npy, npy, batch=npz, stride_b=plane,
                  stride_c=plane)
        dev.dgemm(Vzt, w1, w0, npz, plane, npz, npz, plane, plane, scale=self.Dinv)
        # V (.)
        dev.dgemm(Vz, w0, w1, npz, plane, npz, npz, plane, plane)
        dev.dgemm(Vx, w1, w0, npx, npy, npx, npx, npy, npy, batch=npz, stride_b=plane,
                  stride_c=plane)
        dev.dgemm(w0, Vyt, out, npz * npx, npy, npy, npy, npy, npy)
        return out

    def _solve_torch(self, b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        npz, npx, npy = self.shape
        Vz, Vx, Vy = self.V
        Vzt, Vxt, Vyt = self.Vt
        T = torch.matmul(b.view(npz * npx, npy), Vy)                       # V_y^T along y
        T = torch.matmul(Vxt, T.view(npz, npx, npy))                       # V_x^T along x
        T = torch.matmul(Vzt, T.view(npz, npx * npy)).view(npz, npx, npy)  # V_z^T along z
        T.mul_(self.Dinv)
        T = torch.matmul(Vz, T.view(npz, npx * npy))
        T = torch.matmul(Vx, T.view(npz, npx, npy))
        if out is None:
            return torch.matmul(T.view(npz * npx, npy), Vyt).view(-1)
        torch.matmul(T.view(npz * npx, npy), Vyt, out=out.view(npz * npx, npy))
        return out
