"""Device-side objects built on the C ABI (``include/sktopt_b200.h``).

PyTorch is used only as plumbing here: it owns the device buffers (fp64 /
int32 / uint8 CUDA tensors) and the current stream; all arithmetic is done by
the hand-written kernels in ``csrc/`` through ctypes.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import lib as _lib

F64 = torch.float64
I32 = torch.int32
U8 = torch.uint8


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "sktopt (B200 build) needs a CUDA device; there is no CPU fallback"
        )


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_dev(a, dtype=F64, device=None):
    """numpy / torch -> contiguous CUDA tensor of `dtype`."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device or "cuda", dtype=dtype).contiguous()
    np_dtype = {F64: np.float64, I32: np.int32, U8: np.uint8}[dtype]
    arr = np.ascontiguousarray(np.asarray(a), dtype=np_dtype)
    return torch.from_numpy(arr).to(device or "cuda")


def element_classes(p: np.ndarray, t: np.ndarray, rel_tol: float = 1e-11):
    """Group elements with congruent (translated) geometry.

    Returns (elem_class int32 (n_elem,) or None, class_rep int32 (n_class,)).
    ``None`` means "every element is its own class" (unstructured mesh).
    The unit element matrix depends on geometry only (material factors out of
    the reference's forms, fea/composer.py:80-98,136-141), so one matrix per
    class is enough.
    """
    ne = t.shape[1]
    x0 = p[:, t[0]]
    d = p[:, t[1:]] - x0[:, None, :]
    scale = float(np.abs(d).max())
    q = np.rint(d.reshape(-1, ne) / (scale * rel_tol)).astype(np.int64)
    rng = np.random.default_rng(12345)
    h1 = np.zeros(ne, dtype=np.uint64)
    h2 = np.zeros(ne, dtype=np.uint64)
    mult1 = rng.integers(1, 2**63 - 1, size=q.shape[0], dtype=np.uint64) | np.uint64(1)
    mult2 = rng.integers(1, 2**63 - 1, size=q.shape[0], dtype=np.uint64) | np.uint64(1)
    qu = q.view(np.uint64)
    with np.errstate(over="ignore"):
        for r in range(q.shape[0]):
            h1 += qu[r] * mult1[r]
            h2 += (qu[r] ^ np.uint64(0x9E3779B97F4A7C15)) * mult2[r]
    keys = np.empty(ne, dtype=[("a", np.uint64), ("b", np.uint64)])
    keys["a"] = h1
    keys["b"] = h2
    _, rep, inv = np.unique(keys, return_index=True, return_inverse=True)
    n_class = rep.size
    if n_class > max(64, ne // 4):
        return None, np.arange(ne, dtype=np.int32)
    # number classes by first occurrence so the result is order-stable
    order = np.argsort(rep, kind="stable")
    remap = np.empty(n_class, dtype=np.int64)
    remap[order] = np.arange(n_class)
    return remap[inv].astype(np.int32), rep[order].astype(np.int32)


class DeviceMesh:
    """Device-resident connectivity structures of one mesh (``sktb_mesh``)."""

    def __init__(self, mesh, device: int | None = None):
        require_cuda()
        self.lib = _lib.load()
        self.mesh = mesh
        self.device = torch.cuda.current_device() if device is None else device
        self.nen = mesh.t.shape[0]
        self.elem_type = 0 if self.nen == 8 else 1
        self.n_elem = int(mesh.t.shape[1])
        self.n_nodes = int(mesh.p.shape[1])
        conn = np.ascontiguousarray(mesh.t, dtype=np.int32)
        coords = np.ascontiguousarray(mesh.p, dtype=np.float64)
        handle = C.c_void_p()
        _lib.check(
            self.lib.sktb_mesh_create(
                C.byref(handle), self.elem_type, self.n_elem, self.n_nodes,
                conn.ctypes.data_as(C.c_void_p), coords.ctypes.data_as(C.c_void_p),
                self.device,
            )
        )
        self.handle = handle
        self.node_nnz = int(self.lib.sktb_mesh_node_nnz(handle))
        cls, rep = element_classes(coords, conn)
        self.class_rep_h = np.ascontiguousarray(rep, dtype=np.int32)
        self.n_class = int(rep.size)
        self.elem_class = None if cls is None else to_dev(cls, I32)
        self._unit_ke = {}
        self._patterns = {}

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_mesh_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- graph / pattern -----------------------------------------------------
    def node_graph(self):
        rp = np.empty(self.n_nodes + 1, dtype=np.int32)
        ci = np.empty(self.node_nnz, dtype=np.int32)
        _lib.check(
            self.lib.sktb_mesh_node_graph_h(
                self.handle, rp.ctypes.data_as(C.c_void_p), ci.ctypes.data_as(C.c_void_p)
            )
        )
        return rp, ci

    def dof_pattern(self, dpn: int):
        """(row_ptr, col_idx) int32 CUDA tensors of the dpn-dof CSR pattern."""
        if dpn not in self._patterns:
            n = dpn * self.n_nodes
            rp = torch.empty(n + 1, dtype=I32, device="cuda")
            ci = torch.empty(dpn * dpn * self.node_nnz, dtype=I32, device="cuda")
            _lib.check(
                self.lib.sktb_mesh_dof_pattern(self.handle, dpn, _ptr(rp), _ptr(ci), _stream())
            )
            self._patterns[dpn] = (rp, ci)
        return self._patterns[dpn]

    def dof_pattern_rows(self, dpn: int, n0: int, n1: int):
        """Local CSR pattern of the rows of nodes [n0, n1) (global columns)."""
        rp_h, _ = self.node_graph_cached()
        nnz = dpn * dpn * int(rp_h[n1] - rp_h[n0])
        rp = torch.empty(dpn * (n1 - n0) + 1, dtype=I32, device="cuda")
        ci = torch.empty(nnz, dtype=I32, device="cuda")
        _lib.check(
            self.lib.sktb_mesh_dof_pattern_rows(self.handle, dpn, int(n0), int(n1), _ptr(rp), _ptr(ci), _stream())
        )
        return rp, ci

    def node_graph_cached(self):
        if getattr(self, "_graph_h", None) is None:
            self._graph_h = self.node_graph()
        return self._graph_h

    def assemble_rows(self, dpn, n0, n1, unit_ke, scale=None, dir_mask=None, out=None,
                      per_element=False, ke_base: int = 0):
        """Rows of the nodes [n0, n1).  ``per_element=True``: ``unit_ke`` holds one
        matrix per element, starting with element ``ke_base`` (a slab of the
        element range: only the elements touching the rows are read)."""
        nde = dpn * self.nen
        ke_ptr = _ptr(unit_ke)
        if per_element and ke_base:
            ke_ptr = C.c_void_p(unit_ke.data_ptr() - int(ke_base) * nde * nde * 8)
        _lib.check(
            self.lib.sktb_assemble_rows(
                self.handle, dpn, int(n0), int(n1), ke_ptr,
                None if per_element else _ptr(self.elem_class),
                _ptr(scale), _ptr(dir_mask), _ptr(out), _stream(),
            )
        )
        return out

    # -- unit element matrices ----------------------------------------------
    def unit_ke(self, kind: int, X: np.ndarray, W: np.ndarray, nu: float = 0.0):
        key = (kind, float(nu), X.shape[1], float(W.sum()), float(X.sum()))
        if key not in self._unit_ke:
            dpn = 3 if kind == 0 else 1
            nde = self.nen * dpn
            out = torch.empty((self.n_class, nde, nde), dtype=F64, device="cuda")
            Xc = np.ascontiguousarray(X, dtype=np.float64)
            Wc = np.ascontiguousarray(W, dtype=np.float64)
            _lib.check(
                self.lib.sktb_unit_ke(
                    self.handle, kind, float(nu), int(Xc.shape[1]),
                    Xc.ctypes.data_as(C.c_void_p), Wc.ctypes.data_as(C.c_void_p),
                    self.n_class, self.class_rep_h.ctypes.data_as(C.c_void_p),
                    _ptr(out), _stream(),
                )
            )
            self._unit_ke[key] = out
        return self._unit_ke[key]

    # -- kernels ---------------------------------------------------------------
    def assemble(self, dpn, unit_ke, scale=None, dir_mask=None, out=None,
                 per_element=False):
        """``per_element=True``: ``unit_ke`` holds one matrix per element (the
        class table is bypassed), e.g. Galerkin coarse element matrices."""
        if out is None:
            out = torch.empty(dpn * dpn * self.node_nnz, dtype=F64, device="cuda")
        cls = None if per_element else self.elem_class
        _lib.check(
            self.lib.sktb_assemble(
                self.handle, dpn, _ptr(unit_ke), _ptr(cls), _ptr(scale),
                _ptr(dir_mask), _ptr(out), _stream(),
            )
        )
        return out

    def element_energy(self, dpn, unit_ke, scale, u, out=None):
        if out is None:
            out = torch.empty(self.n_elem, dtype=F64, device="cuda")
        if (dpn == 3 and self.nen == 8 and self.elem_class is not None and self.n_class == 1
                and os.environ.get("SKTOPT_B200_ENERGY_UNIFORM", "1") != "0"):
            # one geometry class: thread-per-element kernel, Ke0 as kernel parameters
            key = ("ke_host", unit_ke.data_ptr())
            if key not in self._unit_ke:
                self._unit_ke[key] = np.ascontiguousarray(
                    unit_ke[0].detach().cpu().numpy().reshape(576), dtype=np.float64)
            ke_h = self._unit_ke[key]
            _lib.check(
                self.lib.sktb_element_energy_hex_uniform(
                    self.handle, ke_h.ctypes.data_as(C.c_void_p), _ptr(scale), _ptr(u),
                    _ptr(out), _stream()))
            return out
        _lib.check(
            self.lib.sktb_element_energy(
                self.handle, dpn, _ptr(unit_ke), _ptr(self.elem_class), _ptr(scale),
                _ptr(u), _ptr(out), _stream(),
            )
        )
        return out

    def element_bilinear(self, dpn, unit_ke, scale, u, v, factor=1.0, out=None):
        """out[e] = factor * scale[e] * u_e^T Ke0 v_e (scale may be None)."""
        if out is None:
            out = torch.empty(self.n_elem, dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_element_bilinear(
                self.handle, dpn, _ptr(unit_ke), _ptr(self.elem_class), _ptr(scale),
                _ptr(u), _ptr(v), float(factor), _ptr(out), _stream(),
            )
        )
        return out

    def heat_exchange_local(self, tables, rho_node, T, p, q, h, T_env,
                            want_num=True, want_local=True):
        """(den_e [n_elem], num_e [n_elem] | None, local [nen, n_elem] | None)."""
        N, G, dx, _ = tables
        nq = N.shape[1]
        den = torch.empty(self.n_elem, dtype=F64, device="cuda")
        num = torch.empty(self.n_elem, dtype=F64, device="cuda") if want_num else None
        local = (torch.empty((self.nen, self.n_elem), dtype=F64, device="cuda")
                 if want_local else None)
        _lib.check(
            self.lib.sktb_heat_exchange_local(
                self.handle, nq, _ptr(self.elem_class), _ptr(N), _ptr(G), _ptr(dx),
                _ptr(rho_node), _ptr(T), float(p), float(q), float(h), float(T_env),
                _ptr(den), _ptr(num), _ptr(local), _stream(),
            )
        )
        return den, num, local

    # -- heat: quadrature tables and Robin kernels -----------------------------
    def geom_tables(self, X: np.ndarray, W: np.ndarray):
        """(N [cls,q,a], G [cls,q,a,3], dx [cls,q]) CUDA tensors."""
        key = ("geom", X.shape[1], float(W.sum()), float(X.sum()))
        if key not in self._unit_ke:
            nq = int(X.shape[1])
            N = torch.empty((self.n_class, nq, self.nen), dtype=F64, device="cuda")
            G = torch.empty((self.n_class, nq, self.nen, 3), dtype=F64, device="cuda")
            dx = torch.empty((self.n_class, nq), dtype=F64, device="cuda")
            Xc = np.ascontiguousarray(X, dtype=np.float64)
            Wc = np.ascontiguousarray(W, dtype=np.float64)
            _lib.check(
                self.lib.sktb_geom_tables(
                    self.handle, nq, Xc.ctypes.data_as(C.c_void_p),
                    Wc.ctypes.data_as(C.c_void_p), self.n_class,
                    self.class_rep_h.ctypes.data_as(C.c_void_p), _ptr(N), _ptr(G),
                    _ptr(dx), _stream(),
                )
            )
            Mq = torch.empty((self.n_class, nq, self.nen, self.nen), dtype=F64, device="cuda")
            _lib.check(
                self.lib.sktb_unit_qp_mass(self.handle, nq, self.n_class, _ptr(N), _ptr(dx), _ptr(Mq), _stream())
            )
            self._unit_ke[key] = (N, G, dx, Mq)
        return self._unit_ke[key]

    def robin_virtual_scale(self, tables, rho_node, h, p, q, out=None):
        N, G, dx, _ = tables
        nq = N.shape[1]
        if out is None:
            out = torch.empty((nq, self.n_elem), dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_robin_virtual_scale(
                self.handle, nq, _ptr(self.elem_class), _ptr(N), _ptr(G), _ptr(dx),
                _ptr(rho_node), float(h), float(p), float(q), _ptr(out), _stream(),
            )
        )
        return out

    def assemble_terms(self, unit, scale, out=None):
        n_terms = scale.shape[0]
        if out is None:
            out = torch.empty(self.node_nnz, dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_assemble_terms(
                self.handle, int(n_terms), _ptr(unit), _ptr(self.elem_class), _ptr(scale),
                _ptr(out), _stream(),
            )
        )
        return out

    def robin_explicit_local(self, tables, rho_node, T, h, T_env, p, q, out=None):
        N, G, dx, _ = tables
        nq = N.shape[1]
        if out is None:
            out = torch.empty((self.nen, self.n_elem), dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_robin_explicit_local(
                self.handle, nq, _ptr(self.elem_class), _ptr(N), _ptr(G), _ptr(dx),
                _ptr(rho_node), _ptr(T), float(h), float(T_env), float(p), float(q),
                _ptr(out), _stream(),
            )
        )
        return out

    def local_to_nodes(self, local, divisor=None, out=None):
        if out is None:
            out = torch.empty(self.n_nodes, dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_local_to_nodes(self.handle, _ptr(local), _ptr(divisor), _ptr(out), _stream())
        )
        return out

    def e2n_wsum(self, w):
        out = torch.empty(self.n_nodes, dtype=F64, device="cuda")
        _lib.check(self.lib.sktb_e2n_wsum(self.handle, _ptr(w), _ptr(out), _stream()))
        return out

    def e2n(self, w, val, design_u8, fixed_value, wsum, out=None):
        if out is None:
            out = torch.empty(self.n_nodes, dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_e2n(
                self.handle, _ptr(w), _ptr(val), _ptr(design_u8), float(fixed_value),
                _ptr(wsum), _ptr(out), _stream(),
            )
        )
        return out

    def n2e_mean(self, x, clamp_max0=False, out=None):
        if out is None:
            out = torch.empty(self.n_elem, dtype=F64, device="cuda")
        _lib.check(
            self.lib.sktb_n2e_mean(self.handle, _ptr(x), int(bool(clamp_max0)), _ptr(out), _stream())
        )
        return out


_MESH_CACHE: dict = {}


def device_mesh(mesh) -> DeviceMesh:
    """One DeviceMesh per host mesh object (keyed by identity, kept alive)."""
    key = id(mesh)
    ent = _MESH_CACHE.get(key)
    if ent is None or ent[0] is not mesh:
        ent = (mesh, DeviceMesh(mesh))
        _MESH_CACHE[key] = ent
    return ent[1]


# ---------------------------------------------------------------- CSR / PCG --
def spmv(row_ptr, col_idx, vals, x, dpn_hint, out=None):
    n = row_ptr.numel() - 1
    if out is None:
        out = torch.empty(n, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_spmv(n, dpn_hint, _ptr(row_ptr), _ptr(col_idx), _ptr(vals), _ptr(x), _ptr(out), _stream())
    )
    return out


def spmv_bsr3(node_ptr, node_col, vals, x, out=None):
    """y = A x with block column indices (3 dofs per node, CSR value layout)."""
    n_nodes = node_ptr.numel() - 1
    if out is None:
        out = torch.empty(3 * n_nodes, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_spmv_bsr3(n_nodes, _ptr(node_ptr), _ptr(node_col), _ptr(vals), _ptr(x), _ptr(out), _stream())
    )
    return out


def bsr3_inv_diag(node_ptr, node_col, vals, out=None, node0: int = 0):
    n_nodes = node_ptr.numel() - 1
    if out is None:
        out = torch.empty(3 * n_nodes, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_bsr3_inv_diag(n_nodes, int(node0), _ptr(node_ptr), _ptr(node_col), _ptr(vals), _ptr(out), _stream())
    )
    return out


def spmv_bsr3_tma(node_ptr, node_col, vals, x, max_deg, out=None):
    """Bulk-async pipelined variant of ``spmv_bsr3`` (max_deg <= 27)."""
    n_nodes = node_ptr.numel() - 1
    if out is None:
        out = torch.empty(3 * n_nodes, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_spmv_bsr3_tma(n_nodes, int(node_col.numel()), int(max_deg), _ptr(node_ptr), _ptr(node_col), _ptr(vals), _ptr(x), _ptr(out), _stream())
    )
    return out


def to_f32(v, out=None):
    """Single-precision copy of an fp64 device vector (``sktb_f64_to_f32``)."""
    if out is None:
        out = torch.empty(v.numel(), dtype=torch.float32, device="cuda")
    _lib.check(_lib.load().sktb_f64_to_f32(int(v.numel()), _ptr(v), C.c_void_p(out.data_ptr()),
                                           _stream()))
    return out


def spmv_bsr3_tma_f32(node_ptr, node_col, vals32, x, max_deg, out=None):
    """``spmv_bsr3_tma`` with single-precision values (fp64 x, y, accumulation)."""
    n_nodes = node_ptr.numel() - 1
    if out is None:
        out = torch.empty(3 * n_nodes, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_spmv_bsr3_tma_f32(n_nodes, int(node_col.numel()), int(max_deg), _ptr(node_ptr), _ptr(node_col), C.c_void_p(vals32.data_ptr()), _ptr(x), _ptr(out), _stream())
    )
    return out


def csr_enforce(row_ptr, col_idx, vals, mask_u8):
    n = row_ptr.numel() - 1
    _lib.check(
        _lib.load().sktb_csr_enforce(n, _ptr(row_ptr), _ptr(col_idx), _ptr(vals), _ptr(mask_u8), _stream())
    )
    return vals


def csr_inv_diag(row_ptr, col_idx, vals, out=None, row0: int = 0):
    n = row_ptr.numel() - 1
    if out is None:
        out = torch.empty(n, dtype=F64, device="cuda")
    _lib.check(
        _lib.load().sktb_csr_inv_diag_rows(n, int(row0), _ptr(row_ptr), _ptr(col_idx), _ptr(vals), _ptr(out), _stream())
    )
    return out


class PcgSolver:
    """Device-resident Jacobi-PCG workspace (``sktb_pcg``)."""

    def __init__(self, n_rows: int, device: int | None = None, comm=None,
                 n_global: int | None = None, row0: int = 0, halo=None):
        """``comm`` / ``n_global`` / ``row0`` / ``halo`` select the row-sharded
        variant (``sktb_pcg_create_dist``); ``halo`` comes from
        ``dist.build_halo``."""
        require_cuda()
        self.lib = _lib.load()
        self.n = int(n_rows)
        dev = torch.cuda.current_device() if device is None else device
        h = C.c_void_p()
        if comm is None:
            _lib.check(self.lib.sktb_pcg_create(C.byref(h), self.n, dev))
        else:
            peers, send_off, send_idx, recv_off, recv_idx = halo
            vp = lambda a: a.ctypes.data_as(C.c_void_p)
            self._halo_keep = (np.ascontiguousarray(peers, dtype=np.int32),
                               np.ascontiguousarray(send_off, dtype=np.int64),
                               np.ascontiguousarray(send_idx, dtype=np.int32),
                               np.ascontiguousarray(recv_off, dtype=np.int64),
                               np.ascontiguousarray(recv_idx, dtype=np.int32))
            pe, so, si, ro, ri = self._halo_keep
            _lib.check(self.lib.sktb_pcg_create_dist(
                C.byref(h), comm.handle, int(n_global), int(row0), self.n,
                int(pe.size), vp(pe), vp(so), vp(si), vp(ro), vp(ri), dev))
        self.comm = comm
        self.handle = h
        self.last_iters = 0
        self.last_converged = True
        self.last_relres = 0.0
        self.total_iters = 0

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_pcg_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_slab_halo(self, plane_dofs: int, prev_rank: int, next_rank: int):
        """z-slab halo: whole planes exchanged with the previous / next rank."""
        _lib.check(self.lib.sktb_pcg_set_slab_halo(self.handle, int(plane_dofs),
                                                   int(prev_rank), int(next_rank)))

    def set_first_batch(self, n: int):
        """First convergence poll of the next solve after ``n`` iterations."""
        _lib.check(self.lib.sktb_pcg_set_first_batch(self.handle, int(max(n, 0))))

    def set_profile(self, every_n: int):
        _lib.check(self.lib.sktb_pcg_set_profile(self.handle, int(every_n)))

    def get_profile(self):
        """(sum of sampled SpMV durations in ms, number of samples)."""
        ms = C.c_double()
        cnt = C.c_int64()
        _lib.check(
            self.lib.sktb_pcg_get_profile(
                self.handle, C.cast(C.byref(ms), C.c_void_p), C.cast(C.byref(cnt), C.c_void_p))
        )
        return float(ms.value), int(cnt.value)

    def solve(self, row_ptr, col_idx, vals, inv_diag, b, x, dpn_hint, rtol=1e-8,
              maxiter=1000, use_x0=False, check_every=32, block3=False,
              max_deg=0, mg=None):
        """``block3=True``: ``row_ptr`` / ``col_idx`` are the node-level graph
        (one column per 3x3 block), ``vals`` keeps the CSR layout."""
        info = (C.c_int32 * 2)()
        relres = C.c_double()
        tail = (_ptr(inv_diag), _ptr(b), _ptr(x), int(bool(use_x0)), float(rtol),
                int(maxiter), int(check_every), C.cast(info, C.c_void_p),
                C.cast(C.byref(relres), C.c_void_p), _stream())
        if mg is not None:
            _lib.check(self.lib.sktb_pcg_solve_bsr3_mg(
                self.handle, mg.handle, _ptr(row_ptr), _ptr(col_idx), int(col_idx.numel()),
                int(max_deg), _ptr(vals), *tail))
        elif block3:
            _lib.check(self.lib.sktb_pcg_solve_bsr3(
                self.handle, _ptr(row_ptr), _ptr(col_idx), int(col_idx.numel()),
                int(max_deg), _ptr(vals), *tail))
        else:
            _lib.check(self.lib.sktb_pcg_solve(
                self.handle, dpn_hint, _ptr(row_ptr), _ptr(col_idx), _ptr(vals), *tail))
        self.last_iters = int(info[0])
        self.last_converged = bool(info[1])
        self.last_relres = float(relres.value)
        self.total_iters += self.last_iters
        return x

    def solve_smg(self, smg, b, x, rtol=1e-8, maxiter=300, use_x0=False, check_every=2):
        """PCG on the scalar stencil operator of ``smg`` (``ScalarMultigrid``, set up
        from the enforced CSR matrix), preconditioned by its V-cycle."""
        info = (C.c_int32 * 2)()
        relres = C.c_double()
        _lib.check(self.lib.sktb_pcg_solve_smg(
            self.handle, smg.handle, _ptr(b), _ptr(x), int(bool(use_x0)), float(rtol),
            int(maxiter), int(check_every), C.cast(info, C.c_void_p),
            C.cast(C.byref(relres), C.c_void_p), _stream()))
        self.last_iters = int(info[0])
        self.last_converged = bool(info[1])
        self.last_relres = float(relres.value)
        self.total_iters += self.last_iters
        return x

    def solve_grid(self, gridop, inv_diag, b, x, rtol=1e-8, maxiter=1000, use_x0=False,
                   check_every=32, mg=None):
        """PCG on the matrix-free grid operator (``GridOp``)."""
        info = (C.c_int32 * 2)()
        relres = C.c_double()
        _lib.check(self.lib.sktb_pcg_solve_grid(
            self.handle, None if mg is None else mg.handle, gridop.handle,
            _ptr(inv_diag), _ptr(b), _ptr(x), int(bool(use_x0)), float(rtol),
            int(maxiter), int(check_every), C.cast(info, C.c_void_p),
            C.cast(C.byref(relres), C.c_void_p), _stream()))
        self.last_iters = int(info[0])
        self.last_converged = bool(info[1])
        self.last_relres = float(relres.value)
        self.total_iters += self.last_iters
        return x


class GridOp:
    """Matrix-free operator on a uniform hexahedral tensor grid (``csrc/gridop.cu``).

    ``np_axes``: nodes per axis (x, y, z); ``ke0``: (8 dpn, 8 dpn) unit element
    matrix in the mesh's local vertex order; ``bits``: (8, 3) 0/1 offsets of the
    local vertices; ``dir_mask``: per-dof uint8 (host) or None; ``dpn``: dofs per
    node (3: elasticity, 1: scalar)."""

    def __init__(self, np_axes, ke0: np.ndarray, bits: np.ndarray, dir_mask, dpn: int = 3):
        require_cuda()
        self.lib = _lib.load()
        self.dpn = int(dpn)
        self.np_axes = np.ascontiguousarray(np_axes, dtype=np.int32)
        npx, npy, npz = (int(v) for v in self.np_axes)
        self.n_nodes = npx * npy * npz
        code = bits[:, 0] + 2 * bits[:, 1] + 4 * bits[:, 2]
        if sorted(code.tolist()) != list(range(8)):
            raise ValueError("local vertices are not the 8 corners of a box")
        loc = np.empty(8, dtype=np.int64)
        loc[code] = np.arange(8)
        perm = (dpn * loc[:, None] + np.arange(dpn)[None, :]).ravel()
        nd = 8 * dpn
        ke_cc = np.ascontiguousarray(np.asarray(ke0, dtype=np.float64).reshape(nd, nd)[np.ix_(perm, perm)])
        self.dmask = self.node_flags(dir_mask)
        h = C.c_void_p()
        _lib.check(self.lib.sktb_gridop_create(
            C.byref(h), self.dpn, self.np_axes.ctypes.data_as(C.c_void_p),
            ke_cc.ctypes.data_as(C.c_void_p), torch.cuda.current_device()))
        self.handle = h
        self.scale = None
        self._fields = False

    def node_flags(self, dir_mask):
        """Per-node uint8 device flags: bits 0..dpn-1 fixed dofs, bit 3 fixed dof
        in the 27-neighbourhood."""
        npx, npy, npz = (int(v) for v in self.np_axes)
        if dir_mask is None:
            return to_dev(np.zeros(self.n_nodes, dtype=np.uint8), U8)
        m = np.asarray(dir_mask, dtype=np.uint8).reshape(-1, self.dpn)
        own = np.zeros(self.n_nodes, dtype=np.uint8)
        for i in range(self.dpn):
            own |= (m[:, i] != 0).astype(np.uint8) << i
        g = (own != 0).reshape(npz, npx, npy)
        pad = np.pad(g, 1)
        near = np.zeros_like(g)
        for dz in range(3):
            for dx in range(3):
                for dy in range(3):
                    near |= pad[dz:dz + npz, dx:dx + npx, dy:dy + npy]
        return to_dev(own | (near.ravel().astype(np.uint8) << 3), U8)

    @property
    def tile_shape(self):
        t = (C.c_int32 * 3)()
        _lib.check(self.lib.sktb_gridop_tile_shape(self.handle, C.cast(t, C.c_void_p)))
        return tuple(int(v) for v in t)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_gridop_destroy(h)
            except Exception:
                pass
            self.handle = None

    def set_scale(self, scale=None, dmask=None):
        """Element scale (device fp64, None = 1 for the scalar operator) and,
        optionally, another per-node flag array (``node_flags``)."""
        self.scale = scale          # keep the tensors alive
        if dmask is not None:
            self.dmask = dmask
        _lib.check(self.lib.sktb_gridop_set_fields(
            self.handle, None if scale is None else _ptr(scale), _ptr(self.dmask)))
        self._fields = True

    def apply(self, x, node0: int = 0, n_nodes: int | None = None, out=None):
        n_nodes = self.n_nodes - node0 if n_nodes is None else int(n_nodes)
        if out is None:
            out = torch.empty(self.dpn * n_nodes, dtype=F64, device="cuda")
        _lib.check(self.lib.sktb_gridop_apply(self.handle, int(node0), n_nodes, _ptr(x),
                                              _ptr(out), _stream()))
        return out

    def inv_diag(self, node0: int = 0, n_nodes: int | None = None, out=None):
        n_nodes = self.n_nodes - node0 if n_nodes is None else int(n_nodes)
        if out is None:
            out = torch.empty(self.dpn * n_nodes, dtype=F64, device="cuda")
        _lib.check(self.lib.sktb_gridop_inv_diag(self.handle, int(node0), n_nodes, _ptr(out),
                                                 _stream()))
        return out


# ------------------------------------------------------- elementwise kernels --
def interpolate_modulus(rho, E0, Emin, p, ramp=False, out=None):
    if out is None:
        out = torch.empty_like(rho)
    _lib.check(
        _lib.load().sktb_interpolate_modulus(rho.numel(), _ptr(rho), float(E0), float(Emin), float(p), int(ramp), _ptr(out), _stream())
    )
    return out


def dc_drho(rho_proj, energy, E0, Emin, p, ramp=False, dH=None, out=None):
    if out is None:
        out = torch.empty_like(rho_proj)
    _lib.check(
        _lib.load().sktb_dc_drho(rho_proj.numel(), _ptr(rho_proj), _ptr(energy), float(E0), float(Emin), float(p), int(ramp), _ptr(dH), _ptr(out), _stream())
    )
    return out


def heaviside(x, beta, eta, out=None, dH=None):
    _lib.check(
        _lib.load().sktb_heaviside(x.numel(), _ptr(x), float(beta), float(eta), _ptr(out), _ptr(dH), _stream())
    )
    return out, dH


def oc_candidate(dC, rho_e, lmid, eps, eta, move_limit, rho_min, rho_max, sr_min,
                 sr_max, design_idx, scaling_rate, rho_cand, rho_full_cand):
    _lib.check(
        _lib.load().sktb_oc_candidate(
            dC.numel(), _ptr(dC), _ptr(rho_e), float(lmid), float(eps), float(eta),
            float(move_limit), float(rho_min), float(rho_max), float(sr_min),
            float(sr_max), _ptr(design_idx), _ptr(scaling_rate), _ptr(rho_cand),
            _ptr(rho_full_cand), _stream(),
        )
    )


def logmoc_update(rho, dL, eta, move_limit, rho_min, rho_max, clip, scaling_rate,
                  clip_lower, clip_upper):
    _lib.check(
        _lib.load().sktb_logmoc_update(
            rho.numel(), _ptr(rho), _ptr(dL), float(eta), float(move_limit),
            float(rho_min), float(rho_max), float(clip), _ptr(scaling_rate),
            _ptr(clip_lower), _ptr(clip_upper), _stream(),
        )
    )


def reduce_wsum(a, idx=None, w=None) -> float:
    n = idx.numel() if idx is not None else a.numel()
    out = C.c_double()
    _lib.check(
        _lib.load().sktb_reduce_wsum_h(n, _ptr(a), _ptr(idx), _ptr(w), C.cast(C.byref(out), C.c_void_p), _stream())
    )
    return float(out.value)


def reduce_stats(a, idx=None):
    """(min, mean, max, std) of a[idx]."""
    n = idx.numel() if idx is not None else a.numel()
    out = (C.c_double * 4)()
    _lib.check(
        _lib.load().sktb_reduce_stats_h(n, _ptr(a), _ptr(idx), C.cast(out, C.c_void_p), _stream())
    )
    return float(out[0]), float(out[1]), float(out[2]), float(out[3])


def reduce_absmax(a) -> float:
    out = C.c_double()
    _lib.check(
        _lib.load().sktb_reduce_absmax_h(a.numel(), _ptr(a), C.cast(C.byref(out), C.c_void_p), _stream())
    )
    return float(out.value)


def dot(a, b) -> float:
    out = C.c_double()
    _lib.check(
        _lib.load().sktb_dot_h(a.numel(), _ptr(a), _ptr(b), C.cast(C.byref(out), C.c_void_p), _stream())
    )
    return float(out.value)


_PCT_WORK: dict = {}


def abs_percentile(a, q: float) -> float:
    n = a.numel()
    key = (a.device.index, n)
    work = _PCT_WORK.get(key)
    if work is None:
        work = torch.empty(n + 1024, dtype=torch.int64, device="cuda")
        _PCT_WORK.clear()
        _PCT_WORK[key] = work
    out = C.c_double()
    _lib.check(
        _lib.load().sktb_abs_percentile_h(n, _ptr(a), float(q), _ptr(work), C.cast(C.byref(out), C.c_void_p), _stream())
    )
    return float(out.value)


def gather(src, idx, out=None):
    if out is None:
        out = torch.empty(idx.numel(), dtype=F64, device="cuda")
    _lib.check(_lib.load().sktb_gather(idx.numel(), _ptr(src), _ptr(idx), _ptr(out), _stream()))
    return out


def scatter(src, idx, dst):
    _lib.check(_lib.load().sktb_scatter(idx.numel(), _ptr(src), _ptr(idx), _ptr(dst), _stream()))
    return dst


def axpby(a, x, b, y):
    _lib.check(_lib.load().sktb_axpby(x.numel(), float(a), _ptr(x), float(b), _ptr(y), _stream()))
    return y


def affine(a, x, b, y, c, out):
    """out = a*x + b*y + c  (y may be None)."""
    _lib.check(
        _lib.load().sktb_affine(x.numel(), float(a), _ptr(x), float(b), _ptr(y), float(c), _ptr(out), _stream())
    )
    return out


def fill(out, value: float = 0.0):
    """out[:] = value (contiguous fp64 tensor)."""
    _lib.check(_lib.load().sktb_fill_abs(out.numel(), None, float(value), _ptr(out), _stream()))
    return out


def absval(x, out=None):
    """out = |x|."""
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.load().sktb_fill_abs(x.numel(), _ptr(x), 0.0, _ptr(out), _stream()))
    return out


def hadamard(a, x, y, out):
    """out = a*x*y."""
    _lib.check(_lib.load().sktb_hadamard(x.numel(), float(a), _ptr(x), _ptr(y), _ptr(out), _stream()))
    return out


def kkt_residual(rho, g, dv, coef, lo, hi):
    """(max |g + coef*dv| over lo < rho < hi, count of interior entries)."""
    out = (C.c_double * 2)()
    _lib.check(
        _lib.load().sktb_kkt_residual_h(rho.numel(), _ptr(rho), _ptr(g), _ptr(dv), float(coef), float(lo), float(hi), C.cast(out, C.c_void_p), _stream())
    )
    return float(out[0]), int(out[1])


def reduce_maxdiff(a, b, idx=None) -> float:
    n = idx.numel() if idx is not None else a.numel()
    out = C.c_double()
    _lib.check(
        _lib.load().sktb_reduce_maxdiff_h(n, _ptr(a), _ptr(b), _ptr(idx), C.cast(C.byref(out), C.c_void_p), _stream())
    )
    return float(out.value)


def enforce_rhs(b, t, mask_u8, xD, out=None):
    if out is None:
        out = torch.empty_like(b)
    _lib.check(
        _lib.load().sktb_enforce_rhs(b.numel(), _ptr(b), _ptr(t), _ptr(mask_u8), _ptr(xD), _ptr(out), _stream())
    )
    return out


def launch_count() -> int:
    """Kernels launched by libsktopt_b200 in this process so far."""
    return int(_lib.load().sktb_launch_count())


def fp64_peak_tflops(iters: int = 4096, reps: int = 5, const_operand: bool = False) -> float:
    """Measured FP64 FMA throughput of this GPU (DFMA-chain probe, best of reps).
    ``const_operand``: the multiplier comes from the constant bank (uniform
    register operand), as in the grid operator."""
    out = torch.empty(148 * 8 * 256, dtype=F64, device="cuda")
    flops = C.c_int64()
    lib = _lib.load()
    fn = lib.sktb_fp64_probe_const if const_operand else lib.sktb_fp64_probe
    best = 0.0
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(fn(int(iters), _ptr(out), C.cast(C.byref(flops), C.c_void_p), _stream()))
        b.record()
        torch.cuda.synchronize()
        best = max(best, flops.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    return best


def flush_l2(scratch):
    _lib.check(_lib.load().sktb_flush_l2(_ptr(scratch), scratch.numel() * scratch.element_size(), _stream()))


def dgemm(A, B, out, M, N, K, lda, ldb, ldc, batch=1, stride_a=0, stride_b=0, stride_c=0,
          scale=None, stride_s=0):
    """out[b] = A[b] . B[b] (* scale[b] elementwise); row-major, strides in elements
    (``csrc/dgemm.cu``)."""
    _lib.check(_lib.load().sktb_dgemm_batched(
        int(M), int(N), int(K), _ptr(A), int(lda), int(stride_a), _ptr(B), int(ldb),
        int(stride_b), _ptr(out), int(ldc), int(stride_c), int(batch), _ptr(scale),
        int(stride_s), _stream()))
    return out
