"""Multi-GPU plumbing: one process per GPU, NCCL over NVLink.

``torch.distributed`` is used only to ship the 128-byte ncclUniqueId from rank 0
to the other ranks; every collective on the data path (halo exchange of the
search direction, all-reduce of the PCG dot products, all-gather of the
solution) is issued from C++ on the solver's stream (``csrc/comm.cu``,
``csrc/pcg.cu``).

Sharding model: the global CSR is split into contiguous *node* ranges balanced
by non-zeros; rank r assembles and owns the rows of its nodes, column indices
stay global, and the search direction is a full-length vector whose ghost slots
are refreshed before every SpMV.  Element-wise work and the scalar Helmholtz
filter are replicated (they are ~3 % of an iteration), so every rank holds the
same densities and takes the same control-flow decisions.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as _lib

_DEFAULT = None


class Comm:
    def __init__(self, rank: int, world: int, device: int, unique_id: bytes):
        self.lib = _lib.load()
        self.rank, self.world, self.device = rank, world, device
        self.unique_id = bytes(unique_id)
        self.p2p = False
        h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        _lib.check(self.lib.sktb_comm_create(C.byref(h), C.cast(buf, C.c_void_p),
                                             rank, world, device))
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_comm_destroy(h)
            except Exception:
                pass
            self.handle = None

    def setup_peer_arena(self):
        """Symmetric arena for peer-memory halo exchanges (``csrc/comm.cuh``): NCCL
        transport only (one GPU per rank), opt-in with ``SKTOPT_B200_P2P_HALO=1``.
        Measured on 4 x B200 (C5, profiles/r2_bench_4gpu*.json): 80.8 ms / step with
        the peer-memory pulls against 77.3 ms with grouped ncclSend / ncclRecv -- the
        publish / pull / acknowledge handshake costs two NVLink round trips per
        exchange, NCCL's FIFO one -- so NCCL stays the default.
        ``SKTOPT_B200_ARENA_MB`` sizes the arena (default 3072).  Every rank must
        succeed, else none uses it."""
        import os
        import torch.distributed as dist
        self.p2p = False
        if self.unique_id[:4] == b"SHM:" or os.environ.get("SKTOPT_B200_P2P_HALO", "0") != "1":
            return
        mb = int(os.environ.get("SKTOPT_B200_ARENA_MB", "3072"))
        buf = C.create_string_buffer(64)
        ok = self.lib.sktb_comm_arena_create(self.handle, C.c_int64(mb << 20),
                                             C.cast(buf, C.c_void_p)) == 0
        handles = [None] * self.world
        dist.all_gather_object(handles, (ok, buf.raw))
        if not all(h[0] for h in handles):
            return
        cat = C.create_string_buffer(b"".join(h[1] for h in handles), 64 * self.world)
        ok = self.lib.sktb_comm_arena_open(self.handle, C.cast(cat, C.c_void_p)) == 0
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        self.p2p = all(flags)
        if not self.p2p:
            raise RuntimeError("peer arenas could not be mapped on every rank "
                               "(set SKTOPT_B200_P2P_HALO=0 to use NCCL send / recv)")

    def arena_status(self):
        """(bytes used, exchanges done, error flag) of the peer arena."""
        used, epoch, err = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(self.lib.sktb_comm_arena_status(
            self.handle, C.cast(C.byref(used), C.c_void_p), C.cast(C.byref(epoch), C.c_void_p),
            C.cast(C.byref(err), C.c_void_p)))
        return int(used.value), int(epoch.value), int(err.value)

    def allreduce_sum(self, src, dst=None):
        dst = src if dst is None else dst
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.sktb_comm_allreduce_sum(
            self.handle, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()),
            src.numel(), stream))
        return dst

    def allgatherv(self, buf, counts, displs):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        d = np.ascontiguousarray(displs, dtype=np.int64)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.sktb_comm_allgatherv(
            self.handle, C.c_void_p(buf.data_ptr()), c.ctypes.data_as(C.c_void_p),
            d.ctypes.data_as(C.c_void_p), stream))
        return buf


def make_unique_id(transport: str = "nccl") -> bytes:
    """128-byte communicator id: an ncclUniqueId, or ``SHM:<random>`` for the
    host shared-memory transport (several ranks on ONE GPU: correctness runs of
    the sharded paths on a one-GPU box; ``csrc/comm.cuh``)."""
    if transport == "shm":
        import secrets
        return (b"SHM:" + secrets.token_hex(16).encode()).ljust(128, b"\0")
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().sktb_comm_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def transport() -> str:
    """``SKTOPT_B200_COMM`` = nccl | shm; default: shm when the torch.distributed
    job itself runs on gloo (ranks sharing a GPU), else nccl."""
    import os
    import torch.distributed as dist
    want = os.environ.get("SKTOPT_B200_COMM", "auto").lower()
    if want in ("nccl", "shm"):
        return want
    return "shm" if dist.get_backend() == "gloo" else "nccl"


def default_comm():
    """The communicator of the current ``torch.distributed`` job, or None when
    running on one GPU."""
    global _DEFAULT
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return None
    if _DEFAULT is None:
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [make_unique_id(transport()) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        _DEFAULT = Comm(rank, world, torch.cuda.current_device(), box[0])
        _DEFAULT.setup_peer_arena()
    return _DEFAULT


def is_io_rank() -> bool:
    """True on the one process that owns the run directory (rank 0 of the
    ``torch.distributed`` job, or the only process).  Every rank computes the
    same densities and histories; only this one writes them."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank() == 0
    return True


def io_barrier():
    """Other ranks wait here until the I/O rank has prepared the run directory."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def reset_default_comm():
    global _DEFAULT
    _DEFAULT = None


# ------------------------------------------------------------ partitioning --
def partition_nodes(node_ptr: np.ndarray, world: int) -> np.ndarray:
    """Contiguous node ranges with (nearly) equal non-zero counts.
    Returns the world+1 range boundaries."""
    n_nodes = node_ptr.size - 1
    total = int(node_ptr[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(node_ptr, target, side="left"))
        k = min(max(k, cuts[-1] + 1), n_nodes - (world - r))
        cuts.append(k)
    cuts.append(n_nodes)
    return np.asarray(cuts, dtype=np.int64)


def partition_planes(n_planes: int, world: int) -> np.ndarray:
    """Contiguous ranges of whole z-planes, as even as possible.  Returns the
    world+1 plane boundaries."""
    return np.round(np.linspace(0, n_planes, world + 1)).astype(np.int64)


def build_halo(node_ptr: np.ndarray, node_col: np.ndarray, cuts: np.ndarray,
               rank: int, dpn: int):
    """Halo description of ``rank`` in dof indices.

    Returns (peers, send_off, send_idx, recv_off, recv_idx): for peer i the
    owned global dofs it needs from us and the global dofs we receive from it,
    both sorted ascending (the graph is symmetric, so each side can derive its
    lists locally and they agree)."""
    n0, n1 = int(cuts[rank]), int(cuts[rank + 1])
    s, e = int(node_ptr[n0]), int(node_ptr[n1])
    cols = node_col[s:e].astype(np.int64)
    rows = np.repeat(np.arange(n0, n1, dtype=np.int64), np.diff(node_ptr[n0:n1 + 1]))
    outside = (cols < n0) | (cols >= n1)
    cols_o, rows_o = cols[outside], rows[outside]
    owner = np.searchsorted(cuts, cols_o, side="right") - 1
    peers = np.unique(owner)
    send_off, recv_off = [0], [0]
    send_idx, recv_idx = [], []
    expand = lambda nodes: (dpn * nodes[:, None] + np.arange(dpn)[None, :]).ravel()
    for pr in peers:
        sel = owner == pr
        ghosts = np.unique(cols_o[sel])           # their nodes we read
        mine = np.unique(rows_o[sel])             # our nodes they read (symmetry)
        recv_idx.append(expand(ghosts))
        send_idx.append(expand(mine))
        recv_off.append(recv_off[-1] + dpn * ghosts.size)
        send_off.append(send_off[-1] + dpn * mine.size)
    cat = lambda xs: (np.concatenate(xs) if xs else np.zeros(0, dtype=np.int64)).astype(np.int32)
    return (peers.astype(np.int32), np.asarray(send_off, dtype=np.int64), cat(send_idx),
            np.asarray(recv_off, dtype=np.int64), cat(recv_idx))
