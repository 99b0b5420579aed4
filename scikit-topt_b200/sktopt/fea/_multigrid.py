"""Host-side set-up of the geometric multigrid preconditioner (``csrc/mg.cu``).

Stands in for ``pyamg.smoothed_aggregation_solver(K).aspreconditioner()`` of the
reference's ``cg_pyamg`` path (``fea/solver_elastic.py:94-104``) on
tensor-product hexahedral meshes (``create_box_hex``): the grid hierarchy halves
the cell counts per level, prolongation is trilinear, and the coarse operators
are exact Galerkin products formed element-wise on the device every time K(rho)
changes.  A preconditioner only changes the iteration count, not the converged
solution, so parity tests are unaffected.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from sktopt._b200 import device as dev
from sktopt._b200 import lib as _lib
from sktopt._fem import MeshHex


def detect_tensor_grid(mesh):
    """(xs, ys, zs) if ``mesh`` is exactly ``MeshHex.init_tensor(xs, ys, zs)``
    (same node numbering and connectivity), else None."""
    if not isinstance(mesh, MeshHex):
        return None
    p = mesh.p
    xs, ys, zs = (np.unique(p[d]) for d in range(3))
    if xs.size * ys.size * zs.size != p.shape[1] or min(xs.size, ys.size, zs.size) < 2:
        return None
    ref = MeshHex.init_tensor(xs, ys, zs)
    if ref.t.shape != mesh.t.shape or not np.array_equal(ref.t, mesh.t):
        return None
    if not np.array_equal(ref.p, p):
        return None
    return xs, ys, zs


def lattice_graph(nx: int, ny: int, nz: int):
    """Node graph (CSR, columns ascending, int32) of the 27-point lattice stencil
    on nx x ny x nz nodes, node = iy + ny*ix + ny*nx*iz."""
    iz, ix, iy = np.meshgrid(np.arange(nz), np.arange(nx), np.arange(ny), indexing="ij")
    iz, ix, iy = iz.ravel(), ix.ravel(), iy.ravel()
    cols, ok = [], []
    for dz in (-1, 0, 1):
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                jz, jx, jy = iz + dz, ix + dx, iy + dy
                ok.append((jz >= 0) & (jz < nz) & (jx >= 0) & (jx < nx) & (jy >= 0) & (jy < ny))
                cols.append(jy + ny * (jx + nx * jz))
    ok = np.stack(ok, axis=1)
    cols = np.stack(cols, axis=1)
    rp = np.concatenate([[0], np.cumsum(ok.sum(axis=1))]).astype(np.int32)
    return rp, cols[ok].astype(np.int32)


def detect_lattice(node_ptr, node_col, n_nodes: int):
    """(nx, ny, nz) node counts if the node graph is a sub-graph of the 27-point
    stencil of a lattice numbered like ``init_tensor`` (node = iy + ny*ix +
    ny*nx*iz, at least 3 nodes along x and y), else None.  Only the topology
    matters: the geometry may be jittered or graded and the elements may be
    hexahedra or the Kuhn tetrahedra of ``MeshTet.init_tensor``."""
    rp = np.asarray(node_ptr, dtype=np.int64)
    ci = np.asarray(node_col, dtype=np.int64)
    n = int(n_nodes)
    if n < 27 or ci.size == 0:
        return None
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    pos = np.unique((ci - rows)[ci > rows])
    if pos.size == 0 or pos[0] != 1:
        return None

    def split(v, m):
        """v = r + m*q with r in {-1, 0, 1}, or None"""
        q = (v + 1) // m
        r = v - m * q
        return (q, r) if np.all(np.abs(r) <= 1) else None

    big = pos[pos > 1]
    if big.size == 0:
        return None
    for ny in (int(big[0]), int(big[0]) + 1):
        if ny < 3 or n % ny:
            continue
        sp = split(pos, ny)
        if sp is None:
            continue
        m = np.unique(sp[0][sp[0] > 0])
        if m.size == 0 or m[0] != 1:
            continue
        mb = m[m > 1]
        cand = [n // ny] if mb.size == 0 else [int(mb[0]), int(mb[0]) + 1]
        for nx in cand:
            if nx < 3 or (n // ny) % nx:
                continue
            nz = n // (ny * nx)
            if split(m, nx) is None:
                continue
            # full check: every edge joins lattice neighbours
            ay, ar = rows % ny, rows // ny
            by, br = ci % ny, ci // ny
            if (np.all(np.abs(ay - by) <= 1) and np.all(np.abs(ar % nx - br % nx) <= 1)
                    and np.all(np.abs(ar // nx - br // nx) <= 1)):
                return int(nx), int(ny), int(nz)
    return None


def coarse_index_map(n_cells: int) -> np.ndarray:
    """Fine node index of every coarse node along one axis."""
    nc = (n_cells + 1) // 2
    return np.minimum(2 * np.arange(nc + 1), n_cells)


def axis_tables(n_cells: int):
    """1-D linear interpolation between the nested node sets of one axis.

    Returns (c0, c1, w0, w1) indexed by fine node and (fT (3, nc+1), wT (3, nc+1))
    indexed by coarse node (-1 = empty slot)."""
    fmap = coarse_index_map(n_cells)
    nc = fmap.size - 1
    c0 = np.zeros(n_cells + 1, dtype=np.int32)
    c1 = np.zeros(n_cells + 1, dtype=np.int32)
    w0 = np.zeros(n_cells + 1)
    w1 = np.zeros(n_cells + 1)
    coarse_of = {int(f): i for i, f in enumerate(fmap)}
    for i in range(n_cells + 1):
        if i in coarse_of:
            c0[i] = c1[i] = coarse_of[i]
            w0[i] = 1.0
        else:                       # midpoint of a full coarse cell
            c0[i] = coarse_of[i - 1]
            c1[i] = coarse_of[i + 1]
            w0[i] = w1[i] = 0.5
    fT = np.full((3, nc + 1), -1, dtype=np.int32)
    wT = np.zeros((3, nc + 1))
    fill = np.zeros(nc + 1, dtype=int)
    for i in range(n_cells + 1):
        for c, w in ((c0[i], w0[i]), (c1[i], w1[i])):
            if w != 0.0:
                fT[fill[c], c] = i
                wT[fill[c], c] = w
                fill[c] += 1
    return c0, c1, w0, w1, fT, wT


_W1D = {
    (0, 0): np.array([[1.0, 0.0], [0.5, 0.5]]),   # full parent, first child
    (0, 1): np.array([[0.5, 0.5], [0.0, 1.0]]),   # full parent, second child
    (1, 0): np.array([[1.0, 0.0], [0.0, 1.0]]),   # half parent (single child)
    (1, 1): np.zeros((2, 2)),
}


def vertex_bits(mesh) -> np.ndarray:
    """(8, 3) 0/1 offsets (x, y, z) of the local vertices, read off element 0."""
    xe = mesh.p[:, mesh.t[:, 0]]
    lo = xe.min(axis=1, keepdims=True)
    hi = xe.max(axis=1, keepdims=True)
    return (xe > 0.5 * (lo + hi)).T.astype(int)


def q_tables(bits: np.ndarray) -> np.ndarray:
    """Qtab[type][child][a][A]: weight of parent vertex A in child vertex a.
    type = tx + 2 ty + 4 tz (t = 1: the parent has a single child along that
    axis), child = cx + 2 cy + 4 cz."""
    Q = np.zeros((8, 8, 8, 8))
    for ty_ in range(8):
        t = (ty_ & 1, (ty_ >> 1) & 1, (ty_ >> 2) & 1)
        for ch in range(8):
            c = (ch & 1, (ch >> 1) & 1, (ch >> 2) & 1)
            W = [_W1D[(t[d], c[d])] for d in range(3)]
            for a in range(8):
                for A in range(8):
                    Q[ty_, ch, a, A] = np.prod([W[d][bits[a, d], bits[A, d]] for d in range(3)])
    return Q


def child_tables(fine_cells, coarse_cells):
    """child[c][E] (fine element id or -1) and ptype[E] for every coarse
    element; element id = ey + ny*ex + ny*nx*ez on both levels."""
    nx, ny, nz = fine_cells
    cx_, cy_, cz_ = coarse_cells
    Ez, Ex, Ey = np.meshgrid(np.arange(cz_), np.arange(cx_), np.arange(cy_), indexing="ij")
    Ex, Ey, Ez = Ex.ravel(), Ey.ravel(), Ez.ravel()      # ordered by element id
    child = np.full((8, Ex.size), -1, dtype=np.int32)
    for ch in range(8):
        c = (ch & 1, (ch >> 1) & 1, (ch >> 2) & 1)
        fx, fy, fz = 2 * Ex + c[0], 2 * Ey + c[1], 2 * Ez + c[2]
        ok = (fx < nx) & (fy < ny) & (fz < nz)
        child[ch, ok] = (fy + ny * fx + ny * nx * fz)[ok]
    half = lambda E, n: (2 * E + 1 >= n).astype(np.uint8)
    ptype = (half(Ex, nx) + 2 * half(Ey, ny) + 4 * half(Ez, nz)).astype(np.uint8)
    return child, ptype


def chebyshev_coefficients(lam_max: float, lam_min: float, degree: int):
    """(c1, c2) of the Chebyshev iteration for a spectrum in [lam_min, lam_max]
    (Saad, Iterative Methods, Alg. 12.1): d_k = c1[k] d_{k-1} + c2[k] z_k."""
    theta, delta = 0.5 * (lam_max + lam_min), 0.5 * (lam_max - lam_min)
    sigma = theta / delta
    c1, c2 = np.zeros(degree), np.zeros(degree)
    rho = 1.0 / sigma
    c2[0] = 1.0 / theta
    for k in range(1, degree):
        rho_new = 1.0 / (2.0 * sigma - rho)
        c1[k] = rho_new * rho
        c2[k] = 2.0 * rho_new / delta
        rho = rho_new
    return c1, c2


def plan_slab_sharding(coords, plane_cuts0, world: int, rank: int, shard_min: int):
    """z-slab ownership of every multigrid level, pure host logic.

    ``coords[l]`` = (xs, ys, zs) node coordinates of level l; ``plane_cuts0`` = the
    world+1 node-plane boundaries of level 0.  A coarse plane belongs to the rank
    that owns the coincident fine plane.  Assembled levels with at least
    ``shard_min`` nodes (and >= 2 planes on every rank) are sharded: a rank stores
    the rows of its planes and computes the Galerkin element matrices of its slab
    only (plus the element planes the coarser levels' products need).  The first
    replicated level gets its element matrices from slab-wise products that are
    all-gathered (an even split of its element planes).

    Returns (shard, first_replicated, gather_plan, plane_cuts): ``shard[l]`` is
    None (level 0 and replicated levels) or a dict with the owned node range
    [node0, node1), ``plane`` nodes per plane, ``prev`` / ``next`` ranks and the
    element range [elem0, elem1) whose matrices the rank computes."""
    L = len(coords)
    shard = [None] * L
    zc = [np.asarray(plane_cuts0, dtype=np.int64)]           # plane cuts per level
    for l in range(1, L):
        nz_f = coords[l - 1][2].size - 1
        fmap = coarse_index_map(nz_f)                        # fine plane of coarse plane
        owner = np.searchsorted(zc[-1], fmap, side="right") - 1
        zc.append(np.searchsorted(owner, np.arange(world + 1), side="left").astype(np.int64))
    last = 0
    for l in range(1, L - 1):
        xs, ys, zs = coords[l]
        if xs.size * ys.size * zs.size < shard_min or np.diff(zc[l]).min() < 2:
            break
        last = l
    if last == 0:
        return shard, 1, None, zc
    lr = last + 1
    cz_lr = coords[lr][2].size - 1
    a = np.round(np.linspace(0, cz_lr, world + 1)).astype(np.int64)
    plane_e = lambda l: (coords[l][0].size - 1) * (coords[l][1].size - 1)
    gather_plan = dict(level=lr, cuts=a, plane_elems=plane_e(lr))
    need_lo, need_hi = int(a[rank]), int(a[rank + 1])
    for l in range(last, 0, -1):
        cz = coords[l][2].size - 1
        z0, z1 = int(zc[l][rank]), int(zc[l][rank + 1])
        lo, hi = max(z0 - 1, 0), min(z1, cz)                 # element planes touching my nodes
        if need_hi > need_lo:                                # + the children of what level l+1 needs
            lo, hi = min(lo, 2 * need_lo), max(hi, min(2 * need_hi, cz))
        npl = coords[l][0].size * coords[l][1].size
        shard[l] = dict(node0=z0 * npl, node1=z1 * npl, plane=npl,
                        prev=rank - 1 if rank > 0 else -1,
                        next=rank + 1 if rank < world - 1 else -1,
                        elem0=lo * plane_e(l), elem1=hi * plane_e(l))
        need_lo, need_hi = lo, hi
    return shard, lr, gather_plan, zc


class Multigrid:
    """Grid hierarchy + Galerkin set-up for one elasticity engine."""

    MIN_FINE_NODES = 1500
    DENSE_MAX_DOFS = 160
    FP32_LEVEL_MIN_NODES = 50000

    def __init__(self, engine, axes, omega: float | None = None, nu_coarse: int = 30,
                 coarsest_max_cells: int = 6, algebraic: bool = False):
        self.lib = _lib.load()
        self.eng = engine
        # algebraic = True: lattice topology only (any geometry / element type);
        # level 0 is the engine's assembled node-block matrix and the coarse
        # operators are P^T A P formed matrix to matrix (csrc/galerkin_bsr.cu);
        # ``axes`` then only carries the node counts
        self.algebraic = bool(algebraic)
        self.omega_auto = omega is None
        self.omega = 0.5 if omega is None else float(omega)
        nu_coarse = int(os.environ.get("SKTOPT_B200_MG_NU_COARSE", nu_coarse))
        self.nu_coarse = int(nu_coarse)
        self.lambda_max = None
        omega = self.omega
        xs, ys, zs = axes
        coords = [(xs, ys, zs)]
        while True:
            cx, cy, cz = (c.size - 1 for c in coords[-1])
            # stop once the coarsest level fits the dense exact solve
            # (csrc/mg.cu: kDenseMax dofs) or cannot be halved any more
            if 3 * (cx + 1) * (cy + 1) * (cz + 1) <= self.DENSE_MAX_DOFS or max(cx, cy, cz) <= 1:
                break
            coords.append(tuple(a[coarse_index_map(a.size - 1)] for a in coords[-1]))
        self.coords = coords
        self.n_levels = len(coords)
        if self.n_levels < 2:
            raise ValueError("grid too small for a multigrid hierarchy")
        if not self.algebraic:
            fine_mesh = engine.basis.mesh
            bits = vertex_bits(fine_mesh)
            self.Qtab = dev.to_dev(q_tables(bits).ravel())
        h = C.c_void_p()
        _lib.check(self.lib.sktb_mg_create(C.byref(h), self.n_levels, torch.cuda.current_device()))
        self.handle = h
        _lib.check(self.lib.sktb_mg_set_params(h, float(omega), int(nu_coarse)))
        # level-0 products of the V-cycle in single precision (matrix-free level 0
        # only; SKTOPT_B200_MG_FP32=0 keeps them in fp64)
        self.fp32 = os.environ.get("SKTOPT_B200_MG_FP32", "1") != "0"
        _lib.check(self.lib.sktb_mg_set_precision(h, int(self.fp32)))
        # one cooperative kernel for the levels <= 1024 nodes; measured equal to the
        # launch-per-operation path at C2 (grid barriers cost what launches cost),
        # so it is opt-in
        self.fused_tail = os.environ.get("SKTOPT_B200_MG_FUSED_TAIL", "0") == "1"
        _lib.check(self.lib.sktb_mg_set_fused_tail(h, int(self.fused_tail)))
        # assembled levels with at least FP32_LEVEL_MIN_NODES nodes keep a
        # single-precision copy of their values for the V-cycle's products: such a
        # level is an HBM stream (C2 level 1: 0.26 GB read twice per cycle), the
        # copy halves it; accumulation, diagonal and damping stay fp64
        # (SKTOPT_B200_MG_FP32_LEVELS=0 disables)
        self.fp32_levels = (self.fp32
                            and os.environ.get("SKTOPT_B200_MG_FP32_LEVELS", "1") != "0")
        self._vals32_l0 = None
        self._fp32_min_nodes = int(os.environ.get("SKTOPT_B200_MG_FP32_MIN_NODES",
                                                  self.FP32_LEVEL_MIN_NODES))
        self.levels = [None]          # level 0 lives in the engine
        self.transfers = []
        mask_f = engine.dir_mask.cpu().numpy()
        self._plan_sharding()
        for l in range(1, self.n_levels):
            cxs, cys, czs = coords[l]
            fine_cells = tuple(c.size - 1 for c in coords[l - 1])
            coarse_cells = tuple(c.size - 1 for c in coords[l])
            if self.algebraic:
                dm = None
                rp_h, ci_h = lattice_graph(cxs.size, cys.size, czs.size)
                n_nodes_c, n_elem_c = int(cxs.size * cys.size * czs.size), 0
            else:
                mesh_c = MeshHex.init_tensor(cxs, cys, czs)
                if not np.array_equal(vertex_bits(mesh_c), bits):
                    raise RuntimeError("coarse and fine meshes disagree on local vertex order")
                dm = dev.DeviceMesh(mesh_c)
                rp_h, ci_h = dm.node_graph()
                n_nodes_c, n_elem_c = dm.n_nodes, dm.n_elem
                child, ptype = child_tables(fine_cells, coarse_cells)
            sh = self.shard[l]
            if sh is not None:
                # owned rows only (global columns); element matrices of the slab
                n0, n1 = sh["node0"], sh["node1"]
                s_, e_ = int(rp_h[n0]), int(rp_h[n1])
                rp_loc, ci_loc = rp_h[n0:n1 + 1] - s_, ci_h[s_:e_]
                n_ke = sh["elem1"] - sh["elem0"]
            else:
                rp_loc, ci_loc, n0, n1 = rp_h, ci_h, 0, n_nodes_c
                n_ke = n_elem_c
            # Dirichlet mask: a coarse dof is fixed iff the coincident fine dof is
            fm = [coarse_index_map(n) for n in fine_cells]
            npx_f, npy_f = fine_cells[0] + 1, fine_cells[1] + 1
            Iz, Ix, Iy = np.meshgrid(np.arange(coarse_cells[2] + 1), np.arange(coarse_cells[0] + 1),
                                     np.arange(coarse_cells[1] + 1), indexing="ij")
            fnode = (fm[1][Iy] + npy_f * fm[0][Ix] + npy_f * npx_f * fm[2][Iz]).ravel()
            mask_c = mask_f.reshape(-1, 3)[fnode].ravel().copy()
            lvl = dict(
                dm=dm, n_nodes=n_nodes_c, n_elem=n_elem_c, node0=n0, node1=n1,
                node_ptr=dev.to_dev(rp_loc, dev.I32), node_col=dev.to_dev(ci_loc, dev.I32),
                max_deg=int(np.diff(rp_loc).max()),
                vals=torch.empty(9 * ci_loc.size, dtype=dev.F64, device="cuda"),
                inv_diag=torch.empty(3 * (n1 - n0), dtype=dev.F64, device="cuda"),
                mask=dev.to_dev(mask_c, dev.U8),
            )
            if self.fp32_levels and (n1 - n0) >= self._fp32_min_nodes:
                lvl["vals32"] = torch.empty(9 * ci_loc.size, dtype=torch.float32, device="cuda")
            if not self.algebraic:
                lvl.update(ke=torch.empty((n_ke, 576), dtype=dev.F64, device="cuda"),
                           child=dev.to_dev(child, dev.I32), ptype=dev.to_dev(ptype, dev.U8))
            self.levels.append(lvl)
            # transfer tables fine (l-1) -> coarse (l), concatenated [x | y | z]
            tabs = [axis_tables(n) for n in fine_cells]
            cat = lambda k, dt: np.ascontiguousarray(np.concatenate([t[k] for t in tabs]).astype(dt))
            catT = lambda k, dt: np.ascontiguousarray(np.concatenate([t[k] for t in tabs], axis=1).astype(dt))
            tr = dict(
                c0=dev.to_dev(cat(0, np.int32), dev.I32), c1=dev.to_dev(cat(1, np.int32), dev.I32),
                w0=dev.to_dev(cat(2, np.float64)), w1=dev.to_dev(cat(3, np.float64)),
                fT=dev.to_dev(catT(4, np.int32).ravel(), dev.I32), wT=dev.to_dev(catT(5, np.float64).ravel()),
                fnp=np.array([n + 1 for n in fine_cells], dtype=np.int32),
                cnp=np.array([n + 1 for n in coarse_cells], dtype=np.int32),
            )
            self.transfers.append(tr)
            _lib.check(self.lib.sktb_mg_set_transfer(
                h, l - 1, tr["fnp"].ctypes.data_as(C.c_void_p), tr["cnp"].ctypes.data_as(C.c_void_p),
                dev._ptr(tr["c0"]), dev._ptr(tr["c1"]), dev._ptr(tr["w0"]), dev._ptr(tr["w1"]),
                dev._ptr(tr["fT"]), dev._ptr(tr["wT"])))
            mask_f = mask_c
        # smoothing sweeps per level: "a,b,c,..." for levels 0,1,2,... (last value
        # repeats).  Levels 0/1 carry the cost, the cheap coarse levels get more sweeps
        # (measured at C2: 34 -> 22 PCG iterations)
        # (lattice hierarchy, assembled level 0: the products are cheap next to the
        # ~50 small launches of a cycle, two sweeps everywhere pay: C3 44.6 -> 37.1 ms)
        sweeps = [int(v) for v in os.environ.get(
            "SKTOPT_B200_MG_SWEEPS", "2,2,2,3" if self.algebraic else "1,1,2,3").split(",")]
        self.sweeps = [sweeps[min(l, len(sweeps) - 1)] for l in range(self.n_levels)]
        for l, nu in enumerate(self.sweeps):
            _lib.check(self.lib.sktb_mg_set_level_sweeps(h, l, nu))
        self.cheb_alpha = float(os.environ.get("SKTOPT_B200_MG_CHEB_ALPHA", "0"))
        self.setup_count = 0
        # level 0 -> 1 tables T[cls][type][c] = Q_c^T Ke0[cls] Q_c (host, once)
        self.T01 = None
        if self.algebraic:
            return
        ke0 = engine.unit_ke.cpu().numpy().reshape(-1, 24, 24)
        Q = q_tables(bits)
        if ke0.shape[0] <= 16:
            T = np.zeros((ke0.shape[0], 8, 8, 24, 24))
            eye3 = np.eye(3)
            for k in range(ke0.shape[0]):
                for ty_ in range(8):
                    for ch in range(8):
                        Qv = np.kron(Q[ty_, ch], eye3)          # (24 child dofs, 24 parent dofs)
                        T[k, ty_, ch] = Qv.T @ ke0[k] @ Qv
            self.T01 = dev.to_dev(T.ravel())

    # ------------------------------------------------------------ sharding --
    SHARD_MIN_NODES = 300000

    def _plan_sharding(self):
        """z-slab ownership of every level: see ``plan_slab_sharding``."""
        eng = self.eng
        self.shard = [None] * self.n_levels
        self.first_replicated = 1
        self.gather_plan = None
        comm = eng.comm
        if comm is None or getattr(eng, "plane_cuts", None) is None or self.algebraic:
            return
        shard_min = int(os.environ.get("SKTOPT_B200_MG_SHARD_MIN", self.SHARD_MIN_NODES))
        self.shard, self.first_replicated, self.gather_plan, self.plane_cuts = \
            plan_slab_sharding(self.coords, eng.plane_cuts, comm.world, comm.rank, shard_min)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_mg_destroy(h)
            except Exception:
                pass
            self.handle = None

    def _lambda_max_level(self, l: int, iters: int = 12) -> float:
        """Largest eigenvalue of D^-1 A of level ``l`` by power iteration.  Level
        0 runs inside the PCG workspace (it may be row-sharded); the replicated
        coarse levels use plain device kernels.  Every rank gets the same value."""
        if l == 0 and self.eng.matrix_free:
            eng = self.eng
            out = C.c_double()
            _lib.check(self.lib.sktb_pcg_lambda_max_grid(
                eng.pcg.handle, eng.gridop.handle, dev._ptr(eng.inv_diag), int(iters),
                C.cast(C.byref(out), C.c_void_p), dev._stream()))
            return float(out.value)
        if l == 0:
            eng = self.eng
            out = C.c_double()
            _lib.check(self.lib.sktb_pcg_lambda_max_bsr3(
                eng.pcg.handle, dev._ptr(eng.node_ptr_loc), dev._ptr(eng.node_col_loc),
                int(eng.node_col_loc.numel()), int(eng.max_deg), dev._ptr(eng.vals),
                dev._ptr(eng.inv_diag), int(iters), C.cast(C.byref(out), C.c_void_p),
                dev._stream()))
            return float(out.value)
        lv = self.levels[l]
        if self.shard[l] is not None:
            return self._lambda_max_sharded(l, iters)
        n = 3 * lv["n_nodes"]
        g = torch.Generator(device="cuda")
        g.manual_seed(1234)
        v = torch.rand(n, dtype=dev.F64, device="cuda", generator=g) - 0.5
        w = torch.empty_like(v)
        lam = 1.0
        for _ in range(iters):
            v /= float(np.sqrt(dev.dot(v, v)))
            dev.spmv_bsr3(lv["node_ptr"], lv["node_col"], lv["vals"], v, out=w)
            dev.hadamard(1.0, w, lv["inv_diag"], w)
            lam = dev.dot(v, w)
            v, w = w, v
        return float(lam)

    def _lambda_max_sharded(self, l: int, iters: int) -> float:
        """Power iteration on a z-slab-sharded level: products over the owned rows
        (ghost planes exchanged inside ``sktb_mg_level_apply``), dots all-reduced.
        The start vector is a function of the global index, so every rank count
        gives the same estimate."""
        lv, sh, eng = self.levels[l], self.shard[l], self.eng
        n_glob = 3 * lv["n_nodes"]
        lo, hi = 3 * sh["node0"], 3 * sh["node1"]
        idx = torch.arange(n_glob, dtype=torch.int64, device="cuda")
        v = ((idx * 2654435761) % 1048573).to(dev.F64) / 1048573.0 - 0.5
        w = torch.empty(hi - lo, dtype=dev.F64, device="cuda")
        pack = torch.empty(1, dtype=dev.F64, device="cuda")

        def gdot(a, b):
            pack[0] = dev.dot(a, b)
            return float(eng.comm.allreduce_sum(pack)[0])
        lam = 1.0
        for _ in range(iters):
            own = v[lo:hi]
            own /= float(np.sqrt(gdot(own, own)))
            _lib.check(self.lib.sktb_mg_level_apply(self.handle, l, eng.pcg.handle, dev._ptr(v),
                                                    dev._ptr(w), dev._stream()))
            dev.hadamard(1.0, w, lv["inv_diag"], w)
            lam = gdot(own, w)
            own.copy_(w)
        return float(lam)

    def _galerkin_level(self, l: int, st):
        """Element matrices of level ``l`` for the element range this rank needs."""
        eng, lib, lv = self.eng, self.lib, self.levels[l]
        sh = self.shard[l]
        gp = self.gather_plan
        ke_out = lv["ke"]
        if sh is not None:
            e_lo, e_hi = sh["elem0"], sh["elem1"]
        elif gp is not None and gp["level"] == l:
            r = eng.comm.rank
            e_lo, e_hi = (int(gp["cuts"][r]) * gp["plane_elems"],
                          int(gp["cuts"][r + 1]) * gp["plane_elems"])
            ke_out = lv["ke"][e_lo:]           # written at its global position
        else:
            e_lo, e_hi = 0, lv["n_elem"]
        if e_hi > e_lo:
            if l == 1 and self.T01 is not None:
                _lib.check(lib.sktb_elem_combine_range(
                    lv["n_elem"], e_lo, e_hi, dev._ptr(lv["child"]), dev._ptr(lv["ptype"]),
                    dev._ptr(self.T01), dev._ptr(eng.dm.elem_class), dev._ptr(eng.scale),
                    dev._ptr(ke_out), st))
            elif l == 1:
                _lib.check(lib.sktb_elem_restrict_range(
                    lv["n_elem"], e_lo, e_hi, 0, dev._ptr(lv["child"]), dev._ptr(lv["ptype"]),
                    dev._ptr(self.Qtab), None, dev._ptr(eng.unit_ke), dev._ptr(eng.dm.elem_class),
                    dev._ptr(eng.scale), dev._ptr(ke_out), st))
            else:
                fsh = self.shard[l - 1]
                _lib.check(lib.sktb_elem_restrict_range(
                    lv["n_elem"], e_lo, e_hi, 0 if fsh is None else fsh["elem0"],
                    dev._ptr(lv["child"]), dev._ptr(lv["ptype"]), dev._ptr(self.Qtab),
                    dev._ptr(self.levels[l - 1]["ke"]), None, None, None, dev._ptr(ke_out), st))
        if sh is None and gp is not None and gp["level"] == l:
            # slab-wise products -> whole level on every rank
            cnt = np.diff(gp["cuts"]) * gp["plane_elems"] * 576
            eng.comm.allgatherv(lv["ke"].view(-1), cnt, np.concatenate([[0], np.cumsum(cnt)[:-1]]))

    def _push_vals32(self, l: int, handle=None, convert: bool = True):
        """Refresh the single-precision copy of level ``l`` and hand it to the
        hierarchy (after ``sktb_mg_set_level``, which forgets it)."""
        lv = self.levels[l]
        v32 = lv.get("vals32")
        if v32 is None:
            return
        if convert:
            _lib.check(self.lib.sktb_f64_to_f32(int(v32.numel()), dev._ptr(lv["vals"]),
                                                C.c_void_p(v32.data_ptr()), dev._stream()))
        _lib.check(self.lib.sktb_mg_set_level_vals32(
            self.handle if handle is None else handle, l, C.c_void_p(v32.data_ptr())))

    def _push_vals32_level0(self, handle=None, convert: bool = True):
        """Assembled level 0 (one GPU): the V-cycle's sweeps and residual products
        stream a single-precision copy of K, the PCG's own product stays fp64."""
        eng = self.eng
        if (not self.fp32_levels or eng.sharded
                or (eng.node1 - eng.node0) < self._fp32_min_nodes):
            return
        v = eng.vals
        if self._vals32_l0 is None or self._vals32_l0.numel() != v.numel():
            self._vals32_l0 = torch.empty(v.numel(), dtype=torch.float32, device="cuda")
            convert = True
        if convert:
            _lib.check(self.lib.sktb_f64_to_f32(int(v.numel()), dev._ptr(v),
                                                C.c_void_p(self._vals32_l0.data_ptr()),
                                                dev._stream()))
        _lib.check(self.lib.sktb_mg_set_level_vals32(
            self.handle if handle is None else handle, 0,
            C.c_void_p(self._vals32_l0.data_ptr())))

    def _galerkin_algebraic(self, l: int, st):
        """vals of level ``l`` = P^T A_{l-1} P, matrix to matrix."""
        eng, lv, tr = self.eng, self.levels[l], self.transfers[l - 1]
        if l == 1:
            fp, fc, fv, fm = eng.node_ptr_loc, eng.node_col_loc, eng.vals, eng.dir_mask
        else:
            f = self.levels[l - 1]
            fp, fc, fv, fm = f["node_ptr"], f["node_col"], f["vals"], f["mask"]
        _lib.check(self.lib.sktb_galerkin_bsr3_lattice(
            tr["fnp"].ctypes.data_as(C.c_void_p), tr["cnp"].ctypes.data_as(C.c_void_p),
            dev._ptr(tr["c0"]), dev._ptr(tr["c1"]), dev._ptr(tr["w0"]), dev._ptr(tr["w1"]),
            dev._ptr(tr["fT"]), dev._ptr(tr["wT"]), dev._ptr(fp), dev._ptr(fc), dev._ptr(fv),
            dev._ptr(fm), dev._ptr(lv["node_ptr"]), dev._ptr(lv["node_col"]),
            dev._ptr(lv["mask"]), dev._ptr(lv["vals"]), st))

    def setup(self):
        """Galerkin coarse operators for the engine's current modulus field;
        call after the engine assembled level 0 and its inverse diagonal."""
        eng = self.eng

        st = dev._stream()
        lib = self.lib
        _lib.check(lib.sktb_mg_set_level0_range(self.handle, int(eng.node0), int(eng.dm.n_nodes)))
        sh0 = getattr(eng, "slab", None)
        if sh0 is not None:
            _lib.check(lib.sktb_mg_set_level_slab(
                self.handle, 0, int(eng.node0), int(eng.dm.n_nodes), int(sh0["plane"]),
                int(sh0["prev"]), int(sh0["next"])))
        if eng.matrix_free:
            _lib.check(lib.sktb_mg_set_level0_grid(
                self.handle, eng.gridop.handle, int(eng.node1 - eng.node0),
                dev._ptr(eng.inv_diag), dev._ptr(eng.dir_mask)))
        else:
            _lib.check(lib.sktb_mg_set_level(
                self.handle, 0, int(eng.node1 - eng.node0), int(eng.node_col_loc.numel()),
                int(eng.max_deg), dev._ptr(eng.node_ptr_loc), dev._ptr(eng.node_col_loc),
                dev._ptr(eng.vals), dev._ptr(eng.inv_diag), dev._ptr(eng.dir_mask)))
            self._push_vals32_level0()
        for l in range(1, self.n_levels):
            lv = self.levels[l]
            sh = self.shard[l]
            if self.algebraic:
                self._galerkin_algebraic(l, st)
                dev.bsr3_inv_diag(lv["node_ptr"], lv["node_col"], lv["vals"], out=lv["inv_diag"])
                _lib.check(lib.sktb_mg_set_level(
                    self.handle, l, lv["node1"] - lv["node0"], int(lv["node_col"].numel()),
                    lv["max_deg"], dev._ptr(lv["node_ptr"]), dev._ptr(lv["node_col"]),
                    dev._ptr(lv["vals"]), dev._ptr(lv["inv_diag"]), dev._ptr(lv["mask"])))
                self._push_vals32(l)
                continue
            self._galerkin_level(l, st)
            if sh is not None:
                lv["dm"].assemble_rows(3, sh["node0"], sh["node1"], lv["ke"], scale=None,
                                       dir_mask=lv["mask"], out=lv["vals"], per_element=True,
                                       ke_base=sh["elem0"])
                dev.bsr3_inv_diag(lv["node_ptr"], lv["node_col"], lv["vals"], out=lv["inv_diag"],
                                  node0=sh["node0"])
                _lib.check(lib.sktb_mg_set_level_slab(
                    self.handle, l, sh["node0"], lv["n_nodes"], sh["plane"], sh["prev"],
                    sh["next"]))
            else:
                lv["dm"].assemble(3, lv["ke"], scale=None, dir_mask=lv["mask"], out=lv["vals"],
                                  per_element=True)
                dev.bsr3_inv_diag(lv["node_ptr"], lv["node_col"], lv["vals"], out=lv["inv_diag"])
            _lib.check(lib.sktb_mg_set_level(
                self.handle, l, lv["node1"] - lv["node0"], int(lv["node_col"].numel()),
                lv["max_deg"], dev._ptr(lv["node_ptr"]), dev._ptr(lv["node_col"]),
                dev._ptr(lv["vals"]), dev._ptr(lv["inv_diag"]), dev._ptr(lv["mask"])))
            self._push_vals32(l)
        _lib.check(lib.sktb_mg_factor_coarsest(self.handle, st))
        if self.omega_auto and self.setup_count == 0:
            # per-level damping: omega_l * lambda_max_l ~ 1.75, inside the
            # stability bound 2 (the power iteration approaches lambda_max from
            # below, hence the 1.03).  Done once: lambda_max(D^-1 A) depends on
            # the discretisation far more than on the density field.
            self.lambda_max = []
            for l in range(self.n_levels):
                lam = self._lambda_max_level(l)
                self.lambda_max.append(lam)
                _lib.check(self.lib.sktb_mg_set_level_omega(
                    self.handle, l, float(1.75 / (1.03 * lam))))
                if self.cheb_alpha > 0.0 and 1 <= l < self.n_levels - 1 and self.sweeps[l] >= 2:
                    c1, c2 = chebyshev_coefficients(1.1 * lam, 1.1 * lam / self.cheb_alpha,
                                                    self.sweeps[l])
                    _lib.check(self.lib.sktb_mg_set_level_cheby(
                        self.handle, l, self.sweeps[l], c1.ctypes.data_as(C.c_void_p),
                        c2.ctypes.data_as(C.c_void_p)))
        self.setup_count += 1

    def vcycle(self, r, z=None):
        if z is None:
            z = torch.empty_like(r)
        wait = getattr(self.eng, "_wait_mg", None)
        if wait is not None:
            wait()                 # a set-up running on the engine's side stream
        _lib.check(self.lib.sktb_mg_vcycle(self.handle, dev._ptr(r), dev._ptr(z), dev._stream()))
        return z


class MultigridWorkspace:
    """A second set of V-cycle work vectors on the level operators of ``parent``
    (assembled level 0, one GPU): load cases solved concurrently on their own CUDA
    streams each need their own iterates and right-hand sides, the operators,
    masks, transfer tables, damping and the exact coarsest-level inverse are the
    parent's.  ``refresh`` must follow every ``parent.setup()``."""

    def __init__(self, parent: "Multigrid"):
        if parent.eng.matrix_free or parent.eng.sharded or parent.cheb_alpha > 0.0:
            raise ValueError("workspace clones need an assembled, unsharded hierarchy")
        self.parent = parent
        self.lib = lib = parent.lib
        h = C.c_void_p()
        _lib.check(lib.sktb_mg_create(C.byref(h), parent.n_levels, torch.cuda.current_device()))
        self.handle = h
        _lib.check(lib.sktb_mg_set_params(h, float(parent.omega), int(parent.nu_coarse)))
        _lib.check(lib.sktb_mg_set_precision(h, int(parent.fp32)))
        _lib.check(lib.sktb_mg_set_fused_tail(h, 0))
        for l, tr in enumerate(parent.transfers):
            _lib.check(lib.sktb_mg_set_transfer(
                h, l, tr["fnp"].ctypes.data_as(C.c_void_p), tr["cnp"].ctypes.data_as(C.c_void_p),
                dev._ptr(tr["c0"]), dev._ptr(tr["c1"]), dev._ptr(tr["w0"]), dev._ptr(tr["w1"]),
                dev._ptr(tr["fT"]), dev._ptr(tr["wT"])))
        for l, nu in enumerate(parent.sweeps):
            _lib.check(lib.sktb_mg_set_level_sweeps(h, l, nu))
        self.setup_count = -1

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_mg_destroy(h)
            except Exception:
                pass
            self.handle = None

    def refresh(self):
        par, lib, h = self.parent, self.lib, self.handle
        if self.setup_count == par.setup_count:
            return
        eng = par.eng
        _lib.check(lib.sktb_mg_set_level0_range(h, int(eng.node0), int(eng.dm.n_nodes)))
        _lib.check(lib.sktb_mg_set_level(
            h, 0, int(eng.node1 - eng.node0), int(eng.node_col_loc.numel()), int(eng.max_deg),
            dev._ptr(eng.node_ptr_loc), dev._ptr(eng.node_col_loc), dev._ptr(eng.vals),
            dev._ptr(eng.inv_diag), dev._ptr(eng.dir_mask)))
        par._push_vals32_level0(handle=h, convert=False)
        for l in range(1, par.n_levels):
            lv = par.levels[l]
            _lib.check(lib.sktb_mg_set_level(
                h, l, lv["node1"] - lv["node0"], int(lv["node_col"].numel()), lv["max_deg"],
                dev._ptr(lv["node_ptr"]), dev._ptr(lv["node_col"]), dev._ptr(lv["vals"]),
                dev._ptr(lv["inv_diag"]), dev._ptr(lv["mask"])))
            par._push_vals32(l, handle=h, convert=False)
        _lib.check(lib.sktb_mg_share_coarsest(h, par.handle))
        if par.lambda_max is not None:
            for l, lam in enumerate(par.lambda_max):
                _lib.check(lib.sktb_mg_set_level_omega(h, l, float(1.75 / (1.03 * lam))))
        self.setup_count = par.setup_count


class ScalarMultigrid:
    """Geometric multigrid for a scalar operator on a tensor hexahedral grid
    (``csrc/mg_scalar.cu``): stands in for the reference's sparse LU of the heat
    system (``fea/solver_heat.py:191-192``) as the preconditioner of the device
    PCG.  The level-0 operator is whatever enforced CSR matrix the caller
    assembled (conduction + real and virtual Robin terms); the coarse operators
    are its algebraic Galerkin products, rebuilt by ``setup`` whenever it changes."""

    MIN_FINE_NODES = 3000
    DENSE_MAX_NODES = 160

    def __init__(self, engine, axes):
        self.lib = _lib.load()
        self.eng = engine
        coords = [tuple(axes)]
        while True:
            cells = [c.size - 1 for c in coords[-1]]
            n = int(np.prod([c.size for c in coords[-1]]))
            if n <= self.DENSE_MAX_NODES or max(cells) <= 1:
                break
            coords.append(tuple(a[coarse_index_map(a.size - 1)] for a in coords[-1]))
        if len(coords) < 2 or int(np.prod([c.size for c in coords[-1]])) > self.DENSE_MAX_NODES:
            raise ValueError("grid not suited to the scalar multigrid hierarchy")
        self.coords = coords
        self.n_levels = len(coords)
        np_h = np.ascontiguousarray([[c.size for c in lv] for lv in coords], dtype=np.int32)
        h = C.c_void_p()
        _lib.check(self.lib.sktb_smg_create(C.byref(h), self.n_levels,
                                            np_h.ctypes.data_as(C.c_void_p),
                                            torch.cuda.current_device()))
        self.handle = h
        self.np_h = np_h
        self._keep = []
        mask = engine.dir_mask.cpu().numpy().astype(np.uint8)
        for l in range(self.n_levels):
            m_d = dev.to_dev(mask, dev.U8)
            self._keep.append(m_d)
            _lib.check(self.lib.sktb_smg_set_mask(h, l, dev._ptr(m_d)))
            if l + 1 == self.n_levels:
                break
            fine_cells = [c.size - 1 for c in coords[l]]
            coarse_np = [c.size for c in coords[l + 1]]
            tabs = [axis_tables(n) for n in fine_cells]
            cat = lambda k, dt: np.ascontiguousarray(np.concatenate([t[k] for t in tabs]).astype(dt))
            catT = lambda k, dt: np.ascontiguousarray(
                np.concatenate([t[k] for t in tabs], axis=1).astype(dt))
            tr = [dev.to_dev(cat(0, np.int32), dev.I32), dev.to_dev(cat(1, np.int32), dev.I32),
                  dev.to_dev(cat(2, np.float64)), dev.to_dev(cat(3, np.float64)),
                  dev.to_dev(catT(4, np.int32).ravel(), dev.I32),
                  dev.to_dev(catT(5, np.float64).ravel())]
            self._keep.append(tr)
            _lib.check(self.lib.sktb_smg_set_transfer(h, l, *[dev._ptr(t) for t in tr]))
            # coarse node fixed iff the coincident fine node is
            fm = [coarse_index_map(n) for n in fine_cells]
            npx_f, npy_f = fine_cells[0] + 1, fine_cells[1] + 1
            Iz, Ix, Iy = np.meshgrid(np.arange(coarse_np[2]), np.arange(coarse_np[0]),
                                     np.arange(coarse_np[1]), indexing="ij")
            fnode = (fm[1][Iy] + npy_f * fm[0][Ix] + npy_f * npx_f * fm[2][Iz]).ravel()
            mask = mask[fnode].copy()
        sweeps = os.environ.get("SKTOPT_B200_SMG_SWEEPS")
        if sweeps:
            sw = [int(v) for v in sweeps.split(",")]
            for l in range(self.n_levels):
                _lib.check(self.lib.sktb_smg_set_level_sweeps(h, l, sw[min(l, len(sw) - 1)]))
        self.setup_count = 0

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and h.value:
            try:
                self.lib.sktb_smg_destroy(h)
            except Exception:
                pass
            self.handle = None

    def setup(self, row_ptr, col_idx, vals):
        _lib.check(self.lib.sktb_smg_setup_csr(self.handle, dev._ptr(row_ptr), dev._ptr(col_idx),
                                               dev._ptr(vals), dev._stream()))
        self.setup_count += 1

    def vcycle(self, r, z=None):
        if z is None:
            z = torch.empty_like(r)
        _lib.check(self.lib.sktb_smg_vcycle(self.handle, dev._ptr(r), dev._ptr(z), dev._stream()))
        return z

    def apply(self, level, x):
        n = int(np.prod(self.np_h[level]))
        y = torch.empty(n, dtype=dev.F64, device="cuda")
        _lib.check(self.lib.sktb_smg_apply(self.handle, int(level), dev._ptr(x), dev._ptr(y),
                                           dev._stream()))
        return y

    def level_values(self, level):
        n = int(np.prod(self.np_h[level]))
        out = torch.empty((27, n), dtype=dev.F64, device="cuda")
        _lib.check(self.lib.sktb_smg_level_values(self.handle, int(level), dev._ptr(out),
                                                  dev._stream()))
        return out
