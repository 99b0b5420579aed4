// Matrix-free operators for trilinear hexahedra on a tensor grid with ONE
// geometry class (uniform spacing: create_box_hex).
//
//   A = sum_e s_e Ae0     =>     (A x)_n = sum_{e ∋ n} s_e Ae0[a(n,e), :] x_e
//
// DPN = 3: elasticity, A = K(rho), s_e = E(rho_e), Ae0 the 24x24 unit stiffness.
// DPN = 1: scalar operators (Helmholtz filter M + r^2 K, conduction), Ae0 8x8,
//          s_e optional (1 where the element exists).
//
// The assembled elasticity operator streams 8.44 bytes per non-zero (2.1 GB per
// product at 1M elements); here a product reads x (8 DPN B/node), s (8 B per
// element) and writes y, so it is bound by the FP64 pipe instead of HBM.
//
// Tiled kernel: a CTA owns a TY x TX x TZ brick of nodes (one thread per node,
// tile shape chosen on the host to fit the grid), stages the brick's halo of x
// (Dirichlet dofs and out-of-grid nodes as 0) and of s (0 for elements outside
// the grid) in shared memory, and every thread then walks its 27 neighbours
// plane by plane: each neighbour value is read once (LDS) and feeds the <= 8
// elements that contain both nodes.  The Ae0 coefficients are kernel parameters
// (constant bank -> uniform registers), so the inner loop is DFMA + LDCU only;
// boundaries and Dirichlet conditions cost nothing in the inner loop.  Gather
// formulation: deterministic, no atomics.  Node / element numbering are
// MeshHex.init_tensor's: node = iy + npy (ix + npx iz), element = ey + ny (ex +
// nx ez); coefficient rows/cols are ordered by corner code cx + 2 cy + 4 cz.
//
// Dirichlet dofs reproduce what csr_enforce builds (identity rows/columns):
// fixed inputs are read as 0, fixed outputs pass x through.  dmask[n] bit i =
// dof i of node n fixed (bit 3 = a fixed dof somewhere in the 27-neighbourhood,
// used by the untiled kernel only).
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

template <int DPN>
struct GridParams {
  double ke[64 * DPN * DPN];
  float kef[64 * DPN * DPN];  // same, rounded: single-precision products of the V-cycle
  int32_t npx, npy, npz;
  int32_t ty, tx, tz;     // tile shape (nodes)
  int32_t nty, ntx, ntz;  // tiles per axis
  const double *scale;    // per element, may be null (DPN = 1: 1.0)
  const uint8_t *dmask;   // per node
};

// DPN = 1, uniform coefficient: the element sums collapse to a 27-point stencil.
// Kept split by the y-position of the contributing elements so that the y-ends
// of a grid line (which sit in the middle of a warp) need no branch:
//   w0[(dz+1)*3+(dx+1)][k]: elements BELOW the node in y (oy = 0), neighbour dy = k-1
//   w1[(dz+1)*3+(dx+1)][k]: elements ABOVE (oy = 1), neighbour dy = k
struct ScalarStencil {
  double w0[9][2];
  double w1[9][2];
};
struct sktb_gridop {
  int dpn = 3;
  GridParams<3> P3;
  GridParams<1> P1;
  ScalarStencil W1;
  int device = 0;
  int64_t n_nodes = 0;
  bool fields_set = false;
  bool direct = false;  // untiled kernel (DPN = 3 only; default there)
  bool split = false;   // two warps per node (SKTB_GRIDOP_SPLIT=1)
  bool shfl = true;     // y-neighbours by warp shuffle (SKTB_GRIDOP_SHFL=0: all from L1)
  bool scalar_direct = true;  // DPN = 1: untiled stencil kernel (SKTB_GRIDOP_SCALAR_TILED=1: tiled)
  size_t smem = 0;
  // two-nodes-per-thread kernel (DPN = 3): coefficient tables in global memory,
  // staged in shared memory by every CTA (SKTB_GRIDOP_X2=0 disables)
  bool x2 = true;
  // x2 with the element modulus folded into the inputs (6 accumulators per node
  // pair, 16 warps per SM): 1 = single-precision products only (default: the
  // V-cycle's level-0 products, 0.607 -> 0.590 ms per cycle at C2), 2 = fp64 too
  // (SKTB_GRIDOP_X2S=1; measured SLOWER there: 0.089-0.095 vs 0.081 ms, the extra
  // 28 % of FP64 instructions outweigh the doubled occupancy), 0 = never (=0)
  int x2s = 1;
  double *kt_d = nullptr;
  float *ktf_d = nullptr;
};

static ScalarStencil make_scalar_stencil(const double *ke) {
  ScalarStencil W;
  for (int i = 0; i < 9; ++i) W.w0[i][0] = W.w0[i][1] = W.w1[i][0] = W.w1[i][1] = 0.0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int o = 0; o < 8; ++o) {
          const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
          const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
          if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
          const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
          const int cb = bx + 2 * by + 4 * bz;
          const int line = (dz + 1) * 3 + (dx + 1);
          if (oy == 0)
            W.w0[line][by] += ke[ca * 8 + cb];  // by = dy + 1: dy = -1 -> 0, dy = 0 -> 1
          else
            W.w1[line][by] += ke[ca * 8 + cb];  // by = dy: dy = 0 -> 0, dy = +1 -> 1
        }
  return W;
}

// node -> (ix, iy, iz) with 32-bit unsigned arithmetic (node counts are < 2^31:
// connectivity is int32); the 64-bit div/mod the compiler emits otherwise costs
// ~200 instructions per node, more than the whole scalar stencil
__device__ __forceinline__ void node_coords(int64_t n, int npx, int npy, int &ix, int &iy,
                                            int &iz) {
  const unsigned n32 = (unsigned)n;
  const unsigned t = n32 / (unsigned)npy;
  iy = (int)(n32 - t * (unsigned)npy);
  const unsigned z = t / (unsigned)npx;
  ix = (int)(t - z * (unsigned)npx);
  iz = (int)z;
}

__device__ __forceinline__ int clampi(int v, int hi) {
  return v < 0 ? 0 : (v > hi ? hi : v);
}

constexpr int kMaxHalo = 1100;  // (TY+2)(TX+2)(TZ+2) bound enforced by the host
constexpr int kMaxElemTile = 640;

// ------------------------------------------------------------- tiled kernel --
template <int DPN, bool DOT>
__global__ void __launch_bounds__(kBlock, 2)
    grid_apply_tiled_kernel(const __grid_constant__ GridParams<DPN> P, int64_t node0,
                            int64_t n_loc, int tz_lo, int tz_cnt,
                            const double *__restrict__ x, double *__restrict__ y,
                            const double *__restrict__ dotv, double *partials,
                            unsigned int *ticket, double *dot_out,
                            const PcgScalars *S) {
  if (S && S->rr <= S->tol2) return;
  extern __shared__ double sm[];
  const int TY = P.ty, TX = P.tx, TZ = P.tz;
  const int HY = TY + 2, HX = TX + 2, HZ = TZ + 2;
  const int EY = TY + 1, EX = TX + 1, EZ = TZ + 1;
  const int nH = DPN * HY * HX * HZ, nE = EY * EX * EZ;
  double *su = sm;
  double *sE = sm + nH;
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int t = threadIdx.x;
  const bool active = t < TY * TX * TZ;
  const int lt = active ? t : 0;
  const int lty = lt % TY, ltx = (lt / TY) % TX, ltz = lt / (TY * TX);

  // staging descriptors: the tile shape is fixed, so which halo entry a thread
  // copies in pass k never changes -- decode it once
  constexpr int KU = (DPN * kMaxHalo + kBlock - 1) / kBlock;
  constexpr int KE = (kMaxElemTile + kBlock - 1) / kBlock;
  unsigned du[KU], de[KE];
#pragma unroll
  for (int k = 0; k < KU; ++k) {
    const int idx = t + kBlock * k;
    if (idx < nH) {
      const int line = idx / (DPN * HY), q = idx - line * (DPN * HY);
      const int hy = q / DPN, c = q - hy * DPN;
      const int hz = line / HX, hx = line - hz * HX;
      du[k] = (unsigned)hy | ((unsigned)hx << 8) | ((unsigned)hz << 16) | ((unsigned)c << 24);
    } else {
      du[k] = 0xffffffffu;
    }
  }
#pragma unroll
  for (int k = 0; k < KE; ++k) {
    const int idx = t + kBlock * k;
    if (idx < nE) {
      const int ez = idx / (EY * EX), r = idx - ez * (EY * EX);
      const int ex = r / EY, ey = r - ex * EY;
      de[k] = (unsigned)ey | ((unsigned)ex << 8) | ((unsigned)ez << 16);
    } else {
      de[k] = 0xffffffffu;
    }
  }
  // smem addresses of this thread's stencil: 9 (dx, dz) lines + the element corner
  const double *ul[9];
#pragma unroll
  for (int dz = 0; dz < 3; ++dz)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
      ul[dz * 3 + dx] = su + DPN * (((ltz + dz) * HX + ltx + dx) * HY + lty);
  const double *el = sE + ((ltz * EX + ltx) * EY + lty);

  double dot = 0.0;
  const int n_tiles = P.nty * P.ntx * tz_cnt;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tyi = tile % P.nty;
    const int rest = tile / P.nty;
    const int txi = rest % P.ntx;
    const int tzi = tz_lo + rest / P.ntx;
    const int y0 = tyi * TY, x0 = txi * TX, z0 = tzi * TZ;
    // ---- stage the halo of x (masked) and of the element scale; all loads of
    // the tile are issued before the first use (clamped addresses, select after)
    {
      double v[KU];
      unsigned mb[KU];
      bool ok[KU];
#pragma unroll
      for (int k = 0; k < KU; ++k) {
        const unsigned d = du[k];
        const int gy = y0 - 1 + (int)(d & 255u), gx = x0 - 1 + (int)((d >> 8) & 255u),
                  gz = z0 - 1 + (int)((d >> 16) & 255u);
        const unsigned c = (d >> 24) & 3u;
        ok[k] = d != 0xffffffffu && gy >= 0 && gy < npy && gx >= 0 && gx < npx && gz >= 0 &&
                gz < npz;
        const int64_t m = ok[k] ? gy + (int64_t)npy * (gx + (int64_t)npx * gz) : 0;
        mb[k] = (P.dmask[m] >> c) & 1u;
        v[k] = __ldg(&x[DPN * m + (ok[k] ? c : 0u)]);
      }
#pragma unroll
      for (int k = 0; k < KU; ++k)
        if (du[k] != 0xffffffffu) su[t + kBlock * k] = (ok[k] && !mb[k]) ? v[k] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < KE; ++k) {
      const unsigned d = de[k];
      if (d != 0xffffffffu) {
        const int ey = y0 - 1 + (int)(d & 255u), ex = x0 - 1 + (int)((d >> 8) & 255u),
                  ez = z0 - 1 + (int)((d >> 16) & 255u);
        const bool in = ey >= 0 && ey < ny && ex >= 0 && ex < nx && ez >= 0 && ez < nz;
        double v = in ? 1.0 : 0.0;
        if (P.scale) v = in ? __ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : 0.0;
        sE[t + kBlock * k] = v;
      }
    }
    __syncthreads();
    // ---- this thread's node
    double pe[8][DPN];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int i = 0; i < DPN; ++i) pe[o][i] = 0.0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const double *line = ul[(dz + 1) * 3 + dx + 1];
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          double u[DPN];
#pragma unroll
          for (int j = 0; j < DPN; ++j) u[j] = line[DPN * (dy + 1) + j];
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
            const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
            if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
            const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
            const int cb = bx + 2 * by + 4 * bz;
#pragma unroll
            for (int i = 0; i < DPN; ++i)
#pragma unroll
              for (int j = 0; j < DPN; ++j)
                pe[o][i] = fma(P.ke[(DPN * ca + i) * (8 * DPN) + DPN * cb + j], u[j], pe[o][i]);
          }
        }
      }
    }
    double out[DPN];
#pragma unroll
    for (int i = 0; i < DPN; ++i) out[i] = 0.0;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const double E = el[(((o >> 2) * EX + (o & 1)) * EY) + ((o >> 1) & 1)];
#pragma unroll
      for (int i = 0; i < DPN; ++i) out[i] = fma(E, pe[o][i], out[i]);
    }
    const int gy = y0 + lty, gx = x0 + ltx, gz = z0 + ltz;
    if (active && gy < npy && gx < npx && gz < npz) {
      const int64_t n = gy + (int64_t)npy * (gx + (int64_t)npx * gz);
      const int64_t r = n - node0;
      if (r >= 0 && r < n_loc) {
        const unsigned dm = P.dmask[n];
#pragma unroll
        for (int i = 0; i < DPN; ++i) {
          if ((dm >> i) & 1u) out[i] = x[DPN * n + i];
          y[DPN * r + i] = out[i];
          if (DOT) dot = fma(out[i], dotv[DPN * r + i], dot);
        }
      }
    }
    __syncthreads();
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// --------------------------------------- untiled kernel (DPN = 3, default) --
// One thread per node straight from global memory / L1.  FAST: the node is
// interior in x and z and no node of its neighbourhood carries a Dirichlet dof
// (a y-boundary neighbour wraps into the adjacent grid line, harmless: its
// elements have E = 0).
#ifdef SKTB_KE_SMEM
#define KE_SRC(k) ske[k]
#define KE_ARG , const double *__restrict__ ske
#define KE_PASS , ske
#else
#define KE_SRC(k) P.ke[k]
#define KE_ARG
#define KE_PASS
#endif

template <bool FAST>
__device__ __forceinline__ void hexgrid_node_rows(const GridParams<3> &P, int64_t n,
                                                  int ix, int iy, int iz, unsigned dm,
                                                  const double *__restrict__ x,
                                                  double (&out)[3] KE_ARG) {
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  double E[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    const int ex = ix - 1 + (o & 1), ey = iy - 1 + ((o >> 1) & 1), ez = iz - 1 + (o >> 2);
    const bool ok = FAST ? (ey >= 0 && ey < ny)
                         : (ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz);
    E[o] = ok ? __ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : 0.0;
  }
  double pe[8][3];
#pragma unroll
  for (int o = 0; o < 8; ++o) pe[o][0] = pe[o][1] = pe[o][2] = 0.0;
  const double *xc = x + 3 * n;
  const int64_t sx = 3 * (int64_t)npy, sz = 3 * (int64_t)npy * npx;
#pragma unroll
  for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const double *line;
      if (FAST) {
        line = xc + dx * sx + dz * sz;
      } else {
        const int kx = clampi(ix + dx, npx - 1), kz = clampi(iz + dz, npz - 1);
        line = x + 3 * ((int64_t)npy * (kx + (int64_t)npx * kz));
      }
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        double u0, u1, u2;
        if (FAST) {
          u0 = __ldg(line + 3 * dy);
          u1 = __ldg(line + 3 * dy + 1);
          u2 = __ldg(line + 3 * dy + 2);
        } else {
          const int ky = clampi(iy + dy, npy - 1);
          u0 = __ldg(line + 3 * ky);
          u1 = __ldg(line + 3 * ky + 1);
          u2 = __ldg(line + 3 * ky + 2);
          if (dm & 8u) {
            const unsigned mb = P.dmask[(line - x) / 3 + ky];
            if (mb & 1u) u0 = 0.0;
            if (mb & 2u) u1 = 0.0;
            if (mb & 4u) u2 = 0.0;
          }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
          const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
          if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
          const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
          const int cb = bx + 2 * by + 4 * bz;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = (3 * ca + i) * 24 + 3 * cb;
            pe[o][i] = fma(KE_SRC(k), u0, pe[o][i]);
            pe[o][i] = fma(KE_SRC(k + 1), u1, pe[o][i]);
            pe[o][i] = fma(KE_SRC(k + 2), u2, pe[o][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double a = 0.0;
#pragma unroll
    for (int o = 0; o < 8; ++o) a = fma(E[o], pe[o][i], a);
    out[i] = a;
  }
  if (!FAST && (dm & 7u)) {
    if (dm & 1u) out[0] = xc[0];
    if (dm & 2u) out[1] = xc[1];
    if (dm & 4u) out[2] = xc[2];
  }
}

template <bool DOT>
__global__ void __launch_bounds__(kBlock, 2)
    hexgrid_apply_kernel(const __grid_constant__ GridParams<3> P, int64_t node0,
                         int64_t n_loc, const double *__restrict__ x,
                         double *__restrict__ y, const double *__restrict__ dotv,
                         double *partials, unsigned int *ticket, double *dot_out,
                         const PcgScalars *S) {
  if (S && S->rr <= S->tol2) return;
#ifdef SKTB_KE_SMEM
  __shared__ double ske[576];
  for (int i = threadIdx.x; i < 576; i += blockDim.x) ske[i] = P.ke[i];
  __syncthreads();
#endif
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  double dot = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_loc;
       r += stride) {
    const int64_t n = node0 + r;
    int ix, iy, iz;
    node_coords(n, npx, npy, ix, iy, iz);
    const unsigned dm = P.dmask[n];
    double out[3];
    if (ix > 0 && ix < npx - 1 && iz > 0 && iz < npz - 1 && !(dm & 8u))
      hexgrid_node_rows<true>(P, n, ix, iy, iz, dm, x, out KE_PASS);
    else
      hexgrid_node_rows<false>(P, n, ix, iy, iz, dm, x, out KE_PASS);
    y[3 * r] = out[0];
    y[3 * r + 1] = out[1];
    y[3 * r + 2] = out[2];
    if (DOT)
      dot += out[0] * dotv[3 * r] + out[1] * dotv[3 * r + 1] + out[2] * dotv[3 * r + 2];
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// ------------------- shuffle kernel (DPN = 3): y-neighbours from the warp --
// Same per-node work as the untiled kernel, but of the 27 neighbours only the 9
// with dy = 0 are loaded; dy = -1 / +1 come from the adjacent lanes by shuffle
// (consecutive lanes own consecutive nodes of a grid line), the two edge lanes
// fetch theirs.  The interleaved (x, y, z) layout makes every warp load touch
// 6-7 cache lines, so this cuts the L1 look-ups per node row from ~530 to ~230.
// A lane whose neighbour lane sits on another grid line (iy = 0 or npy-1) gets
// a meaningless value there; it only feeds elements with E = 0.
//
// T = float: the products are formed in single precision (vectors stay fp64 in
// memory); used for the two level-0 products inside the multigrid V-cycle, where
// the result only steers a preconditioner.  MODE 1 fuses the damped-Jacobi
// update: y = x + omega * dinv * (b - A x).
template <typename T>
__device__ __forceinline__ T grid_coef(const GridParams<3> &P, int k);
template <>
__device__ __forceinline__ double grid_coef<double>(const GridParams<3> &P, int k) {
  return P.ke[k];
}
template <>
__device__ __forceinline__ float grid_coef<float>(const GridParams<3> &P, int k) {
  return P.kef[k];
}

template <typename T, int MODE, bool DOT>
__global__ void __launch_bounds__(kBlock, 2)
    hexgrid_apply_shfl_kernel(const __grid_constant__ GridParams<3> P, int64_t node0,
                              int64_t n_loc, const double *__restrict__ x,
                              double *__restrict__ y, const double *__restrict__ dotv,
                              double *partials, unsigned int *ticket, double *dot_out,
                              const PcgScalars *S, const double *__restrict__ sm_b,
                              const double *__restrict__ sm_dinv, double sm_omega) {
  if (S && S->rr <= S->tol2) return;
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int lane = threadIdx.x & 31;
  double dot = 0.0;
  const int64_t n_pad = (n_loc + 31) / 32 * 32;  // whole warps run every trip (shuffles)
  // blocked assignment: a CTA walks a CONTIGUOUS run of grid lines, so the
  // dx = +-1 lines of one trip are the dx = 0 lines of the next and come from L1
  const int64_t trips = (n_pad + kBlock - 1) / kBlock;
  const int64_t per_cta = (trips + gridDim.x - 1) / gridDim.x;
  const int64_t r_end = min(n_pad, (int64_t)(blockIdx.x + 1) * per_cta * kBlock);
  for (int64_t r = (int64_t)blockIdx.x * per_cta * kBlock + threadIdx.x; r < r_end;
       r += kBlock) {
    const bool live = r < n_loc;
    // a dead tail lane still serves its left neighbour's dy = +1 shuffle: it
    // takes the node that follows in the grid (not owned by this rank); past the
    // end of the grid it mirrors the last node, whose dy = +1 feeds nothing
    const int64_t n = node0 + r;
    const int64_t n_total = (int64_t)npx * npy * npz;
    const int64_t nc = n < n_total ? n : n_total - 1;
    int ix, iy, iz;
    node_coords(nc, npx, npy, ix, iy, iz);
    const unsigned dm = P.dmask[nc];
    T E[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int ex = ix - 1 + (o & 1), ey = iy - 1 + ((o >> 1) & 1), ez = iz - 1 + (o >> 2);
      const bool ok = ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz;
      E[o] = ok ? (T)__ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : (T)0;
    }
    T pe[8][3];
#pragma unroll
    for (int o = 0; o < 8; ++o) pe[o][0] = pe[o][1] = pe[o][2] = (T)0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int kx = clampi(ix + dx, npx - 1), kz = clampi(iz + dz, npz - 1);
        const int64_t mc = (int64_t)npy * (kx + (int64_t)npx * kz) + iy;
        const double *cp = x + 3 * mc;
        T u[3][3];  // [dy + 1][component]
#pragma unroll
        for (int j = 0; j < 3; ++j) u[1][j] = (T)__ldg(cp + j);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          u[0][j] = __shfl_up_sync(0xffffffffu, u[1][j], 1);
          u[2][j] = __shfl_down_sync(0xffffffffu, u[1][j], 1);
        }
        if (lane == 0 && iy > 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) u[0][j] = (T)__ldg(cp - 3 + j);
        }
        if (lane == 31 && iy < npy - 1) {
#pragma unroll
          for (int j = 0; j < 3; ++j) u[2][j] = (T)__ldg(cp + 3 + j);
        }
        if (dm & 8u) {  // a fixed dof somewhere around: mask the inputs
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int ky = iy + dy;
            if (ky < 0 || ky >= npy) continue;
            const unsigned mb = P.dmask[mc + dy];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if ((mb >> j) & 1u) u[dy + 1][j] = (T)0;
          }
        }
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
            const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
            if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
            const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
            const int cb = bx + 2 * by + 4 * bz;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              const int k = (3 * ca + i) * 24 + 3 * cb;
              pe[o][i] = fma(grid_coef<T>(P, k), u[dy + 1][0], pe[o][i]);
              pe[o][i] = fma(grid_coef<T>(P, k + 1), u[dy + 1][1], pe[o][i]);
              pe[o][i] = fma(grid_coef<T>(P, k + 2), u[dy + 1][2], pe[o][i]);
            }
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        T acc = (T)0;
#pragma unroll
        for (int o = 0; o < 8; ++o) acc = fma(E[o], pe[o][i], acc);
        double a = (double)acc;
        if ((dm >> i) & 1u) a = x[3 * n + i];
        if (MODE == 1)
          a = x[3 * n + i] + sm_omega * sm_dinv[3 * r + i] * (sm_b[3 * r + i] - a);
        y[3 * r + i] = a;
        if (DOT) dot = fma(a, dotv[3 * r + i], dot);
      }
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// ------------- x2 kernel (DPN = 3): two nodes per thread, coefficients in smem --
// The shuffle kernel above takes every stiffness coefficient as a constant-bank
// operand (LDCU -> uniform register): 390 LDCU per node row and DFMAs that B200
// issues at half rate with a uniform operand (sktb_fp64_probe_const), 0.33 of the
// FP64 peak.  Here
//  * the 64 valid (dz, dx, dy, element) 3x3 coefficient blocks sit in shared
//    memory, padded to 16-byte groups (double: 10 per block, float: 12), and
//    are read with LDS.128 broadcasts into REGISTERS: full-rate DFMA;
//  * a thread owns the node pair (2 ixp, iy, iz), (2 ixp + 1, iy, iz): the same
//    coefficient block serves both nodes (node A with neighbour column dx, node
//    B with column dx + 1), so every coefficient load feeds two FMAs, and the 4
//    neighbour columns of a plane are loaded once for the pair (12 column loads
//    per pair instead of 18);
//  * y-neighbours still come from the adjacent lanes by shuffle.
// Work items are (iy, ixp, iz) over WHOLE z-planes (single GPU: all of them; a
// slab-sharded rank: its planes).  Same boundary treatment as the shuffle
// kernel: indices are clamped, elements outside the grid have modulus 0.
struct GridX2 {
  int32_t npx, npy, npz;
  const double *scale;
  const uint8_t *dmask;
  const double *kt;
  const float *ktf;
};
template <typename T> struct X2Tab;
template <> struct X2Tab<double> {
  static constexpr int S = 10;
  __device__ static __forceinline__ void load9(const double *p, double (&c)[9]) {
    const double2 a = *reinterpret_cast<const double2 *>(p);
    const double2 b = *reinterpret_cast<const double2 *>(p + 2);
    const double2 d = *reinterpret_cast<const double2 *>(p + 4);
    const double2 e = *reinterpret_cast<const double2 *>(p + 6);
    c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y; c[4] = d.x; c[5] = d.y;
    c[6] = e.x; c[7] = e.y; c[8] = p[8];
  }
};
template <> struct X2Tab<float> {
  static constexpr int S = 12;
  __device__ static __forceinline__ void load9(const float *p, float (&c)[9]) {
    const float4 a = *reinterpret_cast<const float4 *>(p);
    const float4 b = *reinterpret_cast<const float4 *>(p + 4);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y;
    c[6] = b.z; c[7] = b.w; c[8] = p[8];
  }
};
constexpr int kX2Slots = 216;  // ((dz+1) 3 + (dx+1)) 3 + (dy+1)) 8 + o

#ifndef SKTB_X2_BLOCK
#define SKTB_X2_BLOCK 256
#define SKTB_X2_MINB 1
#endif
constexpr int kX2Block = SKTB_X2_BLOCK;
static_assert(kX2Block == kBlock, "grid_reduce folds kBlock / 32 warp partials per CTA");
template <typename T, int MODE, bool DOT>
__global__ void __launch_bounds__(kX2Block, SKTB_X2_MINB)
    hexgrid_apply_x2_kernel(const GridX2 P, int z0, int nzs, int64_t node0,
                            const double *__restrict__ x, double *__restrict__ y,
                            const double *__restrict__ dotv, double *partials,
                            unsigned int *ticket, double *dot_out, const PcgScalars *S,
                            const double *__restrict__ sm_b,
                            const double *__restrict__ sm_dinv, double sm_omega) {
  if (S && S->rr <= S->tol2) return;
  constexpr int TS = X2Tab<T>::S;
  __shared__ __align__(16) T kt[kX2Slots * TS];
  {
    const T *src = sizeof(T) == 8 ? (const T *)P.kt : (const T *)P.ktf;
    for (int i = threadIdx.x; i < kX2Slots * TS; i += blockDim.x) kt[i] = src[i];
  }
  __syncthreads();
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int npxp = (npx + 1) >> 1;
#ifdef SKTB_X2_SHFL
  const int lane = threadIdx.x & 31;
#endif
  const unsigned n_items = (unsigned)npy * (unsigned)npxp * (unsigned)nzs;
  const unsigned n_pad = (n_items + 31u) & ~31u;
  const unsigned trips = (n_pad + kX2Block - 1) / kX2Block;
  const unsigned per_cta = (trips + gridDim.x - 1) / gridDim.x;
  const unsigned w_end = min(n_pad, (blockIdx.x + 1) * per_cta * kX2Block);
  double dot = 0.0;
  for (unsigned w = blockIdx.x * per_cta * kX2Block + threadIdx.x; w < w_end; w += kX2Block) {
    const bool live = w < n_items;
    const unsigned wc = live ? w : n_items - 1;
    const unsigned t1 = wc / (unsigned)npy;
    const int iy = (int)(wc - t1 * (unsigned)npy);
    const unsigned t2 = t1 / (unsigned)npxp;
    const int ixA = 2 * (int)(t1 - t2 * (unsigned)npxp);
    const int iz = z0 + (int)t2;
    const bool hasB = ixA + 1 < npx;
    const int nA = iy + npy * (ixA + npx * iz);
    const int nB = hasB ? nA + npy : nA;
    const unsigned dmA = P.dmask[nA], dmB = P.dmask[nB];
    const bool near_fixed = ((dmA | dmB) & 8u) != 0;
    // moduli of the 3 x 2 x 2 elements around the pair (0 outside the grid)
    T Eg[3][2][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ex = ixA - 1 + a, ey = iy - 1 + b, ez = iz - 1 + c;
          const bool ok = ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz;
          Eg[a][b][c] = ok ? (T)__ldg(&P.scale[ey + ny * (ex + nx * ez)]) : (T)0;
        }
    T peA[8][3], peB[8][3];
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int i = 0; i < 3; ++i) peA[o][i] = peB[o][i] = (T)0;
    // 32-bit index arithmetic (3 n_nodes < 2^31 is checked at creation)
    int cxo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cxo[c] = npy * clampi(ixA - 1 + c, npx - 1) + iy;
    const int lo = iy > 0 ? -3 : 0, hi = iy < npy - 1 ? 3 : 0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
      const int zo = npy * npx * clampi(iz + dz, npz - 1);
      T col[4][3][3];  // [column ixA-1 .. ixA+2][dy + 1][component]
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int mc = zo + cxo[c];
        const double *cp = x + 3 * mc;
#ifdef SKTB_X2_SHFL
#pragma unroll
        for (int j = 0; j < 3; ++j) col[c][1][j] = (T)__ldg(cp + j);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          col[c][0][j] = __shfl_up_sync(0xffffffffu, col[c][1][j], 1);
          col[c][2][j] = __shfl_down_sync(0xffffffffu, col[c][1][j], 1);
        }
        if (lane == 0 && iy > 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) col[c][0][j] = (T)__ldg(cp - 3 + j);
        }
        if (lane == 31 && iy < npy - 1) {
#pragma unroll
          for (int j = 0; j < 3; ++j) col[c][2][j] = (T)__ldg(cp + 3 + j);
        }
#else
        // the nine values (iy-1 .. iy+1) x 3 components are contiguous in memory;
        // the line ends are clamped (their values only feed elements of modulus 0)
        {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            col[c][0][j] = (T)__ldg(cp + lo + j);
            col[c][1][j] = (T)__ldg(cp + j);
            col[c][2][j] = (T)__ldg(cp + hi + j);
          }
        }
#endif
        if (near_fixed) {  // a fixed dof somewhere around: mask the inputs
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int ky = iy + dy;
            if (ky < 0 || ky >= npy) continue;
            const unsigned mb = P.dmask[mc + dy];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if ((mb >> j) & 1u) col[c][dy + 1][j] = (T)0;
          }
        }
      }
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
            const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
            if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
            T c9[9];
            X2Tab<T>::load9(&kt[((((dz + 1) * 3 + (dx + 1)) * 3 + (dy + 1)) * 8 + o) * TS], c9);
            // j outermost: six independent accumulators between two FMAs of a chain
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
              for (int i = 0; i < 3; ++i) peA[o][i] = fma(c9[3 * i + j], col[dx + 1][dy + 1][j], peA[o][i]);
#pragma unroll
              for (int i = 0; i < 3; ++i) peB[o][i] = fma(c9[3 * i + j], col[dx + 2][dy + 1][j], peB[o][i]);
            }
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        if (nb == 1 && !hasB) break;
        const int n = nb ? nB : nA;
        const unsigned dm = nb ? dmB : dmA;
        const int r = n - (int)node0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          T acc = (T)0;
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const T E = Eg[(o & 1) + nb][(o >> 1) & 1][o >> 2];
            acc = fma(E, nb ? peB[o][i] : peA[o][i], acc);
          }
          double a = (double)acc;
          if ((dm >> i) & 1u) a = x[3 * n + i];
          if (MODE == 1)
            a = x[3 * n + i] + sm_omega * sm_dinv[3 * r + i] * (sm_b[3 * r + i] - a);
          y[3 * r + i] = a;
          if (DOT) dot = fma(a, dotv[3 * r + i], dot);
        }
      }
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// ------- x2s kernel: the x2 kernel with the element modulus folded into the INPUT --
// The x2 kernel keeps one partial product per (node, element) until the end
// (y_i = sum_o E_o pe[o][i]): 48 accumulators per node pair, 255 registers, 8
// warps per SM, and the FP64 pipe waits on latencies (0.37-0.43 of its peak).
// Here every neighbour value is scaled by the modulus of the element it is seen
// through (s = E_o x_m, 3 multiplies) and added straight into the 3 outputs of the
// node: 6 (x 2 for independent chains) accumulators per pair, only two neighbour
// columns live at a time, 128 registers, 16 warps per SM; 1536 instead of 1200
// FP64 instructions per pair (+28 %).  Same tables, same work split, same boundary
// treatment and epilogue as the x2 kernel.  Measured at C2 (B200): fp64 0.089-0.095
// ms against 0.081 ms for x2 (the FP64 pipe is ~50 % instead of 43 % busy, not
// enough to pay for the extra instructions), so fp64 products keep x2; the fp32
// products of the V-cycle gain (0.607 -> 0.590 ms per cycle) and use this kernel.
template <typename T, int MODE, bool DOT>
__global__ void __launch_bounds__(kX2Block, 2)
    hexgrid_apply_x2s_kernel(const GridX2 P, int z0, int nzs, int64_t node0,
                             const double *__restrict__ x, double *__restrict__ y,
                             const double *__restrict__ dotv, double *partials,
                             unsigned int *ticket, double *dot_out, const PcgScalars *S,
                             const double *__restrict__ sm_b,
                             const double *__restrict__ sm_dinv, double sm_omega) {
  if (S && S->rr <= S->tol2) return;
  constexpr int TS = X2Tab<T>::S;
  __shared__ __align__(16) T kt[kX2Slots * TS];
  {
    const T *src = sizeof(T) == 8 ? (const T *)P.kt : (const T *)P.ktf;
    for (int i = threadIdx.x; i < kX2Slots * TS; i += blockDim.x) kt[i] = src[i];
  }
  __syncthreads();
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int npxp = (npx + 1) >> 1;
  const unsigned n_items = (unsigned)npy * (unsigned)npxp * (unsigned)nzs;
  const unsigned n_pad = (n_items + 31u) & ~31u;
  const unsigned trips = (n_pad + kX2Block - 1) / kX2Block;
  const unsigned per_cta = (trips + gridDim.x - 1) / gridDim.x;
  const unsigned w_end = min(n_pad, (blockIdx.x + 1) * per_cta * kX2Block);
  double dot = 0.0;
  for (unsigned w = blockIdx.x * per_cta * kX2Block + threadIdx.x; w < w_end; w += kX2Block) {
    const bool live = w < n_items;
    const unsigned wc = live ? w : n_items - 1;
    const unsigned t1 = wc / (unsigned)npy;
    const int iy = (int)(wc - t1 * (unsigned)npy);
    const unsigned t2 = t1 / (unsigned)npxp;
    const int ixA = 2 * (int)(t1 - t2 * (unsigned)npxp);
    const int iz = z0 + (int)t2;
    const bool hasB = ixA + 1 < npx;
    const int nA = iy + npy * (ixA + npx * iz);
    const int nB = hasB ? nA + npy : nA;
    const unsigned dmA = P.dmask[nA], dmB = P.dmask[nB];
    const bool near_fixed = ((dmA | dmB) & 8u) != 0;
    T Eg[3][2][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ex = ixA - 1 + a, ey = iy - 1 + b, ez = iz - 1 + c;
          const bool ok = ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz;
          Eg[a][b][c] = ok ? (T)__ldg(&P.scale[ey + ny * (ex + nx * ez)]) : (T)0;
        }
    T yA[2][3], yB[2][3];   // two independent chains per output
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int i = 0; i < 3; ++i) yA[h][i] = yB[h][i] = (T)0;
    int cxo[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) cxo[c] = npy * clampi(ixA - 1 + c, npx - 1) + iy;
    const int lo = iy > 0 ? -3 : 0, hi = iy < npy - 1 ? 3 : 0;
    // the plane loop stays rolled: unrolled, ptxas hoists the loads of all three
    // planes and spills ~700 B per thread at 128 registers
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
      const int zo = npy * npx * clampi(iz + dz, npz - 1);
      T col[4][3][3];  // [column ixA-1 .. ixA+2][dy + 1][component]; two live at a time
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int mc = zo + cxo[c];
        const double *cp = x + 3 * mc;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          col[c][0][j] = (T)__ldg(cp + lo + j);
          col[c][1][j] = (T)__ldg(cp + j);
          col[c][2][j] = (T)__ldg(cp + hi + j);
        }
        if (near_fixed) {
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int ky = iy + dy;
            if (ky < 0 || ky >= npy) continue;
            const unsigned mb = P.dmask[mc + dy];
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if ((mb >> j) & 1u) col[c][dy + 1][j] = (T)0;
          }
        }
        if (c == 0) continue;
        // node A sees column c-1 as dx, node B column c: blocks (dz, dx = c - 2)
        const int dx = c - 2;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
            const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
            if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
            T c9[9];
            X2Tab<T>::load9(&kt[((((dz + 1) * 3 + (dx + 1)) * 3 + (dy + 1)) * 8 + o) * TS], c9);
            const T EA = Eg[ox][oy][oz], EB = Eg[ox + 1][oy][oz];
            T sA[3], sB[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              sA[j] = EA * col[dx + 1][dy + 1][j];
              sB[j] = EB * col[dx + 2][dy + 1][j];
            }
#ifdef SKTB_X2S_SINGLE_CHAIN
            const int h = 0;
#else
            const int h = (o ^ (o >> 1) ^ (o >> 2)) & 1;
#endif
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
              for (int i = 0; i < 3; ++i) yA[h][i] = fma(c9[3 * i + j], sA[j], yA[h][i]);
#pragma unroll
              for (int i = 0; i < 3; ++i) yB[h][i] = fma(c9[3 * i + j], sB[j], yB[h][i]);
            }
          }
        }
      }
    }
    if (live) {
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        if (nb == 1 && !hasB) break;
        const int n = nb ? nB : nA;
        const unsigned dm = nb ? dmB : dmA;
        const int r = n - (int)node0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double a = nb ? (double)(yB[0][i] + yB[1][i]) : (double)(yA[0][i] + yA[1][i]);
          if ((dm >> i) & 1u) a = x[3 * n + i];
          if (MODE == 1)
            a = x[3 * n + i] + sm_omega * sm_dinv[3 * r + i] * (sm_b[3 * r + i] - a);
          y[3 * r + i] = a;
          if (DOT) dot = fma(a, dotv[3 * r + i], dot);
        }
      }
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// whole planes only; returns -1 when the range is not eligible
template <typename T, int MODE>
static int launch_x2(const sktb_gridop *op, int64_t node0, int64_t n_nodes, const double *x,
                     double *y, const double *dotv, ReduceScratch *rs, double *dot_out,
                     const PcgScalars *S, const double *b, const double *dinv, double omega,
                     cudaStream_t st) {
  const GridParams<3> &G = op->P3;
  const int64_t plane = (int64_t)G.npx * G.npy;
  if (!op->x2 || !op->kt_d || node0 % plane || n_nodes % plane) return -1;
  GridX2 P{G.npx, G.npy, G.npz, G.scale, G.dmask, op->kt_d, op->ktf_d};
  const int z0 = (int)(node0 / plane), nzs = (int)(n_nodes / plane);
  const int64_t items = (int64_t)G.npy * ((G.npx + 1) / 2) * nzs;
  const int64_t trips = (items + kX2Block - 1) / kX2Block;
  const int64_t cap = (int64_t)kNumSM * SKTB_X2_MINB;
  const int g = (int)(trips < cap ? trips : cap);
  if (op->x2s == 2 || (op->x2s == 1 && sizeof(T) == 4)) {
    const int64_t cap2 = (int64_t)kNumSM * 2;   // 128 registers: two CTAs per SM
    const int g2 = (int)(trips < cap2 ? trips : cap2);
    if (dotv)
      hexgrid_apply_x2s_kernel<T, MODE, true><<<g2, kX2Block, 0, st>>>(
          P, z0, nzs, node0, x, y, dotv, rs->partials, rs->ticket, dot_out, S, b, dinv, omega);
    else
      hexgrid_apply_x2s_kernel<T, MODE, false><<<g2, kX2Block, 0, st>>>(
          P, z0, nzs, node0, x, y, nullptr, nullptr, nullptr, nullptr, S, b, dinv, omega);
    SKTB_KERNEL_OK();
    return 0;
  }
  if (dotv)
    hexgrid_apply_x2_kernel<T, MODE, true><<<g, kX2Block, 0, st>>>(
        P, z0, nzs, node0, x, y, dotv, rs->partials, rs->ticket, dot_out, S, b, dinv, omega);
  else
    hexgrid_apply_x2_kernel<T, MODE, false><<<g, kX2Block, 0, st>>>(
        P, z0, nzs, node0, x, y, nullptr, nullptr, nullptr, nullptr, S, b, dinv, omega);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------ split kernel (DPN = 3): two warps per node --
// The untiled kernel holds 24 accumulators + 8 moduli per thread (128
// registers, 16 warps/SM) and is latency bound.  Here a warp PAIR owns 32
// nodes: the even warp takes each node's four lower elements (oz = 0, neighbour
// planes dz = -1, 0), the odd warp the four upper ones (oz = 1, planes 0, +1):
// 12 accumulators per thread, twice the resident warps.  The odd warp hands its
// three partial sums over through shared memory (named barrier per pair), the
// even warp adds, applies the Dirichlet pass-through and stores.  Fixed
// summation order: deterministic.
template <int H, bool FAST>
__device__ __forceinline__ void hexgrid_node_half(const GridParams<3> &P, int64_t n,
                                                  int ix, int iy, int iz, unsigned dm,
                                                  const double *__restrict__ x,
                                                  double (&out)[3]) {
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  double E[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int ex = ix - 1 + (q & 1), ey = iy - 1 + (q >> 1), ez = iz - 1 + H;
    const bool ok = FAST ? (ey >= 0 && ey < ny)
                         : (ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz);
    E[q] = ok ? __ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : 0.0;
  }
  double pe[4][3];
#pragma unroll
  for (int q = 0; q < 4; ++q) pe[q][0] = pe[q][1] = pe[q][2] = 0.0;
  const double *xc = x + 3 * n;
  const int64_t sx = 3 * (int64_t)npy, sz = 3 * (int64_t)npy * npx;
#pragma unroll
  for (int dzi = 0; dzi < 2; ++dzi) {
    const int dz = H - 1 + dzi;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const double *line;
      if (FAST) {
        line = xc + dx * sx + dz * sz;
      } else {
        const int kx = clampi(ix + dx, npx - 1), kz = clampi(iz + dz, npz - 1);
        line = x + 3 * ((int64_t)npy * (kx + (int64_t)npx * kz));
      }
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        double u0, u1, u2;
        if (FAST) {
          u0 = __ldg(line + 3 * dy);
          u1 = __ldg(line + 3 * dy + 1);
          u2 = __ldg(line + 3 * dy + 2);
        } else {
          const int ky = clampi(iy + dy, npy - 1);
          u0 = __ldg(line + 3 * ky);
          u1 = __ldg(line + 3 * ky + 1);
          u2 = __ldg(line + 3 * ky + 2);
          if (dm & 8u) {
            const unsigned mb = P.dmask[(line - x) / 3 + ky];
            if (mb & 1u) u0 = 0.0;
            if (mb & 2u) u1 = 0.0;
            if (mb & 4u) u2 = 0.0;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ox = q & 1, oy = q >> 1, oz = H;
          const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
          if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
          const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
          const int cb = bx + 2 * by + 4 * bz;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const int k = (3 * ca + i) * 24 + 3 * cb;
            pe[q][i] = fma(P.ke[k], u0, pe[q][i]);
            pe[q][i] = fma(P.ke[k + 1], u1, pe[q][i]);
            pe[q][i] = fma(P.ke[k + 2], u2, pe[q][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) a = fma(E[q], pe[q][i], a);
    out[i] = a;
  }
}

template <bool DOT>
__global__ void __launch_bounds__(kBlock, 4)
    hexgrid_apply_split_kernel(const __grid_constant__ GridParams<3> P, int64_t node0,
                               int64_t n_loc, const double *__restrict__ x,
                               double *__restrict__ y, const double *__restrict__ dotv,
                               double *partials, unsigned int *ticket, double *dot_out,
                               const PcgScalars *S) {
  if (S && S->rr <= S->tol2) return;
  __shared__ double sp[kBlock / 64][3][32];
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pair = w >> 1, h = w & 1;
  double dot = 0.0;
  const int64_t stride = (int64_t)gridDim.x * (kBlock / 2);
  const int64_t n_pad = (n_loc + 31) / 32 * 32;  // both warps of a pair run the same trips
  for (int64_t r = ((int64_t)blockIdx.x * (kBlock / 64) + pair) * 32 + lane; r < n_pad;
       r += stride) {
    const bool live = r < n_loc;
    const int64_t n = node0 + (live ? r : n_loc - 1);
    int ix, iy, iz;
    node_coords(n, npx, npy, ix, iy, iz);
    const unsigned dm = P.dmask[n];
    const bool fast = ix > 0 && ix < npx - 1 && iz > 0 && iz < npz - 1 && !(dm & 8u);
    double out[3];
    if (h == 0) {
      if (fast)
        hexgrid_node_half<0, true>(P, n, ix, iy, iz, dm, x, out);
      else
        hexgrid_node_half<0, false>(P, n, ix, iy, iz, dm, x, out);
    } else {
      if (fast)
        hexgrid_node_half<1, true>(P, n, ix, iy, iz, dm, x, out);
      else
        hexgrid_node_half<1, false>(P, n, ix, iy, iz, dm, x, out);
      sp[pair][0][lane] = out[0];
      sp[pair][1][lane] = out[1];
      sp[pair][2][lane] = out[2];
    }
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
    if (h == 0 && live) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double v = out[i] + sp[pair][i][lane];
        if ((dm >> i) & 1u) v = x[3 * n + i];
        y[3 * r + i] = v;
        if (DOT) dot = fma(v, dotv[3 * r + i], dot);
      }
    }
    asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// out[DPN r + i] = 1 / A_ii (1 at fixed dofs)
template <int DPN>
__global__ void __launch_bounds__(kBlock)
    grid_inv_diag_kernel(const __grid_constant__ GridParams<DPN> P, int64_t node0,
                         int64_t n_loc, double *__restrict__ out) {
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_loc;
       r += stride) {
    const int64_t n = node0 + r;
    int ix, iy, iz;
    node_coords(n, npx, npy, ix, iy, iz);
    double d[DPN];
#pragma unroll
    for (int i = 0; i < DPN; ++i) d[i] = 0.0;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
      const int ex = ix - 1 + ox, ey = iy - 1 + oy, ez = iz - 1 + oz;
      const bool ok = ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz;
      const double E =
          ok ? (P.scale ? __ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : 1.0)
             : 0.0;
      const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
#pragma unroll
      for (int i = 0; i < DPN; ++i) d[i] += E * P.ke[(DPN * ca + i) * (8 * DPN + 1)];
    }
    const unsigned dm = P.dmask[n];
#pragma unroll
    for (int i = 0; i < DPN; ++i) out[DPN * r + i] = ((dm >> i) & 1u) ? 1.0 : 1.0 / d[i];
  }
}

// ------------------------------------- scalar stencil kernel (DPN = 1) ------
// One thread per node, consecutive threads = consecutive nodes along y, blocked
// assignment (a CTA walks a contiguous run of grid lines, so the neighbour
// lines of one trip are L1 hits of the next).  Nodes whose eight elements all
// exist in x and z, with no fixed node around and a uniform coefficient (scale
// == NULL: the Helmholtz operator M + r^2 K), take the 36-FMA stencil W with
// plain offset addressing: 27 coalesced loads, no shuffles, no lane
// predicates (a first version with shuffled y-neighbours spent ~600
// instructions per 36 FMA on lane fix-ups and convergence barriers).  Grid
// faces, the neighbourhood of fixed nodes and variable coefficients take the
// general per-element sum (64 FMA + 8 for the scale).  Fixed nodes: inputs read
// as 0, outputs pass x through (identity rows/columns).
template <bool DOT>
__global__ void __launch_bounds__(kBlock, 3)
    scalar_grid_apply_kernel(const __grid_constant__ GridParams<1> P,
                             const __grid_constant__ ScalarStencil W, int64_t node0,
                             int64_t n_loc, const double *__restrict__ x,
                             double *__restrict__ y, const double *__restrict__ dotv,
                             double *partials, unsigned int *ticket, double *dot_out,
                             const PcgScalars *S) {
  if (S && S->rr <= S->tol2) return;
  const int npx = P.npx, npy = P.npy, npz = P.npz;
  const int nx = npx - 1, ny = npy - 1, nz = npz - 1;
  const int plane = npx * npy;
  double dot = 0.0;
  const int64_t trips = (n_loc + kBlock - 1) / kBlock;
  const int64_t per_cta = (trips + gridDim.x - 1) / gridDim.x;
  const int64_t r_end = min(n_loc, (int64_t)(blockIdx.x + 1) * per_cta * kBlock);
  for (int64_t r = (int64_t)blockIdx.x * per_cta * kBlock + threadIdx.x; r < r_end;
       r += kBlock) {
    const int64_t n = node0 + r;
    int ix, iy, iz;
    node_coords(n, npx, npy, ix, iy, iz);
    const unsigned dm = P.dmask[n];
    double a;
    if (!P.scale && !(dm & 8u) && ix > 0 && ix < npx - 1 && iz > 0 && iz < npz - 1) {
      const double *xc = x + n;
      const int dl = iy > 0 ? 1 : 0, dr = iy < npy - 1 ? 1 : 0;  // stay on the line at its ends
      double a0 = 0.0, a1 = 0.0;  // elements below / above the node in y
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int l = (dz + 1) * 3 + (dx + 1);
          // at the y-ends of a line the missing neighbour is replaced by the node
          // itself; it only feeds the accumulator that is dropped below
          const double *p = xc + (dz * plane + dx * npy);
          const double u0 = __ldg(p - dl), u1 = __ldg(p), u2 = __ldg(p + dr);
          a0 = fma(W.w0[l][0], u0, a0);
          a0 = fma(W.w0[l][1], u1, a0);
          a1 = fma(W.w1[l][0], u1, a1);
          a1 = fma(W.w1[l][1], u2, a1);
        }
      }
      a = (iy > 0 ? a0 : 0.0) + (iy < npy - 1 ? a1 : 0.0);
    } else {
      double pe[8];  // per-element sums
#pragma unroll
      for (int o = 0; o < 8; ++o) pe[o] = 0.0;
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx)
#pragma unroll
          for (int dy = -1; dy <= 1; ++dy) {
            const int kx = ix + dx, ky = iy + dy, kz = iz + dz;
            if (kx < 0 || kx >= npx || ky < 0 || ky >= npy || kz < 0 || kz >= npz) continue;
            const int m = ky + npy * (kx + npx * kz);
            double u = __ldg(x + m);
            if ((dm & 8u) && (P.dmask[m] & 1u)) u = 0.0;  // fixed input
#pragma unroll
            for (int o = 0; o < 8; ++o) {
              const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
              const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
              if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
              const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
              const int cb = bx + 2 * by + 4 * bz;
              pe[o] = fma(P.ke[ca * 8 + cb], u, pe[o]);
            }
          }
      a = 0.0;
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int ex = ix - 1 + (o & 1), ey = iy - 1 + ((o >> 1) & 1), ez = iz - 1 + (o >> 2);
        const bool ok = ex >= 0 && ex < nx && ey >= 0 && ey < ny && ez >= 0 && ez < nz;
        const double E =
            ok ? (P.scale ? __ldg(&P.scale[ey + (int64_t)ny * (ex + (int64_t)nx * ez)]) : 1.0)
               : 0.0;
        a = fma(E, pe[o], a);
      }
    }
    if (dm & 1u) a = x[n];
    y[r] = a;
    if (DOT) dot = fma(a, dotv[r], dot);
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

// ------------------------------------------------------------------ launch --
template <int DPN>
static int launch_tiled(const sktb_gridop *op, const GridParams<DPN> &P, int64_t node0,
                        int64_t n_nodes, const double *x, double *y, const double *dotv,
                        ReduceScratch *rs, double *dot_out, const PcgScalars *S,
                        cudaStream_t st) {
  // tiles whose z range meets the owned nodes (all of them when not sharded)
  const int64_t plane = (int64_t)P.npx * P.npy;
  const int z_lo = (int)(node0 / plane), z_hi = (int)((node0 + n_nodes - 1) / plane);
  const int tz_lo = z_lo / P.tz, tz_cnt = z_hi / P.tz - tz_lo + 1;
  const int64_t n_tiles = (int64_t)P.nty * P.ntx * tz_cnt;
  const int grid = (int)(n_tiles < 2 * kNumSM ? n_tiles : 2 * kNumSM);
  if (dotv)
    grid_apply_tiled_kernel<DPN, true><<<grid, kBlock, op->smem, st>>>(
        P, node0, n_nodes, tz_lo, tz_cnt, x, y, dotv, rs->partials, rs->ticket, dot_out, S);
  else
    grid_apply_tiled_kernel<DPN, false><<<grid, kBlock, op->smem, st>>>(
        P, node0, n_nodes, tz_lo, tz_cnt, x, y, nullptr, nullptr, nullptr, nullptr, S);
  SKTB_KERNEL_OK();
  return 0;
}

int launch_hexgrid_apply(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                         const double *x, double *y, const double *dotv,
                         ReduceScratch *rs, double *dot_out, const PcgScalars *S,
                         cudaStream_t st) {
  if (op->dpn == 1 && op->scalar_direct) {
    const int g = grid_for(n_nodes, kBlock, 3);
    if (dotv)
      scalar_grid_apply_kernel<true><<<g, kBlock, 0, st>>>(
          op->P1, op->W1, node0, n_nodes, x, y, dotv, rs->partials, rs->ticket, dot_out, S);
    else
      scalar_grid_apply_kernel<false><<<g, kBlock, 0, st>>>(
          op->P1, op->W1, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr, S);
    SKTB_KERNEL_OK();
    return 0;
  }
  if (op->dpn == 1)
    return launch_tiled<1>(op, op->P1, node0, n_nodes, x, y, dotv, rs, dot_out, S, st);
  if (!op->direct)
    return launch_tiled<3>(op, op->P3, node0, n_nodes, x, y, dotv, rs, dot_out, S, st);
  if (op->shfl) {
    {
      const int rc = launch_x2<double, 0>(op, node0, n_nodes, x, y, dotv, rs, dot_out, S, nullptr,
                                          nullptr, 0.0, st);
      if (rc != -1) return rc;
    }
    const int g = grid_for(n_nodes, kBlock, 2);
    if (dotv)
      hexgrid_apply_shfl_kernel<double, 0, true><<<g, kBlock, 0, st>>>(
          op->P3, node0, n_nodes, x, y, dotv, rs->partials, rs->ticket, dot_out, S, nullptr,
          nullptr, 0.0);
    else
      hexgrid_apply_shfl_kernel<double, 0, false><<<g, kBlock, 0, st>>>(
          op->P3, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr, S, nullptr,
          nullptr, 0.0);
    SKTB_KERNEL_OK();
    return 0;
  }
  if (op->split) {
    const int sgrid = grid_for(n_nodes, kBlock / 2, 16);
    if (dotv)
      hexgrid_apply_split_kernel<true><<<sgrid, kBlock, 0, st>>>(
          op->P3, node0, n_nodes, x, y, dotv, rs->partials, rs->ticket, dot_out, S);
    else
      hexgrid_apply_split_kernel<false><<<sgrid, kBlock, 0, st>>>(
          op->P3, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr, S);
    SKTB_KERNEL_OK();
    return 0;
  }
  const int grid = grid_for(n_nodes, kBlock, 16);
  if (dotv)
    hexgrid_apply_kernel<true><<<grid, kBlock, 0, st>>>(
        op->P3, node0, n_nodes, x, y, dotv, rs->partials, rs->ticket, dot_out, S);
  else
    hexgrid_apply_kernel<false><<<grid, kBlock, 0, st>>>(
        op->P3, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr, S);
  SKTB_KERNEL_OK();
  return 0;
}

int gridop_dpn(const sktb_gridop *op) { return op->dpn; }

// products of the multigrid V-cycle: optional single precision, optional fused
// damped-Jacobi update y = x + omega dinv (b - A x).  Returns -1 when this
// operator has no such kernel (the caller then composes it from plain products).
int launch_hexgrid_apply_ex(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                            const double *x, double *y, bool fp32, const double *b,
                            const double *dinv, double omega, cudaStream_t st) {
  if (op->dpn != 3 || !op->shfl || !op->direct) return -1;
  {
    int rc;
    if (b)
      rc = fp32 ? launch_x2<float, 1>(op, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr,
                                      b, dinv, omega, st)
                : launch_x2<double, 1>(op, node0, n_nodes, x, y, nullptr, nullptr, nullptr,
                                       nullptr, b, dinv, omega, st);
    else
      rc = fp32 ? launch_x2<float, 0>(op, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr,
                                      nullptr, nullptr, 0.0, st)
                : launch_x2<double, 0>(op, node0, n_nodes, x, y, nullptr, nullptr, nullptr,
                                       nullptr, nullptr, nullptr, 0.0, st);
    if (rc != -1) return rc;
  }
  const int g = grid_for(n_nodes, kBlock, 2);
#define SKTB_EX(T, MODE)                                                               \
  hexgrid_apply_shfl_kernel<T, MODE, false><<<g, kBlock, 0, st>>>(                     \
      op->P3, node0, n_nodes, x, y, nullptr, nullptr, nullptr, nullptr, nullptr, b, dinv, \
      omega)
  if (b) {
    if (fp32) SKTB_EX(float, 1); else SKTB_EX(double, 1);
  } else {
    if (fp32) SKTB_EX(float, 0); else SKTB_EX(double, 0);
  }
#undef SKTB_EX
  SKTB_KERNEL_OK();
  return 0;
}

// tile shape: as many of the 256 threads busy as possible, small halo
template <int DPN>
static void choose_tiles(GridParams<DPN> &P, size_t *smem) {
  const double a = DPN == 3 ? 1300.0 : 150.0;  // per-thread compute vs staging weight
  double best = 1e300;
  for (int ty = 4; ty <= 64; ++ty)
    for (int tx = 1; tx <= 16; ++tx)
      for (int tz = 1; tz <= 8; ++tz) {
        const int T = ty * tx * tz;
        if (T > kBlock || T < 128) continue;
        const int halo = (ty + 2) * (tx + 2) * (tz + 2);
        const int et = (ty + 1) * (tx + 1) * (tz + 1);
        if (halo > kMaxHalo || et > kMaxElemTile) continue;
        const int64_t nt = (int64_t)((P.npy + ty - 1) / ty) * ((P.npx + tx - 1) / tx) *
                           ((P.npz + tz - 1) / tz);
        const double cost = (double)nt * (a + 25.0 * (DPN * halo + et) / (double)kBlock);
        if (cost < best) {
          best = cost;
          P.ty = ty;
          P.tx = tx;
          P.tz = tz;
        }
      }
  P.nty = (P.npy + P.ty - 1) / P.ty;
  P.ntx = (P.npx + P.tx - 1) / P.tx;
  P.ntz = (P.npz + P.tz - 1) / P.tz;
  *smem = sizeof(double) * ((size_t)DPN * (P.ty + 2) * (P.tx + 2) * (P.tz + 2) +
                            (size_t)(P.ty + 1) * (P.tx + 1) * (P.tz + 1));
}

extern "C" int sktb_gridop_create(sktb_gridop **out, int dpn, const int32_t *np_h,
                                  const double *ke_cc_h, int device) {
  SKTB_REQUIRE(out && np_h && ke_cc_h, "null argument");
  SKTB_REQUIRE(dpn == 1 || dpn == 3, "dofs per node must be 1 or 3");
  SKTB_REQUIRE(np_h[0] >= 2 && np_h[1] >= 2 && np_h[2] >= 2, "grid needs >= 1 cell per axis");
  sktb_gridop *op = new sktb_gridop();
  op->dpn = dpn;
  op->device = device;
  op->n_nodes = (int64_t)np_h[0] * np_h[1] * np_h[2];
  if (op->n_nodes >= ((int64_t)1 << 31) / 3) {
    delete op;
    SKTB_REQUIRE(false, "grid too large for 32-bit node indexing");
  }
  // the 3-dof product is latency bound either way and measures faster straight
  // from L1 (0.100 ms vs 0.128 ms at 1M nodes); SKTB_GRIDOP_TILED=1 forces the
  // shared-memory variant, which is the only one for the scalar operator
  const char *env = getenv("SKTB_GRIDOP_TILED");
  op->direct = dpn == 3 && !(env && env[0] == '1');
  const char *env2 = getenv("SKTB_GRIDOP_SPLIT");
  op->split = env2 && env2[0] == '1';
  const char *env3 = getenv("SKTB_GRIDOP_SHFL");
  op->shfl = !(env3 && env3[0] == '0');
  auto fill = [&](auto &P, int nke) {
    for (int i = 0; i < nke; ++i) P.ke[i] = ke_cc_h[i];
    for (int i = 0; i < nke; ++i) P.kef[i] = (float)ke_cc_h[i];
    P.npx = np_h[0];
    P.npy = np_h[1];
    P.npz = np_h[2];
    P.scale = nullptr;
    P.dmask = nullptr;
  };
  SKTB_CUDA_OK(cudaSetDevice(device));
  if (dpn == 3) {
    fill(op->P3, 576);
    choose_tiles<3>(op->P3, &op->smem);
    {
      // coefficient blocks of the x2 kernel, slot = (((dz+1) 3 + dx+1) 3 + dy+1) 8 + o
      const char *envx = getenv("SKTB_GRIDOP_X2");
      op->x2 = !(envx && envx[0] == '0');
      const char *envs = getenv("SKTB_GRIDOP_X2S");
      if (envs) op->x2s = envs[0] == '1' ? 2 : (envs[0] == '0' ? 0 : 1);
      std::vector<double> kt((size_t)kX2Slots * 10, 0.0);
      std::vector<float> ktf((size_t)kX2Slots * 12, 0.f);
      for (int dz = -1; dz <= 1; ++dz)
        for (int dx = -1; dx <= 1; ++dx)
          for (int dy = -1; dy <= 1; ++dy)
            for (int o = 0; o < 8; ++o) {
              const int ox = o & 1, oy = (o >> 1) & 1, oz = o >> 2;
              const int bx = dx + 1 - ox, by = dy + 1 - oy, bz = dz + 1 - oz;
              if (bx < 0 || bx > 1 || by < 0 || by > 1 || bz < 0 || bz > 1) continue;
              const int ca = (1 - ox) + 2 * (1 - oy) + 4 * (1 - oz);
              const int cb = bx + 2 * by + 4 * bz;
              const int slot = (((dz + 1) * 3 + (dx + 1)) * 3 + (dy + 1)) * 8 + o;
              for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                  const double v = ke_cc_h[(3 * ca + i) * 24 + 3 * cb + j];
                  kt[(size_t)slot * 10 + 3 * i + j] = v;
                  ktf[(size_t)slot * 12 + 3 * i + j] = (float)v;
                }
            }
      SKTB_CUDA_OK(cudaMalloc(&op->kt_d, sizeof(double) * kt.size()));
      SKTB_CUDA_OK(cudaMalloc(&op->ktf_d, sizeof(float) * ktf.size()));
      SKTB_CUDA_OK(cudaMemcpy(op->kt_d, kt.data(), sizeof(double) * kt.size(),
                              cudaMemcpyHostToDevice));
      SKTB_CUDA_OK(cudaMemcpy(op->ktf_d, ktf.data(), sizeof(float) * ktf.size(),
                              cudaMemcpyHostToDevice));
    }
    SKTB_CUDA_OK(cudaFuncSetAttribute(grid_apply_tiled_kernel<3, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    SKTB_CUDA_OK(cudaFuncSetAttribute(grid_apply_tiled_kernel<3, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  } else {
    fill(op->P1, 64);
    choose_tiles<1>(op->P1, &op->smem);
    op->W1 = make_scalar_stencil(op->P1.ke);
    const char *env4 = getenv("SKTB_GRIDOP_SCALAR_TILED");
    op->scalar_direct = !(env4 && env4[0] == '1');
  }
  *out = op;
  return 0;
}

extern "C" void sktb_gridop_destroy(sktb_gridop *op) {
  if (!op) return;
  cudaFree(op->kt_d);
  cudaFree(op->ktf_d);
  delete op;
}

extern "C" int sktb_gridop_set_fields(sktb_gridop *op, const double *scale,
                                      const uint8_t *dmask) {
  SKTB_REQUIRE(op && dmask, "null argument");
  SKTB_REQUIRE(scale || op->dpn == 1, "the elasticity operator needs the element moduli");
  op->P3.scale = op->P1.scale = scale;
  op->P3.dmask = op->P1.dmask = dmask;
  op->fields_set = true;
  return 0;
}

extern "C" int sktb_gridop_tile_shape(const sktb_gridop *op, int32_t *tile_h) {
  SKTB_REQUIRE(op && tile_h, "null argument");
  tile_h[0] = op->dpn == 3 ? op->P3.tx : op->P1.tx;
  tile_h[1] = op->dpn == 3 ? op->P3.ty : op->P1.ty;
  tile_h[2] = op->dpn == 3 ? op->P3.tz : op->P1.tz;
  return 0;
}

bool gridop_ready(const sktb_gridop *op) { return op && op->fields_set; }

extern "C" int sktb_gridop_apply(const sktb_gridop *op, int64_t node0,
                                 int64_t n_nodes, const double *x, double *y,
                                 void *stream) {
  SKTB_REQUIRE(gridop_ready(op) && x && y, "null argument");
  SKTB_REQUIRE(node0 >= 0 && n_nodes > 0 && node0 + n_nodes <= op->n_nodes, "bad node range");
  return launch_hexgrid_apply(op, node0, n_nodes, x, y, nullptr, nullptr, nullptr,
                              nullptr, (cudaStream_t)stream);
}

extern "C" int sktb_gridop_inv_diag(const sktb_gridop *op, int64_t node0,
                                    int64_t n_nodes, double *out, void *stream) {
  SKTB_REQUIRE(gridop_ready(op) && out, "null argument");
  SKTB_REQUIRE(node0 >= 0 && n_nodes > 0 && node0 + n_nodes <= op->n_nodes, "bad node range");
  if (op->dpn == 3)
    grid_inv_diag_kernel<3><<<grid_for(n_nodes), kBlock, 0, (cudaStream_t)stream>>>(
        op->P3, node0, n_nodes, out);
  else
    grid_inv_diag_kernel<1><<<grid_for(n_nodes), kBlock, 0, (cudaStream_t)stream>>>(
        op->P1, node0, n_nodes, out);
  SKTB_KERNEL_OK();
  return 0;
}
