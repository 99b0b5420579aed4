// Geometric multigrid preconditioner for the 3-dof elasticity operator on
// tensor-product hexahedral grids ("AMG-smoothed PCG" of the north star, built
// on the grid hierarchy because the benchmark meshes are box grids).
//
//  * Hierarchy: every level halves the cell counts (ceil), coarse nodes are a
//    subset of the fine nodes, prolongation is trilinear interpolation.
//  * Coarse operators are exact Galerkin products A_{l+1} = P^T A_l P, formed
//    element-wise: a coarse element matrix is sum_c Q_c^T Ke_child Q_c over its
//    (up to 8) children (elem_restrict_kernel), then gathered into the level's
//    CSR by the ordinary assembly kernel (mesh.cu).  Level-0 children are
//    E_e * Ke0[class].
//  * V(1,1) cycle with damped Jacobi; residual and restriction are fused.
//  * Dirichlet dofs are masked on every level (coarse dof fixed iff the
//    coincident fine dof is fixed), so M^-1 stays symmetric positive definite.
#include <cooperative_groups.h>

#include <vector>

#include <cstdlib>
#include <utility>

#include "comm.cuh"
#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

constexpr int kTailMaxLevels = 8;   // fused coarse tail (mg_tail_kernel)
constexpr int kTailMaxNodes = 1024;
constexpr int kTailGrid = 148;

struct MgLevel {
  int64_t n_nodes = 0, n_blocks = 0;
  int max_deg = 0;
  const int32_t *node_ptr = nullptr, *node_col = nullptr;
  const double *vals = nullptr, *inv_diag = nullptr;
  // single-precision copy of vals (caller-owned, may be null): the V-cycle's
  // products on this level stream it instead (sktb_mg_set_level_vals32)
  const float *vals32 = nullptr;
  const uint8_t *mask = nullptr;  // per dof, may be null
  const sktb_gridop *gop = nullptr;  // level 0 only: matrix-free operator
  double *dense_inv = nullptr;       // coarsest level only: dense inverse (owned unless shared)
  bool dense_shared = false;         // dense_inv belongs to another hierarchy (sktb_mg_share_coarsest)
  int dense_n = 0;                   // 0: not factored
  double *x = nullptr, *b = nullptr, *tmp = nullptr;  // owned work vectors
  // level 0 only: rows owned by this rank (node0 = 0, n_nodes = n_global when
  // the operator is not sharded); x is always full length (n_global nodes)
  int64_t node0 = 0, n_global = 0;
  // z-slab sharding of this level (level 0 and the large assembled levels):
  // whole node planes per rank, ghost planes exchanged with prev / next
  bool sharded = false;
  int64_t plane = 0;         // nodes per z-plane
  int prev = -1, next = -1;  // neighbour ranks (-1: none)
  double *res = nullptr;     // full-length residual for a sharded -> sharded restriction (owned)
  double omega = 0.0;  // damping of this level's Jacobi smoother (0: use the global one)
  int nu = 1;          // pre- and post-smoothing sweeps on this level
  bool cheb = false;   // Chebyshev polynomial smoother of degree nu instead of nu Jacobi sweeps
  double cheb_c1[8] = {0}, cheb_c2[8] = {0};
  double *d = nullptr; // Chebyshev direction (owned)
  double *x2 = nullptr; // second iterate of the fused tail kernel (owned)
  // transfer to the next coarser level (tensor grid tables, device)
  int32_t fnp[3] = {0, 0, 0}, cnp[3] = {0, 0, 0};  // nodes per axis (x, y, z)
  const int32_t *ax_c0 = nullptr, *ax_c1 = nullptr;  // [fnx | fny | fnz]
  const double *ax_w0 = nullptr, *ax_w1 = nullptr;
  const int32_t *axT_f = nullptr;  // [3 slots][cnx | cny | cnz], -1 = none
  const double *axT_w = nullptr;
};

struct sktb_mg {
  int device = 0;
  std::vector<MgLevel> lv;
  double omega = 0.5;
  int nu_coarse = 20;
  bool fp32_level0 = false;  // single-precision products on a matrix-free level 0
  bool fused_tail = false;   // levels <= kTailMaxNodes nodes in one cooperative kernel
  bool fused_sweeps = true;  // Jacobi sweeps of levels >= 1 fused into the product kernel
  // the restriction into a replicated level also writes that level's first Jacobi step
  // (SKTB_MG_FUSE_RESTRICT_JACOBI=0 keeps the separate kernel)
  bool fuse_restrict_jacobi = true;
};

extern "C" int sktb_mg_create(sktb_mg **out, int n_levels, int device) {
  SKTB_REQUIRE(out && n_levels >= 1 && n_levels <= 16, "bad argument");
  sktb_mg *m = new sktb_mg();
  m->device = device;
  m->lv.resize(n_levels);
  const char *env = getenv("SKTB_MG_FUSED_SWEEPS");
  m->fused_sweeps = !(env && env[0] == '0');
  const char *env2 = getenv("SKTB_MG_FUSE_RESTRICT_JACOBI");
  m->fuse_restrict_jacobi = !(env2 && env2[0] == '0');
  *out = m;
  return 0;
}

extern "C" void sktb_mg_destroy(sktb_mg *m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (auto &l : m->lv) {
    dev_free(l.x);
    cudaFree(l.b);
    cudaFree(l.tmp);
    if (!l.dense_shared) cudaFree(l.dense_inv);
    cudaFree(l.d);
    dev_free(l.x2);
    dev_free(l.res);
  }
  delete m;
}

extern "C" int sktb_mg_set_params(sktb_mg *m, double omega, int nu_coarse) {
  SKTB_REQUIRE(m && omega > 0.0 && omega < 2.0 && nu_coarse >= 0, "bad argument");
  m->omega = omega;
  m->nu_coarse = nu_coarse;
  return 0;
}

extern "C" int sktb_mg_set_level_sweeps(sktb_mg *m, int level, int nu) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size() && nu >= 1 && nu <= 8,
               "bad argument");
  m->lv[level].nu = nu;
  return 0;
}

// Chebyshev smoother of degree nu on one level (>= 1): d_k = c1[k] d_{k-1} +
// c2[k] D^-1 r_k, x += d_k (c1[0] is ignored); replaces the nu Jacobi sweeps
extern "C" int sktb_mg_set_level_cheby(sktb_mg *m, int level, int nu, const double *c1_h,
                                       const double *c2_h) {
  SKTB_REQUIRE(m && level >= 1 && level < (int)m->lv.size() && nu >= 1 && nu <= 8 && c1_h &&
                   c2_h,
               "bad argument");
  MgLevel &l = m->lv[level];
  SKTB_REQUIRE(l.n_nodes > 0, "set the level before its smoother");
  SKTB_CUDA_OK(cudaSetDevice(m->device));
  if (!l.d) SKTB_CUDA_OK(cudaMalloc(&l.d, sizeof(double) * 3 * l.n_nodes));
  l.nu = nu;
  l.cheb = true;
  for (int k = 0; k < nu; ++k) {
    l.cheb_c1[k] = c1_h[k];
    l.cheb_c2[k] = c2_h[k];
  }
  return 0;
}

extern "C" int sktb_mg_set_precision(sktb_mg *m, int fp32_level0) {
  SKTB_REQUIRE(m, "null argument");
  m->fp32_level0 = fp32_level0 != 0;
  return 0;
}

extern "C" int sktb_mg_set_fused_tail(sktb_mg *m, int on) {
  SKTB_REQUIRE(m, "null argument");
  m->fused_tail = on != 0;
  return 0;
}

extern "C" int sktb_mg_set_level_omega(sktb_mg *m, int level, double omega) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size() && omega > 0.0 && omega < 2.0,
               "bad argument");
  m->lv[level].omega = omega;
  return 0;
}

extern "C" int sktb_mg_set_level(sktb_mg *m, int level, int64_t n_nodes,
                                 int64_t n_blocks, int max_deg,
                                 const int32_t *node_ptr,
                                 const int32_t *node_col, const double *vals,
                                 const double *inv_diag, const uint8_t *mask) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size(), "bad level");
  SKTB_REQUIRE(node_ptr && node_col && vals && inv_diag && n_nodes > 0, "null argument");
  SKTB_CUDA_OK(cudaSetDevice(m->device));
  MgLevel &l = m->lv[level];
  if (l.n_global < n_nodes) l.n_global = n_nodes;
  if (!l.sharded && level > 0) l.node0 = 0;
  if (l.n_nodes != n_nodes || !l.x) {
    dev_free(l.x);
    cudaFree(l.b);
    cudaFree(l.tmp);
    if (l.sharded) {
      if (dev_alloc_exchangeable(&l.x, (size_t)3 * l.n_global)) return 1;
    } else {
      SKTB_CUDA_OK(cudaMalloc(&l.x, sizeof(double) * 3 * l.n_global));
      SKTB_CUDA_OK(cudaMemset(l.x, 0, sizeof(double) * 3 * l.n_global));
    }
    SKTB_CUDA_OK(cudaMalloc(&l.b, sizeof(double) * 3 * n_nodes));
    SKTB_CUDA_OK(cudaMalloc(&l.tmp, sizeof(double) * 3 * n_nodes));
    dev_free(l.x2);
    l.x2 = nullptr;
    if (level > 0 && l.sharded) {
      // second full-length iterate of the fused Jacobi sweeps
      if (dev_alloc_exchangeable(&l.x2, (size_t)3 * l.n_global)) return 1;
    } else if (level > 0 && n_nodes <= kTailMaxNodes) {
      SKTB_CUDA_OK(cudaMalloc(&l.x2, sizeof(double) * 3 * n_nodes));
    }
  }
  l.n_nodes = n_nodes;
  l.gop = nullptr;
  l.dense_n = 0;
  l.n_blocks = n_blocks;
  l.max_deg = max_deg;
  l.node_ptr = node_ptr;
  l.node_col = node_col;
  l.vals = vals;
  l.vals32 = nullptr;
  l.inv_diag = inv_diag;
  l.mask = mask;
  return 0;
}

// single-precision copy of the level's values (same layout; sktb_f64_to_f32) for
// the V-cycle's products; call after sktb_mg_set_level, which forgets it
extern "C" int sktb_mg_set_level_vals32(sktb_mg *m, int level, const float *vals32) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size(), "bad level");
  SKTB_REQUIRE(m->lv[level].vals, "set the level before its single-precision copy");
  m->lv[level].vals32 = vals32;
  return 0;
}

// matrix-free level 0 (gridop.cu) instead of assembled values
extern "C" int sktb_mg_set_level0_grid(sktb_mg *m, const sktb_gridop *op,
                                       int64_t n_nodes, const double *inv_diag,
                                       const uint8_t *mask) {
  SKTB_REQUIRE(m && gridop_ready(op) && inv_diag && n_nodes > 0, "null argument");
  SKTB_CUDA_OK(cudaSetDevice(m->device));
  MgLevel &l = m->lv[0];
  if (l.n_global < n_nodes) l.n_global = n_nodes;
  if (l.n_nodes != n_nodes || !l.x) {
    dev_free(l.x);
    cudaFree(l.b);
    cudaFree(l.tmp);
    if (l.sharded) {
      if (dev_alloc_exchangeable(&l.x, (size_t)3 * l.n_global)) return 1;
    } else {
      SKTB_CUDA_OK(cudaMalloc(&l.x, sizeof(double) * 3 * l.n_global));
      SKTB_CUDA_OK(cudaMemset(l.x, 0, sizeof(double) * 3 * l.n_global));
    }
    SKTB_CUDA_OK(cudaMalloc(&l.b, sizeof(double) * 3 * n_nodes));
    SKTB_CUDA_OK(cudaMalloc(&l.tmp, sizeof(double) * 3 * n_nodes));
  }
  l.n_nodes = n_nodes;
  l.gop = op;
  l.node_ptr = l.node_col = nullptr;
  l.vals = nullptr;
  l.inv_diag = inv_diag;
  l.mask = mask;
  return 0;
}

// level 0 of a row-sharded operator: this rank owns the nodes
// [node0, node0 + n_owned) of n_global; call before sktb_mg_set_level(0, ...)
extern "C" int sktb_mg_set_level0_range(sktb_mg *m, int64_t node0,
                                        int64_t n_global) {
  SKTB_REQUIRE(m && node0 >= 0 && n_global > 0, "bad argument");
  MgLevel &l = m->lv[0];
  if (l.n_global != n_global) {
    dev_free(l.x);
    l.x = nullptr;
  }
  l.node0 = node0;
  l.n_global = n_global;
  return 0;
}

// z-slab sharding of one level: this rank owns the nodes [node0, node0 +
// n_owned) (whole planes of `plane_nodes` nodes) of n_global; prev / next are
// the ranks owning the adjacent planes (-1: none).  Call before
// sktb_mg_set_level / sktb_mg_set_level0_grid of that level.  plane_nodes = 0
// marks the level as replicated again.
extern "C" int sktb_mg_set_level_slab(sktb_mg *m, int level, int64_t node0,
                                      int64_t n_global, int64_t plane_nodes,
                                      int prev_rank, int next_rank) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size() && node0 >= 0 && n_global > 0 &&
                   plane_nodes >= 0,
               "bad argument");
  MgLevel &l = m->lv[level];
  if (l.n_global != n_global || l.sharded != (plane_nodes > 0)) {
    dev_free(l.x);
    l.x = nullptr;
    dev_free(l.res);
    l.res = nullptr;
  }
  l.node0 = node0;
  l.n_global = n_global;
  l.sharded = plane_nodes > 0;
  l.plane = plane_nodes;
  l.prev = prev_rank;
  l.next = next_rank;
  return 0;
}

extern "C" int sktb_mg_set_transfer(sktb_mg *m, int level, const int32_t *fine_np_h,
                                    const int32_t *coarse_np_h,
                                    const int32_t *ax_c0, const int32_t *ax_c1,
                                    const double *ax_w0, const double *ax_w1,
                                    const int32_t *axT_f, const double *axT_w) {
  SKTB_REQUIRE(m && level >= 0 && level + 1 < (int)m->lv.size(), "bad level");
  SKTB_REQUIRE(fine_np_h && coarse_np_h && ax_c0 && ax_c1 && ax_w0 && ax_w1 &&
                   axT_f && axT_w,
               "null argument");
  MgLevel &l = m->lv[level];
  for (int a = 0; a < 3; ++a) {
    l.fnp[a] = fine_np_h[a];
    l.cnp[a] = coarse_np_h[a];
  }
  l.ax_c0 = ax_c0;
  l.ax_c1 = ax_c1;
  l.ax_w0 = ax_w0;
  l.ax_w1 = ax_w1;
  l.axT_f = axT_f;
  l.axT_w = axT_w;
  return 0;
}

#define GS(i, n)                                                       \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x,     \
               _st = (int64_t)gridDim.x * blockDim.x;                  \
       i < (n); i += _st)

// ------------------------------------------------ Galerkin element matrices --
// out[E] = sum_c Q_c^T K_child Q_c (Q_c = trilinear weights x I3).  One CTA per
// coarse element, three output entries per thread.  Per child: the 24x24 child
// matrix is staged in shared memory (coalesced), T = K_child (Q x I3) is formed
// there (8 FMAs per entry), then out += (Q x I3)^T T (8 FMAs per entry): two
// small dense products from shared memory instead of a sparse triple sum over
// strided global loads.
__global__ void __launch_bounds__(192)
    elem_restrict_kernel(int64_t n_coarse, const int32_t *__restrict__ child,
                         const uint8_t *__restrict__ ptype,
                         const double *__restrict__ Qtab,
                         const double *__restrict__ fine_ke,
                         const double *__restrict__ unit,
                         const int32_t *__restrict__ cls,
                         const double *__restrict__ scale,
                         double *__restrict__ out, int64_t e_lo, int64_t fine_base) {
  // coarse elements [e_lo, e_lo + gridDim.x); out holds them from e_lo on,
  // fine_ke holds the fine elements from fine_base on (slab-sharded set-up)
  const int64_t E = e_lo + blockIdx.x;
  const int type = ptype[E];
  __shared__ int32_t ch[8];
  __shared__ double Q[8][64];     // Q[c][a * 8 + A]: weight of parent vertex A in child vertex a
  __shared__ double Kc[24][25];   // padded: column reads of the second product
  __shared__ double T[24][25];
  if (threadIdx.x < 8) ch[threadIdx.x] = child[(int64_t)threadIdx.x * n_coarse + E];
  for (int k = threadIdx.x; k < 512; k += blockDim.x)
    Q[k >> 6][k & 63] = Qtab[(type * 8 + (k >> 6)) * 64 + (k & 63)];
  double acc[3] = {0.0, 0.0, 0.0};
  __syncthreads();
  for (int cc = 0; cc < 8; ++cc) {
    const int32_t ce = ch[cc];
    if (ce < 0) continue;  // uniform across the CTA
    const double *src;
    double sc = 1.0;
    if (fine_ke) {
      src = fine_ke + ((int64_t)ce - fine_base) * 576;
    } else {
      src = unit + (int64_t)(cls ? cls[ce] : 0) * 576;
      sc = scale[ce];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ent = threadIdx.x + 192 * k;
      Kc[ent / 24][ent % 24] = sc * __ldg(&src[ent]);
    }
    __syncthreads();
    // T[r][3B + j] = sum_b Kc[r][3b + j] Q[b][B]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ent = threadIdx.x + 192 * k;
      const int r = ent / 24, c = ent - 24 * r;
      const int B = c / 3, j = c - 3 * B;
      double t = 0.0;
#pragma unroll
      for (int bb = 0; bb < 8; ++bb) t = fma(Kc[r][3 * bb + j], Q[cc][bb * 8 + B], t);
      T[r][c] = t;
    }
    __syncthreads();
    // out[3A + i][c] += sum_a Q[a][A] T[3a + i][c]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int ent = threadIdx.x + 192 * k;
      const int r = ent / 24, c = ent - 24 * r;
      const int A = r / 3, i = r - 3 * A;
      double t = acc[k];
#pragma unroll
      for (int aa = 0; aa < 8; ++aa) t = fma(Q[cc][aa * 8 + A], T[3 * aa + i][c], t);
      acc[k] = t;
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) out[(E - e_lo) * 576 + threadIdx.x + 192 * k] = acc[k];
}

// Level 0 -> 1 fast path: children are scale[e] * Ke0[class], so the Galerkin
// element matrix is a linear combination of precomputed tables
//   T[(cls * 8 + type) * 8 + c] = Q_c^T Ke0[cls] Q_c   (576 doubles each)
__global__ void __launch_bounds__(192)
    elem_combine_kernel(int64_t n_coarse, const int32_t *__restrict__ child,
                        const uint8_t *__restrict__ ptype,
                        const double *__restrict__ T,
                        const int32_t *__restrict__ cls,
                        const double *__restrict__ scale,
                        double *__restrict__ out, int64_t e_lo, int64_t e_hi) {
  // persistent CTAs, one coarse element per trip, three entries per thread; the
  // tables of the common case (class 0, full 2x2x2 parent) live in registers
  double t0[8][3];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) t0[c][k] = __ldg(&T[(int64_t)c * 576 + threadIdx.x + 192 * k]);
  __shared__ double s_sc[2][8];
  __shared__ int s_tab[2][8];  // table index (cls * 8 + type) * 8 + c, -1 = no child
  int buf = 0;
  for (int64_t E = e_lo + blockIdx.x; E < e_hi; E += gridDim.x, buf ^= 1) {
    if (threadIdx.x < 8) {
      const int c = threadIdx.x;
      const int32_t ce = child[(int64_t)c * n_coarse + E];
      s_sc[buf][c] = ce >= 0 ? scale[ce] : 0.0;
      s_tab[buf][c] = ce >= 0 ? ((cls ? cls[ce] : 0) * 8 + ptype[E]) * 8 + c : -1;
    }
    __syncthreads();  // (double-buffered: one barrier per trip)
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int tab = s_tab[buf][c];
      if (tab < 0) continue;
      const double sc = s_sc[buf][c];
      if (tab == c) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[k] = fma(sc, t0[c][k], acc[k]);
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k)
          acc[k] = fma(sc, __ldg(&T[(int64_t)tab * 576 + threadIdx.x + 192 * k]), acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[(E - e_lo) * 576 + threadIdx.x + 192 * k] = acc[k];
  }
}

extern "C" int sktb_elem_combine_range(int64_t n_coarse, int64_t e_lo, int64_t e_hi,
                                       const int32_t *child, const uint8_t *ptype,
                                       const double *T, const int32_t *cls,
                                       const double *scale, double *out, void *stream) {
  SKTB_REQUIRE(child && ptype && T && scale && out && n_coarse > 0, "null argument");
  SKTB_REQUIRE(0 <= e_lo && e_lo < e_hi && e_hi <= n_coarse, "bad element range");
  const int64_t cap = (int64_t)kNumSM * 10, ne = e_hi - e_lo;
  elem_combine_kernel<<<(unsigned)(ne < cap ? ne : cap), 192, 0, (cudaStream_t)stream>>>(
      n_coarse, child, ptype, T, cls, scale, out, e_lo, e_hi);
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_elem_combine(int64_t n_coarse, const int32_t *child,
                                 const uint8_t *ptype, const double *T,
                                 const int32_t *cls, const double *scale,
                                 double *out, void *stream) {
  return sktb_elem_combine_range(n_coarse, 0, n_coarse, child, ptype, T, cls, scale, out,
                                 stream);
}

extern "C" int sktb_elem_restrict_range(int64_t n_coarse, int64_t e_lo, int64_t e_hi,
                                        int64_t fine_base, const int32_t *child,
                                        const uint8_t *ptype, const double *Qtab,
                                        const double *fine_ke, const double *unit,
                                        const int32_t *cls, const double *scale,
                                        double *out, void *stream) {
  SKTB_REQUIRE(child && ptype && Qtab && out && n_coarse > 0, "null argument");
  SKTB_REQUIRE(fine_ke || (unit && scale), "need fine element matrices or (unit, scale)");
  SKTB_REQUIRE(0 <= e_lo && e_lo < e_hi && e_hi <= n_coarse, "bad element range");
  elem_restrict_kernel<<<(unsigned)(e_hi - e_lo), 192, 0, (cudaStream_t)stream>>>(
      n_coarse, child, ptype, Qtab, fine_ke, unit, cls, scale, out, e_lo, fine_base);
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_elem_restrict(int64_t n_coarse, const int32_t *child,
                                  const uint8_t *ptype, const double *Qtab,
                                  const double *fine_ke, const double *unit,
                                  const int32_t *cls, const double *scale,
                                  double *out, void *stream) {
  return sktb_elem_restrict_range(n_coarse, 0, n_coarse, 0, child, ptype, Qtab, fine_ke, unit,
                                  cls, scale, out, stream);
}

// ------------------------------------------------------------ level kernels --

__global__ void __launch_bounds__(kBlock)
    mg_jacobi0_kernel(int64_t n, double omega, const double *__restrict__ dinv,
                      const double *__restrict__ b, double *__restrict__ x) {
  GS(i, n) x[i] = omega * dinv[i] * b[i];
}
__global__ void __launch_bounds__(kBlock)
    mg_jacobi_kernel(int64_t n, double omega, const double *__restrict__ dinv,
                     const double *__restrict__ b, const double *__restrict__ Ax,
                     double *__restrict__ x) {
  GS(i, n) x[i] += omega * dinv[i] * (b[i] - Ax[i]);
}

// Chebyshev step: d = c1 d + c2 D^-1 (b - Ax) ; x = (from_zero ? d : x + d)
__global__ void __launch_bounds__(kBlock)
    mg_cheby_kernel(int64_t n, double c1, double c2, const double *__restrict__ dinv,
                    const double *__restrict__ b, const double *__restrict__ Ax,
                    int from_zero, double *__restrict__ d, double *__restrict__ x) {
  GS(i, n) {
    const double r = Ax ? b[i] - Ax[i] : b[i];
    const double di = (c1 != 0.0 ? c1 * d[i] : 0.0) + c2 * dinv[i] * r;
    d[i] = di;
    x[i] = from_zero ? di : x[i] + di;
  }
}

// b_c = mask_c * P^T (b_f - Ax_f) ; one thread per coarse node (large levels:
// 136k coarse nodes at C2 take 29 us this way, 61 us with a warp per node)
__global__ void __launch_bounds__(kBlock)
    mg_restrict_thread_kernel(int cnx, int cny, int cnz, int fnx, int fny, int fnz,
                              const int32_t *__restrict__ axT_f,
                              const double *__restrict__ axT_w,
                              const double *__restrict__ bf,
                              const double *__restrict__ Axf,
                              const uint8_t *__restrict__ mask_c,
                              double *__restrict__ bc, int64_t f_lo, int64_t f_hi,
                              int64_t c_lo, int64_t c_hi, int64_t c_base,
                              const double *__restrict__ jd, double jom,
                              double *__restrict__ jx) {
  // coarse nodes [c_lo, c_hi) are produced, written at bc[3 (I - c_base)];
  // bf / Axf hold the fine rows [f_lo, f_hi) (Axf may be null: bf is a residual)
  const int tot = cnx + cny + cnz;
  GS(Ii, c_hi - c_lo) {
    const int64_t I = c_lo + Ii;
    const int iy = (int)(I % cny);
    const int ix = (int)((I / cny) % cnx);
    const int iz = (int)(I / ((int64_t)cny * cnx));
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int sz = 0; sz < 3; ++sz) {
      const int fz = axT_f[sz * tot + cnx + cny + iz];
      if (fz < 0) continue;
      const double wz = axT_w[sz * tot + cnx + cny + iz];
      for (int sx = 0; sx < 3; ++sx) {
        const int fx = axT_f[sx * tot + ix];
        if (fx < 0) continue;
        const double wx = axT_w[sx * tot + ix] * wz;
        for (int sy = 0; sy < 3; ++sy) {
          const int fy = axT_f[sy * tot + cnx + iy];
          if (fy < 0) continue;
          const double w = axT_w[sy * tot + cnx + iy] * wx;
          const int64_t fn = (int64_t)fy + (int64_t)fny * fx + (int64_t)fny * fnx * fz;
          if (fn < f_lo || fn >= f_hi) continue;  // another rank's fine node
          const int64_t f = 3 * (fn - f_lo);      // bf / Axf hold the owned rows
          if (Axf) {
            a0 += w * (bf[f] - Axf[f]);
            a1 += w * (bf[f + 1] - Axf[f + 1]);
            a2 += w * (bf[f + 2] - Axf[f + 2]);
          } else {
            a0 += w * bf[f];
            a1 += w * bf[f + 1];
            a2 += w * bf[f + 2];
          }
        }
      }
    }
    const int64_t o = 3 * I, q = 3 * (I - c_base);
    a0 = (mask_c && mask_c[o]) ? 0.0 : a0;
    a1 = (mask_c && mask_c[o + 1]) ? 0.0 : a1;
    a2 = (mask_c && mask_c[o + 2]) ? 0.0 : a2;
    bc[q] = a0;
    bc[q + 1] = a1;
    bc[q + 2] = a2;
    if (jx) {  // first damped-Jacobi step of the coarse level from a zero iterate
      jx[q] = jom * jd[q] * a0;
      jx[q + 1] = jom * jd[q + 1] * a1;
      jx[q + 2] = jom * jd[q + 2] * a2;
    }
  }
}

// r[own rows of the full-length vector] = b - Ax (b, Ax: owned rows)
__global__ void __launch_bounds__(kBlock)
    mg_residual_kernel(int64_t n, const double *__restrict__ b,
                       const double *__restrict__ Ax, double *__restrict__ r) {
  GS(i, n) r[i] = b[i] - Ax[i];
}

// b_c = mask_c * P^T (b_f - Ax_f) ; one WARP per coarse node, one lane per
// fine node of its 3x3x3 stencil (fixed-order shuffle reduction: deterministic)
__global__ void __launch_bounds__(kBlock)
    mg_restrict_kernel(int cnx, int cny, int cnz, int fnx, int fny, int fnz,
                       const int32_t *__restrict__ axT_f,
                       const double *__restrict__ axT_w,
                       const double *__restrict__ bf,
                       const double *__restrict__ Axf,
                       const uint8_t *__restrict__ mask_c,
                       double *__restrict__ bc, int64_t f_lo, int64_t f_hi,
                       int64_t c_lo, int64_t c_hi, int64_t c_base,
                       const double *__restrict__ jd, double jom, double *__restrict__ jx) {
  const int tot = cnx + cny + cnz;
  const int lane = threadIdx.x & 31;
  const int sy = lane % 3, sx = (lane / 3) % 3, sz = lane / 9;  // lanes >= 27 idle
  const int64_t wstride = (int64_t)gridDim.x * (kBlock / 32);
  for (int64_t I = c_lo + (int64_t)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5); I < c_hi;
       I += wstride) {
    const int iy = (int)(I % cny);
    const int ix = (int)((I / cny) % cnx);
    const int iz = (int)(I / ((int64_t)cny * cnx));
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    if (lane < 27) {
      const int fz = axT_f[sz * tot + cnx + cny + iz];
      const int fx = axT_f[sx * tot + ix];
      const int fy = axT_f[sy * tot + cnx + iy];
      if (fz >= 0 && fx >= 0 && fy >= 0) {
        const int64_t fn = (int64_t)fy + (int64_t)fny * fx + (int64_t)fny * fnx * fz;
        if (fn >= f_lo && fn < f_hi) {          // else: another rank's fine node
          const double w = axT_w[sz * tot + cnx + cny + iz] * axT_w[sx * tot + ix] *
                           axT_w[sy * tot + cnx + iy];
          const int64_t f = 3 * (fn - f_lo);    // bf / Axf hold the owned rows
          a0 = w * (Axf ? bf[f] - Axf[f] : bf[f]);
          a1 = w * (Axf ? bf[f + 1] - Axf[f + 1] : bf[f + 1]);
          a2 = w * (Axf ? bf[f + 2] - Axf[f + 2] : bf[f + 2]);
        }
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      const int64_t o = 3 * I, q = 3 * (I - c_base);
      a0 = (mask_c && mask_c[o]) ? 0.0 : a0;
      a1 = (mask_c && mask_c[o + 1]) ? 0.0 : a1;
      a2 = (mask_c && mask_c[o + 2]) ? 0.0 : a2;
      bc[q] = a0;
      bc[q + 1] = a1;
      bc[q + 2] = a2;
      if (jx) {  // first damped-Jacobi step of the coarse level from a zero iterate
        jx[q] = jom * jd[q] * a0;
        jx[q + 1] = jom * jd[q + 1] * a1;
        jx[q + 2] = jom * jd[q + 2] * a2;
      }
    }
  }
}

// x_f += mask_f * P x_c ; one thread per fine node
__global__ void __launch_bounds__(kBlock)
    mg_prolong_kernel(int cnx, int cny, int cnz, int fnx, int fny, int fnz,
                      const int32_t *__restrict__ c0, const int32_t *__restrict__ c1,
                      const double *__restrict__ w0, const double *__restrict__ w1,
                      const double *__restrict__ xc,
                      const uint8_t *__restrict__ mask_f,
                      double *__restrict__ xf, int64_t f_lo, int64_t f_hi) {
  GS(Fi, f_hi - f_lo) {
    const int64_t F = f_lo + Fi;
    const int iy = (int)(F % fny);
    const int ix = (int)((F / fny) % fnx);
    const int iz = (int)(F / ((int64_t)fny * fnx));
    const int cx[2] = {c0[ix], c1[ix]};
    const double wx[2] = {w0[ix], w1[ix]};
    const int cy[2] = {c0[fnx + iy], c1[fnx + iy]};
    const double wy[2] = {w0[fnx + iy], w1[fnx + iy]};
    const int cz[2] = {c0[fnx + fny + iz], c1[fnx + fny + iz]};
    const double wz[2] = {w0[fnx + fny + iz], w1[fnx + fny + iz]};
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int kz = 0; kz < 2; ++kz) {
      if (wz[kz] == 0.0) continue;
#pragma unroll
      for (int kx = 0; kx < 2; ++kx) {
        if (wx[kx] == 0.0) continue;
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
          if (wy[ky] == 0.0) continue;
          const double w = wz[kz] * wx[kx] * wy[ky];
          const int64_t c = 3 * ((int64_t)cy[ky] + (int64_t)cny * cx[kx] +
                                 (int64_t)cny * cnx * cz[kz]);
          a0 += w * xc[c];
          a1 += w * xc[c + 1];
          a2 += w * xc[c + 2];
        }
      }
    }
    const int64_t o = 3 * F;
    if (!(mask_f && mask_f[o])) xf[o] += a0;
    if (!(mask_f && mask_f[o + 1])) xf[o + 1] += a1;
    if (!(mask_f && mask_f[o + 2])) xf[o + 2] += a2;
  }
}

// Coarsest level: all damped-Jacobi sweeps in ONE single-CTA kernel (the level
// has a few hundred nodes; 2 x nu launches of ~4 us each would dominate it).
// x = omega D^-1 b, then nu times x += omega D^-1 (b - A x).
__global__ void __launch_bounds__(1024)
    mg_coarse_solve_kernel(int n_nodes, const int32_t *__restrict__ node_ptr,
                           const int32_t *__restrict__ node_col,
                           const double *__restrict__ vals,
                           const double *__restrict__ dinv,
                           const double *__restrict__ b, double omega, int nu,
                           double *__restrict__ x, double *__restrict__ tmp) {
  const int n = 3 * n_nodes;
  for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = omega * dinv[i] * b[i];
  __syncthreads();
  for (int s = 0; s < nu; ++s) {
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
      const int nd = r / 3, ri = r - 3 * nd;
      const int32_t s0 = node_ptr[nd], deg = node_ptr[nd + 1] - s0;
      const double *vp = vals + (int64_t)9 * s0 + (int64_t)ri * 3 * deg;
      double acc = 0.0;
      for (int k = 0; k < deg; ++k) {
        const double *xb = x + 3 * node_col[s0 + k];
        acc += vp[3 * k] * xb[0] + vp[3 * k + 1] * xb[1] + vp[3 * k + 2] * xb[2];
      }
      tmp[r] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      x[i] += omega * dinv[i] * (b[i] - tmp[i]);
    __syncthreads();
  }
}

// Exact coarsest-level solve: the (enforced, SPD) operator of the last level is
// expanded to a dense n x n matrix in shared memory and inverted in place by
// Gauss-Jordan elimination (diagonal pivots) once per set-up; the V-cycle then
// applies x = A^-1 b as one dense product.  Single CTA, n <= kDenseMax.
constexpr int kDenseMax = 160;

__global__ void __launch_bounds__(1024)
    mg_dense_invert_kernel(int n_nodes, const int32_t *__restrict__ node_ptr,
                           const int32_t *__restrict__ node_col,
                           const double *__restrict__ vals,
                           double *__restrict__ inv) {
  extern __shared__ double sm[];
  const int n = 3 * n_nodes;
  double *A = sm;            // n x n
  double *col = sm + n * n;  // n
  __shared__ double piv_inv;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) A[i] = 0.0;
  __syncthreads();
  for (int r = threadIdx.x; r < n; r += blockDim.x) {
    const int nd = r / 3, ri = r - 3 * nd;
    const int32_t s0 = node_ptr[nd], deg = node_ptr[nd + 1] - s0;
    const double *vp = vals + (int64_t)9 * s0 + (int64_t)ri * 3 * deg;
    for (int k = 0; k < deg; ++k) {
      const int c = 3 * node_col[s0 + k];
      A[r * n + c] = vp[3 * k];
      A[r * n + c + 1] = vp[3 * k + 1];
      A[r * n + c + 2] = vp[3 * k + 2];
    }
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) col[i] = A[i * n + k];
    if (threadIdx.x == 0) piv_inv = 1.0 / A[k * n + k];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x)
      A[k * n + j] = ((j == k) ? 1.0 : A[k * n + j]) * piv_inv;
    __syncthreads();
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
      const int i = e / n, j = e - i * n;
      if (i == k) continue;
      const double old = (j == k) ? 0.0 : A[e];
      A[e] = old - col[i] * A[k * n + j];
    }
    __syncthreads();
  }
  // symmetrised copy (the exact inverse is symmetric; rounding is not)
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    inv[e] = 0.5 * (A[e] + A[j * n + i]);
  }
}

// x = Ainv b ; one warp per row
__global__ void __launch_bounds__(kBlock)
    mg_dense_apply_kernel(int n, const double *__restrict__ Ainv,
                          const double *__restrict__ b, double *__restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  if (r >= n) return;
  double a = 0.0;
  for (int j = lane; j < n; j += 32) a += Ainv[(int64_t)r * n + j] * b[j];
  a = warp_sum(a);
  if (lane == 0) x[r] = a;
}

// factor the coarsest level (call after its values are set); levels too large
// for the dense path keep the damped-Jacobi sweeps
extern "C" int sktb_mg_factor_coarsest(sktb_mg *m, void *stream) {
  SKTB_REQUIRE(m && m->lv.size() >= 2, "bad argument");
  MgLevel &l = m->lv.back();
  SKTB_REQUIRE(l.node_ptr && l.vals, "coarsest level not set");
  const int n = (int)(3 * l.n_nodes);
  if (n > kDenseMax) {
    l.dense_n = 0;
    return 0;
  }
  SKTB_REQUIRE(!l.dense_shared, "the coarsest level borrows another hierarchy's inverse");
  SKTB_CUDA_OK(cudaSetDevice(m->device));
  if (!l.dense_inv)
    SKTB_CUDA_OK(cudaMalloc(&l.dense_inv, sizeof(double) * kDenseMax * kDenseMax));
  const size_t smem = sizeof(double) * ((size_t)n * n + n);
  static bool attr_set = false;
  if (!attr_set) {
    SKTB_CUDA_OK(cudaFuncSetAttribute(mg_dense_invert_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(sizeof(double) * (kDenseMax * kDenseMax + kDenseMax))));
    attr_set = true;
  }
  mg_dense_invert_kernel<<<1, 1024, smem, (cudaStream_t)stream>>>(
      (int)l.n_nodes, l.node_ptr, l.node_col, l.vals, l.dense_inv);
  SKTB_KERNEL_OK();
  l.dense_n = n;
  return 0;
}

// A second V-cycle workspace on the same level operators (concurrent load cases:
// one hierarchy of work vectors per CUDA stream) borrows the exact inverse of the
// coarsest level instead of factoring it again.  Call after sktb_mg_set_level of
// dst's coarsest level and after src was factored; the caller orders the streams.
extern "C" int sktb_mg_share_coarsest(sktb_mg *dst, const sktb_mg *src) {
  SKTB_REQUIRE(dst && src && dst != src && dst->lv.size() == src->lv.size() &&
                   dst->lv.size() >= 2,
               "bad argument");
  MgLevel &d = dst->lv.back();
  const MgLevel &c = src->lv.back();
  SKTB_REQUIRE(d.n_nodes == c.n_nodes && d.vals == c.vals, "the coarsest levels differ");
  SKTB_REQUIRE(d.dense_shared || !d.dense_inv, "dst already owns an inverse");
  d.dense_inv = c.dense_inv;
  d.dense_n = c.dense_n;
  d.dense_shared = true;
  return 0;
}

// ------------------------------------------------ fused coarse tail kernel --
// The levels below a few thousand nodes are launch-latency bound (every kernel
// is 3-5 us of ramp-up for < 1 us of work) and they need the most smoothing
// sweeps.  One cooperative kernel walks all of them down and up with grid-wide
// barriers in place of launches: sweeps are SpMV + Jacobi fused through a
// ping-pong pair of iterates (one barrier per sweep).
namespace cg = cooperative_groups;

struct TailLevel {
  int n_nodes;
  const int32_t *node_ptr, *node_col;
  const double *vals, *dinv;
  const uint8_t *mask;
  double *x, *x2, *b, *tmp;
  double omega;
  int nu;
  int fnp[3], cnp[3];
  const int32_t *c0, *c1;
  const double *w0, *w1;
  const int32_t *fT;
  const double *wT;
  const double *dense_inv;
  int dense_n;
};
struct TailParams {
  int n_levels;
  TailLevel lv[kTailMaxLevels];
};

// row r of A x by one warp: lanes stride the row's 3 deg entries (coalesced
// values), all loads of the row in flight together
__device__ __forceinline__ double tail_row(const TailLevel &L, int r, int lane,
                                           const double *__restrict__ x) {
  const int nd = r / 3, ri = r - 3 * nd;
  const int32_t s0 = L.node_ptr[nd], deg = L.node_ptr[nd + 1] - s0;
  const double *vp = L.vals + (int64_t)9 * s0 + (int64_t)ri * 3 * deg;
  double acc = 0.0;
#pragma unroll 3
  for (int e = lane; e < 3 * deg; e += 32) {
    const int k = e / 3, j = e - 3 * k;
    acc += vp[e] * x[3 * L.node_col[s0 + k] + j];
  }
  return warp_sum(acc);
}

// dst = src + omega D^-1 (b - A src)   (or tmp = A src when residual_only)
__device__ __forceinline__ void tail_sweep(const TailLevel &L, const double *src, double *dst,
                                           bool residual_only, int gtid, int gsize) {
  const int n = 3 * L.n_nodes;
  const int lane = threadIdx.x & 31;
  for (int r = gtid >> 5; r < n; r += gsize >> 5) {
    const double ax = tail_row(L, r, lane, src);
    if (lane == 0) {
      if (residual_only)
        dst[r] = ax;
      else
        dst[r] = src[r] + L.omega * L.dinv[r] * (L.b[r] - ax);
    }
  }
}

__global__ void __launch_bounds__(kBlock)
    mg_tail_kernel(const __grid_constant__ TailParams P) {
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsize = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const int gwarp = gtid >> 5, nwarp = gsize >> 5;
  const int NL = P.n_levels;
  // ---- down
  for (int k = 0; k < NL; ++k) {
    const TailLevel &L = P.lv[k];
    const int n = 3 * L.n_nodes;
    if (k == NL - 1) {
      if (L.dense_n == n) {
        for (int r = gwarp; r < n; r += nwarp) {
          double a = 0.0;
          for (int j = lane; j < n; j += 32) a += L.dense_inv[(int64_t)r * n + j] * L.b[j];
          a = warp_sum(a);
          if (lane == 0) L.x[r] = a;
        }
      } else {  // no exact solve available: nu sweeps from zero
        double *cur = (L.nu & 1) ? L.x : L.x2, *nxt = (L.nu & 1) ? L.x2 : L.x;
        for (int i = gtid; i < n; i += gsize) cur[i] = L.omega * L.dinv[i] * L.b[i];
        grid.sync();
        for (int s = 1; s < L.nu; ++s) {
          tail_sweep(L, cur, nxt, false, gtid, gsize);
          grid.sync();
          double *t = cur;
          cur = nxt;
          nxt = t;
        }
      }
      grid.sync();
      break;
    }
    // the level's iterate ends in L.x after (nu - 1) + nu swaps: start in x2
    double *cur = L.x2, *nxt = L.x;
    for (int i = gtid; i < n; i += gsize) cur[i] = L.omega * L.dinv[i] * L.b[i];
    grid.sync();
    for (int s = 1; s < L.nu; ++s) {
      tail_sweep(L, cur, nxt, false, gtid, gsize);
      grid.sync();
      double *t = cur;
      cur = nxt;
      nxt = t;
    }
    tail_sweep(L, cur, L.tmp, true, gtid, gsize);
    grid.sync();
    // restrict b - A x into the next level (warp per coarse node, 27 lanes)
    const TailLevel &Cn = P.lv[k + 1];
    {
      const int cnx = L.cnp[0], cny = L.cnp[1], cnz = L.cnp[2];
      const int fnx = L.fnp[0], fny = L.fnp[1];
      const int tot = cnx + cny + cnz;
      const int sy = lane % 3, sx = (lane / 3) % 3, sz = lane / 9;
      for (int I = gwarp; I < Cn.n_nodes; I += nwarp) {
        const int iy = I % cny, ix = (I / cny) % cnx, iz = I / (cny * cnx);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        if (lane < 27) {
          const int fz = L.fT[sz * tot + cnx + cny + iz];
          const int fx = L.fT[sx * tot + ix];
          const int fy = L.fT[sy * tot + cnx + iy];
          if (fz >= 0 && fx >= 0 && fy >= 0) {
            const double w = L.wT[sz * tot + cnx + cny + iz] * L.wT[sx * tot + ix] *
                             L.wT[sy * tot + cnx + iy];
            const int f = 3 * (fy + fny * (fx + fnx * fz));
            a0 = w * (L.b[f] - L.tmp[f]);
            a1 = w * (L.b[f + 1] - L.tmp[f + 1]);
            a2 = w * (L.b[f + 2] - L.tmp[f + 2]);
          }
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        a2 = warp_sum(a2);
        if (lane == 0) {
          const int o = 3 * I;
          Cn.b[o] = (Cn.mask && Cn.mask[o]) ? 0.0 : a0;
          Cn.b[o + 1] = (Cn.mask && Cn.mask[o + 1]) ? 0.0 : a1;
          Cn.b[o + 2] = (Cn.mask && Cn.mask[o + 2]) ? 0.0 : a2;
        }
      }
    }
    grid.sync();
  }
  // ---- up
  for (int k = NL - 2; k >= 0; --k) {
    const TailLevel &L = P.lv[k];
    const TailLevel &Cn = P.lv[k + 1];
    const int n = 3 * L.n_nodes;
    // where the down sweep left this level's iterate
    double *cur = ((L.nu - 1) & 1) ? L.x : L.x2;
    double *nxt = ((L.nu - 1) & 1) ? L.x2 : L.x;
    {
      const int cnx = L.cnp[0], cny = L.cnp[1];
      const int fnx = L.fnp[0], fny = L.fnp[1];
      for (int F = gtid; F < L.n_nodes; F += gsize) {
        const int iy = F % fny, ix = (F / fny) % fnx, iz = F / (fny * fnx);
        const int cx[2] = {L.c0[ix], L.c1[ix]};
        const double wx[2] = {L.w0[ix], L.w1[ix]};
        const int cy[2] = {L.c0[fnx + iy], L.c1[fnx + iy]};
        const double wy[2] = {L.w0[fnx + iy], L.w1[fnx + iy]};
        const int cz[2] = {L.c0[fnx + fny + iz], L.c1[fnx + fny + iz]};
        const double wz[2] = {L.w0[fnx + fny + iz], L.w1[fnx + fny + iz]};
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int kz = 0; kz < 2; ++kz) {
          if (wz[kz] == 0.0) continue;
#pragma unroll
          for (int kx = 0; kx < 2; ++kx) {
            if (wx[kx] == 0.0) continue;
#pragma unroll
            for (int ky = 0; ky < 2; ++ky) {
              if (wy[ky] == 0.0) continue;
              const double w = wz[kz] * wx[kx] * wy[ky];
              const int c = 3 * (cy[ky] + cny * (cx[kx] + cnx * cz[kz]));
              a0 += w * Cn.x[c];
              a1 += w * Cn.x[c + 1];
              a2 += w * Cn.x[c + 2];
            }
          }
        }
        const int o = 3 * F;
        if (!(L.mask && L.mask[o])) cur[o] += a0;
        if (!(L.mask && L.mask[o + 1])) cur[o + 1] += a1;
        if (!(L.mask && L.mask[o + 2])) cur[o + 2] += a2;
      }
    }
    grid.sync();
    for (int s = 0; s < L.nu; ++s) {
      tail_sweep(L, cur, nxt, false, gtid, gsize);
      grid.sync();
      double *t = cur;
      cur = nxt;
      nxt = t;
    }
    (void)n;
  }
}

static int level_spmv(const MgLevel &l, const double *x, double *y, cudaStream_t st,
                      bool fp32 = false) {
  if (l.gop) {
    if (fp32) {
      int rc = launch_hexgrid_apply_ex(l.gop, l.node0, l.n_nodes, x, y, true, nullptr,
                                       nullptr, 0.0, st);
      if (rc != -1) return rc;
    }
    return launch_hexgrid_apply(l.gop, l.node0, l.n_nodes, x, y, nullptr, nullptr,
                                nullptr, nullptr, st);
  }
  int rc = -1;
  if (l.n_nodes >= kTmaMinNodes && l.vals32)
    rc = launch_spmv_bsr3_tma_f32(l.n_nodes, l.n_blocks, l.max_deg, l.node_ptr, l.node_col,
                                  l.vals32, x, y, JacobiEpi(), st);
  if (rc == -1 && l.n_nodes >= kTmaMinNodes)
    rc = launch_spmv_bsr3_tma(l.n_nodes, l.n_blocks, l.max_deg, l.node_ptr, l.node_col, l.vals,
                              x, y, nullptr, nullptr, nullptr, nullptr, st);
  if (rc != -1) return rc;
  return launch_spmv_bsr3(l.n_nodes, l.node_ptr, l.node_col, l.vals, x, y,
                          nullptr, nullptr, nullptr, nullptr, st);
}

// One damped-Jacobi sweep x <- x + omega D^-1 (b - A x) of an assembled level
// (k >= 1) as ONE kernel: the product writes the new iterate into the level's
// second buffer through the fused epilogue, then the two buffers swap roles
// (replicated level: x <-> tmp; z-slab-sharded level: the two full-length
// iterates x <-> x2, ghost planes refreshed by the caller before the sweep).
// Returns -1 when the level has no such kernel (caller: product + update).
static int level_sweep(MgLevel &l, const double *b, double omega, cudaStream_t st) {
  if (l.gop || !l.vals) return -1;
  if (!l.sharded && (l.node0 != 0 || l.n_global != l.n_nodes)) return -1;
  if (l.sharded && !l.x2) return -1;
  JacobiEpi epi;
  epi.b = b;
  epi.dinv = l.inv_diag;
  epi.omega = omega;
  double *y = l.tmp;
  if (l.sharded) {
    epi.xo = l.x + 3 * l.node0;
    y = l.x2 + 3 * l.node0;
  }
  int rc = -1;
  if (l.n_nodes >= kTmaMinNodes && l.vals32)
    rc = launch_spmv_bsr3_tma_f32(l.n_nodes, l.n_blocks, l.max_deg, l.node_ptr, l.node_col,
                                  l.vals32, l.x, y, epi, st);
  if (rc == -1 && l.n_nodes >= kTmaMinNodes)
    rc = launch_spmv_bsr3_tma_jacobi(l.n_nodes, l.n_blocks, l.max_deg, l.node_ptr, l.node_col,
                                     l.vals, l.x, y, epi, st);
  if (rc == -1)
    rc = launch_spmv_bsr3_jacobi(l.n_nodes, l.node_ptr, l.node_col, l.vals, l.x, y, epi, st);
  if (rc) return rc;
  if (l.sharded)
    std::swap(l.x, l.x2);
  else
    std::swap(l.x, l.tmp);
  return 0;
}

// first level of the fused tail (L: none)
static int tail_start(const sktb_mg *m) {
  const int L = (int)m->lv.size();
  if (!m->fused_tail) return L;
  int k0 = L;
  for (int k = L - 1; k >= 1; --k) {
    const MgLevel &l = m->lv[k];
    if (l.n_nodes > kTailMaxNodes || !l.x2 || !l.vals || l.cheb || l.sharded ||
        L - k > kTailMaxLevels)
      break;
    k0 = k;
  }
  return k0;
}

static int launch_tail(sktb_mg *m, int k0, cudaStream_t st) {
  const int L = (int)m->lv.size();
  TailParams P;
  P.n_levels = L - k0;
  for (int k = k0; k < L; ++k) {
    const MgLevel &l = m->lv[k];
    TailLevel &t = P.lv[k - k0];
    t.n_nodes = (int)l.n_nodes;
    t.node_ptr = l.node_ptr;
    t.node_col = l.node_col;
    t.vals = l.vals;
    t.dinv = l.inv_diag;
    t.mask = l.mask;
    t.x = l.x;
    t.x2 = l.x2;
    t.b = l.b;
    t.tmp = l.tmp;
    t.omega = l.omega > 0.0 ? l.omega : m->omega;
    t.nu = (k == L - 1) ? (m->nu_coarse > 0 ? m->nu_coarse : 1) : l.nu;
    for (int a = 0; a < 3; ++a) {
      t.fnp[a] = l.fnp[a];
      t.cnp[a] = l.cnp[a];
    }
    t.c0 = l.ax_c0;
    t.c1 = l.ax_c1;
    t.w0 = l.ax_w0;
    t.w1 = l.ax_w1;
    t.fT = l.axT_f;
    t.wT = l.axT_w;
    t.dense_inv = l.dense_inv;
    t.dense_n = l.dense_n;
  }
  void *args[] = {(void *)&P};
  SKTB_CUDA_OK(cudaLaunchCooperativeKernel((void *)mg_tail_kernel, dim3(kTailGrid),
                                           dim3(kBlock), args, 0, st));
  SKTB_COUNT(1);
  return 0;
}

// ghost planes of a sharded level's full-length vector (no-op when replicated)
static int level_halo(const MgLevel &l, double *vfull, sktb_pcg *dist, cudaStream_t st) {
  if (!dist || !pcg_is_dist(dist)) return 0;
  if (l.sharded && l.plane > 0)
    return slab_halo_exchange(pcg_comm(dist), vfull, 3 * l.node0, 3 * l.n_nodes, 3 * l.plane,
                              l.prev, l.next, st);
  if (l.node0 != 0 || l.n_global != l.n_nodes)  // index-list halo of the PCG (level 0)
    return pcg_halo_exchange(dist, vfull, st);
  return 0;
}

// b_c = mask_c P^T (b_f - A x_f) for the coarse level c of fine level l
// x0_done (may be null): set when the kernel also wrote the coarse level's first
// Jacobi step x_c = om_c D_c^-1 b_c (one launch less per level and cycle); only for a
// replicated coarse level whose right-hand side is final after this kernel
static int level_restrict(MgLevel &l, MgLevel &c, const double *b, sktb_pcg *dist,
                          cudaStream_t st, bool *x0_done = nullptr, double om_c = 0.0) {
  const bool part = l.node0 != 0 || l.n_global != l.n_nodes;  // fine rows are a slab
  const double *bf = b, *Axf = l.tmp;
  int64_t f_lo = part ? l.node0 : 0, f_hi = part ? l.node0 + l.n_nodes : l.n_nodes;
  int64_t c_lo = 0, c_hi = c.n_nodes, c_base = 0;
  if (c.sharded) {
    // sharded -> sharded: the owned coarse rows need the fine residual on one
    // ghost plane each side: r = b - Ax into the full-length buffer, exchange
    if (!l.res) {
      if (dev_alloc_exchangeable(&l.res, (size_t)3 * l.n_global)) return 1;
    }
    mg_residual_kernel<<<grid_for(3 * l.n_nodes), kBlock, 0, st>>>(
        3 * l.n_nodes, b, l.tmp, l.res + 3 * l.node0);
    SKTB_COUNT(1);
    if (level_halo(l, l.res, dist, st)) return 1;
    bf = l.res;
    Axf = nullptr;
    f_lo = 0;
    f_hi = l.n_global;
    c_lo = c.node0;
    c_hi = c.node0 + c.n_nodes;
    c_base = c.node0;
  }
  const int64_t nc = c_hi - c_lo;
  const bool reduce_after = part && !c.sharded && dist;
  const bool fuse = x0_done && !c.sharded && !reduce_after && !c.cheb && c.inv_diag && c.x &&
                    c.dense_n != 3 * c.n_nodes && om_c > 0.0;
  const double *jd = fuse ? c.inv_diag : nullptr;
  double *jx = fuse ? c.x : nullptr;
  if (nc > 20000)
    mg_restrict_thread_kernel<<<grid_for(nc), kBlock, 0, st>>>(
        l.cnp[0], l.cnp[1], l.cnp[2], l.fnp[0], l.fnp[1], l.fnp[2], l.axT_f, l.axT_w, bf, Axf,
        c.mask, c.b, f_lo, f_hi, c_lo, c_hi, c_base, jd, om_c, jx);
  else
    mg_restrict_kernel<<<grid_for(nc * 32, kBlock, 16), kBlock, 0, st>>>(
        l.cnp[0], l.cnp[1], l.cnp[2], l.fnp[0], l.fnp[1], l.fnp[2], l.axT_f, l.axT_w, bf, Axf,
        c.mask, c.b, f_lo, f_hi, c_lo, c_hi, c_base, jd, om_c, jx);
  SKTB_COUNT(1);
  if (x0_done) *x0_done = fuse;
  // sharded fine level, replicated coarse level: every rank summed its own fine
  // rows into the whole coarse vector
  if (part && !c.sharded && dist) {
    phase_mark(l.gop || l.node0 >= 0 ? (l.sharded && l.vals ? PH_VC_L1 : PH_VC_L0) : PH_VC_L0, st);
    if (pcg_allreduce_vec(dist, c.b, 3 * c.n_nodes, st)) return 1;
    phase_mark(PH_VC_TRANS, st);
  }
  return 0;
}

// z = M^-1 r : V cycle.  Level 0 and the large assembled levels may be z-slab
// sharded (their x is a full-length vector whose ghost planes are refreshed
// before every product); the remaining coarse levels are replicated.
int mg_vcycle(sktb_mg *m, const double *r, double *z, cudaStream_t st,
              sktb_pcg *dist) {
  const int L = (int)m->lv.size();
  const int k_tail = tail_start(m);
  MgLevel &l0 = m->lv[0];
  bool x0_done = false;   // the restriction into level k already wrote its first Jacobi step
  // damping of level k + 1 when the restriction into it may do so (0: it may not)
  auto om_next = [&](int k) {
    if (k + 1 >= L || k + 1 == k_tail || !m->fuse_restrict_jacobi) return 0.0;
    const MgLevel &c = m->lv[k + 1];
    return c.omega > 0.0 ? c.omega : m->omega;
  };
  // downward sweep
  for (int k = 0; k < L; ++k) {
    if (k == k_tail) {
      if (launch_tail(m, k_tail, st)) return 1;
      break;
    }
    MgLevel &l = m->lv[k];
    const int64_t n = 3 * l.n_nodes;
    const double *b = (k == 0) ? r : l.b;
    const int g = grid_for(n);
    const double om = l.omega > 0.0 ? l.omega : m->omega;
#define XOWN (l.x + 3 * l.node0)
    if (k == L - 1 && k > 0 && l.dense_n == n) {
      mg_dense_apply_kernel<<<(int)((n + kBlock / 32 - 1) / (kBlock / 32)), kBlock, 0, st>>>(
          (int)n, l.dense_inv, b, XOWN);
      SKTB_COUNT(1);
      break;
    }
    if (l.cheb && k > 0 && k < L - 1 && !l.sharded) {
      // Chebyshev pre-smoothing of degree nu from a zero initial guess
      mg_cheby_kernel<<<g, kBlock, 0, st>>>(n, 0.0, l.cheb_c2[0], l.inv_diag, b, nullptr, 1,
                                           l.d, l.x);
      SKTB_COUNT(1);
      for (int s = 1; s < l.nu; ++s) {
        if (level_spmv(l, l.x, l.tmp, st)) return 1;
        mg_cheby_kernel<<<g, kBlock, 0, st>>>(n, l.cheb_c1[s], l.cheb_c2[s], l.inv_diag, b,
                                             l.tmp, 0, l.d, l.x);
        SKTB_COUNT(1);
      }
      if (level_spmv(l, l.x, l.tmp, st)) return 1;
      if (level_restrict(l, m->lv[k + 1], b, dist, st, &x0_done, om_next(k))) return 1;
      continue;
    }
    if (!x0_done) {
      mg_jacobi0_kernel<<<g, kBlock, 0, st>>>(n, om, l.inv_diag, b, XOWN);
      SKTB_COUNT(1);
    }
    x0_done = false;
    if (k == L - 1 && k > 0 && l.n_nodes <= 4096) {
      // (jacobi0 above is redone inside; harmless and keeps the code uniform)
      mg_coarse_solve_kernel<<<1, 1024, 0, st>>>((int)l.n_nodes, l.node_ptr,
                                                 l.node_col, l.vals, l.inv_diag,
                                                 b, om, m->nu_coarse, l.x, l.tmp);
      SKTB_COUNT(1);
    } else if (k == L - 1) {
      for (int s = 0; s < m->nu_coarse; ++s) {
        if (level_halo(l, l.x, dist, st)) return 1;
        if (level_spmv(l, l.x, l.tmp, st)) return 1;
        mg_jacobi_kernel<<<g, kBlock, 0, st>>>(n, om, l.inv_diag, b, l.tmp, XOWN);
        SKTB_COUNT(1);
      }
    } else {
      for (int s = 1; s < l.nu; ++s) {  // extra pre-smoothing sweeps
        if (level_halo(l, l.x, dist, st)) return 1;
        if (m->fused_sweeps && (k > 0 || !l.gop)) {
          const int rc = level_sweep(l, b, om, st);
          if (rc == 0) continue;
          if (rc != -1) return rc;
        }
        if (level_spmv(l, l.x, l.tmp, st)) return 1;
        mg_jacobi_kernel<<<g, kBlock, 0, st>>>(n, om, l.inv_diag, b, l.tmp, XOWN);
        SKTB_COUNT(1);
      }
      if (level_halo(l, l.x, dist, st)) return 1;
      if (level_spmv(l, l.x, l.tmp, st, k == 0 && m->fp32_level0)) return 1;
      if (level_restrict(l, m->lv[k + 1], b, dist, st, &x0_done, om_next(k))) return 1;
      phase_mark(k == 0 ? PH_VC_L0 : (l.sharded ? PH_VC_L1 : PH_VC_COARSE), st);
    }
  }
  phase_mark(PH_VC_COARSE, st);
  // upward sweep
  bool z_done = false;
  for (int k = (k_tail < L ? k_tail - 1 : L - 2); k >= 0; --k) {
    MgLevel &l = m->lv[k];
    MgLevel &c = m->lv[k + 1];
    const int64_t n = 3 * l.n_nodes;
    const double *b = (k == 0) ? r : l.b;
    const int64_t lo = l.node0, hi = l.node0 + l.n_nodes;
    const double om = l.omega > 0.0 ? l.omega : m->omega;
    if (level_halo(c, c.x, dist, st)) return 1;  // ghost planes of a sharded coarse level
    mg_prolong_kernel<<<grid_for(hi - lo), kBlock, 0, st>>>(
        l.cnp[0], l.cnp[1], l.cnp[2], l.fnp[0], l.fnp[1], l.fnp[2], l.ax_c0,
        l.ax_c1, l.ax_w0, l.ax_w1, c.x, l.mask, l.x, lo, hi);
    SKTB_COUNT(1);
    if (l.cheb && k > 0 && !l.sharded) {
      for (int s = 0; s < l.nu; ++s) {
        if (level_spmv(l, l.x, l.tmp, st)) return 1;
        mg_cheby_kernel<<<grid_for(n), kBlock, 0, st>>>(n, s ? l.cheb_c1[s] : 0.0,
                                                       l.cheb_c2[s], l.inv_diag, b, l.tmp, 0,
                                                       l.d, l.x);
        SKTB_COUNT(1);
      }
      continue;
    }
    for (int s = 1; s < l.nu; ++s) {  // extra post-smoothing sweeps
      if (level_halo(l, l.x, dist, st)) return 1;
      if (m->fused_sweeps && (k > 0 || !l.gop)) {
        const int rc = level_sweep(l, b, om, st);
        if (rc == 0) continue;
        if (rc != -1) return rc;
      }
      if (level_spmv(l, l.x, l.tmp, st)) return 1;
      mg_jacobi_kernel<<<grid_for(n), kBlock, 0, st>>>(n, om, l.inv_diag, b, l.tmp, XOWN);
      SKTB_COUNT(1);
    }
    if (level_halo(l, l.x, dist, st)) return 1;
    if (m->fused_sweeps && (k > 0 || !l.gop)) {  // last post-smoothing sweep
      const int rc = level_sweep(l, b, om, st);
      if (rc == 0) {
        phase_mark(l.sharded ? PH_VC_UP1 : PH_VC_COARSE, st);
        continue;
      }
      if (rc != -1) return rc;
    }
    if (k == 0 && l.gop) {
      // fused post-smoothing straight into z: z = x + om D^-1 (r - A x)
      int rc = launch_hexgrid_apply_ex(l.gop, l.node0, l.n_nodes, l.x, z, m->fp32_level0,
                                       b, l.inv_diag, om, st);
      if (rc == 0) {
        z_done = true;
        phase_mark(PH_VC_UP0, st);
        continue;
      }
      if (rc != -1) return rc;
    }
    if (level_spmv(l, l.x, l.tmp, st)) return 1;
    mg_jacobi_kernel<<<grid_for(n), kBlock, 0, st>>>(n, om, l.inv_diag, b, l.tmp, XOWN);
    SKTB_COUNT(1);
  }
#undef XOWN
  if (!z_done)
    SKTB_CUDA_OK(cudaMemcpyAsync(z, l0.x + 3 * l0.node0, sizeof(double) * 3 * l0.n_nodes,
                                 cudaMemcpyDeviceToDevice, st));
  SKTB_KERNEL_CHECK();
  return 0;
}

extern "C" int sktb_mg_vcycle(sktb_mg *m, const double *r, double *z, void *stream) {
  SKTB_REQUIRE(m && r && z, "null argument");
  for (auto &l : m->lv) SKTB_REQUIRE(l.node_ptr || l.gop, "multigrid level not set");
  return mg_vcycle(m, r, z, (cudaStream_t)stream, nullptr);
}

// y[owned rows] = A_level x for a full-length x (ghost planes of a sharded
// level are refreshed first; `dist` may be null on one GPU).  Used by the
// host-driven power iteration that sets the per-level damping.
extern "C" int sktb_mg_level_apply(sktb_mg *m, int level, sktb_pcg *dist, double *x_full,
                                   double *y_own, void *stream) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size() && x_full && y_own, "bad argument");
  MgLevel &l = m->lv[level];
  SKTB_REQUIRE(l.node_ptr || l.gop, "multigrid level not set");
  cudaStream_t st = (cudaStream_t)stream;
  if (level_halo(l, x_full, dist, st)) return 1;
  return level_spmv(l, x_full, y_own, st);
}
