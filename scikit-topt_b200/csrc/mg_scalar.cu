// Geometric multigrid + PCG for SCALAR operators on tensor-product hexahedral
// grids (heat conduction with Robin terms, reference fea/solver_heat.py:136-253
// where the reference factorises K with a sparse LU; any 27-point operator).
//
// Storage: "stencil-diagonal" (DIA) format.  On a tensor grid every row of a
// Q1 operator couples the 27 nodes (dz, dx, dy) in {-1,0,1}^3, so the matrix is
// 27 arrays of n doubles, vals[k][i], k = 9 (dz+1) + 3 (dx+1) + (dy+1), with
// node i = iy + npy (ix + npx iz).  No column indices, every load coalesced:
// 216 B per row and product (CSR: 324 B), entries outside the grid are 0.
//
//  * level 0 is converted from the caller's enforced CSR matrix (whatever
//    terms it holds: conduction, real and virtual Robin) every set-up;
//  * coarse operators are ALGEBRAIC Galerkin products A_c = P^T A_f P with
//    trilinear P, formed stencil to stencil by one kernel per level (one thread
//    per coarse row, 27 accumulators in shared memory), so they follow the
//    fine operator exactly, boundary terms included;
//  * V-cycle: damped Jacobi (fused into the product: one kernel per sweep),
//    residual + restriction, exact dense solve on the coarsest level;
//  * fixed (Dirichlet) nodes are identity rows on every level (a coarse node
//    is fixed iff the coincident fine node is), P has zero rows / columns there.
#include <vector>

#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

namespace {

constexpr int kDenseMaxS = 160;  // coarsest level: dense inverse up to this many nodes

struct SLevel {
  int np[3] = {0, 0, 0};  // nodes per axis (x, y, z)
  int64_t n = 0;
  double *vals = nullptr;   // [27][n] (owned)
  double *dinv = nullptr;   // [n] (owned)
  const uint8_t *mask = nullptr;  // [n] fixed nodes, caller-owned, may be null
  double *x = nullptr, *x2 = nullptr, *b = nullptr, *tmp = nullptr;  // owned
  double omega = 0.6;
  int nu = 1;
  // transfer to the next coarser level (device tables, caller-owned; same
  // layout as sktb_mg_set_transfer)
  int fnp[3] = {0, 0, 0}, cnp[3] = {0, 0, 0};
  const int32_t *c0 = nullptr, *c1 = nullptr;
  const double *w0 = nullptr, *w1 = nullptr;
  const int32_t *fT = nullptr;
  const double *wT = nullptr;
  double *dense_inv = nullptr;  // coarsest level (owned)
  int dense_n = 0;
};

}  // namespace

struct sktb_smg {
  int device = 0;
  std::vector<SLevel> lv;
  bool omega_set = false;
  double *scal = nullptr;    // device scalar for the power iteration
  double *scal_h = nullptr;  // pinned
  double *partials = nullptr;
  unsigned int *ticket = nullptr;
  int nu_default[8] = {1, 1, 2, 2, 3, 3, 3, 3};
};

#define GS(i, n)                                                       \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x,     \
               _st = (int64_t)gridDim.x * blockDim.x;                  \
       i < (n); i += _st)

// ------------------------------------------------------------------ kernels --
// CSR (sorted or not) -> DIA; every column must lie in the 27-neighbourhood
__global__ void __launch_bounds__(kBlock)
    csr_to_dia_kernel(int64_t n, int npx, int npy, const int32_t *__restrict__ rp,
                      const int32_t *__restrict__ ci, const double *__restrict__ v,
                      double *__restrict__ dia, int *bad) {
  GS(i, n) {
#pragma unroll
    for (int k = 0; k < 27; ++k) dia[(int64_t)k * n + i] = 0.0;
    const int iy = (int)(i % npy), ix = (int)((i / npy) % npx), iz = (int)(i / ((int64_t)npy * npx));
    for (int32_t e = rp[i]; e < rp[i + 1]; ++e) {
      const int64_t c = ci[e];
      const int dy = (int)(c % npy) - iy, dx = (int)((c / npy) % npx) - ix,
                dz = (int)(c / ((int64_t)npy * npx)) - iz;
      if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || dz < -1 || dz > 1) {
        *bad = 1;
        continue;
      }
      dia[(int64_t)(9 * (dz + 1) + 3 * (dx + 1) + (dy + 1)) * n + i] = v[e];
    }
  }
}

// MODE 0: y = A x ; 1: y = x + omega dinv (b - A x) ; 2: y = b - A x.
// DOT: also publishes sum_i dotv[i] y[i] (PCG: p.Ap).
template <int MODE, bool DOT>
__global__ void __launch_bounds__(kBlock)
    dia_apply_kernel(int64_t n, int npx, int npy, const double *__restrict__ A,
                     const double *__restrict__ x, double *__restrict__ y,
                     const double *__restrict__ b, const double *__restrict__ dinv,
                     double omega, const double *__restrict__ dotv, double *partials,
                     unsigned int *ticket, double *dot_out, const PcgScalars *S) {
  if (S && S->rr <= S->tol2) return;
  const int64_t plane = (int64_t)npx * npy;
  double dot = 0.0;
  GS(i, n) {
    double acc = 0.0;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int64_t base = i + dz * plane + (int64_t)dx * npy;
        const int k = 9 * (dz + 1) + 3 * (dx + 1);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          int64_t j = base + dy;
          j = j < 0 ? 0 : (j >= n ? n - 1 : j);  // coefficient is 0 outside the grid
          acc = fma(__ldcs(&A[(int64_t)(k + dy + 1) * n + i]), __ldg(&x[j]), acc);
        }
      }
    double out = acc;
    if (MODE == 1) out = x[i] + omega * dinv[i] * (b[i] - acc);
    if (MODE == 2) out = b[i] - acc;
    y[i] = out;
    if (DOT) dot += dotv[i] * out;
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

__global__ void __launch_bounds__(kBlock)
    dia_inv_diag_kernel(int64_t n, const double *__restrict__ A, double *__restrict__ dinv) {
  GS(i, n) dinv[i] = 1.0 / A[(int64_t)13 * n + i];
}

__global__ void __launch_bounds__(kBlock)
    s_jacobi0_kernel(int64_t n, double omega, const double *__restrict__ dinv,
                     const double *__restrict__ b, double *__restrict__ x) {
  GS(i, n) x[i] = omega * dinv[i] * b[i];
}

// A_c = P^T A_f P, stencil to stencil.  One thread per coarse node I; the 27
// accumulators live in shared memory ([27][blockDim], conflict free).
constexpr int kGalBlock = 128;
__global__ void __launch_bounds__(kGalBlock)
    galerkin_dia_kernel(int cnx, int cny, int cnz, int fnx, int fny, int fnz,
                        const int32_t *__restrict__ fT, const double *__restrict__ wT,
                        const int32_t *__restrict__ c0, const int32_t *__restrict__ c1,
                        const double *__restrict__ w0, const double *__restrict__ w1,
                        const double *__restrict__ Af, const uint8_t *__restrict__ mask_f,
                        const uint8_t *__restrict__ mask_c, double *__restrict__ Ac) {
  __shared__ double acc[27][kGalBlock];
  const int64_t nc = (int64_t)cnx * cny * cnz, nf = (int64_t)fnx * fny * fnz;
  const int tot = cnx + cny + cnz;
  const int64_t I = (int64_t)blockIdx.x * kGalBlock + threadIdx.x;
  if (I >= nc) return;
  const int t = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 27; ++k) acc[k][t] = 0.0;
  const int Iy = (int)(I % cny), Ix = (int)((I / cny) % cnx), Iz = (int)(I / ((int64_t)cny * cnx));
  if (mask_c && mask_c[I]) {
    for (int k = 0; k < 27; ++k) Ac[(int64_t)k * nc + I] = (k == 13) ? 1.0 : 0.0;
    return;
  }
  for (int sz = 0; sz < 3; ++sz) {
    const int fz = fT[sz * tot + cnx + cny + Iz];
    if (fz < 0) continue;
    const double wz = wT[sz * tot + cnx + cny + Iz];
    for (int sx = 0; sx < 3; ++sx) {
      const int fx = fT[sx * tot + Ix];
      if (fx < 0) continue;
      const double wzx = wz * wT[sx * tot + Ix];
      for (int sy = 0; sy < 3; ++sy) {
        const int fy = fT[sy * tot + cnx + Iy];
        if (fy < 0) continue;
        const int64_t i = fy + (int64_t)fny * (fx + (int64_t)fnx * fz);
        if (mask_f && mask_f[i]) continue;  // zero row of P
        const double wi = wzx * wT[sy * tot + cnx + Iy];
        // row i of A_f times P, scattered by coarse offset
        for (int dz = -1; dz <= 1; ++dz) {
          const int jz = fz + dz;
          if (jz < 0 || jz >= fnz) continue;
          const int pz[2] = {c0[fnx + fny + jz], c1[fnx + fny + jz]};
          const double vz[2] = {w0[fnx + fny + jz], w1[fnx + fny + jz]};
          for (int dx = -1; dx <= 1; ++dx) {
            const int jx = fx + dx;
            if (jx < 0 || jx >= fnx) continue;
            const int px[2] = {c0[jx], c1[jx]};
            const double vx[2] = {w0[jx], w1[jx]};
            for (int dy = -1; dy <= 1; ++dy) {
              const int jy = fy + dy;
              if (jy < 0 || jy >= fny) continue;
              const double a = Af[(int64_t)(9 * (dz + 1) + 3 * (dx + 1) + (dy + 1)) * nf + i];
              if (a == 0.0) continue;
              const int64_t j = jy + (int64_t)fny * (jx + (int64_t)fnx * jz);
              if (mask_f && mask_f[j]) continue;  // zero row of P (column side)
              const int py[2] = {c0[fnx + jy], c1[fnx + jy]};
              const double vy[2] = {w0[fnx + jy], w1[fnx + jy]};
              const double wa = wi * a;
#pragma unroll
              for (int kz = 0; kz < 2; ++kz) {
                if (vz[kz] == 0.0) continue;
#pragma unroll
                for (int kx = 0; kx < 2; ++kx) {
                  if (vx[kx] == 0.0) continue;
#pragma unroll
                  for (int ky = 0; ky < 2; ++ky) {
                    if (vy[ky] == 0.0) continue;
                    const int oz = pz[kz] - Iz, ox = px[kx] - Ix, oy = py[ky] - Iy;
                    // |o| <= 1 by construction of the nested grids
                    acc[9 * (oz + 1) + 3 * (ox + 1) + (oy + 1)][t] +=
                        wa * vz[kz] * vx[kx] * vy[ky];
                  }
                }
              }
            }
          }
        }
      }
    }
  }
  // fixed coarse columns drop out (zero column of P)
  for (int k = 0; k < 27; ++k) {
    double v = acc[k][t];
    if (mask_c && k != 13) {
      const int oz = k / 9 - 1, ox = (k / 3) % 3 - 1, oy = k % 3 - 1;
      const int Jz = Iz + oz, Jx = Ix + ox, Jy = Iy + oy;
      if (Jz >= 0 && Jz < cnz && Jx >= 0 && Jx < cnx && Jy >= 0 && Jy < cny &&
          mask_c[Jy + (int64_t)cny * (Jx + (int64_t)cnx * Jz)])
        v = 0.0;
    }
    Ac[(int64_t)k * nc + I] = v;
  }
}

// b_c = mask_c P^T r_f ; one thread per coarse node
__global__ void __launch_bounds__(kBlock)
    s_restrict_kernel(int cnx, int cny, int cnz, int fnx, int fny,
                      const int32_t *__restrict__ fT, const double *__restrict__ wT,
                      const double *__restrict__ rf, const uint8_t *__restrict__ mask_c,
                      double *__restrict__ bc) {
  const int64_t nc = (int64_t)cnx * cny * cnz;
  const int tot = cnx + cny + cnz;
  GS(I, nc) {
    const int Iy = (int)(I % cny), Ix = (int)((I / cny) % cnx),
              Iz = (int)(I / ((int64_t)cny * cnx));
    double a = 0.0;
    for (int sz = 0; sz < 3; ++sz) {
      const int fz = fT[sz * tot + cnx + cny + Iz];
      if (fz < 0) continue;
      const double wz = wT[sz * tot + cnx + cny + Iz];
      for (int sx = 0; sx < 3; ++sx) {
        const int fx = fT[sx * tot + Ix];
        if (fx < 0) continue;
        const double wzx = wz * wT[sx * tot + Ix];
        for (int sy = 0; sy < 3; ++sy) {
          const int fy = fT[sy * tot + cnx + Iy];
          if (fy < 0) continue;
          a += wzx * wT[sy * tot + cnx + Iy] * rf[fy + (int64_t)fny * (fx + (int64_t)fnx * fz)];
        }
      }
    }
    bc[I] = (mask_c && mask_c[I]) ? 0.0 : a;
  }
}

// x_f += mask_f P x_c ; one thread per fine node
__global__ void __launch_bounds__(kBlock)
    s_prolong_kernel(int cnx, int cny, int fnx, int fny, int fnz,
                     const int32_t *__restrict__ c0, const int32_t *__restrict__ c1,
                     const double *__restrict__ w0, const double *__restrict__ w1,
                     const double *__restrict__ xc, const uint8_t *__restrict__ mask_f,
                     double *__restrict__ xf) {
  const int64_t nf = (int64_t)fnx * fny * fnz;
  GS(F, nf) {
    if (mask_f && mask_f[F]) continue;
    const int iy = (int)(F % fny), ix = (int)((F / fny) % fnx),
              iz = (int)(F / ((int64_t)fny * fnx));
    const int cx[2] = {c0[ix], c1[ix]};
    const double wx[2] = {w0[ix], w1[ix]};
    const int cy[2] = {c0[fnx + iy], c1[fnx + iy]};
    const double wy[2] = {w0[fnx + iy], w1[fnx + iy]};
    const int cz[2] = {c0[fnx + fny + iz], c1[fnx + fny + iz]};
    const double wz[2] = {w0[fnx + fny + iz], w1[fnx + fny + iz]};
    double a = 0.0;
#pragma unroll
    for (int kz = 0; kz < 2; ++kz)
#pragma unroll
      for (int kx = 0; kx < 2; ++kx)
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
          const double w = wz[kz] * wx[kx] * wy[ky];
          if (w != 0.0) a += w * xc[cy[ky] + (int64_t)cny * (cx[kx] + (int64_t)cnx * cz[kz])];
        }
    xf[F] += a;
  }
}

// coarsest level: dense Gauss-Jordan inverse in shared memory (single CTA)
__global__ void __launch_bounds__(1024)
    s_dense_invert_kernel(int n, int npx, int npy, const double *__restrict__ dia,
                          double *__restrict__ inv) {
  extern __shared__ double sm[];
  double *A = sm, *col = sm + n * n;
  __shared__ double piv_inv;
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) A[e] = 0.0;
  __syncthreads();
  const int plane = npx * npy;
  for (int e = threadIdx.x; e < 27 * n; e += blockDim.x) {
    const int k = e / n, i = e - k * n;
    const double v = dia[(int64_t)k * n + i];
    if (v == 0.0) continue;
    const int j = i + (k / 9 - 1) * plane + ((k / 3) % 3 - 1) * npy + (k % 3 - 1);
    if (j >= 0 && j < n) A[i * n + j] = v;
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
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    inv[e] = 0.5 * (A[e] + A[j * n + i]);
  }
}

__global__ void __launch_bounds__(kBlock)
    s_dense_apply_kernel(int n, const double *__restrict__ Ainv, const double *__restrict__ b,
                         double *__restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
  if (r >= n) return;
  double a = 0.0;
  for (int j = lane; j < n; j += 32) a += Ainv[(int64_t)r * n + j] * b[j];
  a = warp_sum(a);
  if (lane == 0) x[r] = a;
}

__global__ void __launch_bounds__(kBlock)
    s_dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
                 double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  GS(i, n) v[0] += a[i] * b[i];
  grid_reduce<1>(v, partials, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
    s_scale_kernel(int64_t n, double a, const double *__restrict__ x,
                   const double *__restrict__ d, double *__restrict__ y) {
  GS(i, n) y[i] = a * x[i] * (d ? d[i] : 1.0);
}
__global__ void __launch_bounds__(kBlock) s_hash_kernel(int64_t n, double *__restrict__ v) {
  GS(i, n) {
    unsigned long long h = (unsigned long long)i * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    v[i] = (double)(h & 0xFFFFFull) / 1048576.0 - 0.5;
  }
}

// ---------------------------------------------------------------------- API --
static int dia_apply(const SLevel &l, int mode, const double *x, double *y, const double *b,
                     double omega, cudaStream_t st) {
  const int g = grid_for(l.n, kBlock, 16);
  if (mode == 0)
    dia_apply_kernel<0, false><<<g, kBlock, 0, st>>>(l.n, l.np[0], l.np[1], l.vals, x, y, nullptr,
                                                    nullptr, 0.0, nullptr, nullptr, nullptr,
                                                    nullptr, nullptr);
  else if (mode == 1)
    dia_apply_kernel<1, false><<<g, kBlock, 0, st>>>(l.n, l.np[0], l.np[1], l.vals, x, y, b, l.dinv,
                                                    omega, nullptr, nullptr, nullptr, nullptr,
                                                    nullptr);
  else
    dia_apply_kernel<2, false><<<g, kBlock, 0, st>>>(l.n, l.np[0], l.np[1], l.vals, x, y, b,
                                                    nullptr, 0.0, nullptr, nullptr, nullptr,
                                                    nullptr, nullptr);
  SKTB_KERNEL_OK();
  return 0;
}

// PCG operator hook (pcg.cu): q = A_0 p (+ p.q) on the level-0 stencil
int smg_apply_level0(const sktb_smg *m, const double *x, double *y, const double *dotv,
                     ReduceScratch *rs, double *dot_out, const PcgScalars *S, cudaStream_t st) {
  const SLevel &l = m->lv[0];
  const int g = grid_for(l.n, kBlock, 16);
  if (dotv)
    dia_apply_kernel<0, true><<<g, kBlock, 0, st>>>(l.n, l.np[0], l.np[1], l.vals, x, y, nullptr,
                                                   nullptr, 0.0, dotv, rs->partials, rs->ticket,
                                                   dot_out, S);
  else
    dia_apply_kernel<0, false><<<g, kBlock, 0, st>>>(l.n, l.np[0], l.np[1], l.vals, x, y, nullptr,
                                                    nullptr, 0.0, nullptr, nullptr, nullptr,
                                                    nullptr, S);
  SKTB_KERNEL_OK();
  return 0;
}
int64_t smg_n(const sktb_smg *m) { return m ? m->lv[0].n : 0; }
const double *smg_inv_diag(const sktb_smg *m) { return m->lv[0].dinv; }

extern "C" int sktb_smg_create(sktb_smg **out, int n_levels, const int32_t *np_h, int device) {
  SKTB_REQUIRE(out && n_levels >= 1 && n_levels <= 16 && np_h, "bad argument");
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_smg *m = new sktb_smg();
  m->device = device;
  m->lv.resize(n_levels);
  for (int k = 0; k < n_levels; ++k) {
    SLevel &l = m->lv[k];
    for (int a = 0; a < 3; ++a) l.np[a] = np_h[3 * k + a];
    l.n = (int64_t)l.np[0] * l.np[1] * l.np[2];
    SKTB_REQUIRE(l.n > 0, "empty level");
    SKTB_CUDA_OK(cudaMalloc(&l.vals, sizeof(double) * 27 * l.n));
    SKTB_CUDA_OK(cudaMalloc(&l.dinv, sizeof(double) * l.n));
    SKTB_CUDA_OK(cudaMalloc(&l.x, sizeof(double) * l.n));
    SKTB_CUDA_OK(cudaMalloc(&l.x2, sizeof(double) * l.n));
    SKTB_CUDA_OK(cudaMalloc(&l.b, sizeof(double) * l.n));
    SKTB_CUDA_OK(cudaMalloc(&l.tmp, sizeof(double) * l.n));
    l.nu = m->nu_default[k < 8 ? k : 7];
  }
  SKTB_REQUIRE(n_levels == 1 || m->lv.back().n <= kDenseMaxS,
               "coarsest level too large for the dense solve");
  SKTB_CUDA_OK(cudaMalloc(&m->scal, sizeof(double) * 4));
  SKTB_CUDA_OK(cudaMallocHost(&m->scal_h, sizeof(double) * 4));
  SKTB_CUDA_OK(cudaMalloc(&m->partials,
                          sizeof(double) * ReduceScratch::kMaxVals * ReduceScratch::kMaxBlocks));
  SKTB_CUDA_OK(cudaMalloc(&m->ticket, sizeof(unsigned int)));
  SKTB_CUDA_OK(cudaMemset(m->ticket, 0, sizeof(unsigned int)));
  *out = m;
  return 0;
}

extern "C" void sktb_smg_destroy(sktb_smg *m) {
  if (!m) return;
  cudaSetDevice(m->device);
  for (auto &l : m->lv) {
    cudaFree(l.vals);
    cudaFree(l.dinv);
    cudaFree(l.x);
    cudaFree(l.x2);
    cudaFree(l.b);
    cudaFree(l.tmp);
    cudaFree(l.dense_inv);
  }
  cudaFree(m->scal);
  cudaFreeHost(m->scal_h);
  cudaFree(m->partials);
  cudaFree(m->ticket);
  delete m;
}

extern "C" int sktb_smg_set_mask(sktb_smg *m, int level, const uint8_t *mask) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size(), "bad level");
  m->lv[level].mask = mask;
  return 0;
}

extern "C" int sktb_smg_set_level_sweeps(sktb_smg *m, int level, int nu) {
  SKTB_REQUIRE(m && level >= 0 && level < (int)m->lv.size() && nu >= 1 && nu <= 8, "bad argument");
  m->lv[level].nu = nu;
  return 0;
}

extern "C" int sktb_smg_set_transfer(sktb_smg *m, int level, const int32_t *ax_c0,
                                     const int32_t *ax_c1, const double *ax_w0,
                                     const double *ax_w1, const int32_t *axT_f,
                                     const double *axT_w) {
  SKTB_REQUIRE(m && level >= 0 && level + 1 < (int)m->lv.size(), "bad level");
  SKTB_REQUIRE(ax_c0 && ax_c1 && ax_w0 && ax_w1 && axT_f && axT_w, "null argument");
  SLevel &l = m->lv[level];
  for (int a = 0; a < 3; ++a) {
    l.fnp[a] = l.np[a];
    l.cnp[a] = m->lv[level + 1].np[a];
  }
  l.c0 = ax_c0;
  l.c1 = ax_c1;
  l.w0 = ax_w0;
  l.w1 = ax_w1;
  l.fT = axT_f;
  l.wT = axT_w;
  return 0;
}

static int smg_lambda_max(sktb_smg *m, SLevel &l, int iters, double *out, cudaStream_t st) {
  const int g = grid_for(l.n);
  auto dot = [&](const double *a, const double *b, double *res) -> int {
    s_dot_kernel<<<g, kBlock, 0, st>>>(l.n, a, b, m->partials, m->ticket, m->scal);
    SKTB_KERNEL_OK();
    SKTB_CUDA_OK(cudaMemcpyAsync(m->scal_h, m->scal, sizeof(double), cudaMemcpyDeviceToHost, st));
    SKTB_CUDA_OK(cudaStreamSynchronize(st));
    *res = m->scal_h[0];
    return 0;
  };
  s_hash_kernel<<<g, kBlock, 0, st>>>(l.n, l.x);
  SKTB_KERNEL_OK();
  double lam = 1.0;
  for (int it = 0; it < iters; ++it) {
    double nrm2 = 0.0;
    if (dot(l.x, l.x, &nrm2)) return 1;
    SKTB_REQUIRE(nrm2 > 0.0, "power iteration broke down");
    s_scale_kernel<<<g, kBlock, 0, st>>>(l.n, 1.0 / sqrt(nrm2), l.x, nullptr, l.x);
    SKTB_KERNEL_OK();
    if (dia_apply(l, 0, l.x, l.tmp, nullptr, 0.0, st)) return 1;
    s_scale_kernel<<<g, kBlock, 0, st>>>(l.n, 1.0, l.tmp, l.dinv, l.tmp);
    SKTB_KERNEL_OK();
    if (dot(l.x, l.tmp, &lam)) return 1;
    SKTB_CUDA_OK(cudaMemcpyAsync(l.x, l.tmp, sizeof(double) * l.n, cudaMemcpyDeviceToDevice, st));
  }
  *out = lam;
  return 0;
}

// level 0 from the caller's (enforced) CSR matrix, then the Galerkin chain
extern "C" int sktb_smg_setup_csr(sktb_smg *m, const int32_t *row_ptr, const int32_t *col_idx,
                                  const double *vals, void *stream) {
  SKTB_REQUIRE(m && row_ptr && col_idx && vals, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  SLevel &l0 = m->lv[0];
  int *bad = (int *)(m->scal + 2);
  SKTB_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), st));
  csr_to_dia_kernel<<<grid_for(l0.n, kBlock, 16), kBlock, 0, st>>>(
      l0.n, l0.np[0], l0.np[1], row_ptr, col_idx, vals, l0.vals, bad);
  SKTB_KERNEL_OK();
  const int L = (int)m->lv.size();
  for (int k = 0; k < L; ++k) {
    SLevel &l = m->lv[k];
    dia_inv_diag_kernel<<<grid_for(l.n), kBlock, 0, st>>>(l.n, l.vals, l.dinv);
    SKTB_KERNEL_OK();
    if (k + 1 < L) {
      SLevel &c = m->lv[k + 1];
      SKTB_REQUIRE(l.fT, "transfer tables not set");
      galerkin_dia_kernel<<<(unsigned)((c.n + kGalBlock - 1) / kGalBlock), kGalBlock, 0, st>>>(
          l.cnp[0], l.cnp[1], l.cnp[2], l.fnp[0], l.fnp[1], l.fnp[2], l.fT, l.wT, l.c0, l.c1,
          l.w0, l.w1, l.vals, l.mask, c.mask, c.vals);
      SKTB_KERNEL_OK();
    }
  }
  if (L > 1) {
    SLevel &l = m->lv.back();
    const int n = (int)l.n;
    if (!l.dense_inv) SKTB_CUDA_OK(cudaMalloc(&l.dense_inv, sizeof(double) * kDenseMaxS * kDenseMaxS));
    static bool attr = false;
    if (!attr) {
      SKTB_CUDA_OK(cudaFuncSetAttribute(s_dense_invert_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * (kDenseMaxS * kDenseMaxS + kDenseMaxS))));
      attr = true;
    }
    s_dense_invert_kernel<<<1, 1024, sizeof(double) * ((size_t)n * n + n), st>>>(
        n, l.np[0], l.np[1], l.vals, l.dense_inv);
    SKTB_KERNEL_OK();
    l.dense_n = n;
  }
  if (!m->omega_set) {
    // per-level damping omega_l = 1.75 / (1.03 lambda_max(D^-1 A_l)), once: it
    // depends on the discretisation far more than on the coefficient field
    for (int k = 0; k + 1 < L || k == 0; ++k) {
      double lam = 2.0;
      if (smg_lambda_max(m, m->lv[k], 12, &lam, st)) return 1;
      m->lv[k].omega = 1.75 / (1.03 * lam);
      if (L == 1) break;
    }
    m->omega_set = true;
  }
  SKTB_CUDA_OK(cudaMemcpyAsync(m->scal_h + 2, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  SKTB_REQUIRE(*(int *)(m->scal_h + 2) == 0,
               "matrix couples nodes outside the 27-point neighbourhood of the grid");
  return 0;
}

// z = M^-1 r : V cycle
int smg_vcycle(sktb_smg *m, const double *r, double *z, cudaStream_t st) {
  const int L = (int)m->lv.size();
  if (L == 1) {  // Jacobi
    s_jacobi0_kernel<<<grid_for(m->lv[0].n), kBlock, 0, st>>>(m->lv[0].n, 1.0, m->lv[0].dinv, r, z);
    SKTB_KERNEL_OK();
    return 0;
  }
  for (int k = 0; k < L; ++k) {
    SLevel &l = m->lv[k];
    const double *b = k == 0 ? r : l.b;
    if (k == L - 1) {
      s_dense_apply_kernel<<<(int)((l.n + kBlock / 32 - 1) / (kBlock / 32)), kBlock, 0, st>>>(
          (int)l.n, l.dense_inv, b, l.x);
      SKTB_KERNEL_OK();
      break;
    }
    s_jacobi0_kernel<<<grid_for(l.n), kBlock, 0, st>>>(l.n, l.omega, l.dinv, b, l.x);
    SKTB_KERNEL_OK();
    for (int s = 1; s < l.nu; ++s) {
      if (dia_apply(l, 1, l.x, l.x2, b, l.omega, st)) return 1;
      std::swap(l.x, l.x2);
    }
    if (dia_apply(l, 2, l.x, l.tmp, b, 0.0, st)) return 1;  // residual
    SLevel &c = m->lv[k + 1];
    s_restrict_kernel<<<grid_for(c.n), kBlock, 0, st>>>(l.cnp[0], l.cnp[1], l.cnp[2], l.fnp[0],
                                                       l.fnp[1], l.fT, l.wT, l.tmp, c.mask, c.b);
    SKTB_KERNEL_OK();
  }
  for (int k = L - 2; k >= 0; --k) {
    SLevel &l = m->lv[k];
    SLevel &c = m->lv[k + 1];
    const double *b = k == 0 ? r : l.b;
    s_prolong_kernel<<<grid_for(l.n), kBlock, 0, st>>>(l.cnp[0], l.cnp[1], l.fnp[0], l.fnp[1],
                                                      l.fnp[2], l.c0, l.c1, l.w0, l.w1, c.x,
                                                      l.mask, l.x);
    SKTB_KERNEL_OK();
    for (int s = 0; s < l.nu; ++s) {
      const bool last = (k == 0 && s == l.nu - 1);
      if (dia_apply(l, 1, l.x, last ? z : l.x2, b, l.omega, st)) return 1;
      if (!last) std::swap(l.x, l.x2);
    }
  }
  return 0;
}

extern "C" int sktb_smg_vcycle(sktb_smg *m, const double *r, double *z, void *stream) {
  SKTB_REQUIRE(m && r && z, "null argument");
  return smg_vcycle(m, r, z, (cudaStream_t)stream);
}

extern "C" int sktb_smg_apply(sktb_smg *m, int level, const double *x, double *y, void *stream) {
  SKTB_REQUIRE(m && x && y && level >= 0 && level < (int)m->lv.size(), "bad argument");
  return dia_apply(m->lv[level], 0, x, y, nullptr, 0.0, (cudaStream_t)stream);
}

extern "C" int sktb_smg_level_values(sktb_smg *m, int level, double *out27n, void *stream) {
  SKTB_REQUIRE(m && out27n && level >= 0 && level < (int)m->lv.size(), "bad argument");
  const SLevel &l = m->lv[level];
  SKTB_CUDA_OK(cudaMemcpyAsync(out27n, l.vals, sizeof(double) * 27 * l.n,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}
