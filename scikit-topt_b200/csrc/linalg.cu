// CSR kernels and the device-resident preconditioned CG (K3-K6).
//
// SpMV layout: CSR with int32 indices / fp64 values.  A group of TPR lanes
// owns one row (TPR = 32 for the 81-nnz elasticity rows, 8 for the 27-nnz
// scalar rows); lanes stride the row so matrix traffic is coalesced and
// streamed once (ld.global.cs), while the gathered vector stays in L1/L2.
//
// PCG: all CG scalars live in device memory; every kernel of an iteration
// tests the convergence flag itself and turns into a no-op once
// ||r||^2 <= tol^2, so the host only polls every `check_every` iterations and
// the result is exactly the iterate at which scipy's criterion first holds.
#include "common.cuh"

using namespace sktb;

// ------------------------------------------------------------- CSR helpers --
__global__ void __launch_bounds__(kBlock)
    csr_enforce_kernel(int64_t n_rows, const int32_t *__restrict__ row_ptr,
                       const int32_t *__restrict__ col_idx,
                       double *__restrict__ vals,
                       const uint8_t *__restrict__ mask) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const bool mr = mask[r];
    const int32_t e = row_ptr[r + 1];
    for (int32_t k = row_ptr[r] + lane; k < e; k += 32) {
      const int32_t c = col_idx[k];
      if (mr || mask[c]) vals[k] = (c == r) ? 1.0 : 0.0;
    }
  }
}

extern "C" int sktb_csr_enforce(int64_t n_rows, const int32_t *row_ptr,
                                const int32_t *col_idx, double *vals,
                                const uint8_t *dir_mask, void *stream) {
  SKTB_REQUIRE(row_ptr && col_idx && vals && dir_mask, "null argument");
  csr_enforce_kernel<<<grid_for(n_rows * 32, kBlock, 16), kBlock, 0,
                       (cudaStream_t)stream>>>(n_rows, row_ptr, col_idx, vals,
                                               dir_mask);
  SKTB_KERNEL_OK();
  return 0;
}

__global__ void __launch_bounds__(kBlock)
    csr_inv_diag_kernel(int64_t n_rows, const int32_t *__restrict__ row_ptr,
                        const int32_t *__restrict__ col_idx,
                        const double *__restrict__ vals,
                        double *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rows; r += nwarps) {
    const int32_t e = row_ptr[r + 1];
    for (int32_t k = row_ptr[r] + lane; k < e; k += 32)
      if (col_idx[k] == r) out[r] = 1.0 / vals[k];
  }
}

extern "C" int sktb_csr_inv_diag(int64_t n_rows, const int32_t *row_ptr,
                                 const int32_t *col_idx, const double *vals,
                                 double *out, void *stream) {
  SKTB_REQUIRE(row_ptr && col_idx && vals && out, "null argument");
  csr_inv_diag_kernel<<<grid_for(n_rows * 32, kBlock, 16), kBlock, 0,
                        (cudaStream_t)stream>>>(n_rows, row_ptr, col_idx, vals,
                                                out);
  SKTB_KERNEL_OK();
  return 0;
}

// --------------------------------------------------------------------- SpMV --
struct PcgScalars {
  double rz;       // r.z of the current iterate
  double pq;       // p.Ap
  double rz_new;   // r.z after the update
  double rr;       // ||r||^2
  double tol2;     // (rtol*||b||)^2
  double bb;       // ||b||^2
  int iters;       // completed iterations
  int pad;
};

__device__ __forceinline__ bool pcg_done(const PcgScalars *S) {
  return S && (S->rr <= S->tol2);
}

template <int TPR>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = TPR / 2; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o, TPR);
  return v;
}

// Predicated load of up to U entries per lane of one row segment [s, e).
template <int TPR, int U>
__device__ __forceinline__ void row_load(const int32_t *__restrict__ col_idx,
                                         const double *__restrict__ vals,
                                         int32_t k, int32_t e, int32_t (&c)[U],
                                         double (&v)[U]) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int32_t kk = k + u * TPR;
    const bool ok = kk < e;
    c[u] = ok ? __ldcs(&col_idx[kk]) : 0;
    v[u] = ok ? __ldcs(&vals[kk]) : 0.0;
  }
}

// y = A x ; if DOT also publishes sum_r dotv[r]*y[r] into *dot_out.
// Each group of TPR lanes owns two consecutive rows per trip so that
// 2*U matrix loads per lane are in flight before the first gather of x.
template <int TPR, int U, bool DOT>
__global__ void __launch_bounds__(kBlock)
    spmv_kernel(int64_t n_rows, const int32_t *__restrict__ row_ptr,
                const int32_t *__restrict__ col_idx,
                const double *__restrict__ vals, const double *__restrict__ x,
                double *__restrict__ y, const double *__restrict__ dotv,
                double *partials, unsigned int *ticket, double *dot_out,
                const PcgScalars *S) {
  if (pcg_done(S)) return;
  const int lane = threadIdx.x & (TPR - 1);
  const int gw = (threadIdx.x & 31) / TPR;  // group within the warp
  const int64_t wg0 =
      ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / TPR;
  const int64_t ngrp = ((int64_t)gridDim.x * blockDim.x) / TPR;
  double dot = 0.0;
  // warp-uniform trip count: shuffles below use the full mask
  for (int64_t gb = wg0; 2 * gb < n_rows; gb += ngrp) {
    const int64_t r0 = 2 * (gb + gw), r1 = r0 + 1;
    int32_t s0 = 0, e0 = 0, e1 = 0;
    if (r0 < n_rows) {
      s0 = __ldg(&row_ptr[r0]);
      e0 = __ldg(&row_ptr[r0 + 1]);
      e1 = (r1 < n_rows) ? __ldg(&row_ptr[r1 + 1]) : e0;
    }
    double acc0 = 0.0, acc1 = 0.0;
    int32_t k0 = s0 + lane, k1 = e0 + lane;
    {
      int32_t c0[U], c1[U];
      double v0[U], v1[U];
      row_load<TPR, U>(col_idx, vals, k0, e0, c0, v0);
      row_load<TPR, U>(col_idx, vals, k1, e1, c1, v1);
#pragma unroll
      for (int u = 0; u < U; ++u) acc0 += v0[u] * __ldg(&x[c0[u]]);
#pragma unroll
      for (int u = 0; u < U; ++u) acc1 += v1[u] * __ldg(&x[c1[u]]);
      k0 += U * TPR;
      k1 += U * TPR;
    }
    // long rows (rare): keep going one row at a time
    for (; k0 < e0; k0 += U * TPR) {
      int32_t c0[U];
      double v0[U];
      row_load<TPR, U>(col_idx, vals, k0, e0, c0, v0);
#pragma unroll
      for (int u = 0; u < U; ++u) acc0 += v0[u] * __ldg(&x[c0[u]]);
    }
    for (; k1 < e1; k1 += U * TPR) {
      int32_t c1[U];
      double v1[U];
      row_load<TPR, U>(col_idx, vals, k1, e1, c1, v1);
#pragma unroll
      for (int u = 0; u < U; ++u) acc1 += v1[u] * __ldg(&x[c1[u]]);
    }
    acc0 = group_sum<TPR>(acc0);
    acc1 = group_sum<TPR>(acc1);
    if (lane == 0 && r0 < n_rows) {
      y[r0] = acc0;
      if (DOT) dot += acc0 * dotv[r0];
      if (r1 < n_rows) {
        y[r1] = acc1;
        if (DOT) dot += acc1 * dotv[r1];
      }
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

static int launch_spmv(int64_t n_rows, int dpn_hint, const int32_t *row_ptr,
                       const int32_t *col_idx, const double *vals,
                       const double *x, double *y, const double *dotv,
                       ReduceScratch *rs, double *dot_out, const PcgScalars *S,
                       cudaStream_t st) {
  if (dpn_hint >= 3) {
    // 81-nnz rows: a full warp per row pair, 3 entries per lane
    const int grid = grid_for((n_rows + 1) / 2 * 32, kBlock, 8);
    if (dotv)
      spmv_kernel<32, 3, true><<<grid, kBlock, 0, st>>>(
          n_rows, row_ptr, col_idx, vals, x, y, dotv, rs->partials, rs->ticket,
          dot_out, S);
    else
      spmv_kernel<32, 3, false><<<grid, kBlock, 0, st>>>(
          n_rows, row_ptr, col_idx, vals, x, y, nullptr, nullptr, nullptr,
          nullptr, S);
  } else {
    // 27-nnz (hex) / ~15-nnz (tet) scalar rows: 8 lanes per row pair
    const int grid = grid_for((n_rows + 1) / 2 * 8, kBlock, 8);
    if (dotv)
      spmv_kernel<8, 4, true><<<grid, kBlock, 0, st>>>(
          n_rows, row_ptr, col_idx, vals, x, y, dotv, rs->partials, rs->ticket,
          dot_out, S);
    else
      spmv_kernel<8, 4, false><<<grid, kBlock, 0, st>>>(
          n_rows, row_ptr, col_idx, vals, x, y, nullptr, nullptr, nullptr,
          nullptr, S);
  }
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_spmv(int64_t n_rows, int dpn_hint, const int32_t *row_ptr,
                         const int32_t *col_idx, const double *vals,
                         const double *x, double *y, void *stream) {
  SKTB_REQUIRE(row_ptr && col_idx && vals && x && y, "null argument");
  return launch_spmv(n_rows, dpn_hint, row_ptr, col_idx, vals, x, y, nullptr,
                     nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------- PCG --
struct sktb_pcg {
  int64_t n = 0;
  int device = 0;
  double *r = nullptr, *z = nullptr, *p = nullptr, *q = nullptr;
  PcgScalars *S = nullptr;    // device
  PcgScalars *S_h = nullptr;  // pinned host
  double *partials = nullptr;
  unsigned int *ticket = nullptr;
  // optional in-situ timing of the SpMV launches (every `prof_every`-th
  // iteration gets a CUDA-event pair on the solver's stream)
  int prof_every = 0;
  static constexpr int kMaxProf = 64;
  cudaEvent_t ev0[kMaxProf], ev1[kMaxProf];
  bool ev_init = false;
  double prof_ms = 0.0;
  long long prof_count = 0;
};

extern "C" int sktb_pcg_set_profile(sktb_pcg *s, int every_n) {
  SKTB_REQUIRE(s, "null argument");
  if (every_n > 0 && !s->ev_init) {
    for (int i = 0; i < sktb_pcg::kMaxProf; ++i) {
      SKTB_CUDA_OK(cudaEventCreate(&s->ev0[i]));
      SKTB_CUDA_OK(cudaEventCreate(&s->ev1[i]));
    }
    s->ev_init = true;
  }
  s->prof_every = every_n;
  s->prof_ms = 0.0;
  s->prof_count = 0;
  return 0;
}

extern "C" int sktb_pcg_get_profile(const sktb_pcg *s, double *ms_sum_h,
                                    int64_t *count_h) {
  SKTB_REQUIRE(s && ms_sum_h && count_h, "null argument");
  *ms_sum_h = s->prof_ms;
  *count_h = s->prof_count;
  return 0;
}

extern "C" int sktb_pcg_create(sktb_pcg **out, int64_t n_rows, int device) {
  SKTB_REQUIRE(out && n_rows > 0, "bad argument");
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_pcg *s = new sktb_pcg();
  s->n = n_rows;
  s->device = device;
  SKTB_CUDA_OK(cudaMalloc(&s->r, sizeof(double) * n_rows));
  SKTB_CUDA_OK(cudaMalloc(&s->z, sizeof(double) * n_rows));
  SKTB_CUDA_OK(cudaMalloc(&s->p, sizeof(double) * n_rows));
  SKTB_CUDA_OK(cudaMalloc(&s->q, sizeof(double) * n_rows));
  SKTB_CUDA_OK(cudaMalloc(&s->S, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMallocHost(&s->S_h, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMalloc(&s->partials, sizeof(double) * ReduceScratch::kMaxVals *
                                            ReduceScratch::kMaxBlocks));
  SKTB_CUDA_OK(cudaMalloc(&s->ticket, sizeof(unsigned int)));
  SKTB_CUDA_OK(cudaMemset(s->ticket, 0, sizeof(unsigned int)));
  *out = s;
  return 0;
}

extern "C" void sktb_pcg_destroy(sktb_pcg *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->r);
  cudaFree(s->z);
  cudaFree(s->p);
  cudaFree(s->q);
  cudaFree(s->S);
  cudaFreeHost(s->S_h);
  cudaFree(s->partials);
  cudaFree(s->ticket);
  delete s;
}

// r = b - q (q = A x0, or r = b when !have_q); z = Minv r; p = z;
// publishes rz, rr, bb and tol2.
__global__ void __launch_bounds__(kBlock)
    pcg_init_kernel(int64_t n, const double *__restrict__ b,
                    const double *__restrict__ q, int have_q,
                    const double *__restrict__ minv, double *__restrict__ r,
                    double *__restrict__ z, double *__restrict__ p,
                    double *__restrict__ x, double rtol, double *partials,
                    unsigned int *ticket, PcgScalars *S) {
  double v[3] = {0.0, 0.0, 0.0};
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const double bi = b[i];
    const double ri = have_q ? bi - q[i] : bi;
    const double zi = minv[i] * ri;
    if (!have_q) x[i] = 0.0;
    r[i] = ri;
    z[i] = zi;
    p[i] = zi;
    v[0] += ri * zi;
    v[1] += ri * ri;
    v[2] += bi * bi;
  }
  __shared__ double res[3];
  if (grid_reduce<3>(v, partials, ticket, res)) {
    if (threadIdx.x == 0) {
      S->rz = res[0];
      S->rr = res[1];
      S->bb = res[2];
      S->tol2 = rtol * rtol * res[2];
      S->pq = 0.0;
      S->rz_new = res[0];
      S->iters = 0;
    }
  }
}

// x += a p ; r -= a q ; z = Minv r ; publishes rz_new, rr ; iters++
__global__ void __launch_bounds__(kBlock)
    pcg_update_kernel(int64_t n, const double *__restrict__ p,
                      const double *__restrict__ q,
                      const double *__restrict__ minv, double *__restrict__ x,
                      double *__restrict__ r, double *__restrict__ z,
                      double *partials, unsigned int *ticket, PcgScalars *S) {
  if (pcg_done(S)) return;
  const double alpha = S->rz / S->pq;
  double v[2] = {0.0, 0.0};
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * q[i];
    const double zi = minv[i] * ri;
    r[i] = ri;
    z[i] = zi;
    v[0] += ri * zi;
    v[1] += ri * ri;
  }
  __shared__ double res[2];
  if (grid_reduce<2>(v, partials, ticket, res)) {
    if (threadIdx.x == 0) {
      S->rz_new = res[0];
      S->rr = res[1];
      S->iters += 1;
    }
  }
}

// p = z + beta p ; the last block rolls rz <- rz_new
__global__ void __launch_bounds__(kBlock)
    pcg_direction_kernel(int64_t n, const double *__restrict__ z,
                         double *__restrict__ p, unsigned int *ticket,
                         PcgScalars *S) {
  if (pcg_done(S)) return;
  const double beta = S->rz_new / S->rz;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = z[i] + beta * p[i];
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) {
      S->rz = S->rz_new;
      *ticket = 0u;
      __threadfence();
    }
  }
}

extern "C" int sktb_pcg_solve(sktb_pcg *s, int dpn_hint, const int32_t *row_ptr,
                              const int32_t *col_idx, const double *vals,
                              const double *inv_diag, const double *b,
                              double *x, int use_x0, double rtol, int maxiter,
                              int check_every, int32_t *info_h,
                              double *relres_h, void *stream) {
  SKTB_REQUIRE(s && row_ptr && col_idx && vals && inv_diag && b && x,
               "null argument");
  SKTB_REQUIRE(maxiter >= 0, "maxiter must be >= 0");
  if (check_every <= 0) check_every = 32;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = s->n;
  ReduceScratch rs;
  rs.partials = s->partials;
  rs.ticket = s->ticket;
  const int vgrid = grid_for(n, kBlock, 8);
  if (use_x0) {
    if (launch_spmv(n, dpn_hint, row_ptr, col_idx, vals, x, s->q, nullptr,
                    nullptr, nullptr, nullptr, st))
      return 1;
  }
  pcg_init_kernel<<<vgrid, kBlock, 0, st>>>(n, b, s->q, use_x0 ? 1 : 0, inv_diag,
                                           s->r, s->z, s->p, x, rtol,
                                           s->partials, s->ticket, s->S);
  SKTB_KERNEL_OK();
  int launched = 0;
  bool done = false;
  int n_ev = 0;
  while (!done) {
    SKTB_CUDA_OK(cudaMemcpyAsync(s->S_h, s->S, sizeof(PcgScalars),
                                 cudaMemcpyDeviceToHost, st));
    SKTB_CUDA_OK(cudaStreamSynchronize(st));
    // events of the previous batch are complete now; a sampled SpMV that ran
    // as a no-op (after convergence) is recognised by iteration index
    for (int i = 0; i < n_ev; ++i) {
      float ms = 0.f;
      SKTB_CUDA_OK(cudaEventElapsedTime(&ms, s->ev0[i], s->ev1[i]));
      s->prof_ms += ms;
      s->prof_count += 1;
    }
    n_ev = 0;
    if (s->S_h->rr <= s->S_h->tol2 || launched >= maxiter) break;
    int batch = maxiter - launched;
    if (batch > check_every) batch = check_every;
    for (int it = 0; it < batch; ++it) {
      // sample only the first iteration of a batch: it is certain to do work
      const bool sample = s->prof_every > 0 && it == 0 &&
                          ((launched / check_every) % s->prof_every == 0) &&
                          n_ev < sktb_pcg::kMaxProf;
      if (sample) SKTB_CUDA_OK(cudaEventRecord(s->ev0[n_ev], st));
      if (launch_spmv(n, dpn_hint, row_ptr, col_idx, vals, s->p, s->q, s->p,
                      &rs, &s->S->pq, s->S, st))
        return 1;
      if (sample) SKTB_CUDA_OK(cudaEventRecord(s->ev1[n_ev++], st));
      pcg_update_kernel<<<vgrid, kBlock, 0, st>>>(n, s->p, s->q, inv_diag, x,
                                                 s->r, s->z, s->partials,
                                                 s->ticket, s->S);
      pcg_direction_kernel<<<vgrid, kBlock, 0, st>>>(n, s->z, s->p, s->ticket,
                                                    s->S);
      SKTB_COUNT(2);
    }
    SKTB_KERNEL_CHECK();
    launched += batch;
  }
  const PcgScalars &h = *s->S_h;
  if (info_h) {
    info_h[0] = h.iters;
    info_h[1] = (h.rr <= h.tol2) ? 1 : 0;
  }
  if (relres_h) *relres_h = (h.bb > 0.0) ? sqrt(h.rr / h.bb) : 0.0;
  return 0;
}

// benchmark utility ---------------------------------------------------------
__global__ void flush_kernel(double *buf, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = buf[i] * 0.5 + 1.0;
}

extern "C" int sktb_flush_l2(void *scratch, int64_t bytes, void *stream) {
  SKTB_REQUIRE(scratch && bytes >= 8, "bad argument");
  flush_kernel<<<kNumSM * 8, kBlock, 0, (cudaStream_t)stream>>>(
      (double *)scratch, bytes / 8);
  SKTB_KERNEL_OK();
  return 0;
}
