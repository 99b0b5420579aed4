// Elementwise sensitivity / projection / density-update kernels (K1, K8,
// K12-K15) and the deterministic single-kernel reductions.
#include <cstring>

#include "common.cuh"

using namespace sktb;

#define GRID_STRIDE(i, n)                                                  \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x,         \
               _st = (int64_t)gridDim.x * blockDim.x;                      \
       i < (n); i += _st)

// ------------------------------------------------------------------- K1 ----
__global__ void __launch_bounds__(kBlock)
    modulus_kernel(int64_t n, const double *__restrict__ rho, double E0,
                   double Emin, double p, int ramp, double *__restrict__ out) {
  GRID_STRIDE(i, n) {
    const double r = rho[i];
    out[i] = ramp ? Emin + (E0 - Emin) * (r / (1.0 + p * (1.0 - r)))
                  : Emin + (E0 - Emin) * pow(r, p);
  }
}
extern "C" int sktb_interpolate_modulus(int64_t n, const double *rho, double E0,
                                        double Emin, double p, int ramp,
                                        double *out, void *stream) {
  SKTB_REQUIRE(rho && out, "null argument");
  modulus_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(
      n, rho, E0, Emin, p, ramp, out);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------- K8 ----
__global__ void __launch_bounds__(kBlock)
    dc_drho_kernel(int64_t n, const double *__restrict__ rho,
                   const double *__restrict__ energy, double E0, double Emin,
                   double p, int ramp, const double *__restrict__ dH,
                   double *__restrict__ out) {
  GRID_STRIDE(i, n) {
    double dE, E;
    const double r = rho[i];
    if (ramp) {
      const double den = 1.0 + p * (1.0 - r);
      dE = (E0 - Emin) * (den - p * r) / (den * den);
      E = Emin + (E0 - Emin) * (r / den);
    } else {
      const double rc = fmax(r, 1e-6);
      dE = p * (E0 - Emin) * pow(rc, p - 1.0);
      E = Emin + (E0 - Emin) * pow(rc, p);
    }
    double g = -2.0 * energy[i] * dE / fmax(E, 1e-12);
    if (dH) g *= dH[i];
    out[i] = g;
  }
}
extern "C" int sktb_dc_drho(int64_t n, const double *rho_proj,
                            const double *energy, double E0, double Emin,
                            double p, int ramp, const double *dH, double *out,
                            void *stream) {
  SKTB_REQUIRE(rho_proj && energy && out, "null argument");
  dc_drho_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(
      n, rho_proj, energy, E0, Emin, p, ramp, dH, out);
  SKTB_KERNEL_OK();
  return 0;
}

// ----------------------------------------------------------------- K13a ----
__global__ void __launch_bounds__(kBlock)
    heaviside_kernel(int64_t n, const double *__restrict__ x, double beta,
                     double eta, double tanh_be, double denom,
                     double *__restrict__ out, double *__restrict__ dH) {
  GRID_STRIDE(i, n) {
    const double a = (x[i] - eta) * beta;
    if (out) out[i] = (tanh_be + tanh(a)) / denom;
    if (dH) {
      const double ch = cosh(a);
      dH[i] = (1.0 / (ch * ch)) * beta / denom;
    }
  }
}
extern "C" int sktb_heaviside(int64_t n, const double *x, double beta,
                              double eta, double *out, double *dH,
                              void *stream) {
  SKTB_REQUIRE(x && (out || dH), "null argument");
  const double tanh_be = tanh(beta * eta);
  const double denom = tanh(beta * eta) + tanh(beta * (1.0 - eta)) + 1e-12;
  heaviside_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(
      n, x, beta, eta, tanh_be, denom, out, dH);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------ K12 ----
__global__ void __launch_bounds__(kBlock)
    oc_candidate_kernel(int64_t n, const double *__restrict__ dC,
                        const double *__restrict__ rho_e, double lmid,
                        double eps, double eta, double move_limit,
                        double rho_min, double rho_max, double sr_min,
                        double sr_max, const int32_t *__restrict__ design_idx,
                        double *__restrict__ scaling_rate,
                        double *__restrict__ rho_cand,
                        double *__restrict__ rho_full_cand) {
  GRID_STRIDE(i, n) {
    double sr = pow(-dC[i] / (lmid + eps), eta);
    // np.clip propagates NaN; fmin/fmax would drop it
    if (sr == sr) sr = fmin(fmax(sr, sr_min), sr_max);
    const double re = rho_e[i];
    const double lo = fmax(re - move_limit, rho_min);
    const double hi = fmin(re + move_limit, rho_max);
    double rc = re * sr;
    if (rc == rc) rc = fmin(fmax(rc, lo), hi);
    scaling_rate[i] = sr;
    rho_cand[i] = rc;
    if (rho_full_cand) rho_full_cand[design_idx[i]] = rc;
  }
}
extern "C" int sktb_oc_candidate(int64_t n_design, const double *dC,
                                 const double *rho_e, double lmid, double eps,
                                 double eta, double move_limit, double rho_min,
                                 double rho_max, double sr_min, double sr_max,
                                 const int32_t *design_idx,
                                 double *scaling_rate, double *rho_cand,
                                 double *rho_full_cand, void *stream) {
  SKTB_REQUIRE(dC && rho_e && scaling_rate && rho_cand, "null argument");
  SKTB_REQUIRE(!rho_full_cand || design_idx, "design_idx required");
  oc_candidate_kernel<<<grid_for(n_design), kBlock, 0, (cudaStream_t)stream>>>(
      n_design, dC, rho_e, lmid, eps, eta, move_limit, rho_min, rho_max, sr_min,
      sr_max, design_idx, scaling_rate, rho_cand, rho_full_cand);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------------------ K14 ----
__global__ void __launch_bounds__(kBlock)
    logmoc_kernel(int64_t n, double *__restrict__ rho,
                  const double *__restrict__ dL, double eta, double move_limit,
                  double rho_min, double rho_max, double clip,
                  double *__restrict__ scaling_rate,
                  double *__restrict__ clip_lower,
                  double *__restrict__ clip_upper) {
  GRID_STRIDE(i, n) {
    double g = dL[i];
    if (g == g) g = fmin(fmax(g, -clip), clip);
    const double r = fmin(fmax(rho[i], rho_min), rho_max);
    const double lr = log(r);
    // the reference recomputes rho as exp(log(rho)) before the division
    const double w = log(move_limit / exp(lr) + 1.0);
    const double lo = lr - w;
    const double hi = lo + 2.0 * w;
    double t = lr - eta * g;
    if (t == t) t = fmin(fmax(t, lo), hi);
    double rn = exp(t);
    if (rn == rn) rn = fmin(fmax(rn, rho_min), rho_max);
    scaling_rate[i] = g;
    clip_lower[i] = lo;
    clip_upper[i] = hi;
    rho[i] = rn;
  }
}
extern "C" int sktb_logmoc_update(int64_t n, double *rho, const double *dL,
                                  double eta, double move_limit, double rho_min,
                                  double rho_max, double clip,
                                  double *scaling_rate, double *clip_lower,
                                  double *clip_upper, void *stream) {
  SKTB_REQUIRE(rho && dL && scaling_rate && clip_lower && clip_upper,
               "null argument");
  logmoc_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(
      n, rho, dL, eta, move_limit, rho_min, rho_max, clip, scaling_rate,
      clip_lower, clip_upper);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------------- small vector ops --
__global__ void __launch_bounds__(kBlock)
    gather_kernel(int64_t n, const double *__restrict__ src,
                  const int32_t *__restrict__ idx, double *__restrict__ dst) {
  GRID_STRIDE(i, n) dst[i] = src[idx[i]];
}
__global__ void __launch_bounds__(kBlock)
    scatter_kernel(int64_t n, const double *__restrict__ src,
                   const int32_t *__restrict__ idx, double *__restrict__ dst) {
  GRID_STRIDE(i, n) dst[idx[i]] = src[i];
}
__global__ void __launch_bounds__(kBlock)
    axpby_kernel(int64_t n, double a, const double *__restrict__ x, double b,
                 double *__restrict__ y) {
  GRID_STRIDE(i, n) y[i] = a * x[i] + b * y[i];
}
__global__ void __launch_bounds__(kBlock)
    enforce_rhs_kernel(int64_t n, const double *__restrict__ b,
                       const double *__restrict__ t,
                       const uint8_t *__restrict__ mask,
                       const double *__restrict__ xD, double *__restrict__ out) {
  GRID_STRIDE(i, n) {
    if (mask[i])
      out[i] = xD ? xD[i] : 0.0;
    else
      out[i] = t ? b[i] - t[i] : b[i];
  }
}
extern "C" int sktb_gather(int64_t n, const double *src, const int32_t *idx,
                           double *dst, void *stream) {
  SKTB_REQUIRE(src && idx && dst, "null argument");
  gather_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, src, idx, dst);
  SKTB_KERNEL_OK();
  return 0;
}
extern "C" int sktb_scatter(int64_t n, const double *src, const int32_t *idx,
                            double *dst, void *stream) {
  SKTB_REQUIRE(src && idx && dst, "null argument");
  scatter_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, src, idx, dst);
  SKTB_KERNEL_OK();
  return 0;
}
extern "C" int sktb_axpby(int64_t n, double a, const double *x, double b,
                          double *y, void *stream) {
  SKTB_REQUIRE(x && y, "null argument");
  axpby_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, a, x, b, y);
  SKTB_KERNEL_OK();
  return 0;
}
__global__ void __launch_bounds__(kBlock)
    affine_kernel(int64_t n, double a, const double *__restrict__ x, double b,
                  const double *__restrict__ y, double c,
                  double *__restrict__ out) {
  GRID_STRIDE(i, n) out[i] = a * x[i] + (y ? b * y[i] : 0.0) + c;
}
__global__ void __launch_bounds__(kBlock)
    hadamard_kernel(int64_t n, double a, const double *__restrict__ x,
                    const double *__restrict__ y, double *__restrict__ out) {
  GRID_STRIDE(i, n) out[i] = a * x[i] * y[i];
}
extern "C" int sktb_affine(int64_t n, double a, const double *x, double b,
                           const double *y, double c, double *out,
                           void *stream) {
  SKTB_REQUIRE(x && out, "null argument");
  affine_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, a, x, b, y,
                                                                  c, out);
  SKTB_KERNEL_OK();
  return 0;
}
extern "C" int sktb_hadamard(int64_t n, double a, const double *x,
                             const double *y, double *out, void *stream) {
  SKTB_REQUIRE(x && y && out, "null argument");
  hadamard_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, a, x, y,
                                                                    out);
  SKTB_KERNEL_OK();
  return 0;
}
__global__ void __launch_bounds__(kBlock)
    fill_abs_kernel(int64_t n, const double *__restrict__ x, double value,
                    double *__restrict__ out) {
  GRID_STRIDE(i, n) out[i] = x ? fabs(x[i]) : value;
}
// out = |x| (x != NULL) or out = value: the two elementwise steps of the loop that
// would otherwise be framework kernels (recorder statistics of |dV|, zeroing the
// full-length sensitivity before the design entries are scattered into it)
extern "C" int sktb_fill_abs(int64_t n, const double *x, double value, double *out,
                             void *stream) {
  SKTB_REQUIRE(out, "null argument");
  fill_abs_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(n, x, value, out);
  SKTB_KERNEL_OK();
  return 0;
}
extern "C" int sktb_enforce_rhs(int64_t n, const double *b, const double *t,
                                const uint8_t *mask, const double *xD,
                                double *out, void *stream) {
  SKTB_REQUIRE(b && mask && out, "null argument");
  enforce_rhs_kernel<<<grid_for(n), kBlock, 0, (cudaStream_t)stream>>>(
      n, b, t, mask, xD, out);
  SKTB_KERNEL_OK();
  return 0;
}

// --------------------------------------------------------------- reductions --
__global__ void __launch_bounds__(kBlock)
    wsum_kernel(int64_t n, const double *__restrict__ a,
                const int32_t *__restrict__ idx, const double *__restrict__ w,
                double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  GRID_STRIDE(i, n) {
    const double ai = idx ? a[idx[i]] : a[i];
    v[0] += w ? ai * w[i] : ai;
  }
  grid_reduce<1>(v, partials, ticket, out);
}
// two-pass statistics: kernel 1 -> min, sum, max ; kernel 2 -> sum (x-mean)^2
__global__ void __launch_bounds__(kBlock)
    stats1_kernel(int64_t n, const double *__restrict__ a,
                  const int32_t *__restrict__ idx, double *partials,
                  unsigned int *ticket, double *out) {
  double v[3] = {1.0 / 0.0, 0.0, -1.0 / 0.0};
  GRID_STRIDE(i, n) {
    const double ai = idx ? a[idx[i]] : a[i];
    v[0] = fmin(v[0], ai);
    v[1] += ai;
    v[2] = fmax(v[2], ai);
  }
  grid_reduce<3, OP_MIN, OP_SUM, OP_MAX>(v, partials, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
    stats2_kernel(int64_t n, const double *__restrict__ a,
                  const int32_t *__restrict__ idx, const double *sum_in,
                  double *partials, unsigned int *ticket, double *out) {
  const double mean = sum_in[1] / (double)n;
  double v[1] = {0.0};
  GRID_STRIDE(i, n) {
    const double d = (idx ? a[idx[i]] : a[i]) - mean;
    v[0] += d * d;
  }
  grid_reduce<1>(v, partials, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
    absmax_kernel(int64_t n, const double *__restrict__ a, double *partials,
                  unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  bool has_nan = false;
  GRID_STRIDE(i, n) {
    const double x = fabs(a[i]);
    if (x != x) has_nan = true;
    v[0] = fmax(v[0], x);
  }
  if (has_nan) v[0] = 1.0 / 0.0;  // surface NaNs loudly (np.max would give NaN)
  grid_reduce<1, OP_MAX>(v, partials, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
    dot_kernel(int64_t n, const double *__restrict__ a,
               const double *__restrict__ b, double *partials,
               unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  GRID_STRIDE(i, n) v[0] += a[i] * b[i];
  grid_reduce<1>(v, partials, ticket, out);
}

// max_i |g[i] + coef*dv[i]| over rho in (lo, hi); count of such i in out[1]
__global__ void __launch_bounds__(kBlock)
    kkt_kernel(int64_t n, const double *__restrict__ rho,
               const double *__restrict__ g, const double *__restrict__ dv,
               double coef, double lo, double hi, double *partials,
               unsigned int *ticket, double *out) {
  double v[2] = {0.0, 0.0};
  GRID_STRIDE(i, n) {
    const double r = rho[i];
    if (r > lo && r < hi) {
      const double d = fabs(g[i] + (dv ? coef * dv[i] : 0.0));
      v[0] = (d != d) ? 1.0 / 0.0 : fmax(v[0], d);
      v[1] += 1.0;
    }
  }
  grid_reduce<2, OP_MAX, OP_SUM>(v, partials, ticket, out);
}
// max_i |a[i] - b[i]|
__global__ void __launch_bounds__(kBlock)
    maxdiff_kernel(int64_t n, const double *__restrict__ a,
                   const double *__restrict__ b, const int32_t *__restrict__ idx,
                   double *partials, unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  GRID_STRIDE(i, n) {
    const int64_t k = idx ? idx[i] : i;
    v[0] = fmax(v[0], fabs(a[k] - b[k]));
  }
  grid_reduce<1, OP_MAX>(v, partials, ticket, out);
}

static int fetch_result(ReduceScratch *rs, int count, double *out_h,
                        cudaStream_t st, int offset = 0) {
  SKTB_CUDA_OK(cudaMemcpyAsync(rs->result_h, rs->result + offset,
                               sizeof(double) * count, cudaMemcpyDeviceToHost,
                               st));
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  for (int i = 0; i < count; ++i) out_h[i] = rs->result_h[i];
  return 0;
}

extern "C" int sktb_reduce_wsum_h(int64_t n, const double *a,
                                  const int32_t *idx, const double *w,
                                  double *out_h, void *stream) {
  SKTB_REQUIRE(a && out_h, "null argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  wsum_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, idx, w, rs->partials,
                                              rs->ticket, rs->result);
  SKTB_KERNEL_OK();
  return fetch_result(rs, 1, out_h, st);
}

extern "C" int sktb_reduce_stats_h(int64_t n, const double *a,
                                   const int32_t *idx, double *out_h,
                                   void *stream) {
  SKTB_REQUIRE(a && out_h && n > 0, "bad argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  stats1_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, idx, rs->partials,
                                                rs->ticket, rs->result);
  stats2_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, idx, rs->result,
                                                rs->partials, rs->ticket,
                                                rs->result + 3);
  SKTB_COUNT(1);
  SKTB_KERNEL_OK();
  double r[4];
  if (fetch_result(rs, 4, r, st)) return 1;
  out_h[0] = r[0];
  out_h[1] = r[1] / (double)n;
  out_h[2] = r[2];
  out_h[3] = sqrt(r[3] / (double)n);
  return 0;
}

extern "C" int sktb_reduce_absmax_h(int64_t n, const double *a, double *out_h,
                                    void *stream) {
  SKTB_REQUIRE(a && out_h, "null argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  absmax_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, rs->partials, rs->ticket,
                                                rs->result);
  SKTB_KERNEL_OK();
  return fetch_result(rs, 1, out_h, st);
}

extern "C" int sktb_kkt_residual_h(int64_t n, const double *rho,
                                   const double *g, const double *dv,
                                   double coef, double lo, double hi,
                                   double *out_h, void *stream) {
  SKTB_REQUIRE(rho && g && out_h, "null argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  kkt_kernel<<<grid_for(n), kBlock, 0, st>>>(n, rho, g, dv, coef, lo, hi,
                                             rs->partials, rs->ticket,
                                             rs->result);
  SKTB_KERNEL_OK();
  return fetch_result(rs, 2, out_h, st);
}

extern "C" int sktb_reduce_maxdiff_h(int64_t n, const double *a,
                                     const double *b, const int32_t *idx,
                                     double *out_h, void *stream) {
  SKTB_REQUIRE(a && b && out_h, "null argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  maxdiff_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, b, idx, rs->partials,
                                                 rs->ticket, rs->result);
  SKTB_KERNEL_OK();
  return fetch_result(rs, 1, out_h, st);
}

extern "C" int sktb_dot_h(int64_t n, const double *a, const double *b,
                          double *out_h, void *stream) {
  SKTB_REQUIRE(a && b && out_h, "null argument");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  dot_kernel<<<grid_for(n), kBlock, 0, st>>>(n, a, b, rs->partials, rs->ticket,
                                             rs->result);
  SKTB_KERNEL_OK();
  return fetch_result(rs, 1, out_h, st);
}

// ------------------------------------------------------------------ K15 ----
// Exact order statistics of |a| by MSB-first radix select on the IEEE bit
// pattern (non-negative doubles order like their uint64 bits): 8 passes of
// 8-bit digits.  State lives in `work`: keys[n] then a small header.
struct SelectState {
  unsigned long long prefix;     // bits decided so far
  unsigned long long rank;       // remaining rank inside the prefix bucket
  unsigned long long rank_next;  // for the (k+1)-th statistic
  unsigned long long result[2];  // key of k-th and (k+1)-th
  unsigned int hist[256];
  unsigned int pad[2];
};

__global__ void __launch_bounds__(kBlock)
    select_init_kernel(int64_t n, const double *__restrict__ a,
                       unsigned long long *__restrict__ keys, SelectState *st,
                       unsigned long long k) {
  GRID_STRIDE(i, n) {
    const double x = fabs(a[i]);
    keys[i] = (unsigned long long)__double_as_longlong(x);
  }
  if (blockIdx.x == 0) {
    if (threadIdx.x < 256) st->hist[threadIdx.x] = 0u;
    if (threadIdx.x == 0) {
      st->prefix = 0ull;
      st->rank = k;
    }
  }
}

// histogram of digit `pass` (0 = most significant byte) among keys matching
// the prefix on the higher bytes
__global__ void __launch_bounds__(kBlock)
    select_hist_kernel(int64_t n, const unsigned long long *__restrict__ keys,
                       SelectState *st, int pass) {
  __shared__ unsigned int sh[256];
  if (threadIdx.x < 256) sh[threadIdx.x] = 0u;
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const unsigned long long prefix = st->prefix;
  const unsigned long long himask =
      pass == 0 ? 0ull : (~0ull) << (shift + 8);
  GRID_STRIDE(i, n) {
    const unsigned long long key = keys[i];
    if ((key & himask) == prefix)
      atomicAdd(&sh[(unsigned int)((key >> shift) & 0xffull)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256 && sh[threadIdx.x])
    atomicAdd(&st->hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void select_scan_kernel(SelectState *st, int pass) {
  // single thread: 256 bins
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int shift = 56 - 8 * pass;
    unsigned long long rank = st->rank, cum = 0ull;
    int d = 0;
    for (; d < 256; ++d) {
      const unsigned long long h = st->hist[d];
      if (rank < cum + h) break;
      cum += h;
    }
    if (d > 255) d = 255;
    st->prefix |= ((unsigned long long)d) << shift;
    st->rank = rank - cum;
    for (int i = 0; i < 256; ++i) st->hist[i] = 0u;
  }
}

// after 8 passes prefix == k-th key.  count(keys <= kth) and min(keys > kth)
__global__ void __launch_bounds__(kBlock)
    select_next_kernel(int64_t n, const unsigned long long *__restrict__ keys,
                       SelectState *st, double *partials, unsigned int *ticket,
                       double *out) {
  const unsigned long long kth = st->prefix;
  double v[2] = {0.0, 1.0 / 0.0};
  GRID_STRIDE(i, n) {
    const unsigned long long key = keys[i];
    if (key <= kth)
      v[0] += 1.0;
    else
      v[1] = fmin(v[1], __longlong_as_double((long long)key));
  }
  grid_reduce<2, OP_SUM, OP_MIN>(v, partials, ticket, out);
}

extern "C" int sktb_abs_percentile_h(int64_t n, const double *a, double q,
                                     void *work, double *out_h, void *stream) {
  SKTB_REQUIRE(a && work && out_h && n > 0, "bad argument");
  SKTB_REQUIRE(q >= 0.0 && q <= 100.0, "percentile must be in [0, 100]");
  ReduceScratch *rs;
  if (reduce_scratch_get(&rs)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long *keys = (unsigned long long *)work;
  SelectState *state = (SelectState *)(keys + n);
  // numpy 'linear': virtual index (n-1)*q/100
  const double vidx = (double)(n - 1) * (q / 100.0);
  long long k = (long long)floor(vidx);
  if (k < 0) k = 0;
  if (k > n - 1) k = n - 1;
  const double frac = vidx - (double)k;
  const int grid = grid_for(n);
  select_init_kernel<<<grid, kBlock, 0, st>>>(n, a, keys, state,
                                              (unsigned long long)k);
  for (int pass = 0; pass < 8; ++pass) {
    select_hist_kernel<<<grid, kBlock, 0, st>>>(n, keys, state, pass);
    select_scan_kernel<<<1, 32, 0, st>>>(state, pass);
  }
  select_next_kernel<<<grid, kBlock, 0, st>>>(n, keys, state, rs->partials,
                                              rs->ticket, rs->result);
  SKTB_COUNT(17);
  SKTB_KERNEL_OK();
  unsigned long long kth_bits = 0;
  SKTB_CUDA_OK(cudaMemcpyAsync(&kth_bits, &state->prefix, sizeof(kth_bits),
                               cudaMemcpyDeviceToHost, st));
  double r[2];
  if (fetch_result(rs, 2, r, st)) return 1;
  double lo;
  memcpy(&lo, &kth_bits, sizeof(double));
  double hi = lo;
  // (k+1)-th statistic: equals kth if there are duplicates covering rank k+1
  if (k + 1 <= n - 1) {
    if ((long long)r[0] > k + 1)
      hi = lo;
    else
      hi = r[1];
  }
  // numpy's lerp: lo + (hi-lo)*frac, with the >= 0.5 branch subtracting from hi
  double res;
  const double diff = hi - lo;
  if (frac >= 0.5)
    res = hi - diff * (1.0 - frac);
  else
    res = lo + diff * frac;
  if (frac == 0.0) res = lo;
  *out_h = res;
  return 0;
}
