// SpMV for the 3-dof-per-node elasticity operator with block (node-level)
// column indices.
//
// The values keep the CSR layout (row 3n+i holds its 3*deg(n) entries
// contiguously, the three rows of a node back to back), but the kernel reads
// ONE int32 column per 3x3 block (the node graph) instead of one per entry:
// 8 + 4/9 bytes per non-zero instead of 12.  One warp owns a node: its
// 9*deg <= 243 values are one contiguous run, loaded with up to 8 coalesced
// loads per lane that are all in flight before the first gather of x.
#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

template <bool DOT>
__global__ void __launch_bounds__(kBlock, 6)
    spmv_bsr3_kernel(int64_t n_nodes, const int32_t *__restrict__ node_ptr,
                     const int32_t *__restrict__ node_col,
                     const double *__restrict__ vals,
                     const double *__restrict__ x, double *__restrict__ y,
                     const double *__restrict__ dotv, double *partials,
                     unsigned int *ticket, double *dot_out,
                     const PcgScalars *S, const JacobiEpi J) {
  if (S && S->rr <= S->tol2) return;
  constexpr int U = 8;  // 8 * 32 = 256 >= 243 values of a 27-neighbour node
  const int lane = threadIdx.x & 31;
  // each CTA walks a CONTIGUOUS chunk of nodes: the x entries gathered for one
  // mesh line are reused from L1 by the neighbouring lines of the same chunk
  const int64_t chunk = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int64_t n_lo = (int64_t)blockIdx.x * chunk;
  const int64_t n_hi = (n_lo + chunk < n_nodes) ? n_lo + chunk : n_nodes;
  double dot = 0.0;
  for (int64_t n = n_lo + (threadIdx.x >> 5); n < n_hi; n += kBlock / 32) {
    const int32_t s0 = __ldg(&node_ptr[n]);
    const int32_t deg = __ldg(&node_ptr[n + 1]) - s0;
    const int32_t w = 3 * deg;          // entries per row
    const int32_t total = 3 * w;        // entries of the node's three rows
    const double *vp = vals + (int64_t)9 * s0;
    const int32_t *cp = node_col + s0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int32_t e0 = lane; e0 < total; e0 += U * 32) {
      double v[U];
      int32_t nc[U];
      // phase 1: matrix stream (values + one block column per entry's block)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t e = e0 + 32 * u;
        const bool ok = e < total;
        const int32_t q = e - ((e >= w) ? w : 0) - ((e >= 2 * w) ? w : 0);
        v[u] = ok ? __ldcs(&vp[e]) : 0.0;
        nc[u] = ok ? __ldg(&cp[q / 3]) : 0;
      }
      // phase 2: gather x and accumulate into the row the entry belongs to
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t e = e0 + 32 * u;
        const int32_t q = e - ((e >= w) ? w : 0) - ((e >= 2 * w) ? w : 0);
        const double t = v[u] * __ldg(&x[3 * nc[u] + (q - 3 * (q / 3))]);
        if (e < w)
          a0 += t;
        else if (e < 2 * w)
          a1 += t;
        else
          a2 += t;
      }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      const int64_t r = 3 * n;
      if (J.b) {  // fused damped-Jacobi sweep
        const double *xo = J.xo ? J.xo : x;
        a0 = xo[r] + J.omega * J.dinv[r] * (J.b[r] - a0);
        a1 = xo[r + 1] + J.omega * J.dinv[r + 1] * (J.b[r + 1] - a1);
        a2 = xo[r + 2] + J.omega * J.dinv[r + 2] * (J.b[r + 2] - a2);
      }
      y[r] = a0;
      y[r + 1] = a1;
      y[r + 2] = a2;
      if (DOT) dot += a0 * dotv[r] + a1 * dotv[r + 1] + a2 * dotv[r + 2];
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

int launch_spmv_bsr3(int64_t n_nodes, const int32_t *node_ptr,
                     const int32_t *node_col, const double *vals,
                     const double *x, double *y, const double *dotv,
                     ReduceScratch *rs, double *dot_out, const PcgScalars *S,
                     cudaStream_t st) {
  const int grid = grid_for(n_nodes * 32, kBlock, 6);
  if (dotv)
    spmv_bsr3_kernel<true><<<grid, kBlock, 0, st>>>(
        n_nodes, node_ptr, node_col, vals, x, y, dotv, rs->partials, rs->ticket,
        dot_out, S, JacobiEpi());
  else
    spmv_bsr3_kernel<false><<<grid, kBlock, 0, st>>>(
        n_nodes, node_ptr, node_col, vals, x, y, nullptr, nullptr, nullptr,
        nullptr, S, JacobiEpi());
  SKTB_KERNEL_OK();
  return 0;
}

int launch_spmv_bsr3_jacobi(int64_t n_nodes, const int32_t *node_ptr,
                            const int32_t *node_col, const double *vals,
                            const double *x, double *y, const JacobiEpi &epi,
                            cudaStream_t st) {
  const int grid = grid_for(n_nodes * 32, kBlock, 6);
  spmv_bsr3_kernel<false><<<grid, kBlock, 0, st>>>(
      n_nodes, node_ptr, node_col, vals, x, y, nullptr, nullptr, nullptr, nullptr,
      nullptr, epi);
  SKTB_KERNEL_OK();
  return 0;
}

// out[3n+i] = 1 / A[3n+i, 3n+i] from the node-block layout
__global__ void __launch_bounds__(kBlock)
    bsr3_inv_diag_kernel(int64_t n_nodes, int64_t node0,
                         const int32_t *__restrict__ node_ptr,
                         const int32_t *__restrict__ node_col,
                         const double *__restrict__ vals,
                         double *__restrict__ out) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; n < n_nodes; n += stride) {
    const int32_t s0 = node_ptr[n], deg = node_ptr[n + 1] - s0;
    for (int32_t s = 0; s < deg; ++s) {
      if (node_col[s0 + s] == n + node0) {
        const double *vp = vals + (int64_t)9 * s0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
          out[3 * n + i] = 1.0 / vp[(int64_t)i * 3 * deg + 3 * s + i];
        break;
      }
    }
  }
}

extern "C" int sktb_bsr3_inv_diag(int64_t n_nodes, int64_t node0,
                                  const int32_t *node_ptr,
                                  const int32_t *node_col, const double *vals,
                                  double *out, void *stream) {
  SKTB_REQUIRE(node_ptr && node_col && vals && out, "null argument");
  bsr3_inv_diag_kernel<<<grid_for(n_nodes), kBlock, 0, (cudaStream_t)stream>>>(
      n_nodes, node0, node_ptr, node_col, vals, out);
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_spmv_bsr3(int64_t n_nodes, const int32_t *node_ptr,
                              const int32_t *node_col, const double *vals,
                              const double *x, double *y, void *stream) {
  SKTB_REQUIRE(node_ptr && node_col && vals && x && y, "null argument");
  return launch_spmv_bsr3(n_nodes, node_ptr, node_col, vals, x, y, nullptr,
                          nullptr, nullptr, nullptr, (cudaStream_t)stream);
}
