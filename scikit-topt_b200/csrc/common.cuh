// Shared helpers for the sktopt_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstdio>
#include <string>

#include "sktopt_b200.h"

namespace sktb {

void set_error(const std::string &msg);

#define SKTB_CUDA_OK(call)                                                   \
  do {                                                                       \
    cudaError_t _e = (call);                                                 \
    if (_e != cudaSuccess) {                                                 \
      sktb::set_error(std::string(#call) + ": " + cudaGetErrorString(_e));   \
      return 1;                                                              \
    }                                                                        \
  } while (0)

// every kernel launch of the library is counted (sktb_launch_count)
extern std::atomic<long long> g_launch_count;
#define SKTB_COUNT(k) sktb::g_launch_count.fetch_add((k), std::memory_order_relaxed)

#define SKTB_KERNEL_OK()                                                     \
  do {                                                                       \
    SKTB_COUNT(1);                                                           \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      sktb::set_error(std::string("kernel launch: ") +                       \
                      cudaGetErrorString(_e));                               \
      return 1;                                                              \
    }                                                                        \
  } while (0)

#define SKTB_KERNEL_CHECK()                                                  \
  do {                                                                       \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      sktb::set_error(std::string("kernel launch: ") +                       \
                      cudaGetErrorString(_e));                               \
      return 1;                                                              \
    }                                                                        \
  } while (0)

#define SKTB_REQUIRE(cond, msg)                                              \
  do {                                                                       \
    if (!(cond)) {                                                           \
      sktb::set_error(msg);                                                  \
      return 2;                                                              \
    }                                                                        \
  } while (0)

constexpr int kNumSM = 148;  // B200
constexpr int kBlock = 256;

inline int grid_for(int64_t n, int block = kBlock, int per_sm = 8) {
  int64_t need = (n + block - 1) / block;
  int64_t cap = (int64_t)kNumSM * per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// Scratch for deterministic grid-wide reductions: per-block partials + ticket.
struct ReduceScratch {
  double *partials = nullptr;  // [kMaxVals][kMaxBlocks]
  unsigned int *ticket = nullptr;
  double *result = nullptr;  // [kMaxVals] device
  double *result_h = nullptr;  // pinned host mirror
  static constexpr int kMaxVals = 4;
  static constexpr int kMaxBlocks = kNumSM * 16;
};
int reduce_scratch_get(ReduceScratch **out);  // per-device singleton

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v = fmin(v, __shfl_down_sync(0xffffffffu, v, o));
  return v;
}

enum ReduceOp { OP_SUM = 0, OP_MAX = 1, OP_MIN = 2 };

template <int OP>
__device__ __forceinline__ double red_combine(double a, double b) {
  if (OP == OP_SUM) return a + b;
  if (OP == OP_MAX) return fmax(a, b);
  return fmin(a, b);
}
template <int OP>
__device__ __forceinline__ double red_warp(double v) {
  if (OP == OP_SUM) return warp_sum(v);
  if (OP == OP_MAX) return warp_max(v);
  return warp_min(v);
}
template <int OP>
__device__ __forceinline__ double red_identity() {
  if (OP == OP_SUM) return 0.0;
  if (OP == OP_MAX) return -1.0 / 0.0;
  return 1.0 / 0.0;
}

// Block-level reduction of NV values followed by a fixed-order reduction of
// the per-block partials in whichever block finishes last ("ticket" pattern).
// Deterministic for a fixed grid.  Returns true in the finishing block after
// `out[0..NV)` has been written (by thread 0).  All threads of the block must
// call it.  blockDim.x must be kBlock.
template <int NV, int OP0 = OP_SUM, int OP1 = OP_SUM, int OP2 = OP_SUM,
          int OP3 = OP_SUM>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials,
                                            unsigned int *ticket, double *out) {
  __shared__ double sm[NV][kBlock / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int ops[4] = {OP0, OP1, OP2, OP3};
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double w = (k == 0)   ? red_warp<OP0>(v[k])
               : (k == 1) ? red_warp<OP1>(v[k])
               : (k == 2) ? red_warp<OP2>(v[k])
                          : red_warp<OP3>(v[k]);
    if (lane == 0) sm[k][wid] = w;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double a = sm[k][0];
      for (int w = 1; w < kBlock / 32; ++w) {
        a = (ops[k] == OP_SUM)   ? a + sm[k][w]
            : (ops[k] == OP_MAX) ? fmax(a, sm[k][w])
                                 : fmin(a, sm[k][w]);
      }
      partials[k * ReduceScratch::kMaxBlocks + blockIdx.x] = a;
    }
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
  // fixed-order: thread t folds partials t, t+B, ...; then smem tree
  __shared__ double fin[NV][kBlock];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double a = (ops[k] == OP_SUM) ? 0.0
               : (ops[k] == OP_MAX) ? -1.0 / 0.0
                                    : 1.0 / 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kBlock) {
      double pv = __ldcg(&partials[k * ReduceScratch::kMaxBlocks + b]);
      a = (ops[k] == OP_SUM)   ? a + pv
          : (ops[k] == OP_MAX) ? fmax(a, pv)
                               : fmin(a, pv);
    }
    fin[k][threadIdx.x] = a;
  }
  __syncthreads();
  for (int s = kBlock / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double a = fin[k][threadIdx.x], b = fin[k][threadIdx.x + s];
        fin[k][threadIdx.x] = (ops[k] == OP_SUM)   ? a + b
                              : (ops[k] == OP_MAX) ? fmax(a, b)
                                                   : fmin(a, b);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) out[k] = fin[k][0];
    *ticket = 0u;
    __threadfence();
  }
  return true;
}

}  // namespace sktb
