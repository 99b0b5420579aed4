// Bulk-async (TMA: cp.async.bulk + mbarrier) pipelined SpMV for the
// 3-dof-per-node elasticity operator with node-block column indices.
//
// ncu on the register-staged kernels (profiles/r1_*) shows they are bound by
// instruction issue, not by bytes: ~410 warp instructions per node row for
// index arithmetic, per-entry column loads and three 5-step fp64 warp
// reductions.  This kernel removes both limits:
//   * the matrix stream is decoupled from the warps: a persistent CTA (one per
//     SM pair of slots) walks tiles of 16 consecutive nodes; a producer warp
//     issues two 1-D bulk copies per tile (the tile's contiguous values,
//     <= 31 KB, and its block columns) into a 3-stage shared-memory ring and
//     signals an mbarrier; two CTAs per SM keep ~130 KB per SM in flight
//     regardless of occupancy;
//   * four threads own one matrix row and walk it from shared memory (no
//     cross-lane reduction besides two shuffles, one column load per 3x3 block,
//     x gathered through the read-only path: the three rows of a node
//     broadcast and neighbouring nodes coalesce) -> ~30 warp instructions per
//     node row.
// Consumers release a stage through an "empty" mbarrier; the producer never
// blocks the consumers with __syncthreads.
#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

namespace {

constexpr int kTile = 16;     // nodes per tile -> 48 rows x 4 threads -> 6 consumer warps
constexpr int kStages = 3;
constexpr int kMaxDeg = 27;   // hex8 node graph
constexpr int kConsumerWarps = 6;
constexpr int kProducerWarp = 6;
constexpr int kColCap = kTile * kMaxDeg + 8;      // int32   (+ alignment slack)
constexpr int kPtrCap = 36;                       // >= 33 slots (one per producer lane + end)
// V = double: the PCG operator; V = float: multigrid levels whose values are kept
// in single precision (half the HBM stream of a V-cycle product; the accumulation
// stays fp64)
template <typename V> struct TmaCfg {
  static constexpr int kAlign = 16 / (int)sizeof(V);              // values per 16 bytes
  static constexpr int kValCap = kTile * 9 * kMaxDeg + 2 * kAlign;  // + alignment slack
  static constexpr int kValBytes = kValCap * (int)sizeof(V);
  static constexpr int kStageBytes = kValBytes + kColCap * 4 + kPtrCap * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStages * 8 + 16;
  static_assert(kValBytes % 16 == 0 && kStageBytes % 16 == 0,
                "stages must stay 16-byte aligned");
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar,
                                                      uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src,
                                         uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}

template <typename V> struct StagePtrs {
  V *vals;
  int32_t *cols;
  int32_t *nptr;
};

template <typename V>
__device__ __forceinline__ StagePtrs<V> stage_ptrs(unsigned char *smem, int s) {
  unsigned char *b = smem + (size_t)s * TmaCfg<V>::kStageBytes;
  StagePtrs<V> p;
  p.vals = (V *)b;
  p.cols = (int32_t *)(b + TmaCfg<V>::kValBytes);
  p.nptr = (int32_t *)(b + TmaCfg<V>::kValBytes + kColCap * 4);
  return p;
}

// producer warp: stage the tile's node_ptr slice and issue its two bulk copies
template <typename V>
__device__ __forceinline__ void issue_tile(
    unsigned char *smem, uint64_t *full, int s, int64_t tile, int64_t n_nodes,
    int64_t n_blocks, const int32_t *__restrict__ node_ptr,
    const int32_t *__restrict__ node_col, const V *__restrict__ vals) {
  constexpr int64_t AL = TmaCfg<V>::kAlign;
  const int lane = threadIdx.x & 31;
  StagePtrs<V> sp = stage_ptrs<V>(smem, s);
  const int64_t n_a = tile * kTile;
  const int64_t n_b = (n_a + kTile < n_nodes) ? n_a + kTile : n_nodes;
  const int nn = (int)(n_b - n_a);
  int32_t my = (lane < nn) ? __ldg(&node_ptr[n_a + lane]) : 0;
  const int32_t s_b = __ldg(&node_ptr[n_b]);
  if (lane >= nn) my = s_b;
  sp.nptr[lane] = my;  // slots nn..31 hold the end pointer
  const int32_t s_a = __shfl_sync(0xffffffffu, my, 0);
  __syncwarp();
  if (lane == 0) {
    // 16-byte aligned windows around the tile's values / block columns
    const int64_t v_lo = ((int64_t)9 * s_a) & ~(AL - 1);
    const int64_t v_hi = ((int64_t)9 * s_b + AL - 1) & ~(AL - 1);
    const int64_t v_end = (int64_t)9 * n_blocks;
    const int64_t v_bulk_hi = v_hi <= v_end ? v_hi : (v_end & ~(AL - 1));
    const int64_t c_lo = (int64_t)s_a - (s_a & 3);
    const int64_t c_hi = ((int64_t)s_b + 3) & ~(int64_t)3;
    const int64_t c_bulk_hi = c_hi <= n_blocks ? c_hi : (n_blocks & ~(int64_t)3);
    const uint32_t vbytes =
        v_bulk_hi > v_lo ? (uint32_t)((v_bulk_hi - v_lo) * (int64_t)sizeof(V)) : 0u;
    const uint32_t cbytes = c_bulk_hi > c_lo ? (uint32_t)((c_bulk_hi - c_lo) * 4) : 0u;
    // tails that would run past the arrays: plain loads (at most 1 / 3 items)
    for (int64_t k = (v_bulk_hi > v_lo ? v_bulk_hi : v_lo); k < (int64_t)9 * s_b; ++k)
      sp.vals[k - v_lo] = vals[k];
    for (int64_t k = (c_bulk_hi > c_lo ? c_bulk_hi : c_lo); k < s_b; ++k)
      sp.cols[k - c_lo] = node_col[k];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(&full[s], vbytes + cbytes);
    if (vbytes) bulk_g2s(sp.vals, vals + v_lo, vbytes, &full[s]);
    if (cbytes) bulk_g2s(sp.cols, node_col + c_lo, cbytes, &full[s]);
  }
}

// CTAs per SM: the single-precision stages are half the size and the kernel is
// bound by the latency of the x gathers, so a third CTA per SM pays there
template <typename V> constexpr int tma_ctas_per_sm() { return sizeof(V) == 4 ? 3 : 2; }

template <bool DOT, typename V = double>
__global__ void __launch_bounds__(kBlock, tma_ctas_per_sm<V>())
    spmv_bsr3_tma_kernel(int64_t n_nodes, int64_t n_blocks,
                         const int32_t *__restrict__ node_ptr,
                         const int32_t *__restrict__ node_col,
                         const V *__restrict__ vals,
                         const double *__restrict__ x, double *__restrict__ y,
                         const double *__restrict__ dotv, double *partials,
                         unsigned int *ticket, double *dot_out,
                         const PcgScalars *S, const JacobiEpi J) {
  if (S && S->rr <= S->tol2) return;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = (uint64_t *)(smem + (size_t)kStages * TmaCfg<V>::kStageBytes);
  uint64_t *empty = full + kStages;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t n_tiles = (n_nodes + kTile - 1) / kTile;
  // contiguous run of tiles per CTA (x stays hot in L1 between mesh lines)
  const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t t_lo = (int64_t)blockIdx.x * per_cta;
  const int64_t t_hi = (t_lo + per_cta < n_tiles) ? t_lo + per_cta : n_tiles;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double dot = 0.0;
  if (wid == kProducerWarp) {
    int stage = 0;
    uint32_t parity = 1;  // a fresh barrier passes a wait on the previous phase
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
      mbar_wait(&empty[stage], parity);
      issue_tile<V>(smem, full, stage, tile, n_nodes, n_blocks, node_ptr, node_col, vals);
      if (++stage == kStages) {
        stage = 0;
        parity ^= 1u;
      }
    }
  } else if (wid < kConsumerWarps) {
    // four threads per row: 8 rows per warp, 48 rows (16 nodes) per tile
    const int quarter = lane & 3;
    const int row_in_tile = wid * 8 + (lane >> 2);
    const int ln = row_in_tile / 3;       // node within the tile
    const int ri = row_in_tile - 3 * ln;  // row within the node
    int stage = 0;
    uint32_t parity = 0;
    for (int64_t tile = t_lo; tile < t_hi; ++tile) {
      mbar_wait(&full[stage], parity);
      StagePtrs<V> sp = stage_ptrs<V>(smem, stage);
      const int64_t n_a = tile * kTile;
      const int nn = (int)((n_a + kTile < n_nodes ? n_a + kTile : n_nodes) - n_a);
      const int32_t s_a = sp.nptr[0];
      const int64_t v_lo = ((int64_t)9 * s_a) & ~((int64_t)TmaCfg<V>::kAlign - 1);
      const int64_t c_lo = (int64_t)s_a - (s_a & 3);
      double acc = 0.0;
      if (ln < nn) {
        const int32_t s0 = sp.nptr[ln];
        const int32_t deg = sp.nptr[ln + 1] - s0;
        const int32_t b0 = (deg * quarter) >> 2;
        const int32_t b1 = (deg * (quarter + 1)) >> 2;
        const V *vp = sp.vals + ((int64_t)9 * s0 - v_lo) + (int64_t)ri * 3 * deg;
        const int32_t *cp = sp.cols + ((int64_t)s0 - c_lo);
        double acc1 = 0.0, acc2 = 0.0;
#pragma unroll 4
        for (int32_t b = b0; b < b1; ++b) {
          const double *xb = x + (int64_t)3 * cp[b];
          const V *vb = vp + 3 * b;
          acc += (double)vb[0] * __ldg(&xb[0]);
          acc1 += (double)vb[1] * __ldg(&xb[1]);
          acc2 += (double)vb[2] * __ldg(&xb[2]);
        }
        acc += acc1 + acc2;
      }
      // this warp has finished reading the stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (quarter == 0 && ln < nn) {
        const int64_t r = 3 * (n_a + ln) + ri;
        if (J.b)  // fused Jacobi sweep
          acc = (J.xo ? J.xo : x)[r] + J.omega * J.dinv[r] * (J.b[r] - acc);
        y[r] = acc;
        if (DOT) dot += acc * dotv[r];
      }
      if (++stage == kStages) {
        stage = 0;
        parity ^= 1u;
      }
    }
  }
  if (DOT) {
    double v[1] = {dot};
    grid_reduce<1>(v, partials, ticket, dot_out);
  }
}

bool g_attr_set[4] = {false, false, false, false};

}  // namespace

// returns 0 on success, -1 if the layout is not eligible (caller falls back)
template <typename V>
static int launch_tma(int64_t n_nodes, int64_t n_blocks, int max_deg,
                      const int32_t *node_ptr, const int32_t *node_col,
                      const V *vals, const double *x, double *y,
                      const double *dotv, ReduceScratch *rs, double *dot_out,
                      const PcgScalars *S, const JacobiEpi &epi, cudaStream_t st) {
  if (max_deg > kMaxDeg || n_nodes < 8 * kTile) return -1;
  constexpr int kSmem = TmaCfg<V>::kSmemBytes;
  const int64_t n_tiles = (n_nodes + kTile - 1) / kTile;
  int64_t g = (int64_t)kNumSM * tma_ctas_per_sm<V>();
  if (g > n_tiles) g = n_tiles;
  const int grid = (int)g;
  const int which = (dotv ? 1 : 0) + (sizeof(V) == 4 ? 2 : 0);
  if (!g_attr_set[which]) {
    if (dotv)
      SKTB_CUDA_OK(cudaFuncSetAttribute(spmv_bsr3_tma_kernel<true, V>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    else
      SKTB_CUDA_OK(cudaFuncSetAttribute(spmv_bsr3_tma_kernel<false, V>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    g_attr_set[which] = true;
  }
  if (dotv)
    spmv_bsr3_tma_kernel<true, V><<<grid, kBlock, kSmem, st>>>(
        n_nodes, n_blocks, node_ptr, node_col, vals, x, y, dotv, rs->partials,
        rs->ticket, dot_out, S, epi);
  else
    spmv_bsr3_tma_kernel<false, V><<<grid, kBlock, kSmem, st>>>(
        n_nodes, n_blocks, node_ptr, node_col, vals, x, y, nullptr, nullptr,
        nullptr, nullptr, S, epi);
  SKTB_KERNEL_OK();
  return 0;
}

int launch_spmv_bsr3_tma(int64_t n_nodes, int64_t n_blocks, int max_deg,
                         const int32_t *node_ptr, const int32_t *node_col,
                         const double *vals, const double *x, double *y,
                         const double *dotv, ReduceScratch *rs, double *dot_out,
                         const PcgScalars *S, cudaStream_t st) {
  return launch_tma<double>(n_nodes, n_blocks, max_deg, node_ptr, node_col, vals, x, y, dotv,
                            rs, dot_out, S, JacobiEpi(), st);
}

int launch_spmv_bsr3_tma_jacobi(int64_t n_nodes, int64_t n_blocks, int max_deg,
                                const int32_t *node_ptr, const int32_t *node_col,
                                const double *vals, const double *x, double *y,
                                const JacobiEpi &epi, cudaStream_t st) {
  return launch_tma<double>(n_nodes, n_blocks, max_deg, node_ptr, node_col, vals, x, y,
                            nullptr, nullptr, nullptr, nullptr, epi, st);
}

// single-precision values (multigrid levels): plain product, or a fused Jacobi
// sweep when epi.b is set
int launch_spmv_bsr3_tma_f32(int64_t n_nodes, int64_t n_blocks, int max_deg,
                             const int32_t *node_ptr, const int32_t *node_col,
                             const float *vals, const double *x, double *y,
                             const JacobiEpi &epi, cudaStream_t st) {
  return launch_tma<float>(n_nodes, n_blocks, max_deg, node_ptr, node_col, vals, x, y,
                           nullptr, nullptr, nullptr, nullptr, epi, st);
}

// out[i] = (float) in[i]
__global__ void __launch_bounds__(kBlock)
    f64_to_f32_kernel(int64_t n, const double *__restrict__ in, float *__restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride) {
    if (i + 1 < n) {
      const double2 v = *reinterpret_cast<const double2 *>(in + i);
      *reinterpret_cast<float2 *>(out + i) = make_float2((float)v.x, (float)v.y);
    } else {
      out[i] = (float)in[i];
    }
  }
}

extern "C" int sktb_f64_to_f32(int64_t n, const double *in, float *out, void *stream) {
  SKTB_REQUIRE(in && out && n >= 0, "null argument");
  SKTB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, "unaligned buffer");
  if (n == 0) return 0;
  f64_to_f32_kernel<<<grid_for((n + 1) / 2), kBlock, 0, (cudaStream_t)stream>>>(n, in, out);
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_spmv_bsr3_tma_f32(int64_t n_nodes, int64_t n_blocks, int max_deg,
                                      const int32_t *node_ptr, const int32_t *node_col,
                                      const float *vals, const double *x, double *y,
                                      void *stream) {
  SKTB_REQUIRE(node_ptr && node_col && vals && x && y, "null argument");
  int rc = launch_spmv_bsr3_tma_f32(n_nodes, n_blocks, max_deg, node_ptr, node_col, vals, x, y,
                                    JacobiEpi(), (cudaStream_t)stream);
  if (rc == -1) {
    sktb::set_error("layout not eligible for the bulk-async SpMV (max_deg > 27 or tiny)");
    return 2;
  }
  return rc;
}

extern "C" int sktb_spmv_bsr3_tma(int64_t n_nodes, int64_t n_blocks, int max_deg,
                                  const int32_t *node_ptr,
                                  const int32_t *node_col, const double *vals,
                                  const double *x, double *y, void *stream) {
  SKTB_REQUIRE(node_ptr && node_col && vals && x && y, "null argument");
  int rc = launch_spmv_bsr3_tma(n_nodes, n_blocks, max_deg, node_ptr, node_col,
                                vals, x, y, nullptr, nullptr, nullptr, nullptr,
                                (cudaStream_t)stream);
  if (rc == -1) {
    sktb::set_error("layout not eligible for the bulk-async SpMV (max_deg > 27 or tiny)");
    return 2;
  }
  return rc;
}
