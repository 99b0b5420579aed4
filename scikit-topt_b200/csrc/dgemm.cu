// Small dense fp64 products for the direct (fast-diagonalisation) Helmholtz
// solve (filters/_fastdiag.py; stands in for the sparse LU of the reference's
// filters/helmholtz_filter_nodal.py:121-157).  The solve is six products of the
// nodal field, viewed as a (npz, npx, npy) array, with the small eigenvector
// matrices of the three axes (K <= a few hundred), so one batched, strided,
// shared-memory-tiled DGEMM kernel covers all of them:
//
//   C[b] = A[b] . B[b]   (* S elementwise, optional),   row-major, b < batch
//
// (16 TM) x (16 TN) output tile per CTA with the register block (TM, TN) chosen
// per product, 16-deep k-slices staged in shared memory.  tcgen05 has no fp64 path and
// the DMMA rate on B200 equals the DFMA rate, so plain FMAs are the right tool.
// Accumulation order is fixed (k ascending): deterministic.
#include "common.cuh"

using namespace sktb;

namespace {

constexpr int BK = 16;

// TM x TN register block per thread, 16 x 16 threads: (16 TM) x (16 TN) output
// tile.  The host picks (TM, TN) per product so that the padded tile grid wastes
// as little as possible on the small dimensions of these products (71, 105, 140
// at C2: a fixed 64 x 64 tile would idle up to 45 % of its threads).
template <int TM, int TN>
__global__ void __launch_bounds__(256)
    dgemm_tile_kernel(int M, int N, int K, const double *__restrict__ A, int lda, int64_t sA,
                      const double *__restrict__ B, int ldb, int64_t sB, double *__restrict__ C,
                      int ldc, int64_t sC, const double *__restrict__ S, int64_t sS) {
  constexpr int BM = 16 * TM, BN = 16 * TN;
  __shared__ double As[BK][BM + 1];
  __shared__ double Bs[BK][BN];
  const int b = blockIdx.z;
  A += (int64_t)b * sA;
  B += (int64_t)b * sB;
  C += (int64_t)b * sC;
  if (S) S += (int64_t)b * sS;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  double acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    // A tile: BM x BK (row-major source: consecutive threads along k)
    for (int e = threadIdx.x; e < BM * BK; e += 256) {
      const int m = e / BK, k = e - m * BK;
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? __ldg(&A[(int64_t)gm * lda + gk]) : 0.0;
    }
    // B tile: BK x BN (consecutive threads along n: coalesced)
    for (int e = threadIdx.x; e < BK * BN; e += 256) {
      const int k = e / BN, n = e - k * BN;
      const int gk = k0 + k, gn = n0 + n;
      Bs[k][n] = (gk < K && gn < N) ? __ldg(&B[(int64_t)gk * ldb + gn]) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      double a[TM], bb[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      double v = acc[i][j];
      if (S) v *= S[(int64_t)gm * ldc + gn];
      C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

template <int TM, int TN>
void launch(dim3 grid, cudaStream_t st, int M, int N, int K, const double *A, int lda,
            int64_t sA, const double *B, int ldb, int64_t sB, double *C, int ldc, int64_t sC,
            const double *S, int64_t sS) {
  dgemm_tile_kernel<TM, TN><<<grid, 256, 0, st>>>(M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, S,
                                                  sS);
}

}  // namespace

extern "C" int sktb_dgemm_batched(int M, int N, int K, const double *A, int lda,
                                  int64_t stride_a, const double *B, int ldb, int64_t stride_b,
                                  double *C, int ldc, int64_t stride_c, int batch,
                                  const double *scale, int64_t stride_s, void *stream) {
  SKTB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "bad argument");
  SKTB_REQUIRE(lda >= K && ldb >= N && ldc >= N, "leading dimension too small");
  // register block (TM, TN) in {3, 4, 5} x {3, 4, 5, 7}: least padded work
  static const int tms[3] = {3, 4, 5}, tns[4] = {3, 4, 5, 7};
  int bm = 4, bn = 4;
  double best = 1e300;
  for (int tm : tms)
    for (int tn : tns) {
      const double pm = (double)((M + 16 * tm - 1) / (16 * tm)) * 16 * tm;
      const double pn = (double)((N + 16 * tn - 1) / (16 * tn)) * 16 * tn;
      // padded flops, mildly preferring larger blocks (more FMAs per shared-memory load)
      const double cost = pm * pn * (1.0 + 0.6 / tm + 0.6 / tn);
      if (cost < best) {
        best = cost;
        bm = tm;
        bn = tn;
      }
    }
  SKTB_REQUIRE(batch <= 65535 && (M + 16 * bm - 1) / (16 * bm) <= 65535, "grid too large");
  dim3 grid((N + 16 * bn - 1) / (16 * bn), (M + 16 * bm - 1) / (16 * bm), batch);
  cudaStream_t st = (cudaStream_t)stream;
#define SKTB_DG(TM, TN)                                                                     \
  if (bm == TM && bn == TN)                                                                 \
  launch<TM, TN>(grid, st, M, N, K, A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c,  \
                 scale, stride_s)
  SKTB_DG(3, 3); SKTB_DG(3, 4); SKTB_DG(3, 5); SKTB_DG(3, 7);
  SKTB_DG(4, 3); SKTB_DG(4, 4); SKTB_DG(4, 5); SKTB_DG(4, 7);
  SKTB_DG(5, 3); SKTB_DG(5, 4); SKTB_DG(5, 5); SKTB_DG(5, 7);
#undef SKTB_DG
  SKTB_KERNEL_OK();
  return 0;
}
