// Small dense fp64 products for the direct (fast-diagonalisation) Helmholtz
// solve (filters/_fastdiag.py; stands in for the sparse LU of the reference's
// filters/helmholtz_filter_nodal.py:121-157).  The solve is six products of the
// nodal field, viewed as a (npz, npx, npy) array, with the small eigenvector
// matrices of the three axes (K <= a few hundred), so one batched, strided,
// shared-memory-tiled DGEMM kernel covers all of them:
//
//   C[b] = A[b] . B[b]   (* S elementwise, optional),   row-major, b < batch
//
// 64 x 64 output tile per CTA, 16-deep k-slices staged in shared memory, 4 x 4
// register block per thread (16 DFMA per 8 LDS).  tcgen05 has no fp64 path and
// the DMMA rate on B200 equals the DFMA rate, so plain FMAs are the right tool.
// Accumulation order is fixed (k ascending): deterministic.
#include "common.cuh"

using namespace sktb;

namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
    dgemm_tile_kernel(int M, int N, int K, const double *__restrict__ A, int lda, int64_t sA,
                      const double *__restrict__ B, int ldb, int64_t sB, double *__restrict__ C,
                      int ldc, int64_t sC, const double *__restrict__ S, int64_t sS) {
  __shared__ double As[BK][BM + 1];
  __shared__ double Bs[BK][BN];
  const int b = blockIdx.z;
  A += (int64_t)b * sA;
  B += (int64_t)b * sB;
  C += (int64_t)b * sC;
  if (S) S += (int64_t)b * sS;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
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
      double a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      double v = acc[i][j];
      if (S) v *= S[(int64_t)gm * ldc + gn];
      C[(int64_t)gm * ldc + gn] = v;
    }
  }
}

}  // namespace

extern "C" int sktb_dgemm_batched(int M, int N, int K, const double *A, int lda,
                                  int64_t stride_a, const double *B, int ldb, int64_t stride_b,
                                  double *C, int ldc, int64_t stride_c, int batch,
                                  const double *scale, int64_t stride_s, void *stream) {
  SKTB_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "bad argument");
  SKTB_REQUIRE(lda >= K && ldb >= N && ldc >= N, "leading dimension too small");
  SKTB_REQUIRE(batch <= 65535 && (M + BM - 1) / BM <= 65535, "grid too large");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  dgemm_tile_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, lda, stride_a, B, ldb,
                                                           stride_b, C, ldc, stride_c, scale,
                                                           stride_s);
  SKTB_KERNEL_OK();
  return 0;
}
