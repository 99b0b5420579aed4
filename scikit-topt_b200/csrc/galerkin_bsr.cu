// Algebraic Galerkin product A_c = P^T A P for node-block (3 dofs per node)
// operators whose nodes form a lattice (node = iy + npy*ix + npy*npx*iz) with
// ANY geometry and element type: jittered / graded hexahedra, Kuhn tetrahedra
// of MeshTet.init_tensor.  P is the trilinear interpolation in index space of
// sktb_mg_set_transfer, so A_c couples the 27 lattice neighbours of a coarse
// node whatever the fine stencil is (15 points for Kuhn tets, 27 for hexes).
//
// It stands in for the set-up of pyamg.smoothed_aggregation_solver(K) in the
// reference's cg_pyamg path (fea/solver_elastic.py:94-100) on meshes where the
// element-wise Galerkin kernels of mg.cu (uniform hexahedra) do not apply.
//
// Gather formulation, no atomics, fixed summation order (bit-reproducible):
// one CTA per coarse node I, nine warps, lane s < 27 of each owns the 3x3 block of
// the coarse neighbour J = I + (dz, dx, dy).  A warp walks a third of a third of
// the <= 27 fine nodes i that interpolate from I and, for each, the blocks (i, j)
// of the fine row; a lane adds w_iI * w_jJ * A_ij when j interpolates from its J.
// All lanes read the same A_ij (one broadcast load per value).
#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

namespace {

struct LatticeTransfer {
  int fnx, fny, fnz, cnx, cny, cnz;
  const int32_t *c0, *c1;   // by fine axis index, [x | y | z]
  const double *w0, *w1;
  const int32_t *fT;        // [3 slots][x | y | z] by coarse axis index, -1 = empty
  const double *wT;
};

// weight of fine axis index f (table offset off) in coarse axis index J
__device__ __forceinline__ double axis_weight(const LatticeTransfer &T, int off, int f, int J) {
  const int a = __ldg(&T.c0[off + f]), b = __ldg(&T.c1[off + f]);
  double w = 0.0;
  if (a == J) w += __ldg(&T.w0[off + f]);
  if (b == J && b != a) w += __ldg(&T.w1[off + f]);
  return w;
}

constexpr int kGalWarps = 9;   // one warp per (sz, sx) pair of fine parent slots

__global__ void __launch_bounds__(32 * kGalWarps)
    galerkin_bsr3_lattice_kernel(const LatticeTransfer T, const int32_t *__restrict__ fptr,
                                 const int32_t *__restrict__ fcol,
                                 const double *__restrict__ fvals,
                                 const uint8_t *__restrict__ fmask,
                                 const int32_t *__restrict__ cptr,
                                 const int32_t *__restrict__ ccol,
                                 const uint8_t *__restrict__ cmask, double *__restrict__ cvals) {
  // CTA = one coarse node I.  Warp w walks the fine parents with (sz, sx) =
  // (w / 3, w % 3) (<= 3 fine nodes, <= 81 blocks); its lanes own the 27 coarse
  // neighbours.  The nine partial blocks of a neighbour are then added in warp
  // order (fixed order: bit-reproducible).  A serial walk of all 27 parents by one
  // warp is latency bound on the small levels (729 dependent steps).
  __shared__ double part[kGalWarps][27][9];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n_coarse = T.cnx * T.cny * T.cnz;
  const int tot = T.cnx + T.cny + T.cnz;
  const int dy = lane % 3 - 1, dx = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
  const int sz = wid / 3, sx = wid % 3;
  for (int I = blockIdx.x; I < n_coarse; I += gridDim.x) {
    const int Iy = I % T.cny, Ix = (I / T.cny) % T.cnx, Iz = I / (T.cny * T.cnx);
    const int Jx = Ix + dx, Jy = Iy + dy, Jz = Iz + dz;
    const bool live = lane < 27 && Jx >= 0 && Jx < T.cnx && Jy >= 0 && Jy < T.cny && Jz >= 0 &&
                      Jz < T.cnz;
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    const int fz = __ldg(&T.fT[sz * tot + T.cnx + T.cny + Iz]);
    const int fx = __ldg(&T.fT[sx * tot + Ix]);
    if (fz >= 0 && fx >= 0) {
      const double wzx = __ldg(&T.wT[sz * tot + T.cnx + T.cny + Iz]) * __ldg(&T.wT[sx * tot + Ix]);
      for (int sy = 0; sy < 3; ++sy) {
        const int fy = __ldg(&T.fT[sy * tot + T.cnx + Iy]);
        if (fy < 0) continue;
        const double wI = wzx * __ldg(&T.wT[sy * tot + T.cnx + Iy]);
        const int i = fy + T.fny * (fx + T.fnx * fz);
        const int s0 = __ldg(&fptr[i]), deg = __ldg(&fptr[i + 1]) - s0;
        const double *vp = fvals + (int64_t)9 * s0;
        bool ri[3] = {false, false, false};   // fixed rows of node i contribute nothing
        if (fmask) {
          ri[0] = fmask[3 * i] != 0;
          ri[1] = fmask[3 * i + 1] != 0;
          ri[2] = fmask[3 * i + 2] != 0;
        }
        if (ri[0] && ri[1] && ri[2]) continue;
        for (int k = 0; k < deg; ++k) {
          const int j = __ldg(&fcol[s0 + k]);
          const int jy = j % T.fny, jr = j / T.fny;
          const int jx = jr % T.fnx, jz = jr / T.fnx;
          double w = 0.0;
          if (live) {
            w = axis_weight(T, T.fnx + T.fny, jz, Jz);
            if (w != 0.0) w *= axis_weight(T, 0, jx, Jx);
            if (w != 0.0) w *= axis_weight(T, T.fnx, jy, Jy);
          }
          if (__ballot_sync(0xffffffffu, w != 0.0) == 0u) continue;
          w *= wI;
          bool cj[3] = {false, false, false};
          if (fmask) {
            cj[0] = fmask[3 * j] != 0;
            cj[1] = fmask[3 * j + 1] != 0;
            cj[2] = fmask[3 * j + 2] != 0;
          }
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            if (ri[a]) continue;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
              if (cj[b]) continue;
              acc[3 * a + b] += w * __ldg(&vp[(int64_t)a * 3 * deg + 3 * k + b]);
            }
          }
        }
      }
    }
    if (lane < 27) {
#pragma unroll
      for (int q = 0; q < 9; ++q) part[wid][lane][q] = acc[q];
    }
    __syncthreads();
    if (threadIdx.x < 243) {
      const int s = threadIdx.x / 9, q = threadIdx.x % 9;
      const int a = q / 3, b = q % 3;
      const int ox = Ix + (s / 3) % 3 - 1, oy = Iy + s % 3 - 1, oz = Iz + s / 9 - 1;
      if (ox >= 0 && ox < T.cnx && oy >= 0 && oy < T.cny && oz >= 0 && oz < T.cnz) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kGalWarps; ++w) v += part[w][s][q];
        const int J = oy + T.cny * (ox + T.cnx * oz);
        const int c0 = __ldg(&cptr[I]), cdeg = __ldg(&cptr[I + 1]) - c0;
        int kc = -1;
        for (int k = 0; k < cdeg; ++k)
          if (__ldg(&ccol[c0 + k]) == J) {
            kc = k;
            break;
          }
        if (kc >= 0) {
          const bool diag = (J == I) && (a == b);
          if (cmask && (cmask[3 * I + a] || cmask[3 * J + b])) v = diag ? 1.0 : 0.0;
          if (diag && !(v > 0.0)) v = 1.0;   // a coarse dof nothing interpolates from
          cvals[(int64_t)9 * c0 + (int64_t)a * 3 * cdeg + 3 * kc + b] = v;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int sktb_galerkin_bsr3_lattice(
    const int32_t *fine_np_h, const int32_t *coarse_np_h, const int32_t *ax_c0,
    const int32_t *ax_c1, const double *ax_w0, const double *ax_w1, const int32_t *axT_f,
    const double *axT_w, const int32_t *fine_node_ptr, const int32_t *fine_node_col,
    const double *fine_vals, const uint8_t *fine_mask, const int32_t *coarse_node_ptr,
    const int32_t *coarse_node_col, const uint8_t *coarse_mask, double *coarse_vals,
    void *stream) {
  SKTB_REQUIRE(fine_np_h && coarse_np_h && ax_c0 && ax_c1 && ax_w0 && ax_w1 && axT_f && axT_w,
               "null transfer table");
  SKTB_REQUIRE(fine_node_ptr && fine_node_col && fine_vals && coarse_node_ptr &&
                   coarse_node_col && coarse_vals,
               "null operator argument");
  LatticeTransfer T;
  T.fnx = fine_np_h[0];
  T.fny = fine_np_h[1];
  T.fnz = fine_np_h[2];
  T.cnx = coarse_np_h[0];
  T.cny = coarse_np_h[1];
  T.cnz = coarse_np_h[2];
  SKTB_REQUIRE(T.fnx > 0 && T.fny > 0 && T.fnz > 0 && T.cnx > 0 && T.cny > 0 && T.cnz > 0,
               "bad lattice size");
  SKTB_REQUIRE((int64_t)T.fnx * T.fny * T.fnz < (int64_t)1 << 30, "lattice too large");
  T.c0 = ax_c0;
  T.c1 = ax_c1;
  T.w0 = ax_w0;
  T.w1 = ax_w1;
  T.fT = axT_f;
  T.wT = axT_w;
  const int64_t n_coarse = (int64_t)T.cnx * T.cny * T.cnz;
  const int64_t cap = (int64_t)kNumSM * 32;
  galerkin_bsr3_lattice_kernel<<<(int)(n_coarse < cap ? n_coarse : cap), 32 * kGalWarps, 0,
                                 (cudaStream_t)stream>>>(
      T, fine_node_ptr, fine_node_col, fine_vals, fine_mask, coarse_node_ptr, coarse_node_col,
      coarse_mask, coarse_vals);
  SKTB_KERNEL_OK();
  return 0;
}
