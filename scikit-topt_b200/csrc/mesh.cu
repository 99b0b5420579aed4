// Mesh-derived structures and the element-batched kernels that use them:
// node graph / contributor lists (host build, device resident), dof pattern
// expansion, unit element matrices, gather assembly (K2), element energy (K7)
// and the Helmholtz transfer operators (K9).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace sktb {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
std::atomic<long long> g_launch_count{0};

static std::mutex g_scratch_mu;
static ReduceScratch g_scratch[16];
static bool g_scratch_init[16] = {false};

int reduce_scratch_get(ReduceScratch **out) {
  int dev = 0;
  SKTB_CUDA_OK(cudaGetDevice(&dev));
  SKTB_REQUIRE(dev >= 0 && dev < 16, "device index out of range");
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  ReduceScratch &s = g_scratch[dev];
  if (!g_scratch_init[dev]) {
    SKTB_CUDA_OK(cudaMalloc(&s.partials, sizeof(double) * ReduceScratch::kMaxVals *
                                            ReduceScratch::kMaxBlocks));
    SKTB_CUDA_OK(cudaMalloc(&s.ticket, sizeof(unsigned int)));
    SKTB_CUDA_OK(cudaMemset(s.ticket, 0, sizeof(unsigned int)));
    SKTB_CUDA_OK(cudaMalloc(&s.result, sizeof(double) * 8));
    SKTB_CUDA_OK(cudaMallocHost(&s.result_h, sizeof(double) * 8));
    g_scratch_init[dev] = true;
  }
  *out = &s;
  return 0;
}

}  // namespace sktb

using namespace sktb;

struct sktb_mesh {
  int elem_type = 0;
  int nen = 0;
  int device = 0;
  int64_t n_elem = 0, n_nodes = 0, node_nnz = 0, n_contrib = 0;
  // device arrays
  int32_t *conn = nullptr;        // [nen][n_elem]
  double *coords = nullptr;       // [3][n_nodes]
  int32_t *node_ptr = nullptr;    // [n_nodes+1]
  int32_t *node_col = nullptr;    // [node_nnz]
  int32_t *n2e_ptr = nullptr;     // [n_nodes+1]
  int32_t *n2e_elem = nullptr;    // [nen*n_elem]
  uint8_t *n2e_loc = nullptr;     // [nen*n_elem]
  int32_t *pair_ptr = nullptr;    // [node_nnz+1]
  int32_t *contrib_elem = nullptr;  // [nen*nen*n_elem]
  uint8_t *contrib_ab = nullptr;    // [nen*nen*n_elem]  a*nen+b
  // host copies of the node graph (for pattern export)
  std::vector<int32_t> node_ptr_h, node_col_h;
};

extern "C" const char *sktb_last_error(void) { return sktb::g_err.c_str(); }
extern "C" int sktb_version(void) { return 100; }
extern "C" int64_t sktb_launch_count(void) {
  return (int64_t)sktb::g_launch_count.load();
}

template <typename T>
static int upload(T **dst, const std::vector<T> &src) {
  size_t bytes = sizeof(T) * (src.size() ? src.size() : 1);
  SKTB_CUDA_OK(cudaMalloc(dst, bytes));
  if (!src.empty())
    SKTB_CUDA_OK(cudaMemcpy(*dst, src.data(), sizeof(T) * src.size(),
                            cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int sktb_mesh_create(sktb_mesh **out, int elem_type, int64_t n_elem,
                                int64_t n_nodes, const int32_t *conn_h,
                                const double *coords_h, int device) {
  SKTB_REQUIRE(out && conn_h && coords_h, "null argument");
  SKTB_REQUIRE(elem_type == SKTB_HEX8 || elem_type == SKTB_TET4,
               "elem_type must be SKTB_HEX8 or SKTB_TET4");
  SKTB_REQUIRE(n_elem > 0 && n_nodes > 0, "empty mesh");
  const int nen = elem_type == SKTB_HEX8 ? 8 : 4;
  SKTB_REQUIRE((int64_t)nen * nen * n_elem < (int64_t)2147483647,
               "mesh too large for int32 contributor indexing");
  SKTB_CUDA_OK(cudaSetDevice(device));
  for (int64_t i = 0; i < (int64_t)nen * n_elem; ++i)
    SKTB_REQUIRE(conn_h[i] >= 0 && conn_h[i] < n_nodes,
                 "connectivity index out of range");

  sktb_mesh *m = new sktb_mesh();
  m->elem_type = elem_type;
  m->nen = nen;
  m->device = device;
  m->n_elem = n_elem;
  m->n_nodes = n_nodes;

  // ---- node -> element adjacency (elements ascending per node)
  std::vector<int32_t> n2e_ptr(n_nodes + 1, 0);
  for (int a = 0; a < nen; ++a)
    for (int64_t e = 0; e < n_elem; ++e) n2e_ptr[conn_h[a * n_elem + e] + 1]++;
  for (int64_t n = 0; n < n_nodes; ++n) n2e_ptr[n + 1] += n2e_ptr[n];
  std::vector<int32_t> n2e_elem((size_t)nen * n_elem);
  std::vector<uint8_t> n2e_loc((size_t)nen * n_elem);
  {
    std::vector<int32_t> fill(n2e_ptr.begin(), n2e_ptr.end() - 1);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int a = 0; a < nen; ++a) {
        int32_t n = conn_h[a * n_elem + e];
        int32_t k = fill[n]++;
        n2e_elem[k] = (int32_t)e;
        n2e_loc[k] = (uint8_t)a;
      }
  }
  // ---- node graph: sorted unique neighbour nodes
  std::vector<int32_t> &node_ptr = m->node_ptr_h;
  std::vector<int32_t> &node_col = m->node_col_h;
  node_ptr.assign(n_nodes + 1, 0);
  node_col.reserve((size_t)n_nodes * (nen == 8 ? 27 : 16));
  {
    std::vector<int32_t> cand;
    for (int64_t n = 0; n < n_nodes; ++n) {
      cand.clear();
      for (int32_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
        int32_t e = n2e_elem[k];
        for (int b = 0; b < nen; ++b) cand.push_back(conn_h[b * n_elem + e]);
      }
      std::sort(cand.begin(), cand.end());
      cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
      node_col.insert(node_col.end(), cand.begin(), cand.end());
      SKTB_REQUIRE(node_col.size() < (size_t)2147483647 / 9,
                   "node graph too large for int32 CSR indexing");
      node_ptr[n + 1] = (int32_t)node_col.size();
    }
  }
  m->node_nnz = (int64_t)node_col.size();
  // ---- contributor lists per (row node, col node) pair
  std::vector<int32_t> pair_ptr(m->node_nnz + 1, 0);
  auto slot_of = [&](int64_t n, int32_t mnode) -> int32_t {
    const int32_t *b = node_col.data() + node_ptr[n];
    const int32_t *e = node_col.data() + node_ptr[n + 1];
    return (int32_t)(std::lower_bound(b, e, mnode) - node_col.data());
  };
  for (int64_t n = 0; n < n_nodes; ++n)
    for (int32_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
      int32_t e = n2e_elem[k];
      for (int b = 0; b < nen; ++b)
        pair_ptr[slot_of(n, conn_h[b * n_elem + e]) + 1]++;
    }
  for (int64_t i = 0; i < m->node_nnz; ++i) pair_ptr[i + 1] += pair_ptr[i];
  m->n_contrib = pair_ptr[m->node_nnz];
  std::vector<int32_t> contrib_elem((size_t)m->n_contrib);
  std::vector<uint8_t> contrib_ab((size_t)m->n_contrib);
  {
    std::vector<int32_t> fill(pair_ptr.begin(), pair_ptr.end() - 1);
    for (int64_t n = 0; n < n_nodes; ++n)
      for (int32_t k = n2e_ptr[n]; k < n2e_ptr[n + 1]; ++k) {
        int32_t e = n2e_elem[k];
        int a = n2e_loc[k];
        for (int b = 0; b < nen; ++b) {
          int32_t s = slot_of(n, conn_h[b * n_elem + e]);
          int32_t c = fill[s]++;
          contrib_elem[c] = e;
          contrib_ab[c] = (uint8_t)(a * nen + b);
        }
      }
  }
  // ---- upload
  std::vector<int32_t> conn_v(conn_h, conn_h + (size_t)nen * n_elem);
  std::vector<double> coords_v(coords_h, coords_h + (size_t)3 * n_nodes);
  int rc = 0;
  rc |= upload(&m->conn, conn_v);
  rc |= upload(&m->coords, coords_v);
  rc |= upload(&m->node_ptr, node_ptr);
  rc |= upload(&m->node_col, node_col);
  rc |= upload(&m->n2e_ptr, n2e_ptr);
  rc |= upload(&m->n2e_elem, n2e_elem);
  rc |= upload(&m->n2e_loc, n2e_loc);
  rc |= upload(&m->pair_ptr, pair_ptr);
  rc |= upload(&m->contrib_elem, contrib_elem);
  rc |= upload(&m->contrib_ab, contrib_ab);
  if (rc) {
    sktb_mesh_destroy(m);
    return 1;
  }
  *out = m;
  return 0;
}

extern "C" void sktb_mesh_destroy(sktb_mesh *m) {
  if (!m) return;
  cudaSetDevice(m->device);
  cudaFree(m->conn);
  cudaFree(m->coords);
  cudaFree(m->node_ptr);
  cudaFree(m->node_col);
  cudaFree(m->n2e_ptr);
  cudaFree(m->n2e_elem);
  cudaFree(m->n2e_loc);
  cudaFree(m->pair_ptr);
  cudaFree(m->contrib_elem);
  cudaFree(m->contrib_ab);
  delete m;
}

extern "C" int64_t sktb_mesh_node_nnz(const sktb_mesh *m) {
  return m ? m->node_nnz : -1;
}

extern "C" int sktb_mesh_node_graph_h(const sktb_mesh *m, int32_t *row_ptr_h,
                                      int32_t *col_idx_h) {
  SKTB_REQUIRE(m && row_ptr_h && col_idx_h, "null argument");
  std::memcpy(row_ptr_h, m->node_ptr_h.data(),
              sizeof(int32_t) * m->node_ptr_h.size());
  std::memcpy(col_idx_h, m->node_col_h.data(),
              sizeof(int32_t) * m->node_col_h.size());
  return 0;
}

// ------------------------------------------------------------ dof pattern --
// rows of the nodes [n_begin, n_end); row_ptr / col_idx are local to that range
template <int D>
__global__ void __launch_bounds__(kBlock)
    dof_pattern_kernel(int64_t n_begin, int64_t n_end,
                       const int32_t *__restrict__ node_ptr,
                       const int32_t *__restrict__ node_col,
                       int32_t *__restrict__ row_ptr,
                       int32_t *__restrict__ col_idx) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t off = (int64_t)D * D * node_ptr[n_begin];
  for (int64_t n = n_begin + warp; n < n_end; n += nwarps) {
    const int32_t s = node_ptr[n], deg = node_ptr[n + 1] - s;
    const int64_t base = (int64_t)D * D * s - off;
    if (lane < D)
      row_ptr[D * (n - n_begin) + lane] = (int32_t)(base + (int64_t)lane * D * deg);
    if (n == n_end - 1 && lane == 0)
      row_ptr[D * (n_end - n_begin)] =
          (int32_t)((int64_t)D * D * node_ptr[n_end] - off);
    for (int e = lane; e < D * D * deg; e += 32) {
      int q = e % (D * deg);
      col_idx[base + e] = D * node_col[s + q / D] + (q % D);
    }
  }
}

extern "C" int sktb_mesh_dof_pattern(const sktb_mesh *m, int dpn,
                                     int32_t *row_ptr, int32_t *col_idx,
                                     void *stream) {
  SKTB_REQUIRE(m && row_ptr && col_idx, "null argument");
  SKTB_REQUIRE(dpn == 1 || dpn == 3, "dpn must be 1 or 3");
  SKTB_REQUIRE((int64_t)dpn * dpn * m->node_nnz < (int64_t)2147483647,
               "nnz exceeds int32 CSR indexing");
  return sktb_mesh_dof_pattern_rows(m, dpn, 0, m->n_nodes, row_ptr, col_idx,
                                    stream);
}

extern "C" int sktb_mesh_dof_pattern_rows(const sktb_mesh *m, int dpn,
                                          int64_t node_begin, int64_t node_end,
                                          int32_t *row_ptr, int32_t *col_idx,
                                          void *stream) {
  SKTB_REQUIRE(m && row_ptr && col_idx, "null argument");
  SKTB_REQUIRE(dpn == 1 || dpn == 3, "dpn must be 1 or 3");
  SKTB_REQUIRE(0 <= node_begin && node_begin < node_end && node_end <= m->n_nodes,
               "bad node range");
  const int64_t nnz_nodes =
      (int64_t)m->node_ptr_h[node_end] - m->node_ptr_h[node_begin];
  SKTB_REQUIRE((int64_t)dpn * dpn * nnz_nodes < (int64_t)2147483647,
               "nnz exceeds int32 CSR indexing");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for((node_end - node_begin) * 32);
  if (dpn == 1)
    dof_pattern_kernel<1><<<grid, kBlock, 0, st>>>(
        node_begin, node_end, m->node_ptr, m->node_col, row_ptr, col_idx);
  else
    dof_pattern_kernel<3><<<grid, kBlock, 0, st>>>(
        node_begin, node_end, m->node_ptr, m->node_col, row_ptr, col_idx);
  SKTB_KERNEL_OK();
  return 0;
}

// ---------------------------------------------------------------- unit Ke --
// Reference-cube corner of hex local vertex a, bit d = coordinate d
// (v0=000 v1=001 v2=010 v3=100 v4=011 v5=101 v6=110 v7=111 as (X,Y,Z)).
__constant__ int c_hex_vx[8] = {0, 0, 0, 1, 0, 1, 1, 1};
__constant__ int c_hex_vy[8] = {0, 0, 1, 0, 1, 0, 1, 1};
__constant__ int c_hex_vz[8] = {0, 1, 0, 0, 1, 1, 0, 1};

__device__ __forceinline__ void hex_shape(int a, double X, double Y, double Z,
                                          double &N, double (&dN)[3]) {
  double fx = c_hex_vx[a] ? X : 1.0 - X, dx = c_hex_vx[a] ? 1.0 : -1.0;
  double fy = c_hex_vy[a] ? Y : 1.0 - Y, dy = c_hex_vy[a] ? 1.0 : -1.0;
  double fz = c_hex_vz[a] ? Z : 1.0 - Z, dz = c_hex_vz[a] ? 1.0 : -1.0;
  N = fx * fy * fz;
  dN[0] = dx * fy * fz;
  dN[1] = fx * dy * fz;
  dN[2] = fx * fy * dz;
}
__device__ __forceinline__ void tet_shape(int a, double X, double Y, double Z,
                                          double &N, double (&dN)[3]) {
  if (a == 0) {
    N = 1.0 - X - Y - Z;
    dN[0] = dN[1] = dN[2] = -1.0;
  } else {
    N = a == 1 ? X : (a == 2 ? Y : Z);
    dN[0] = a == 1 ? 1.0 : 0.0;
    dN[1] = a == 2 ? 1.0 : 0.0;
    dN[2] = a == 3 ? 1.0 : 0.0;
  }
}

// one block per geometry class, one thread per matrix entry
template <int NEN>
__global__ void unit_ke_kernel(int kind, double lam0, double mu0, int nqp,
                               const double *__restrict__ Xq,
                               const double *__restrict__ Wq,
                               const int32_t *__restrict__ class_rep,
                               const int32_t *__restrict__ conn, int64_t n_elem,
                               const double *__restrict__ coords,
                               int64_t n_nodes, double *__restrict__ out) {
  const int D = (kind == SKTB_KE_ELASTIC) ? 3 : 1;
  const int nde = NEN * D;
  __shared__ double xe[NEN][3];
  const int64_t cls = blockIdx.x;
  const int64_t el = class_rep[cls];
  if (threadIdx.x < NEN * 3) {
    int a = threadIdx.x / 3, d = threadIdx.x % 3;
    xe[a][d] = coords[(int64_t)d * n_nodes + conn[(int64_t)a * n_elem + el]];
  }
  __syncthreads();
  for (int ent = threadIdx.x; ent < nde * nde; ent += blockDim.x) {
    const int r = ent / nde, c = ent % nde;
    const int a = r / D, i = r % D, b = c / D, j = c % D;
    double acc = 0.0;
    for (int q = 0; q < nqp; ++q) {
      const double X = Xq[q], Y = Xq[nqp + q], Z = Xq[2 * nqp + q];
      // Jacobian J[d][k] = sum_a x_a[d] dN_a/dX_k
      double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      double Na = 0, Nb = 0, dNa[3], dNb[3];
#pragma unroll
      for (int v = 0; v < NEN; ++v) {
        double N, dN[3];
        if (NEN == 8)
          hex_shape(v, X, Y, Z, N, dN);
        else
          tet_shape(v, X, Y, Z, N, dN);
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
          for (int k = 0; k < 3; ++k) J[d][k] += xe[v][d] * dN[k];
        if (v == a) {
          Na = N;
          dNa[0] = dN[0], dNa[1] = dN[1], dNa[2] = dN[2];
        }
        if (v == b) {
          Nb = N;
          dNb[0] = dN[0], dNb[1] = dN[1], dNb[2] = dN[2];
        }
      }
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
      const double id = 1.0 / det;
      // inverse Jacobian iJ[k][d]
      double iJ[3][3];
      iJ[0][0] = c00 * id;
      iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
      iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
      iJ[1][0] = c01 * id;
      iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
      iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
      iJ[2][0] = c02 * id;
      iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
      iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
      // physical gradients g[d] = sum_k dN[k] * iJ[k][d]
      double ga[3], gb[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        ga[d] = dNa[0] * iJ[0][d] + dNa[1] * iJ[1][d] + dNa[2] * iJ[2][d];
        gb[d] = dNb[0] * iJ[0][d] + dNb[1] * iJ[1][d] + dNb[2] * iJ[2][d];
      }
      const double dx = Wq[q] * fabs(det);
      double f;
      if (kind == SKTB_KE_ELASTIC) {
        // test = (a,i), trial = (b,j):
        // lam tr e(u) tr e(v) + 2 mu e(u):e(v)
        const double gg = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
        f = lam0 * ga[i] * gb[j] + mu0 * ga[j] * gb[i] + (i == j ? mu0 * gg : 0.0);
      } else if (kind == SKTB_KE_LAPLACE) {
        f = ga[0] * gb[0] + ga[1] * gb[1] + ga[2] * gb[2];
      } else {
        f = Na * Nb;
      }
      acc += f * dx;
    }
    out[cls * nde * nde + ent] = acc;
  }
}

extern "C" int sktb_unit_ke(const sktb_mesh *m, int kind, double nu, int nqp,
                            const double *X_h, const double *W_h,
                            int64_t n_class, const int32_t *class_rep_h,
                            double *out, void *stream) {
  SKTB_REQUIRE(m && X_h && W_h && class_rep_h && out, "null argument");
  SKTB_REQUIRE(kind >= 0 && kind <= 2, "unknown unit-Ke kind");
  SKTB_REQUIRE(nqp > 0 && n_class > 0, "empty quadrature or class list");
  cudaStream_t st = (cudaStream_t)stream;
  double *Xd = nullptr, *Wd = nullptr;
  int32_t *rep = nullptr;
  SKTB_CUDA_OK(cudaMalloc(&Xd, sizeof(double) * 3 * nqp));
  SKTB_CUDA_OK(cudaMalloc(&Wd, sizeof(double) * nqp));
  SKTB_CUDA_OK(cudaMalloc(&rep, sizeof(int32_t) * n_class));
  SKTB_CUDA_OK(cudaMemcpyAsync(Xd, X_h, sizeof(double) * 3 * nqp,
                               cudaMemcpyHostToDevice, st));
  SKTB_CUDA_OK(cudaMemcpyAsync(Wd, W_h, sizeof(double) * nqp,
                               cudaMemcpyHostToDevice, st));
  SKTB_CUDA_OK(cudaMemcpyAsync(rep, class_rep_h, sizeof(int32_t) * n_class,
                               cudaMemcpyHostToDevice, st));
  const double lam0 = nu / ((1.0 + nu) * (1.0 - 2.0 * nu));
  const double mu0 = 1.0 / (2.0 * (1.0 + nu));
  const int64_t kMaxGrid = 1 << 30;
  SKTB_REQUIRE(n_class < kMaxGrid, "too many classes");
  if (m->nen == 8)
    unit_ke_kernel<8><<<(unsigned)n_class, 192, 0, st>>>(
        kind, lam0, mu0, nqp, Xd, Wd, rep, m->conn, m->n_elem, m->coords,
        m->n_nodes, out);
  else
    unit_ke_kernel<4><<<(unsigned)n_class, 64, 0, st>>>(
        kind, lam0, mu0, nqp, Xd, Wd, rep, m->conn, m->n_elem, m->coords,
        m->n_nodes, out);
  SKTB_KERNEL_OK();
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(Xd);
  cudaFree(Wd);
  cudaFree(rep);
  return 0;
}

// ---------------------------------------------------------------- assembly --
// One warp per row node; lanes stride over the D*D*deg values of the node's
// D rows in CSR order, so the stores are fully coalesced and every value is
// produced by a fixed-order gather over its contributor list (no atomics).
template <int D, int NEN>
__global__ void __launch_bounds__(kBlock)
    assemble_kernel(int64_t n_begin, int64_t n_end,
                    const int32_t *__restrict__ node_ptr,
                    const int32_t *__restrict__ node_col,
                    const int32_t *__restrict__ pair_ptr,
                    const int32_t *__restrict__ contrib_elem,
                    const uint8_t *__restrict__ contrib_ab,
                    const int32_t *__restrict__ elem_class,
                    const double *__restrict__ unit_ke,
                    const double *__restrict__ scale,
                    const uint8_t *__restrict__ mask,
                    double *__restrict__ vals) {
  constexpr int NDE = D * NEN;
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t off = (int64_t)D * D * node_ptr[n_begin];
  for (int64_t n = n_begin + warp; n < n_end; n += nwarps) {
    const int32_t s0 = node_ptr[n], deg = node_ptr[n + 1] - s0;
    const int64_t base = (int64_t)D * D * s0 - off;
    for (int e = lane; e < D * D * deg; e += 32) {
      const int i = e / (D * deg), q = e - i * (D * deg);
      const int s = q / D, j = q - s * D;
      const int32_t pair = s0 + s;
      double acc = 0.0;
      const int32_t c1 = pair_ptr[pair + 1];
      for (int32_t c = pair_ptr[pair]; c < c1; ++c) {
        const int32_t el = __ldg(&contrib_elem[c]);
        const int ab = __ldg(&contrib_ab[c]);
        const int a = ab / NEN, b = ab - a * NEN;
        const int64_t cls = elem_class ? (int64_t)__ldg(&elem_class[el]) : (int64_t)el;
        const double ke =
            __ldg(&unit_ke[cls * (NDE * NDE) + (D * a + i) * NDE + (D * b + j)]);
        acc += (scale ? __ldg(&scale[el]) : 1.0) * ke;
      }
      if (mask) {
        const int64_t r = (int64_t)D * n + i;
        const int64_t cd = (int64_t)D * node_col[pair] + j;
        if (mask[r] || mask[cd]) acc = (r == cd) ? 1.0 : 0.0;
      }
      vals[base + e] = acc;
    }
  }
}

extern "C" int sktb_assemble(const sktb_mesh *m, int dpn, const double *unit_ke,
                             const int32_t *elem_class, const double *scale,
                             const uint8_t *dir_mask, double *vals,
                             void *stream) {
  SKTB_REQUIRE(m, "null argument");
  return sktb_assemble_rows(m, dpn, 0, m->n_nodes, unit_ke, elem_class, scale,
                            dir_mask, vals, stream);
}

extern "C" int sktb_assemble_rows(const sktb_mesh *m, int dpn,
                                  int64_t node_begin, int64_t node_end,
                                  const double *unit_ke,
                                  const int32_t *elem_class,
                                  const double *scale, const uint8_t *dir_mask,
                                  double *vals, void *stream) {
  SKTB_REQUIRE(m && unit_ke && vals, "null argument");
  SKTB_REQUIRE(dpn == 1 || dpn == 3, "dpn must be 1 or 3");
  SKTB_REQUIRE(0 <= node_begin && node_begin < node_end && node_end <= m->n_nodes,
               "bad node range");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for((node_end - node_begin) * 32, kBlock, 16);
#define LAUNCH(D, NEN)                                                        \
  assemble_kernel<D, NEN><<<grid, kBlock, 0, st>>>(                           \
      node_begin, node_end, m->node_ptr, m->node_col, m->pair_ptr,            \
      m->contrib_elem, m->contrib_ab, elem_class, unit_ke, scale, dir_mask,   \
      vals)
  if (m->nen == 8 && dpn == 3)
    LAUNCH(3, 8);
  else if (m->nen == 8 && dpn == 1)
    LAUNCH(1, 8);
  else if (m->nen == 4 && dpn == 3)
    LAUNCH(3, 4);
  else
    LAUNCH(1, 4);
#undef LAUNCH
  SKTB_KERNEL_OK();
  return 0;
}

// ----------------------------------------------------------- element energy --
// One warp per element: lanes hold the element's dof values, the NDE*NDE
// quadratic form is split over lanes with operands exchanged by shuffles.
template <int D, int NEN>
__global__ void __launch_bounds__(kBlock)
    element_energy_kernel(int64_t n_elem, int64_t n_nodes_unused,
                          const int32_t *__restrict__ conn,
                          const int32_t *__restrict__ elem_class,
                          const double *__restrict__ unit_ke,
                          const double *__restrict__ scale,
                          const double *__restrict__ u,
                          const double *__restrict__ v, double factor,
                          double *__restrict__ out) {
  constexpr int NDE = D * NEN;
  static_assert(NDE <= 32, "element dofs must fit in one warp");
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e = warp; e < n_elem; e += nwarps) {
    double ue = 0.0, ve = 0.0;
    if (lane < NDE) {
      const int a = lane / D, c = lane - a * D;
      const int64_t dof = (int64_t)D * conn[(int64_t)a * n_elem + e] + c;
      ue = u[dof];
      ve = v ? v[dof] : ue;
    }
    // the stiffness / conduction element matrices annihilate translations:
    // subtracting local node 0 removes the cancellation a large mean value
    // (e.g. T ~ 600 K with 0.1 K variation) would cause in u^T Ke u
    ue -= __shfl_sync(0xffffffffu, ue, lane % D);
    ve -= __shfl_sync(0xffffffffu, ve, lane % D);
    if (lane >= NDE) ue = ve = 0.0;
    const int64_t cls = elem_class ? (int64_t)elem_class[e] : e;
    const double *ke = unit_ke + cls * (NDE * NDE);
    double acc = 0.0;
    // warp-uniform trip count (NDE*NDE need not be a multiple of 32: tets)
    for (int base = 0; base < NDE * NDE; base += 32) {
      const int ent = base + lane;
      const bool ok = ent < NDE * NDE;
      const int r = ok ? ent / NDE : 0, c = ok ? ent - r * NDE : 0;
      const double ur = __shfl_sync(0xffffffffu, ue, r);
      const double uc = __shfl_sync(0xffffffffu, ve, c);
      if (ok) acc += ur * __ldg(&ke[ent]) * uc;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[e] = factor * (scale ? scale[e] : 1.0) * acc;
  }
}

static int element_form(const sktb_mesh *m, int dpn, const double *unit_ke,
                        const int32_t *elem_class, const double *scale,
                        const double *u, const double *v, double factor,
                        double *out, void *stream);

extern "C" int sktb_element_energy(const sktb_mesh *m, int dpn,
                                   const double *unit_ke,
                                   const int32_t *elem_class,
                                   const double *scale, const double *u,
                                   double *out, void *stream) {
  return element_form(m, dpn, unit_ke, elem_class, scale, u, nullptr, 0.5, out,
                      stream);
}

extern "C" int sktb_element_bilinear(const sktb_mesh *m, int dpn,
                                     const double *unit_ke,
                                     const int32_t *elem_class,
                                     const double *scale, const double *u,
                                     const double *v, double factor,
                                     double *out, void *stream) {
  SKTB_REQUIRE(v, "null argument");
  return element_form(m, dpn, unit_ke, elem_class, scale, u, v, factor, out,
                      stream);
}

static int element_form(const sktb_mesh *m, int dpn, const double *unit_ke,
                        const int32_t *elem_class, const double *scale,
                        const double *u, const double *v, double factor,
                        double *out, void *stream) {
  SKTB_REQUIRE(m && unit_ke && u && out, "null argument");
  SKTB_REQUIRE(dpn == 1 || dpn == 3, "dpn must be 1 or 3");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->n_elem * 32, kBlock, 16);
#define LAUNCH(D, NEN)                                                        \
  element_energy_kernel<D, NEN><<<grid, kBlock, 0, st>>>(                     \
      m->n_elem, m->n_nodes, m->conn, elem_class, unit_ke, scale, u, v,       \
      factor, out)
  if (m->nen == 8 && dpn == 3)
    LAUNCH(3, 8);
  else if (m->nen == 8 && dpn == 1)
    LAUNCH(1, 8);
  else if (m->nen == 4 && dpn == 3)
    LAUNCH(3, 4);
  else
    LAUNCH(1, 4);
#undef LAUNCH
  SKTB_KERNEL_OK();
  return 0;
}

// K7 on meshes with ONE hexahedral geometry class (every tensor grid): one thread
// per element, the 24x24 unit matrix as kernel parameters (constant bank, like
// the grid operator), upper triangle only: U_e = scale_e sum_i u_i (K_ii u_i / 2
// + sum_{j>i} K_ij u_j) = 300 FMA per element instead of a warp per element.
struct Ke24 {
  double k[576];
};
__global__ void __launch_bounds__(kBlock, 3)
    hex_energy_uniform_kernel(const __grid_constant__ Ke24 K, int64_t n_elem,
                              const int32_t *__restrict__ conn,
                              const double *__restrict__ scale,
                              const double *__restrict__ u, double *__restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_elem; e += stride) {
    double ue[24];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int64_t nd = conn[(int64_t)a * n_elem + e];
#pragma unroll
      for (int c = 0; c < 3; ++c) ue[3 * a + c] = __ldg(&u[3 * nd + c]);
    }
    // the element matrix annihilates translations: remove local node 0 (same
    // cancellation guard as element_energy_kernel)
#pragma unroll
    for (int a = 7; a >= 0; --a)
#pragma unroll
      for (int c = 0; c < 3; ++c) ue[3 * a + c] -= ue[c];
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      double acc = 0.5 * K.k[i * 24 + i] * ue[i];
#pragma unroll
      for (int j = i + 1; j < 24; ++j) acc = fma(K.k[i * 24 + j], ue[j], acc);
      tot = fma(ue[i], acc, tot);
    }
    out[e] = (scale ? scale[e] : 1.0) * tot;
  }
}

extern "C" int sktb_element_energy_hex_uniform(const sktb_mesh *m,
                                               const double *unit_ke_h,
                                               const double *scale,
                                               const double *u, double *out,
                                               void *stream) {
  SKTB_REQUIRE(m && unit_ke_h && u && out, "null argument");
  SKTB_REQUIRE(m->nen == 8, "hexahedral meshes only");
  Ke24 K;
  for (int i = 0; i < 576; ++i) K.k[i] = unit_ke_h[i];
  hex_energy_uniform_kernel<<<grid_for(m->n_elem), kBlock, 0, (cudaStream_t)stream>>>(
      K, m->n_elem, m->conn, scale, u, out);
  SKTB_KERNEL_OK();
  return 0;
}

// ------------------------------------------------- Helmholtz transfer (K9) --
__global__ void __launch_bounds__(kBlock)
    e2n_kernel(int64_t n_nodes, const int32_t *__restrict__ n2e_ptr,
               const int32_t *__restrict__ n2e_elem,
               const double *__restrict__ w, const double *__restrict__ val,
               const uint8_t *__restrict__ design, double fixed_value,
               const double *__restrict__ wsum, double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_nodes; i += stride) {
    double acc = 0.0;
    const int32_t k1 = n2e_ptr[i + 1];
    for (int32_t k = n2e_ptr[i]; k < k1; ++k) {
      const int32_t e = n2e_elem[k];
      double v;
      if (val)
        v = (!design || design[e]) ? val[e] : fixed_value;
      else
        v = 1.0;
      acc += (w ? w[e] : 1.0) * v;
    }
    if (wsum)
      out[i] = acc / wsum[i];
    else
      out[i] = (acc == 0.0) ? 1.0 : acc;  // wsum==0 -> 1 (reference :54)
  }
}

extern "C" int sktb_e2n(const sktb_mesh *m, const double *w, const double *val,
                        const uint8_t *design, double fixed_value,
                        const double *wsum, double *out, void *stream) {
  SKTB_REQUIRE(m && val && wsum && out, "null argument");
  e2n_kernel<<<grid_for(m->n_nodes), kBlock, 0, (cudaStream_t)stream>>>(
      m->n_nodes, m->n2e_ptr, m->n2e_elem, w, val, design, fixed_value,
      wsum, out);
  SKTB_KERNEL_OK();
  return 0;
}

extern "C" int sktb_e2n_wsum(const sktb_mesh *m, const double *w,
                                 double *out, void *stream) {
  SKTB_REQUIRE(m && out, "null argument");
  e2n_kernel<<<grid_for(m->n_nodes), kBlock, 0, (cudaStream_t)stream>>>(
      m->n_nodes, m->n2e_ptr, m->n2e_elem, w, nullptr, nullptr, 0.0, nullptr,
      out);
  SKTB_KERNEL_OK();
  return 0;
}

template <int NEN>
__global__ void __launch_bounds__(kBlock)
    n2e_mean_kernel(int64_t n_elem, const int32_t *__restrict__ conn,
                    const double *__restrict__ x, int clamp_max0,
                    double *__restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_elem; e += stride) {
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < NEN; ++a) acc += __ldg(&x[conn[(int64_t)a * n_elem + e]]);
    acc /= (double)NEN;
    if (clamp_max0) acc = fmin(acc, 0.0);
    out[e] = acc;
  }
}

extern "C" int sktb_n2e_mean(const sktb_mesh *m, const double *x,
                             int clamp_max0, double *out, void *stream) {
  SKTB_REQUIRE(m && x && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (m->nen == 8)
    n2e_mean_kernel<8><<<grid_for(m->n_elem), kBlock, 0, st>>>(
        m->n_elem, m->conn, x, clamp_max0, out);
  else
    n2e_mean_kernel<4><<<grid_for(m->n_elem), kBlock, 0, st>>>(
        m->n_elem, m->conn, x, clamp_max0, out);
  SKTB_KERNEL_OK();
  return 0;
}

// ===================================================================== heat --
// Per-class quadrature tables: N[q][a], physical gradients G[q][a][3] and
// dx[q] = w_q |det J_q|  (what skfem's Basis.interpolate / asm evaluate).
template <int NEN>
__global__ void geom_tables_kernel(int nqp, const double *__restrict__ Xq,
                                   const double *__restrict__ Wq,
                                   const int32_t *__restrict__ class_rep,
                                   const int32_t *__restrict__ conn,
                                   int64_t n_elem,
                                   const double *__restrict__ coords,
                                   int64_t n_nodes, double *__restrict__ Nout,
                                   double *__restrict__ Gout,
                                   double *__restrict__ dxout) {
  __shared__ double xe[NEN][3];
  const int64_t cls = blockIdx.x;
  const int64_t el = class_rep[cls];
  if (threadIdx.x < NEN * 3) {
    int a = threadIdx.x / 3, d = threadIdx.x % 3;
    xe[a][d] = coords[(int64_t)d * n_nodes + conn[(int64_t)a * n_elem + el]];
  }
  __syncthreads();
  for (int q = threadIdx.x; q < nqp; q += blockDim.x) {
    const double X = Xq[q], Y = Xq[nqp + q], Z = Xq[2 * nqp + q];
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    double Nv[NEN], dNv[NEN][3];
#pragma unroll
    for (int v = 0; v < NEN; ++v) {
      if (NEN == 8)
        hex_shape(v, X, Y, Z, Nv[v], dNv[v]);
      else
        tet_shape(v, X, Y, Z, Nv[v], dNv[v]);
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int k = 0; k < 3; ++k) J[d][k] += xe[v][d] * dNv[v][k];
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double id = 1.0 / det;
    double iJ[3][3];
    iJ[0][0] = c00 * id;
    iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    iJ[1][0] = c01 * id;
    iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    iJ[2][0] = c02 * id;
    iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    double *Nq = Nout + (cls * nqp + q) * NEN;
    double *Gq = Gout + (cls * nqp + q) * NEN * 3;
#pragma unroll
    for (int v = 0; v < NEN; ++v) {
      Nq[v] = Nv[v];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        Gq[v * 3 + d] =
            dNv[v][0] * iJ[0][d] + dNv[v][1] * iJ[1][d] + dNv[v][2] * iJ[2][d];
    }
    dxout[cls * nqp + q] = Wq[q] * fabs(det);
  }
}

extern "C" int sktb_geom_tables(const sktb_mesh *m, int nqp, const double *X_h,
                                const double *W_h, int64_t n_class,
                                const int32_t *class_rep_h, double *N_out,
                                double *G_out, double *dx_out, void *stream) {
  SKTB_REQUIRE(m && X_h && W_h && class_rep_h && N_out && G_out && dx_out,
               "null argument");
  SKTB_REQUIRE(nqp > 0 && n_class > 0, "empty quadrature or class list");
  cudaStream_t st = (cudaStream_t)stream;
  double *Xd = nullptr, *Wd = nullptr;
  int32_t *rep = nullptr;
  SKTB_CUDA_OK(cudaMalloc(&Xd, sizeof(double) * 3 * nqp));
  SKTB_CUDA_OK(cudaMalloc(&Wd, sizeof(double) * nqp));
  SKTB_CUDA_OK(cudaMalloc(&rep, sizeof(int32_t) * n_class));
  SKTB_CUDA_OK(cudaMemcpyAsync(Xd, X_h, sizeof(double) * 3 * nqp,
                               cudaMemcpyHostToDevice, st));
  SKTB_CUDA_OK(cudaMemcpyAsync(Wd, W_h, sizeof(double) * nqp,
                               cudaMemcpyHostToDevice, st));
  SKTB_CUDA_OK(cudaMemcpyAsync(rep, class_rep_h, sizeof(int32_t) * n_class,
                               cudaMemcpyHostToDevice, st));
  if (m->nen == 8)
    geom_tables_kernel<8><<<(unsigned)n_class, 64, 0, st>>>(
        nqp, Xd, Wd, rep, m->conn, m->n_elem, m->coords, m->n_nodes, N_out,
        G_out, dx_out);
  else
    geom_tables_kernel<4><<<(unsigned)n_class, 32, 0, st>>>(
        nqp, Xd, Wd, rep, m->conn, m->n_elem, m->coords, m->n_nodes, N_out,
        G_out, dx_out);
  SKTB_KERNEL_OK();
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  cudaFree(Xd);
  cudaFree(Wd);
  cudaFree(rep);
  return 0;
}

// K16a: per (quadrature point, element) weight of the virtual Robin forms
//   s = h rho^p (1-rho)^q |grad rho| dx,  rho interpolated from nodal values
// and, if local_load != NULL, nothing else (the load is T_env * V * 1).
template <int NEN>
__global__ void __launch_bounds__(kBlock)
    robin_virtual_scale_kernel(int64_t n_elem, int nqp,
                               const int32_t *__restrict__ conn,
                               const int32_t *__restrict__ elem_class,
                               const double *__restrict__ Ntab,
                               const double *__restrict__ Gtab,
                               const double *__restrict__ dxtab,
                               const double *__restrict__ rho_n, double h,
                               double p, double q, double *__restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_elem; e += stride) {
    double r[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) r[a] = rho_n[conn[(int64_t)a * n_elem + e]];
    const int64_t cls = elem_class ? (int64_t)elem_class[e] : e;
    for (int k = 0; k < nqp; ++k) {
      const double *Nq = Ntab + (cls * nqp + k) * NEN;
      const double *Gq = Gtab + (cls * nqp + k) * NEN * 3;
      double rq = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        rq += __ldg(&Nq[a]) * r[a];
        g0 += __ldg(&Gq[a * 3 + 0]) * r[a];
        g1 += __ldg(&Gq[a * 3 + 1]) * r[a];
        g2 += __ldg(&Gq[a * 3 + 2]) * r[a];
      }
      const double iface = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
      out[(int64_t)k * n_elem + e] = h * pow(rq, p) * pow(1.0 - rq, q) * iface;
    }
  }
}

extern "C" int sktb_robin_virtual_scale(const sktb_mesh *m, int nqp,
                                        const int32_t *elem_class,
                                        const double *N_tab,
                                        const double *G_tab,
                                        const double *dx_tab,
                                        const double *rho_node, double h,
                                        double p, double q, double *out,
                                        void *stream) {
  SKTB_REQUIRE(m && N_tab && G_tab && dx_tab && rho_node && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->n_elem);
  if (m->nen == 8)
    robin_virtual_scale_kernel<8><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, h,
        p, q, out);
  else
    robin_virtual_scale_kernel<4><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, h,
        p, q, out);
  SKTB_KERNEL_OK();
  return 0;
}

// unit per-quadrature-point mass matrices  Mq[cls][q][a][b] = dx_q N_a N_b
__global__ void unit_qp_mass_kernel(int64_t total, int nqp, int nen,
                                    const double *__restrict__ Ntab,
                                    const double *__restrict__ dxtab,
                                    double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int b = i % nen;
    const int a = (i / nen) % nen;
    const int64_t cq = i / (nen * nen);  // cls*nqp + q
    out[i] = dxtab[cq] * Ntab[cq * nen + a] * Ntab[cq * nen + b];
  }
}

extern "C" int sktb_unit_qp_mass(const sktb_mesh *m, int nqp, int64_t n_class,
                                 const double *N_tab, const double *dx_tab,
                                 double *out, void *stream) {
  SKTB_REQUIRE(m && N_tab && dx_tab && out, "null argument");
  const int64_t total = n_class * nqp * m->nen * m->nen;
  unit_qp_mass_kernel<<<grid_for(total), kBlock, 0, (cudaStream_t)stream>>>(
      total, nqp, m->nen, N_tab, dx_tab, out);
  SKTB_KERNEL_OK();
  return 0;
}

// K16b: scalar gather assembly with several (scale, unit matrix) terms per
// element: vals = sum_e sum_k scale[k][e] * unit[cls(e)][k][a][b]
template <int NEN>
__global__ void __launch_bounds__(kBlock)
    assemble_terms_kernel(int64_t n_nodes, int64_t n_elem, int n_terms,
                          const int32_t *__restrict__ node_ptr,
                          const int32_t *__restrict__ pair_ptr,
                          const int32_t *__restrict__ contrib_elem,
                          const uint8_t *__restrict__ contrib_ab,
                          const int32_t *__restrict__ elem_class,
                          const double *__restrict__ unit,
                          const double *__restrict__ scale,
                          double *__restrict__ vals) {
  int64_t pair = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_pairs = node_ptr[n_nodes];
  for (; pair < n_pairs; pair += stride) {
    double acc = 0.0;
    const int32_t c1 = pair_ptr[pair + 1];
    for (int32_t c = pair_ptr[pair]; c < c1; ++c) {
      const int32_t el = contrib_elem[c];
      const int ab = contrib_ab[c];
      const int64_t cls = elem_class ? (int64_t)elem_class[el] : (int64_t)el;
      const double *u = unit + cls * n_terms * (NEN * NEN) + ab;
      for (int k = 0; k < n_terms; ++k)
        acc += scale[(int64_t)k * n_elem + el] * __ldg(&u[k * (NEN * NEN)]);
    }
    vals[pair] = acc;
  }
}

extern "C" int sktb_assemble_terms(const sktb_mesh *m, int n_terms,
                                   const double *unit,
                                   const int32_t *elem_class,
                                   const double *scale, double *vals,
                                   void *stream) {
  SKTB_REQUIRE(m && unit && scale && vals && n_terms > 0, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->node_nnz, kBlock, 16);
  if (m->nen == 8)
    assemble_terms_kernel<8><<<grid, kBlock, 0, st>>>(
        m->n_nodes, m->n_elem, n_terms, m->node_ptr, m->pair_ptr,
        m->contrib_elem, m->contrib_ab, elem_class, unit, scale, vals);
  else
    assemble_terms_kernel<4><<<grid, kBlock, 0, st>>>(
        m->n_nodes, m->n_elem, n_terms, m->node_ptr, m->pair_ptr,
        m->contrib_elem, m->contrib_ab, elem_class, unit, scale, vals);
  SKTB_KERNEL_OK();
  return 0;
}

// K17a: element-local contributions of the explicit Robin sensitivity form
//   r_e[a] = sum_q dx h ( da |g| phi N_a + a phi (g / max(|g|,1e-12)) . G_a )
//   a = rho^p (1-rho)^q, da = d a / d rho, phi = 2 T_env T - T^2
template <int NEN>
__global__ void __launch_bounds__(kBlock)
    robin_explicit_local_kernel(int64_t n_elem, int nqp,
                                const int32_t *__restrict__ conn,
                                const int32_t *__restrict__ elem_class,
                                const double *__restrict__ Ntab,
                                const double *__restrict__ Gtab,
                                const double *__restrict__ dxtab,
                                const double *__restrict__ rho_n,
                                const double *__restrict__ T, double h,
                                double T_env, double p, double q,
                                double *__restrict__ out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_elem; e += stride) {
    double r[NEN], t[NEN], acc[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      const int32_t nd = conn[(int64_t)a * n_elem + e];
      r[a] = rho_n[nd];
      t[a] = T[nd];
      acc[a] = 0.0;
    }
    const int64_t cls = elem_class ? (int64_t)elem_class[e] : e;
    for (int k = 0; k < nqp; ++k) {
      const double *Nq = Ntab + (cls * nqp + k) * NEN;
      const double *Gq = Gtab + (cls * nqp + k) * NEN * 3;
      const double dx = dxtab[cls * nqp + k];
      double rq = 0.0, tq = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        const double Na = __ldg(&Nq[a]);
        rq += Na * r[a];
        tq += Na * t[a];
        g0 += __ldg(&Gq[a * 3 + 0]) * r[a];
        g1 += __ldg(&Gq[a * 3 + 1]) * r[a];
        g2 += __ldg(&Gq[a * 3 + 2]) * r[a];
      }
      const double iface = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
      const double safe = fmax(iface, 1e-12);
      const double av = pow(rq, p) * pow(1.0 - rq, q);
      const double da = p * pow(rq, p - 1.0) * pow(1.0 - rq, q) -
                        q * pow(rq, p) * pow(1.0 - rq, q - 1.0);
      const double phi = 2.0 * T_env * tq - tq * tq;
      const double c0 = da * iface * phi;
      const double c1 = av * phi / safe;
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        const double gv = g0 * __ldg(&Gq[a * 3 + 0]) + g1 * __ldg(&Gq[a * 3 + 1]) +
                          g2 * __ldg(&Gq[a * 3 + 2]);
        acc[a] += h * (c0 * __ldg(&Nq[a]) + c1 * gv) * dx;
      }
    }
#pragma unroll
    for (int a = 0; a < NEN; ++a) out[(int64_t)a * n_elem + e] = acc[a];
  }
}

extern "C" int sktb_robin_explicit_local(const sktb_mesh *m, int nqp,
                                         const int32_t *elem_class,
                                         const double *N_tab,
                                         const double *G_tab,
                                         const double *dx_tab,
                                         const double *rho_node,
                                         const double *T, double h,
                                         double T_env, double p, double q,
                                         double *out_local, void *stream) {
  SKTB_REQUIRE(m && N_tab && G_tab && dx_tab && rho_node && T && out_local,
               "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->n_elem);
  if (m->nen == 8)
    robin_explicit_local_kernel<8><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, T,
        h, T_env, p, q, out_local);
  else
    robin_explicit_local_kernel<4><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, T,
        h, T_env, p, q, out_local);
  SKTB_KERNEL_OK();
  return 0;
}

// K18: heat-exchange objective pieces on the interface measure |grad rho_n|
// (fea/solver_heat.py:306-324 numerator, :385-392 denominator, :395-413 adjoint
// load).  Per element:
//   den_e   = sum_q |g| dx
//   num_e   = sum_q -T_env h (T_q - T_env) |g| dx                 (T != NULL)
//   local_a = sum_q -T_env heff_q |g| N_a dx, heff_q = sum_b N_b heff_b with
//             the NODAL values heff_b = h rho_b^p (1-rho_b)^q  (local != NULL)
template <int NEN>
__global__ void __launch_bounds__(kBlock)
    heat_exchange_local_kernel(int64_t n_elem, int nqp,
                               const int32_t *__restrict__ conn,
                               const int32_t *__restrict__ elem_class,
                               const double *__restrict__ Ntab,
                               const double *__restrict__ Gtab,
                               const double *__restrict__ dxtab,
                               const double *__restrict__ rho_n,
                               const double *__restrict__ T, double p,
                               double q, double h,
                               double T_env, double *__restrict__ den_out,
                               double *__restrict__ num_out,
                               double *__restrict__ local_out) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; e < n_elem; e += stride) {
    double r[NEN], t[NEN], hf[NEN], acc[NEN];
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      const int32_t nd = conn[(int64_t)a * n_elem + e];
      r[a] = rho_n[nd];
      t[a] = T ? T[nd] : 0.0;
      hf[a] = local_out ? h * pow(r[a], p) * pow(1.0 - r[a], q) : 0.0;
      acc[a] = 0.0;
    }
    const int64_t cls = elem_class ? (int64_t)elem_class[e] : e;
    double den = 0.0, num = 0.0;
    for (int k = 0; k < nqp; ++k) {
      const double *Nq = Ntab + (cls * nqp + k) * NEN;
      const double *Gq = Gtab + (cls * nqp + k) * NEN * 3;
      const double dx = dxtab[cls * nqp + k];
      double tq = 0.0, hq = 0.0, g0 = 0.0, g1 = 0.0, g2 = 0.0;
#pragma unroll
      for (int a = 0; a < NEN; ++a) {
        const double Na = __ldg(&Nq[a]);
        tq += Na * t[a];
        hq += Na * hf[a];
        g0 += __ldg(&Gq[a * 3 + 0]) * r[a];
        g1 += __ldg(&Gq[a * 3 + 1]) * r[a];
        g2 += __ldg(&Gq[a * 3 + 2]) * r[a];
      }
      const double iface = sqrt(g0 * g0 + g1 * g1 + g2 * g2);
      den += iface * dx;
      num += -T_env * h * (tq - T_env) * iface * dx;
      const double w = -T_env * hq * iface * dx;
#pragma unroll
      for (int a = 0; a < NEN; ++a) acc[a] += w * __ldg(&Nq[a]);
    }
    den_out[e] = den;
    if (num_out) num_out[e] = num;
    if (local_out) {
#pragma unroll
      for (int a = 0; a < NEN; ++a) local_out[(int64_t)a * n_elem + e] = acc[a];
    }
  }
}

extern "C" int sktb_heat_exchange_local(const sktb_mesh *m, int nqp,
                                        const int32_t *elem_class,
                                        const double *N_tab,
                                        const double *G_tab,
                                        const double *dx_tab,
                                        const double *rho_node,
                                        const double *T, double p,
                                        double q, double h,
                                        double T_env, double *den_out,
                                        double *num_out, double *local_out,
                                        void *stream) {
  SKTB_REQUIRE(m && N_tab && G_tab && dx_tab && rho_node && den_out,
               "null argument");
  SKTB_REQUIRE(!num_out || T, "numerator needs the temperature field");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->n_elem);
  if (m->nen == 8)
    heat_exchange_local_kernel<8><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, T,
        p, q, h, T_env, den_out, num_out, local_out);
  else
    heat_exchange_local_kernel<4><<<grid, kBlock, 0, st>>>(
        m->n_elem, nqp, m->conn, elem_class, N_tab, G_tab, dx_tab, rho_node, T,
        p, q, h, T_env, den_out, num_out, local_out);
  SKTB_KERNEL_OK();
  return 0;
}

// K17b: nodal sum of element-local vectors local[a][e] (deterministic gather)
__global__ void __launch_bounds__(kBlock)
    local_to_nodes_kernel(int64_t n_nodes, int64_t n_elem,
                          const int32_t *__restrict__ n2e_ptr,
                          const int32_t *__restrict__ n2e_elem,
                          const uint8_t *__restrict__ n2e_loc,
                          const double *__restrict__ local,
                          const double *__restrict__ divisor,
                          double *__restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n_nodes; i += stride) {
    double acc = 0.0;
    const int32_t k1 = n2e_ptr[i + 1];
    for (int32_t k = n2e_ptr[i]; k < k1; ++k)
      acc += local[(int64_t)n2e_loc[k] * n_elem + n2e_elem[k]];
    out[i] = divisor ? acc / divisor[i] : acc;
  }
}

extern "C" int sktb_local_to_nodes(const sktb_mesh *m, const double *local,
                                   const double *divisor, double *out,
                                   void *stream) {
  SKTB_REQUIRE(m && local && out, "null argument");
  local_to_nodes_kernel<<<grid_for(m->n_nodes), kBlock, 0,
                          (cudaStream_t)stream>>>(
      m->n_nodes, m->n_elem, m->n2e_ptr, m->n2e_elem, m->n2e_loc, local,
      divisor, out);
  SKTB_KERNEL_OK();
  return 0;
}


// ------------------------------------------------------ stress at quadrature --
// sigma = 2 mu sym(grad u) + lam tr(sym grad u) I per (element, quadrature point)
// (reference fea/composer.py:444-494, compute_element_stress_tensor /
// stress_tensor_skfem; post-processing, not part of the optimiser loop).
// out[(i*3+j)][e][q], G = physical shape-function gradients [cls][q][a][3].
template <int NEN>
__global__ void __launch_bounds__(kBlock)
    element_stress_kernel(int64_t n_elem, int nqp, const int32_t *__restrict__ conn,
                          const int32_t *__restrict__ cls, const double *__restrict__ G,
                          const double *__restrict__ E, double nu,
                          const double *__restrict__ u, double *__restrict__ out) {
  const int64_t total = n_elem * nqp;
  int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; id < total; id += stride) {
    const int64_t e = id / nqp;
    const int q = (int)(id - e * nqp);
    const int64_t c = cls ? cls[e] : e;
    const double *g = G + ((c * nqp + q) * NEN) * 3;
    double gr[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // du_i / dx_j
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      const int64_t nd = conn[(int64_t)a * n_elem + e];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double ui = u[3 * nd + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) gr[i][j] = fma(ui, g[3 * a + j], gr[i][j]);
      }
    }
    const double Ee = E[e];
    const double lam = nu * Ee / ((1.0 + nu) * (1.0 - 2.0 * nu)), mu = Ee / (2.0 * (1.0 + nu));
    const double tr = gr[0][0] + gr[1][1] + gr[2][2];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double s = mu * (gr[i][j] + gr[j][i]);
        if (i == j) s += lam * tr;
        out[((int64_t)(3 * i + j) * n_elem + e) * nqp + q] = s;
      }
  }
}

extern "C" int sktb_element_stress(const sktb_mesh *m, int nqp, const int32_t *elem_class,
                                   const double *G, const double *E_elem, double nu,
                                   const double *u, double *out, void *stream) {
  SKTB_REQUIRE(m && G && E_elem && u && out && nqp > 0, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(m->n_elem * nqp);
  if (m->nen == 8)
    element_stress_kernel<8><<<grid, kBlock, 0, st>>>(m->n_elem, nqp, m->conn, elem_class, G,
                                                     E_elem, nu, u, out);
  else
    element_stress_kernel<4><<<grid, kBlock, 0, st>>>(m->n_elem, nqp, m->conn, elem_class, G,
                                                     E_elem, nu, u, out);
  SKTB_KERNEL_OK();
  return 0;
}
