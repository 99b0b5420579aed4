// Device-resident Jacobi-preconditioned conjugate gradients (K4-K6), on one
// GPU or row-sharded over several (one process per GPU, NCCL over NVLink).
//
// All CG scalars live in device memory.  Kernels read the global values from
// `S` and publish rank-local sums into `Sloc`; with one GPU both are the same
// struct, with several an in-stream all-reduce maps Sloc -> S.  Every kernel
// tests ||r||^2 <= (rtol ||b||)^2 itself and becomes a no-op after
// convergence (the all-reduce of the unchanged Sloc is then idempotent), so
// the host polls only every `check_every` iterations and the result is the
// iterate at which scipy's criterion first holds.
//
// Sharding: rank r owns the contiguous rows [row0, row0 + n).  The search
// direction p is a full-length vector on every rank; before each SpMV the
// entries other ranks need are packed, exchanged with grouped ncclSend/ncclRecv
// and scattered into the ghost slots of p.  Matrix column indices stay global.
#include <cstdlib>
#include <vector>

#include "comm.cuh"
#include "common.cuh"
#include "linalg.cuh"

using namespace sktb;

struct sktb_pcg {
  int64_t n = 0;         // owned rows
  int64_t n_global = 0;  // length of p
  int64_t row0 = 0;
  int device = 0;
  double *r = nullptr, *z = nullptr, *p = nullptr, *q = nullptr;
  PcgScalars *S = nullptr;     // device, global values
  PcgScalars *Sloc = nullptr;  // device, rank-local sums (== S on one GPU)
  PcgScalars *S_h = nullptr;   // pinned host
  double *partials = nullptr;
  unsigned int *ticket = nullptr;
  // distributed
  sktb_comm *comm = nullptr;
  std::vector<int> peers;
  std::vector<int64_t> send_off, recv_off;
  int32_t *send_idx = nullptr, *recv_idx = nullptr;  // device, global indices
  double *sendbuf = nullptr, *recvbuf = nullptr;
  // z-slab halo (tensor grids): whole node planes, exchanged straight from /
  // into the full-length vector with the previous / next rank; replaces the
  // index lists above when plane > 0
  int64_t slab_plane = 0;
  int slab_prev = -1, slab_next = -1;
  // iterations to run before the first host poll of the NEXT solve (0: check_every)
  int first_batch = 0;
  // in-situ SpMV timing
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

static int pcg_alloc(sktb_pcg *s) {
  SKTB_CUDA_OK(cudaMalloc(&s->r, sizeof(double) * s->n));
  SKTB_CUDA_OK(cudaMalloc(&s->z, sizeof(double) * s->n));
  SKTB_CUDA_OK(cudaMalloc(&s->q, sizeof(double) * s->n));
  if (s->comm) {
    if (dev_alloc_exchangeable(&s->p, (size_t)s->n_global)) return 1;
  } else {
    SKTB_CUDA_OK(cudaMalloc(&s->p, sizeof(double) * s->n_global));
    SKTB_CUDA_OK(cudaMemset(s->p, 0, sizeof(double) * s->n_global));
  }
  SKTB_CUDA_OK(cudaMalloc(&s->S, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMemset(s->S, 0, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMallocHost(&s->S_h, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMalloc(&s->partials, sizeof(double) * ReduceScratch::kMaxVals *
                                            ReduceScratch::kMaxBlocks));
  SKTB_CUDA_OK(cudaMalloc(&s->ticket, sizeof(unsigned int)));
  SKTB_CUDA_OK(cudaMemset(s->ticket, 0, sizeof(unsigned int)));
  s->Sloc = s->S;
  return 0;
}

extern "C" int sktb_pcg_create(sktb_pcg **out, int64_t n_rows, int device) {
  SKTB_REQUIRE(out && n_rows > 0, "bad argument");
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_pcg *s = new sktb_pcg();
  s->n = s->n_global = n_rows;
  s->row0 = 0;
  s->device = device;
  if (pcg_alloc(s)) return 1;
  *out = s;
  return 0;
}

extern "C" int sktb_pcg_create_dist(sktb_pcg **out, sktb_comm *comm,
                                    int64_t n_global, int64_t row0,
                                    int64_t n_local, int n_peers,
                                    const int32_t *peers_h,
                                    const int64_t *send_off_h,
                                    const int32_t *send_idx_h,
                                    const int64_t *recv_off_h,
                                    const int32_t *recv_idx_h, int device) {
  SKTB_REQUIRE(out && comm && n_local > 0 && row0 >= 0 &&
                   row0 + n_local <= n_global,
               "bad argument");
  SKTB_REQUIRE(n_peers == 0 || (peers_h && send_off_h && recv_off_h),
               "null halo description");
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_pcg *s = new sktb_pcg();
  s->n = n_local;
  s->n_global = n_global;
  s->row0 = row0;
  s->device = device;
  s->comm = comm;
  if (pcg_alloc(s)) return 1;
  SKTB_CUDA_OK(cudaMalloc(&s->Sloc, sizeof(PcgScalars)));
  SKTB_CUDA_OK(cudaMemset(s->Sloc, 0, sizeof(PcgScalars)));
  s->peers.assign(peers_h, peers_h + n_peers);
  s->send_off.assign(send_off_h, send_off_h + n_peers + 1);
  s->recv_off.assign(recv_off_h, recv_off_h + n_peers + 1);
  const int64_t ns = n_peers ? s->send_off[n_peers] : 0;
  const int64_t nr = n_peers ? s->recv_off[n_peers] : 0;
  SKTB_CUDA_OK(cudaMalloc(&s->send_idx, sizeof(int32_t) * (ns ? ns : 1)));
  SKTB_CUDA_OK(cudaMalloc(&s->recv_idx, sizeof(int32_t) * (nr ? nr : 1)));
  SKTB_CUDA_OK(cudaMalloc(&s->sendbuf, sizeof(double) * (ns ? ns : 1)));
  SKTB_CUDA_OK(cudaMalloc(&s->recvbuf, sizeof(double) * (nr ? nr : 1)));
  if (ns)
    SKTB_CUDA_OK(cudaMemcpy(s->send_idx, send_idx_h, sizeof(int32_t) * ns,
                            cudaMemcpyHostToDevice));
  if (nr)
    SKTB_CUDA_OK(cudaMemcpy(s->recv_idx, recv_idx_h, sizeof(int32_t) * nr,
                            cudaMemcpyHostToDevice));
  *out = s;
  return 0;
}

extern "C" void sktb_pcg_destroy(sktb_pcg *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->r);
  cudaFree(s->z);
  dev_free(s->p);
  cudaFree(s->q);
  if (s->Sloc != s->S) cudaFree(s->Sloc);
  cudaFree(s->S);
  cudaFreeHost(s->S_h);
  cudaFree(s->partials);
  cudaFree(s->ticket);
  cudaFree(s->send_idx);
  cudaFree(s->recv_idx);
  cudaFree(s->sendbuf);
  cudaFree(s->recvbuf);
  delete s;
}

__device__ __forceinline__ bool done(const PcgScalars *S) {
  return S->rr <= S->tol2;
}

// r = b - q (q = A x0; r = b when !have_q); z = Minv r; p = z; publishes
// rz, rr, bb, tol2 (all linear in the local sums, so one all-reduce suffices)
__global__ void __launch_bounds__(kBlock)
    pcg_init_kernel(int64_t n, const double *__restrict__ b,
                    const double *__restrict__ q, int have_q,
                    const double *__restrict__ minv, double *__restrict__ r,
                    double *__restrict__ z, double *__restrict__ p,
                    double *__restrict__ x, double rtol, double *partials,
                    unsigned int *ticket, PcgScalars *Sloc, PcgScalars *S) {
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
      Sloc->rz = res[0];
      Sloc->rr = res[1];
      Sloc->bb = res[2];
      Sloc->tol2 = rtol * rtol * res[2];
      Sloc->pq = 0.0;
      Sloc->rz_new = res[0];
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
                      double *partials, unsigned int *ticket, PcgScalars *Sloc,
                      PcgScalars *S) {
  if (done(S)) return;
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
      Sloc->rz_new = res[0];
      Sloc->rr = res[1];
      S->iters += 1;
    }
  }
}

// p = z + beta p ; the last block rolls rz <- rz_new
__global__ void __launch_bounds__(kBlock)
    pcg_direction_kernel(int64_t n, const double *__restrict__ z,
                         double *__restrict__ p, unsigned int *ticket,
                         PcgScalars *S) {
  if (done(S)) return;
  const double beta = S->rz_new / S->rz;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = z[i] + beta * p[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      S->rz = S->rz_new;
      *ticket = 0u;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(kBlock)
    halo_pack_kernel(int64_t n, const int32_t *__restrict__ idx,
                     const double *__restrict__ p, double *__restrict__ buf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = p[idx[i]];
}
__global__ void __launch_bounds__(kBlock)
    halo_unpack_kernel(int64_t n, const int32_t *__restrict__ idx,
                       const double *__restrict__ buf, double *__restrict__ p) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[idx[i]] = buf[i];
}

int slab_halo_exchange(sktb_comm *c, double *v, int64_t own0, int64_t n_own,
                       int64_t plane, int prev, int next, cudaStream_t st) {
  if (!c || plane <= 0) return 0;
  {
    const int rc = comm_slab_halo_p2p(c, v, own0, n_own, plane, prev, next, st);
    if (rc != -1) return rc;
  }
  P2POp ops[2];
  int n = 0;
  if (prev >= 0) ops[n++] = {prev, v + own0, plane, v + own0 - plane, plane};
  if (next >= 0)
    ops[n++] = {next, v + own0 + n_own - plane, plane, v + own0 + n_own, plane};
  return comm_p2p(c, n, ops, st);
}

// The host polls the convergence flag every `check_every` iterations; a caller
// that knows how many iterations the previous, similar solve took can ask for
// the first poll to happen only after `n` iterations (one solve, then reset):
// fewer pipeline drains, and with check_every = 1 afterwards no V-cycle runs
// past convergence.
extern "C" int sktb_pcg_set_first_batch(sktb_pcg *s, int n) {
  SKTB_REQUIRE(s && n >= 0, "bad argument");
  s->first_batch = n;
  return 0;
}

extern "C" int sktb_pcg_set_slab_halo(sktb_pcg *s, int64_t plane_dofs, int prev_rank,
                                      int next_rank) {
  SKTB_REQUIRE(s && s->comm && plane_dofs > 0 && plane_dofs <= s->n, "bad argument");
  SKTB_REQUIRE(prev_rank < 0 || s->row0 >= plane_dofs, "no room for the lower ghost plane");
  SKTB_REQUIRE(next_rank < 0 || s->row0 + s->n + plane_dofs <= s->n_global,
               "no room for the upper ghost plane");
  s->slab_plane = plane_dofs;
  s->slab_prev = prev_rank;
  s->slab_next = next_rank;
  return 0;
}

// ghost entries of the full-length vector v <- owners' values
static int halo_exchange(sktb_pcg *s, double *v, cudaStream_t st) {
  if (s->comm && s->slab_plane > 0)
    return slab_halo_exchange(s->comm, v, s->row0, s->n, s->slab_plane, s->slab_prev,
                              s->slab_next, st);
  const int np = (int)s->peers.size();
  if (!s->comm || np == 0) return 0;
  const int64_t ns = s->send_off[np], nr = s->recv_off[np];
  if (ns) {
    halo_pack_kernel<<<grid_for(ns), kBlock, 0, st>>>(ns, s->send_idx, v,
                                                     s->sendbuf);
    SKTB_KERNEL_OK();
  }
  if (comm_exchange(s->comm, np, s->peers.data(), s->sendbuf,
                    s->send_off.data(), s->recvbuf, s->recv_off.data(), st))
    return 1;
  if (nr) {
    halo_unpack_kernel<<<grid_for(nr), kBlock, 0, st>>>(nr, s->recv_idx,
                                                       s->recvbuf, v);
    SKTB_KERNEL_OK();
  }
  return 0;
}

bool pcg_is_dist(const sktb_pcg *s) { return s && s->comm != nullptr; }
sktb_comm *pcg_comm(const sktb_pcg *s) { return s ? s->comm : nullptr; }
int pcg_halo_exchange(sktb_pcg *s, double *full_vec, cudaStream_t st) {
  return halo_exchange(s, full_vec, st);
}
int pcg_allreduce_vec(sktb_pcg *s, double *buf, int64_t n, cudaStream_t st) {
  if (!s || !s->comm) return 0;
  return comm_allreduce_sum(s->comm, buf, buf, n, st);
}

static int reduce_scalars(sktb_pcg *s, double *loc, double *glob, int count,
                          cudaStream_t st) {
  if (!s->comm) return 0;
  return comm_allreduce_sum(s->comm, loc, glob, count, st);
}

// scalar stencil multigrid (mg_scalar.cu)
struct sktb_smg;
int smg_vcycle(sktb_smg *m, const double *r, double *z, cudaStream_t st);
int smg_apply_level0(const sktb_smg *m, const double *x, double *y, const double *dotv,
                     ReduceScratch *rs, double *dot_out, const PcgScalars *S, cudaStream_t st);
int64_t smg_n(const sktb_smg *m);
const double *smg_inv_diag(const sktb_smg *m);

// preconditioner handed to the solver: the elasticity V-cycle or the scalar one
struct PcgPrecond {
  sktb_mg *mg = nullptr;
  sktb_smg *smg = nullptr;
  explicit operator bool() const { return mg || smg; }
  int apply(const double *r, double *z, cudaStream_t st, sktb_pcg *dist) const {
    return mg ? mg_vcycle(mg, r, z, st, dist) : smg_vcycle(smg, r, z, st);
  }
};

// matrix handed to the solver: CSR (kind 0), node-block CSR for 3 dofs per
// node (kind 1: rp/ci index node blocks, vals keep the CSR layout), the
// matrix-free grid operator (kind 2) or the 27-point stencil format of the
// scalar multigrid's level 0 (kind 3)
struct PcgMat {
  int kind;
  int dpn_hint;
  const int32_t *rp;
  const int32_t *ci;
  const double *vals;
  int64_t n_blocks = 0;  // kind 1: number of 3x3 blocks
  int max_deg = 0;       // kind 1: largest number of blocks in a node row
  const sktb_gridop *gop = nullptr;  // kind 2
  int64_t node0 = 0;                 // kind 2: first owned node
  const sktb_smg *smg = nullptr;     // kind 3
};

static int apply_mat(const PcgMat &A, int64_t n, const double *x, double *y,
                     const double *dotv, ReduceScratch *rs, double *dot_out,
                     const PcgScalars *S, cudaStream_t st) {
  if (A.kind == 3) return smg_apply_level0(A.smg, x, y, dotv, rs, dot_out, S, st);
  if (A.kind == 2)
    return launch_hexgrid_apply(A.gop, A.node0, n / gridop_dpn(A.gop), x, y, dotv, rs,
                                dot_out, S, st);
  if (A.kind == 1) {
    int rc = n / 3 < kTmaMinNodes ? -1 : launch_spmv_bsr3_tma(n / 3, A.n_blocks, A.max_deg, A.rp, A.ci,
                                  A.vals, x, y, dotv, rs, dot_out, S, st);
    if (rc != -1) return rc;
    return launch_spmv_bsr3(n / 3, A.rp, A.ci, A.vals, x, y, dotv, rs, dot_out,
                            S, st);
  }
  return launch_spmv(n, A.dpn_hint, A.rp, A.ci, A.vals, x, y, dotv, rs, dot_out,
                     S, st);
}

// rz (and optionally the rolled copy) <- r.z, for a general preconditioner
__global__ void __launch_bounds__(kBlock)
    pcg_rz_kernel(int64_t n, const double *__restrict__ r,
                  const double *__restrict__ z, int set_both, double *partials,
                  unsigned int *ticket, PcgScalars *Sloc, const PcgScalars *S) {
  if (done(S)) return;
  double v[1] = {0.0};
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[0] += r[i] * z[i];
  __shared__ double res[1];
  if (grid_reduce<1>(v, partials, ticket, res)) {
    if (threadIdx.x == 0) {
      Sloc->rz_new = res[0];
      if (set_both) Sloc->rz = res[0];
    }
  }
}

// ------------------------------------------------------------ phase timing --
namespace {
struct PhaseProf {
  bool on = false, init = false;
  std::vector<cudaEvent_t> ev;
  std::vector<int> id;
  size_t n = 0;
} g_ph;
}  // namespace
bool phase_on() {
  if (!g_ph.init) {
    g_ph.init = true;
    const char *e = getenv("SKTB_PHASE_PROF");
    g_ph.on = e && e[0] == '1';
  }
  return g_ph.on;
}
void phase_begin(cudaStream_t st) {
  if (!phase_on()) return;
  g_ph.n = 0;
  phase_mark(-1, st);
}
void phase_mark(int id, cudaStream_t st) {
  if (!phase_on()) return;
  if (g_ph.n == g_ph.ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_ph.ev.push_back(e);
    g_ph.id.push_back(0);
  }
  g_ph.id[g_ph.n] = id;
  cudaEventRecord(g_ph.ev[g_ph.n++], st);
}
void phase_report(int rank, int iters) {
  if (!phase_on() || g_ph.n < 2) return;
  static const char *names[PH_COUNT] = {
      "halo(p)", "A p (+p.q)", "allreduce p.q", "update x,r", "V: level-0 down", "V: halos",
      "V: level-1 (sharded levels >= 1)", "V: transition allreduce", "V: replicated coarse levels",
      "V: level-1 up", "V: level-0 up", "r.z", "allreduce r.z,||r||", "direction p"};
  double sum[PH_COUNT] = {0};
  for (size_t i = 1; i < g_ph.n; ++i) {
    if (g_ph.id[i] < 0) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_ph.ev[i - 1], g_ph.ev[i]) == cudaSuccess) sum[g_ph.id[i]] += ms;
  }
  if (rank != 0) return;
  double tot = 0;
  for (int k = 0; k < PH_COUNT; ++k) tot += sum[k];
  fprintf(stderr, "[phase] solve of %d iterations, %.3f ms in marked phases (%.3f ms / iteration)\n",
          iters, tot, iters ? tot / iters : 0.0);
  for (int k = 0; k < PH_COUNT; ++k)
    if (sum[k] > 0)
      fprintf(stderr, "[phase]   %-36s %8.3f ms  %5.1f %%  %7.1f us/it\n", names[k], sum[k],
              100 * sum[k] / tot, iters ? 1e3 * sum[k] / iters : 0.0);
}

static int pcg_run(sktb_pcg *s, const PcgMat &A, const double *inv_diag,
                   const double *b, double *x, int use_x0, double rtol,
                   int maxiter, int check_every, int32_t *info_h,
                   double *relres_h, void *stream, PcgPrecond mg = PcgPrecond()) {

  SKTB_REQUIRE(s && inv_diag && b && x, "null argument");
  SKTB_REQUIRE(A.kind == 3 ? A.smg != nullptr
                           : A.kind == 2 ? gridop_ready(A.gop) : (A.rp && A.ci && A.vals),
               "null argument");
  SKTB_REQUIRE(maxiter >= 0, "maxiter must be >= 0");
  if (check_every <= 0) check_every = 32;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = s->n;
  double *p_own = s->p + s->row0;  // owned slice of the full-length direction
  ReduceScratch rs;
  rs.partials = s->partials;
  rs.ticket = s->ticket;
  const int vgrid = grid_for(n, kBlock, 8);
  if (use_x0) {
    // q = A x0 needs x0 on the ghost slots: stage it through p
    SKTB_CUDA_OK(cudaMemcpyAsync(p_own, x, sizeof(double) * n,
                                 cudaMemcpyDeviceToDevice, st));
    if (halo_exchange(s, s->p, st)) return 1;
    if (apply_mat(A, n, s->p, s->q, nullptr, nullptr, nullptr, nullptr, st))
      return 1;
  }
  pcg_init_kernel<<<vgrid, kBlock, 0, st>>>(n, b, s->q, use_x0 ? 1 : 0, inv_diag,
                                           s->r, s->z, p_own, x, rtol,
                                           s->partials, s->ticket, s->Sloc,
                                           s->S);
  SKTB_KERNEL_OK();
  // rz, pq, rz_new, rr, tol2, bb are the first six doubles of the struct
  if (reduce_scalars(s, &s->Sloc->rz, &s->S->rz, 6, st)) return 1;
  if (mg) {
    // z = M^-1 r by one V-cycle; p = z; rz = r.z
    if (mg.apply(s->r, s->z, st, s)) return 1;
    SKTB_CUDA_OK(cudaMemcpyAsync(p_own, s->z, sizeof(double) * n,
                                 cudaMemcpyDeviceToDevice, st));
    pcg_rz_kernel<<<vgrid, kBlock, 0, st>>>(n, s->r, s->z, 1, s->partials,
                                           s->ticket, s->Sloc, s->S);
    SKTB_KERNEL_OK();
    // rz, pq (= 0), rz_new
    if (reduce_scalars(s, &s->Sloc->rz, &s->S->rz, 3, st)) return 1;
  }
  int launched = 0;
  int n_ev = 0;
  phase_begin(st);
  while (true) {
    SKTB_CUDA_OK(cudaMemcpyAsync(s->S_h, s->S, sizeof(PcgScalars),
                                 cudaMemcpyDeviceToHost, st));
    SKTB_CUDA_OK(cudaStreamSynchronize(st));
    for (int i = 0; i < n_ev; ++i) {
      float ms = 0.f;
      SKTB_CUDA_OK(cudaEventElapsedTime(&ms, s->ev0[i], s->ev1[i]));
      s->prof_ms += ms;
      s->prof_count += 1;
    }
    n_ev = 0;
    if (launched == 0 && s->S_h->bb == 0.0) {
      // b = 0: the solution is x = 0 whatever the start vector was (scipy's cg
      // returns it at once); with tol2 = 0 a non-zero warm start would
      // otherwise iterate to maxiter.  bb is the global sum: every rank agrees.
      SKTB_CUDA_OK(cudaMemsetAsync(x, 0, sizeof(double) * n, st));
      if (info_h) {
        info_h[0] = 0;
        info_h[1] = 1;
      }
      if (relres_h) *relres_h = 0.0;
      return 0;
    }
    if (s->S_h->rr <= s->S_h->tol2 || launched >= maxiter) break;
    int batch = maxiter - launched;
    const int want = (launched == 0 && s->first_batch > 0) ? s->first_batch : check_every;
    if (batch > want) batch = want;
    for (int it = 0; it < batch; ++it) {
      if (halo_exchange(s, s->p, st)) return 1;
      phase_mark(PH_HALO_P, st);
      // sample only the first iteration of a batch: it is certain to do work
      const bool sample = s->prof_every > 0 && it == 0 &&
                          ((launched / check_every) % s->prof_every == 0) &&
                          n_ev < sktb_pcg::kMaxProf;
      if (sample) SKTB_CUDA_OK(cudaEventRecord(s->ev0[n_ev], st));
      if (apply_mat(A, n, s->p, s->q, p_own, &rs, &s->Sloc->pq, s->S, st))
        return 1;
      if (sample) SKTB_CUDA_OK(cudaEventRecord(s->ev1[n_ev++], st));
      phase_mark(PH_SPMV, st);
      if (reduce_scalars(s, &s->Sloc->pq, &s->S->pq, 1, st)) return 1;
      phase_mark(PH_AR_PQ, st);
      pcg_update_kernel<<<vgrid, kBlock, 0, st>>>(n, p_own, s->q, inv_diag, x,
                                                 s->r, s->z, s->partials,
                                                 s->ticket, s->Sloc, s->S);
      phase_mark(PH_UPDATE, st);
      if (mg) {
        // one all-reduce for (r.z, ||r||^2) after the V-cycle instead of one on
        // each side of it: until then the kernels of the V-cycle see the previous
        // ||r||^2, so at worst the converging iteration runs one idle V-cycle
        if (mg.apply(s->r, s->z, st, s)) return 1;
        pcg_rz_kernel<<<vgrid, kBlock, 0, st>>>(n, s->r, s->z, 0, s->partials,
                                               s->ticket, s->Sloc, s->S);
        SKTB_COUNT(1);
        phase_mark(PH_RZ, st);
      }
      if (reduce_scalars(s, &s->Sloc->rz_new, &s->S->rz_new, 2, st)) return 1;
      phase_mark(PH_AR_RZ, st);
      pcg_direction_kernel<<<vgrid, kBlock, 0, st>>>(n, s->z, p_own, s->ticket,
                                                    s->S);
      SKTB_COUNT(2);
      phase_mark(PH_DIR, st);
    }
    SKTB_KERNEL_CHECK();
    launched += batch;
  }
  s->first_batch = 0;
  const PcgScalars &h = *s->S_h;
  if (mg && phase_on()) phase_report(s->comm ? s->comm->rank : 0, launched);
  if (info_h) {
    info_h[0] = h.iters;
    info_h[1] = (h.rr <= h.tol2) ? 1 : 0;
  }
  if (relres_h) *relres_h = (h.bb > 0.0) ? sqrt(h.rr / h.bb) : 0.0;
  return 0;
}

extern "C" int sktb_pcg_solve(sktb_pcg *s, int dpn_hint, const int32_t *row_ptr,
                              const int32_t *col_idx, const double *vals,
                              const double *inv_diag, const double *b,
                              double *x, int use_x0, double rtol, int maxiter,
                              int check_every, int32_t *info_h,
                              double *relres_h, void *stream) {
  PcgMat A{0, dpn_hint, row_ptr, col_idx, vals};
  return pcg_run(s, A, inv_diag, b, x, use_x0, rtol, maxiter, check_every,
                 info_h, relres_h, stream);
}

extern "C" int sktb_pcg_solve_bsr3(sktb_pcg *s, const int32_t *node_ptr,
                                   const int32_t *node_col, int64_t n_blocks,
                                   int max_deg, const double *vals,
                                   const double *inv_diag, const double *b,
                                   double *x, int use_x0, double rtol,
                                   int maxiter, int check_every,
                                   int32_t *info_h, double *relres_h,
                                   void *stream) {
  SKTB_REQUIRE(s && s->n % 3 == 0, "block solve needs 3 dofs per node");
  PcgMat A{1, 3, node_ptr, node_col, vals, n_blocks, max_deg};
  return pcg_run(s, A, inv_diag, b, x, use_x0, rtol, maxiter, check_every,
                 info_h, relres_h, stream);
}

extern "C" int sktb_pcg_solve_bsr3_mg(sktb_pcg *s, sktb_mg *mg,
                                      const int32_t *node_ptr,
                                      const int32_t *node_col, int64_t n_blocks,
                                      int max_deg, const double *vals,
                                      const double *inv_diag, const double *b,
                                      double *x, int use_x0, double rtol,
                                      int maxiter, int check_every,
                                      int32_t *info_h, double *relres_h,
                                      void *stream) {
  SKTB_REQUIRE(s && mg && s->n % 3 == 0, "block solve needs 3 dofs per node");
  PcgMat A{1, 3, node_ptr, node_col, vals, n_blocks, max_deg};
  PcgPrecond pc;
  pc.mg = mg;
  return pcg_run(s, A, inv_diag, b, x, use_x0, rtol, maxiter, check_every,
                 info_h, relres_h, stream, pc);
}

// ------------------------------------------------ spectral radius estimate --
__global__ void __launch_bounds__(kBlock)
    pw_init_kernel(int64_t n, int64_t row0, double *__restrict__ v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    // deterministic pseudo-random start vector from the GLOBAL index
    unsigned long long h = (unsigned long long)(row0 + i) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 32;
    v[i] = (double)(h & 0xFFFFFull) / 1048576.0 - 0.5;
  }
}
__global__ void __launch_bounds__(kBlock)
    pw_dot_kernel(int64_t n, const double *__restrict__ a,
                  const double *__restrict__ b, double *partials,
                  unsigned int *ticket, double *out) {
  double v[1] = {0.0};
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) v[0] += a[i] * b[i];
  grid_reduce<1>(v, partials, ticket, out);
}
__global__ void __launch_bounds__(kBlock)
    pw_scale_kernel(int64_t n, double a, const double *__restrict__ x,
                    const double *__restrict__ d, double *__restrict__ y) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) y[i] = a * x[i] * (d ? d[i] : 1.0);
}

// lambda_max(D^-1 A) by power iteration on the (possibly row-sharded) operator
static int lambda_max_run(sktb_pcg *s, const PcgMat &A, const double *inv_diag,
                          int iters, double *out_h, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = s->n;
  double *p_own = s->p + s->row0;
  const int vgrid = grid_for(n, kBlock, 8);
  auto gdot = [&](const double *a, const double *b, double *res) -> int {
    pw_dot_kernel<<<vgrid, kBlock, 0, st>>>(n, a, b, s->partials, s->ticket,
                                           &s->Sloc->pq);
    SKTB_KERNEL_OK();
    if (reduce_scalars(s, &s->Sloc->pq, &s->Sloc->pq, 1, st)) return 1;
    SKTB_CUDA_OK(cudaMemcpyAsync(&s->S_h->pq, &s->Sloc->pq, sizeof(double),
                                 cudaMemcpyDeviceToHost, st));
    SKTB_CUDA_OK(cudaStreamSynchronize(st));
    *res = s->S_h->pq;
    return 0;
  };
  pw_init_kernel<<<vgrid, kBlock, 0, st>>>(n, s->row0, p_own);
  SKTB_KERNEL_OK();
  double lam = 1.0;
  for (int it = 0; it < iters; ++it) {
    double nrm2 = 0.0;
    if (gdot(p_own, p_own, &nrm2)) return 1;
    SKTB_REQUIRE(nrm2 > 0.0, "power iteration broke down");
    pw_scale_kernel<<<vgrid, kBlock, 0, st>>>(n, 1.0 / sqrt(nrm2), p_own, nullptr, p_own);
    SKTB_KERNEL_OK();
    if (halo_exchange(s, s->p, st)) return 1;
    if (apply_mat(A, n, s->p, s->q, nullptr, nullptr, nullptr, nullptr, st)) return 1;
    pw_scale_kernel<<<vgrid, kBlock, 0, st>>>(n, 1.0, s->q, inv_diag, s->z);
    SKTB_KERNEL_OK();
    if (gdot(p_own, s->z, &lam)) return 1;
    SKTB_CUDA_OK(cudaMemcpyAsync(p_own, s->z, sizeof(double) * n,
                                 cudaMemcpyDeviceToDevice, st));
  }
  *out_h = lam;
  return 0;
}

extern "C" int sktb_pcg_lambda_max_bsr3(sktb_pcg *s, const int32_t *node_ptr,
                                        const int32_t *node_col,
                                        int64_t n_blocks, int max_deg,
                                        const double *vals,
                                        const double *inv_diag, int iters,
                                        double *out_h, void *stream) {
  SKTB_REQUIRE(s && node_ptr && node_col && vals && inv_diag && out_h && iters > 0,
               "bad argument");
  PcgMat A{1, 3, node_ptr, node_col, vals, n_blocks, max_deg};
  return lambda_max_run(s, A, inv_diag, iters, out_h, stream);
}

// ------------------------------------------- matrix-free grid operator path --
static PcgMat grid_mat(const sktb_pcg *s, const sktb_gridop *op) {
  PcgMat A{2, gridop_dpn(op), nullptr, nullptr, nullptr};
  A.gop = op;
  A.node0 = s->row0 / gridop_dpn(op);
  return A;
}

extern "C" int sktb_pcg_solve_grid(sktb_pcg *s, sktb_mg *mg,
                                   const sktb_gridop *op, const double *inv_diag,
                                   const double *b, double *x, int use_x0,
                                   double rtol, int maxiter, int check_every,
                                   int32_t *info_h, double *relres_h,
                                   void *stream) {
  SKTB_REQUIRE(s && gridop_ready(op), "null argument");
  SKTB_REQUIRE(s->n % gridop_dpn(op) == 0 && s->row0 % gridop_dpn(op) == 0,
               "row range is not a whole number of nodes");
  SKTB_REQUIRE(!mg || gridop_dpn(op) == 3, "multigrid needs the 3-dof operator");
  PcgPrecond pc;
  pc.mg = mg;
  return pcg_run(s, grid_mat(s, op), inv_diag, b, x, use_x0, rtol, maxiter,
                 check_every, info_h, relres_h, stream, pc);
}

// PCG on the scalar stencil operator (level 0 of `smg`, set up from the enforced
// CSR matrix by sktb_smg_setup_csr) preconditioned by its V-cycle
extern "C" int sktb_pcg_solve_smg(sktb_pcg *s, sktb_smg *smg, const double *b, double *x,
                                  int use_x0, double rtol, int maxiter, int check_every,
                                  int32_t *info_h, double *relres_h, void *stream) {
  SKTB_REQUIRE(s && smg && !s->comm, "bad argument (the scalar multigrid runs on one GPU)");
  SKTB_REQUIRE(s->n == smg_n(smg), "workspace and multigrid sizes differ");
  PcgMat A{3, 1, nullptr, nullptr, nullptr};
  A.smg = smg;
  PcgPrecond pc;
  pc.smg = smg;
  return pcg_run(s, A, smg_inv_diag(smg), b, x, use_x0, rtol, maxiter, check_every, info_h,
                 relres_h, stream, pc);
}

extern "C" int sktb_pcg_lambda_max_grid(sktb_pcg *s, const sktb_gridop *op,
                                        const double *inv_diag, int iters,
                                        double *out_h, void *stream) {
  SKTB_REQUIRE(s && gridop_ready(op) && inv_diag && out_h && iters > 0 &&
                   s->n % gridop_dpn(op) == 0,
               "bad argument");
  return lambda_max_run(s, grid_mat(s, op), inv_diag, iters, out_h, stream);
}
