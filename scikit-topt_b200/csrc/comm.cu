// Communicator: NCCL transport (product path) and a host shared-memory
// transport for several ranks on one GPU (see comm.cuh).  Only the handful of
// NCCL entry points the sharded solver needs are bound; enum values are the
// NCCL 2.x ABI.
#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "comm.cuh"
#include "common.cuh"

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void *NcclComm;
constexpr int kNcclFloat64 = 8;  // ncclDataType_t: ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;      // ncclRedOp_t: ncclSum

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId *);
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int);
  int (*CommDestroy)(NcclComm);
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
  int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
  bool ok = false;
};

NcclApi g_api;
std::mutex g_api_mu;

int load_api() {
  std::lock_guard<std::mutex> lk(g_api_mu);
  if (g_api.ok) return 0;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    sktb::set_error(std::string("cannot load libnccl.so.2: ") + dlerror());
    return 1;
  }
#define BIND(field, sym)                                              \
  *(void **)(&g_api.field) = dlsym(h, sym);                           \
  if (!g_api.field) {                                                 \
    sktb::set_error(std::string("libnccl.so.2 lacks symbol ") + sym); \
    return 1;                                                         \
  }
  BIND(GetUniqueId, "ncclGetUniqueId");
  BIND(CommInitRank, "ncclCommInitRank");
  BIND(CommDestroy, "ncclCommDestroy");
  BIND(AllReduce, "ncclAllReduce");
  BIND(AllGather, "ncclAllGather");
  BIND(Broadcast, "ncclBroadcast");
  BIND(Send, "ncclSend");
  BIND(Recv, "ncclRecv");
  BIND(GroupStart, "ncclGroupStart");
  BIND(GroupEnd, "ncclGroupEnd");
  BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
  g_api.ok = true;
  return 0;
}

#define NCCL_OK(call)                                                        \
  do {                                                                       \
    int _r = (call);                                                         \
    if (_r != 0) {                                                           \
      sktb::set_error(std::string(#call) + ": " + g_api.GetErrorString(_r)); \
      return 1;                                                              \
    }                                                                        \
  } while (0)

}  // namespace

// ------------------------------------------------------- shared-memory path --
// Segment: [header | slot 0 | slot 1 | ... ]; a slot is a rank's outbox: a
// directory of (peer, offset, count) records followed by fp64 payload.
struct ShmHeader {
  std::atomic<uint32_t> count;
  std::atomic<uint32_t> gen;
  std::atomic<uint32_t> attached;
};
struct ShmDirEntry {
  int64_t peer, off, cnt;
};
constexpr int kShmMaxDir = 64;
struct ShmSlotHead {
  int64_t n_dir;
  ShmDirEntry dir[kShmMaxDir];
};

struct sktb_shm {
  char name[64];
  int fd = -1;
  size_t bytes = 0, slot_bytes = 0;
  char *base = nullptr;
  int rank = 0, world = 1;
  std::vector<double> host;
  ShmHeader *head() const { return (ShmHeader *)base; }
  ShmSlotHead *slot(int r) const { return (ShmSlotHead *)(base + 4096 + (size_t)r * slot_bytes); }
  double *payload(int r) const { return (double *)((char *)slot(r) + sizeof(ShmSlotHead)); }
  int64_t capacity() const { return (int64_t)((slot_bytes - sizeof(ShmSlotHead)) / sizeof(double)); }
  void barrier() {
    ShmHeader *h = head();
    const uint32_t g = h->gen.load(std::memory_order_acquire);
    if (h->count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world) {
      h->count.store(0, std::memory_order_relaxed);
      h->gen.fetch_add(1, std::memory_order_release);
    } else {
      while (h->gen.load(std::memory_order_acquire) == g) sched_yield();
    }
  }
};

static int shm_open_segment(sktb_comm *c, const char *id128) {
  sktb_shm *s = new sktb_shm();
  s->rank = c->rank;
  s->world = c->world;
  snprintf(s->name, sizeof(s->name), "/sktb_%.40s", id128 + 4);
  const char *mb = getenv("SKTB_SHM_SLOT_MB");
  s->slot_bytes = (size_t)(mb ? atoi(mb) : 64) << 20;
  s->bytes = 4096 + s->slot_bytes * (size_t)c->world;
  s->fd = shm_open(s->name, O_CREAT | O_RDWR, 0600);
  if (s->fd < 0) {
    sktb::set_error("shm_open failed");
    return 1;
  }
  if (ftruncate(s->fd, (off_t)s->bytes) != 0) {
    sktb::set_error("ftruncate of the shared segment failed");
    return 1;
  }
  s->base = (char *)mmap(nullptr, s->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, s->fd, 0);
  if (s->base == MAP_FAILED) {
    sktb::set_error("mmap of the shared segment failed");
    return 1;
  }
  // a fresh segment is zero-filled: count = gen = attached = 0
  ShmHeader *h = s->head();
  h->attached.fetch_add(1);
  while (h->attached.load() < (uint32_t)c->world) sched_yield();
  c->shm = s;
  s->barrier();
  if (c->rank == 0) shm_unlink(s->name);  // everyone is mapped: drop the name
  return 0;
}

static int shm_allreduce(sktb_shm *s, const double *src, double *dst, int64_t count,
                         cudaStream_t st) {
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  const int64_t cap = s->capacity();
  for (int64_t c0 = 0; c0 < count; c0 += cap) {
    const int64_t n = count - c0 < cap ? count - c0 : cap;
    SKTB_CUDA_OK(cudaMemcpy(s->payload(s->rank), src + c0, sizeof(double) * n,
                            cudaMemcpyDeviceToHost));
    s->barrier();
    s->host.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      double a = 0.0;
      for (int r = 0; r < s->world; ++r) a += s->payload(r)[i];  // rank order: deterministic
      s->host[(size_t)i] = a;
    }
    s->barrier();
    SKTB_CUDA_OK(cudaMemcpy(dst + c0, s->host.data(), sizeof(double) * n,
                            cudaMemcpyHostToDevice));
  }
  return 0;
}

static int shm_p2p(sktb_shm *s, int n_ops, const sktb::P2POp *ops, cudaStream_t st) {
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  SKTB_REQUIRE(n_ops <= kShmMaxDir, "too many peers for the shm transport");
  ShmSlotHead *mine = s->slot(s->rank);
  int64_t off = 0;
  mine->n_dir = n_ops;
  for (int i = 0; i < n_ops; ++i) {
    SKTB_REQUIRE(off + ops[i].n_send <= s->capacity(),
                 "halo larger than the shm slot (raise SKTB_SHM_SLOT_MB)");
    mine->dir[i] = {ops[i].peer, off, ops[i].n_send};
    if (ops[i].n_send > 0)
      SKTB_CUDA_OK(cudaMemcpy(s->payload(s->rank) + off, ops[i].send,
                              sizeof(double) * ops[i].n_send, cudaMemcpyDeviceToHost));
    off += ops[i].n_send;
  }
  s->barrier();
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].n_recv <= 0) continue;
    const ShmSlotHead *theirs = s->slot(ops[i].peer);
    bool found = false;
    int seen = 0;  // the k-th op towards a peer pairs with the peer's k-th op towards us
    int want = 0;
    for (int j = 0; j < i; ++j)
      if (ops[j].peer == ops[i].peer) ++want;
    for (int64_t k = 0; k < theirs->n_dir; ++k) {
      if (theirs->dir[k].peer != s->rank) continue;
      if (seen++ != want) continue;
      SKTB_REQUIRE(theirs->dir[k].cnt == ops[i].n_recv, "shm exchange: size mismatch");
      SKTB_CUDA_OK(cudaMemcpy(ops[i].recv, s->payload(ops[i].peer) + theirs->dir[k].off,
                              sizeof(double) * ops[i].n_recv, cudaMemcpyHostToDevice));
      found = true;
      break;
    }
    SKTB_REQUIRE(found, "shm exchange: peer did not post a matching send");
  }
  s->barrier();
  return 0;
}

static int shm_allgatherv(sktb_shm *s, double *buf, const int64_t *counts,
                          const int64_t *displs, cudaStream_t st) {
  SKTB_CUDA_OK(cudaStreamSynchronize(st));
  const int64_t cap = s->capacity();
  int64_t maxc = 0;
  for (int r = 0; r < s->world; ++r) maxc = counts[r] > maxc ? counts[r] : maxc;
  for (int64_t c0 = 0; c0 < maxc; c0 += cap) {
    const int64_t mine = counts[s->rank] - c0;
    if (mine > 0)
      SKTB_CUDA_OK(cudaMemcpy(s->payload(s->rank), buf + displs[s->rank] + c0,
                              sizeof(double) * (mine < cap ? mine : cap),
                              cudaMemcpyDeviceToHost));
    s->barrier();
    for (int r = 0; r < s->world; ++r) {
      const int64_t n = counts[r] - c0;
      if (r == s->rank || n <= 0) continue;
      SKTB_CUDA_OK(cudaMemcpy(buf + displs[r] + c0, s->payload(r),
                              sizeof(double) * (n < cap ? n : cap), cudaMemcpyHostToDevice));
    }
    s->barrier();
  }
  return 0;
}

// ------------------------------------------------- peer-memory slab halos --
namespace {
sktb_comm *g_arena_comm = nullptr;  // one communicator per process owns the arena

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *p >= e; gives up after ~4 s (sets *err) so that a dead neighbour
// cannot hang the GPU
__device__ __forceinline__ bool spin_until(const unsigned long long *p, unsigned long long e,
                                           unsigned int *err) {
  const long long t0 = clock64();
  while (ld_sys(p) < e) {
    // ~60 s at 1.9 GHz (ranks may be seconds apart after host-side set-up work);
    // once the flag is up every later spin gives up at once
    if (*(volatile unsigned int *)err || clock64() - t0 > 120000000000ll) {
      atomicExch(err, 1u);
      return false;
    }
    __nanosleep(40);
  }
  __threadfence_system();
  return true;
}

// One kernel per exchange.  Epoch e is the same on every rank (same call
// sequence).  (1) publish: my planes for e are written (stream order) ->
// ready = e.  (2) for each neighbour: wait for its ready >= e, copy its boundary
// plane into my ghost plane with cache-bypassing loads.  (3) the last block
// acknowledges to the neighbours and waits for their acknowledgements, so that
// no later kernel of this stream overwrites a plane a neighbour still reads.
__global__ void __launch_bounds__(256)
    p2p_halo_kernel(P2PFlags *mine, P2PFlags *fprev, P2PFlags *fnext, unsigned long long e,
                    double *v, const double *vprev, const double *vnext, int64_t own0,
                    int64_t n_own, int64_t plane) {
  __shared__ int ok;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    __threadfence_system();
    st_sys(&mine->ready, e);
  }
  for (int side = 0; side < 2; ++side) {
    P2PFlags *fp = side ? fnext : fprev;
    if (!fp) continue;
    if (threadIdx.x == 0) ok = spin_until(&fp->ready, e, &mine->err) ? 1 : 0;
    __syncthreads();
    if (ok) {
      // previous rank: its last plane sits right below my first one (global
      // indexing); next rank: its first plane right above my last one
      const int64_t off = side ? own0 + n_own : own0 - plane;
      const double *src = (side ? vnext : vprev) + off;
      double *dst = v + off;
      // four independent loads in flight per thread: the copy is bound by the
      // NVLink round trip (~2 us), not by bandwidth
      const int64_t stride = (int64_t)gridDim.x * blockDim.x;
      for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane;
           i += 4 * stride) {
        double t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          t[k] = (i + k * stride < plane) ? __ldcv(src + i + k * stride) : 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (i + k * stride < plane) dst[i + k * stride] = t[k];
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(&mine->done_cnt, 1u);
    if (t == gridDim.x - 1) {
      mine->done_cnt = 0u;
      __threadfence_system();
      if (fprev) st_sys(&fprev->ack[1], e);  // I am the previous rank's "next"
      if (fnext) st_sys(&fnext->ack[0], e);
      if (fprev) spin_until(&mine->ack[0], e, &mine->err);
      if (fnext) spin_until(&mine->ack[1], e, &mine->err);
    }
  }
}
}  // namespace

// my arena + its IPC handle (64 bytes); the launcher all-gathers the handles and
// hands them to sktb_comm_arena_open
extern "C" int sktb_comm_arena_create(sktb_comm *c, int64_t bytes, void *handle64_h) {
  SKTB_REQUIRE(c && bytes >= 4096 && handle64_h, "bad argument");
  SKTB_REQUIRE(!c->shm, "peer-memory halos need one GPU per rank (NCCL transport)");
  SKTB_CUDA_OK(cudaSetDevice(c->device));
  sktb_arena *a = new sktb_arena();
  a->bytes = (size_t)bytes;
  SKTB_CUDA_OK(cudaMalloc((void **)&a->base, a->bytes));
  SKTB_CUDA_OK(cudaMemset(a->base, 0, a->bytes));
  cudaIpcMemHandle_t h;
  SKTB_CUDA_OK(cudaIpcGetMemHandle(&h, a->base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64_h, &h, 64);
  c->arena = a;
  return 0;
}

// handles_h: world x 64 bytes (rank order); maps the previous and the next rank's arena
extern "C" int sktb_comm_arena_open(sktb_comm *c, const void *handles_h) {
  SKTB_REQUIRE(c && c->arena && handles_h, "bad argument");
  SKTB_CUDA_OK(cudaSetDevice(c->device));
  for (int side = 0; side < 2; ++side) {
    const int r = side ? c->rank + 1 : c->rank - 1;
    if (r < 0 || r >= c->world) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles_h + (size_t)64 * r, 64);
    void *p = nullptr;
    SKTB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->arena->peer[side] = (char *)p;
  }
  g_arena_comm = c;
  return 0;
}

extern "C" int sktb_comm_arena_status(const sktb_comm *c, int64_t *used_h, int64_t *epoch_h,
                                      int32_t *err_h) {
  SKTB_REQUIRE(c && used_h && epoch_h && err_h, "null argument");
  *used_h = c->arena ? (int64_t)c->arena->used : 0;
  *epoch_h = c->arena ? (int64_t)c->arena->epoch : 0;
  *err_h = 0;
  if (c->arena) {
    P2PFlags f;
    SKTB_CUDA_OK(cudaMemcpy(&f, c->arena->base, sizeof(f), cudaMemcpyDeviceToHost));
    *err_h = (int32_t)f.err;
  }
  return 0;
}

// -------------------------------------------------------------------- C ABI --
extern "C" int sktb_comm_unique_id(void *id128_h) {
  SKTB_REQUIRE(id128_h, "null argument");
  if (load_api()) return 1;
  NCCL_OK(g_api.GetUniqueId((NcclUniqueId *)id128_h));
  return 0;
}

extern "C" int sktb_comm_create(sktb_comm **out, const void *id128_h, int rank,
                                int world, int device) {
  SKTB_REQUIRE(out && id128_h && world >= 1 && rank >= 0 && rank < world,
               "bad argument");
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_comm *c = new sktb_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  if (memcmp(id128_h, "SHM:", 4) == 0) {
    if (shm_open_segment(c, (const char *)id128_h)) return 1;
    *out = c;
    return 0;
  }
  if (load_api()) return 1;
  NcclUniqueId id;
  memcpy(&id, id128_h, sizeof(id));
  NcclComm comm = nullptr;
  NCCL_OK(g_api.CommInitRank(&comm, world, id, rank));
  c->nccl = comm;
  *out = c;
  return 0;
}

extern "C" void sktb_comm_destroy(sktb_comm *c) {
  if (!c) return;
  if (c->nccl && g_api.ok) g_api.CommDestroy((NcclComm)c->nccl);
  if (c->arena) {
    if (g_arena_comm == c) g_arena_comm = nullptr;
    for (int side = 0; side < 2; ++side)
      if (c->arena->peer[side]) cudaIpcCloseMemHandle(c->arena->peer[side]);
    cudaFree(c->arena->base);
    delete c->arena;
  }
  if (c->shm) {
    munmap(c->shm->base, c->shm->bytes);
    close(c->shm->fd);
    delete c->shm;
  }
  delete c;
}

extern "C" int sktb_comm_rank(const sktb_comm *c) { return c ? c->rank : -1; }
extern "C" int sktb_comm_world(const sktb_comm *c) { return c ? c->world : -1; }

extern "C" int sktb_comm_allreduce_sum(sktb_comm *c, const double *src,
                                       double *dst, int64_t count,
                                       void *stream) {
  SKTB_REQUIRE(c && src && dst && count > 0, "bad argument");
  return sktb::comm_allreduce_sum(c, src, dst, count, (cudaStream_t)stream);
}

extern "C" int sktb_comm_allgatherv(sktb_comm *c, double *buf,
                                    const int64_t *counts_h,
                                    const int64_t *displs_h, void *stream) {
  SKTB_REQUIRE(c && buf && counts_h && displs_h, "null argument");
  return sktb::comm_allgatherv(c, buf, counts_h, displs_h, (cudaStream_t)stream);
}

namespace sktb {

int comm_allreduce_sum(sktb_comm *c, const double *src, double *dst,
                       int64_t count, cudaStream_t st) {
  if (c->shm) return shm_allreduce(c->shm, src, dst, count, st);
  NCCL_OK(g_api.AllReduce(src, dst, (size_t)count, kNcclFloat64, kNcclSum,
                          (NcclComm)c->nccl, st));
  return 0;
}

int dev_alloc_exchangeable(double **out, size_t n_doubles) {
  sktb_comm *c = g_arena_comm;
  const size_t bytes = (sizeof(double) * n_doubles + 255) & ~(size_t)255;
  if (c && c->arena && c->arena->used + bytes <= c->arena->bytes) {
    *out = (double *)(c->arena->base + c->arena->used);
    c->arena->used += bytes;
    return 0;  // (the arena was zero-filled at creation)
  }
  SKTB_CUDA_OK(cudaMalloc((void **)out, sizeof(double) * n_doubles));
  SKTB_CUDA_OK(cudaMemset(*out, 0, sizeof(double) * n_doubles));
  return 0;
}

void dev_free(void *p) {
  sktb_comm *c = g_arena_comm;
  if (p && c && c->arena && (char *)p >= c->arena->base &&
      (char *)p < c->arena->base + c->arena->bytes)
    return;  // bump allocation: released with the arena
  cudaFree(p);
}

int comm_slab_halo_p2p(sktb_comm *c, double *v, int64_t own0, int64_t n_own, int64_t plane,
                       int prev, int next, cudaStream_t st) {
  sktb_arena *a = c ? c->arena : nullptr;
  if (!a || (char *)v < a->base || (char *)v >= a->base + a->bytes) return -1;
  if ((prev >= 0 && !a->peer[0]) || (next >= 0 && !a->peer[1])) return -1;
  const size_t off = (size_t)((char *)v - a->base);
  const unsigned long long e = ++a->epoch;
  P2PFlags *mine = (P2PFlags *)a->base;
  P2PFlags *fp = prev >= 0 ? (P2PFlags *)a->peer[0] : nullptr;
  P2PFlags *fn = next >= 0 ? (P2PFlags *)a->peer[1] : nullptr;
  const double *vp = prev >= 0 ? (const double *)(a->peer[0] + off) : nullptr;
  const double *vn = next >= 0 ? (const double *)(a->peer[1] + off) : nullptr;
  int grid = (int)((plane + 1023) / 1024);  // 4 entries per thread in one pass
  grid = grid < 1 ? 1 : (grid > 2 * sktb::kNumSM ? 2 * sktb::kNumSM : grid);
  p2p_halo_kernel<<<grid, 256, 0, st>>>(mine, fp, fn, e, v, vp, vn, own0, n_own, plane);
  SKTB_KERNEL_OK();
  return 0;
}

int comm_p2p(sktb_comm *c, int n_ops, const P2POp *ops, cudaStream_t st) {
  if (n_ops <= 0) return 0;
  if (c->shm) return shm_p2p(c->shm, n_ops, ops, st);
  NCCL_OK(g_api.GroupStart());
  for (int i = 0; i < n_ops; ++i) {
    if (ops[i].n_send > 0)
      NCCL_OK(g_api.Send(ops[i].send, (size_t)ops[i].n_send, kNcclFloat64,
                         ops[i].peer, (NcclComm)c->nccl, st));
    if (ops[i].n_recv > 0)
      NCCL_OK(g_api.Recv(ops[i].recv, (size_t)ops[i].n_recv, kNcclFloat64,
                         ops[i].peer, (NcclComm)c->nccl, st));
  }
  NCCL_OK(g_api.GroupEnd());
  return 0;
}

int comm_exchange(sktb_comm *c, int n_peers, const int *peers,
                  const double *sendbuf, const int64_t *send_off,
                  double *recvbuf, const int64_t *recv_off, cudaStream_t st) {
  std::vector<P2POp> ops((size_t)n_peers);
  for (int i = 0; i < n_peers; ++i)
    ops[(size_t)i] = {peers[i], sendbuf + send_off[i], send_off[i + 1] - send_off[i],
                      recvbuf + recv_off[i], recv_off[i + 1] - recv_off[i]};
  return comm_p2p(c, n_peers, ops.data(), st);
}

int comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts,
                    const int64_t *displs, cudaStream_t st) {
  if (c->shm) return shm_allgatherv(c->shm, buf, counts, displs, st);
  // equal slices at their natural displacements: one ncclAllGather in place
  bool uniform = true;
  for (int r = 0; r < c->world; ++r)
    if (counts[r] != counts[0] || displs[r] != (int64_t)r * counts[0]) uniform = false;
  if (uniform && counts[0] > 0) {
    NCCL_OK(g_api.AllGather(buf + displs[c->rank], buf, (size_t)counts[0], kNcclFloat64,
                            (NcclComm)c->nccl, st));
    return 0;
  }
  NCCL_OK(g_api.GroupStart());
  for (int r = 0; r < c->world; ++r)
    if (counts[r] > 0)
      NCCL_OK(g_api.Broadcast(buf + displs[r], buf + displs[r],
                              (size_t)counts[r], kNcclFloat64, r,
                              (NcclComm)c->nccl, st));
  NCCL_OK(g_api.GroupEnd());
  return 0;
}

}  // namespace sktb
