// NCCL communicator wrapper (see comm.cuh).  Only the handful of NCCL entry
// points the sharded PCG needs are bound; enum values are NCCL 2.x ABI.
#include <dlfcn.h>

#include <mutex>

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
  if (load_api()) return 1;
  SKTB_CUDA_OK(cudaSetDevice(device));
  sktb_comm *c = new sktb_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
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
  NCCL_OK(g_api.AllReduce(src, dst, (size_t)count, kNcclFloat64, kNcclSum,
                          (NcclComm)c->nccl, st));
  return 0;
}

int comm_exchange(sktb_comm *c, int n_peers, const int *peers,
                  const double *sendbuf, const int64_t *send_off,
                  double *recvbuf, const int64_t *recv_off, cudaStream_t st) {
  NCCL_OK(g_api.GroupStart());
  for (int i = 0; i < n_peers; ++i) {
    const int64_t ns = send_off[i + 1] - send_off[i];
    const int64_t nr = recv_off[i + 1] - recv_off[i];
    if (ns > 0)
      NCCL_OK(g_api.Send(sendbuf + send_off[i], (size_t)ns, kNcclFloat64,
                         peers[i], (NcclComm)c->nccl, st));
    if (nr > 0)
      NCCL_OK(g_api.Recv(recvbuf + recv_off[i], (size_t)nr, kNcclFloat64,
                         peers[i], (NcclComm)c->nccl, st));
  }
  NCCL_OK(g_api.GroupEnd());
  return 0;
}

int comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts,
                    const int64_t *displs, cudaStream_t st) {
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
