// Collectives of the row-sharded solver (one process per GPU).
//
// Two transports behind one interface:
//  * NCCL over NVLink / NVSwitch (the product path): resolved with
//    dlopen/dlsym at run time (the copy torch already loaded), so the library
//    has no link-time NCCL dependency and still loads on a CPU box;
//  * "shm": a POSIX shared-memory mailbox between processes of one host, staged
//    through the host.  NCCL refuses two ranks on one device; this transport
//    lets an N-rank job run on a single GPU, which is how the sharded code
//    paths are verified on one-GPU boxes (tests/test_gpu_dist.py) -- it is a
//    correctness vehicle, never a performance path.  Selected by the first
//    four bytes of the 128-byte id ("SHM:").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct sktb_shm;  // comm.cu

// Symmetric arena for peer-memory halos (NVLink P2P, cudaIpc): every rank
// allocates one arena of the same size and bump-allocates its full-length,
// slab-sharded vectors from it in the same order, so a vector sits at the same
// offset on every rank and the neighbour's copy is peer_base + (ptr - my_base).
struct P2PFlags {                 // first 256 bytes of the arena
  unsigned long long ready;       // last exchange this rank has published
  unsigned long long ack[2];      // [0] written by the previous rank, [1] by the next one
  unsigned int done_cnt;          // block counter of the exchange kernel
  unsigned int err;               // a spin timed out
};
struct sktb_arena {
  char *base = nullptr;           // my arena (device)
  size_t bytes = 0, used = 256;   // header = P2PFlags
  char *peer[2] = {nullptr, nullptr};  // previous / next rank's arena, mapped here
  unsigned long long epoch = 0;   // exchanges done (identical on all ranks)
};

struct sktb_comm {
  void *nccl = nullptr;  // ncclComm_t
  sktb_shm *shm = nullptr;
  sktb_arena *arena = nullptr;  // peer-memory halos (NCCL transport only)
  int rank = 0;
  int world = 1;
  int device = 0;
};

namespace sktb {
// one point-to-point transfer pair with a peer (either count may be 0)
struct P2POp {
  int peer;
  const double *send;
  int64_t n_send;
  double *recv;
  int64_t n_recv;
};
// dst[0..count) = sum over ranks of src[0..count)   (fp64, may alias)
int comm_allreduce_sum(sktb_comm *c, const double *src, double *dst,
                       int64_t count, cudaStream_t st);
// grouped point-to-point exchange of fp64 buffers with n_peers neighbours
int comm_exchange(sktb_comm *c, int n_peers, const int *peers,
                  const double *sendbuf, const int64_t *send_off,
                  double *recvbuf, const int64_t *recv_off, cudaStream_t st);
// the same with one (pointer, count) pair per peer and direction: contiguous
// slabs are sent straight from / received straight into the vectors
int comm_p2p(sktb_comm *c, int n_ops, const P2POp *ops, cudaStream_t st);
// Device memory for a full-length vector that takes part in slab halo exchanges:
// from the current communicator's symmetric arena when there is one (then the
// exchange runs over peer memory), else cudaMalloc.  dev_free handles both.
int dev_alloc_exchangeable(double **out, size_t n_doubles);
void dev_free(void *p);
// ghost planes of v pulled straight from the neighbours' copies over NVLink
// (one kernel: publish, wait for the neighbour, copy, acknowledge); returns -1
// when v is not an arena vector (caller: NCCL send / recv)
int comm_slab_halo_p2p(sktb_comm *c, double *v, int64_t own0, int64_t n_own, int64_t plane,
                       int prev, int next, cudaStream_t st);
// in-place all-gather of variable-sized contiguous slices of `buf`
int comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts,
                    const int64_t *displs, cudaStream_t st);
}  // namespace sktb
