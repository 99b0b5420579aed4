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

struct sktb_comm {
  void *nccl = nullptr;  // ncclComm_t
  sktb_shm *shm = nullptr;
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
// in-place all-gather of variable-sized contiguous slices of `buf`
int comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts,
                    const int64_t *displs, cudaStream_t st);
}  // namespace sktb
