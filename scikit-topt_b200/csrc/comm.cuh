// NCCL plumbing for the row-sharded PCG (one process per GPU).  NCCL is
// resolved with dlopen/dlsym at run time (the copy torch already loaded), so
// the library has no link-time NCCL dependency and still loads on a CPU box.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct sktb_comm {
  void *nccl = nullptr;  // ncclComm_t
  int rank = 0;
  int world = 1;
  int device = 0;
};

namespace sktb {
// dst[0..count) = sum over ranks of src[0..count)   (fp64, may alias)
int comm_allreduce_sum(sktb_comm *c, const double *src, double *dst,
                       int64_t count, cudaStream_t st);
// grouped point-to-point exchange of fp64 buffers with n_peers neighbours
int comm_exchange(sktb_comm *c, int n_peers, const int *peers,
                  const double *sendbuf, const int64_t *send_off,
                  double *recvbuf, const int64_t *recv_off, cudaStream_t st);
// in-place all-gather of variable-sized contiguous slices of `buf`
int comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts,
                    const int64_t *displs, cudaStream_t st);
}  // namespace sktb
