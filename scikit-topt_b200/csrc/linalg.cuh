// Internal interface between the SpMV kernels (linalg.cu) and the PCG driver
// (pcg.cu).
#pragma once
#include "common.cuh"

// Device-resident CG scalars.  Kernels read the (global) values from `S` and
// publish their (rank-local) sums into `Sloc`; on one GPU both are the same
// struct, on several GPUs an all-reduce copies Sloc -> S.
struct PcgScalars {
  double rz;       // r.z of the current iterate
  double pq;       // p.Ap
  double rz_new;   // r.z after the update
  double rr;       // ||r||^2
  double tol2;     // (rtol*||b||)^2
  double bb;       // ||b||^2
  int iters;       // completed iterations
  int pad;
};

// y = A x (+ optional dot with dotv published to *dot_out); no-op once S says
// the solve has converged.
int launch_spmv(int64_t n_rows, int dpn_hint, const int32_t *row_ptr,
                const int32_t *col_idx, const double *vals, const double *x,
                double *y, const double *dotv, sktb::ReduceScratch *rs,
                double *dot_out, const PcgScalars *S, cudaStream_t st);

// same for the node-block layout of the 3-dof elasticity operator (spmv_bsr.cu)
int launch_spmv_bsr3(int64_t n_nodes, const int32_t *node_ptr,
                     const int32_t *node_col, const double *vals,
                     const double *x, double *y, const double *dotv,
                     sktb::ReduceScratch *rs, double *dot_out,
                     const PcgScalars *S, cudaStream_t st);

// bulk-async (TMA) pipelined variant (spmv_bsr_tma.cu); returns -1 when the
// layout is not eligible (max_deg > 27 or too few nodes) so the caller can
// fall back to launch_spmv_bsr3
int launch_spmv_bsr3_tma(int64_t n_nodes, int64_t n_blocks, int max_deg,
                         const int32_t *node_ptr, const int32_t *node_col,
                         const double *vals, const double *x, double *y,
                         const double *dotv, sktb::ReduceScratch *rs,
                         double *dot_out, const PcgScalars *S, cudaStream_t st);

// Damped-Jacobi sweep fused into the product (multigrid smoother):
//   y = x + omega dinv (b - A x)     (y must not alias x)
// launch_*_jacobi pick the same kernels with this epilogue switched on.
struct JacobiEpi {
  const double *b = nullptr;  // null: plain product
  const double *dinv = nullptr;
  double omega = 0.0;
  // rows of x that correspond to the local rows (row-sharded level: x is the
  // full-length vector, xo = x + first owned row); null: xo = x
  const double *xo = nullptr;
};
int launch_spmv_bsr3_jacobi(int64_t n_nodes, const int32_t *node_ptr,
                            const int32_t *node_col, const double *vals,
                            const double *x, double *y, const JacobiEpi &epi,
                            cudaStream_t st);
int launch_spmv_bsr3_tma_jacobi(int64_t n_nodes, int64_t n_blocks, int max_deg,
                                const int32_t *node_ptr, const int32_t *node_col,
                                const double *vals, const double *x, double *y,
                                const JacobiEpi &epi, cudaStream_t st);

// values kept in single precision (large multigrid levels: half the HBM stream,
// fp64 accumulation); epi.b == nullptr: plain product
int launch_spmv_bsr3_tma_f32(int64_t n_nodes, int64_t n_blocks, int max_deg,
                             const int32_t *node_ptr, const int32_t *node_col,
                             const float *vals, const double *x, double *y,
                             const JacobiEpi &epi, cudaStream_t st);

// below ~8k nodes the bulk-async pipeline's fixed start-up (~10 us) exceeds the
// work and the warp-per-node kernel is faster (measured: 2.7k nodes 13 -> 7 us,
// 16k nodes 13 vs 17 us)
constexpr int64_t kTmaMinNodes = 8000;

// matrix-free hexahedral grid operator (gridop.cu): y[local rows] = K x with x
// a full-length vector; same dot / early-exit contract as the SpMV launchers
struct sktb_gridop;
bool gridop_ready(const sktb_gridop *op);
int gridop_dpn(const sktb_gridop *op);
int launch_hexgrid_apply(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                         const double *x, double *y, const double *dotv,
                         sktb::ReduceScratch *rs, double *dot_out,
                         const PcgScalars *S, cudaStream_t st);

int launch_hexgrid_apply_ex(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                            const double *x, double *y, bool fp32, const double *b,
                            const double *dinv, double omega, cudaStream_t st);

// multigrid preconditioner z = M^-1 r (mg.cu)
struct sktb_mg;
struct sktb_pcg;
// `dist` (may be null) supplies the halo exchange / all-reduce of a row-sharded
// level 0; coarser levels are replicated
int mg_vcycle(sktb_mg *m, const double *r, double *z, cudaStream_t st,
              sktb_pcg *dist = nullptr);
// row-sharded helpers implemented by the PCG workspace (pcg.cu)
struct sktb_comm;
bool pcg_is_dist(const sktb_pcg *s);
sktb_comm *pcg_comm(const sktb_pcg *s);
int pcg_halo_exchange(sktb_pcg *s, double *full_vec, cudaStream_t st);
int pcg_allreduce_vec(sktb_pcg *s, double *buf, int64_t n, cudaStream_t st);
// Ghost planes of a z-slab-sharded full-length vector: this rank owns the
// entries [own0, own0 + n_own); `plane` entries are sent to / received from the
// previous and the next rank (-1: none), straight from / into the vector.
int slab_halo_exchange(sktb_comm *c, double *v, int64_t own0, int64_t n_own,
                       int64_t plane, int prev, int next, cudaStream_t st);

// Phase timing of the sharded solve (SKTB_PHASE_PROF=1): CUDA events at the phase
// boundaries of every PCG iteration / V-cycle, summed per phase over a solve and
// printed by rank 0 when the solve ends (profiles/r2_phase_*.txt).  Off: no cost.
enum PhaseId {
  PH_HALO_P = 0, PH_SPMV, PH_AR_PQ, PH_UPDATE, PH_VC_L0, PH_VC_HALO, PH_VC_L1, PH_VC_TRANS,
  PH_VC_COARSE, PH_VC_UP1, PH_VC_UP0, PH_RZ, PH_AR_RZ, PH_DIR, PH_COUNT
};
void phase_mark(int id, cudaStream_t st);   // the time since the previous mark goes to `id`
void phase_begin(cudaStream_t st);
void phase_report(int rank, int iters);
bool phase_on();
