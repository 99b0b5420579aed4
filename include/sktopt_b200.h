/*
 * sktopt_b200.h -- C ABI of the B200-native hot path of scikit-topt's
 * per-iteration FEA + sensitivity loop.
 *
 * The reference (kevin-tofu/scikit-topt, pure Python) has no FFI; this header is
 * the boundary a maintainer would bind with ctypes.  Each
 * entry point names the reference call site it replaces (paths relative to
 * /root/reference/scikit-topt/sktopt/).
 *
 * Conventions
 *  - Every function returns 0 on success, non-zero on failure;
 *    sktb_last_error() returns a thread-local message for the last failure.
 *  - Pointers suffixed _h are HOST pointers, everything else is a DEVICE pointer
 *    (fp64 / int32 / uint8) on the CUDA device the mesh was created on.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    All work is enqueued asynchronously unless the doc says "synchronises".
 *  - Element types: 0 = hex8 (trilinear), 1 = tet4 (linear).  dofs-per-node
 *    (dpn) is 3 for elasticity and 1 for scalar (heat / Helmholtz) problems.
 *    DOF numbering is dpn*node + comp (reference mesh/task_elastic.py:72).
 *  - CSR: int32 row_ptr[n_rows+1], int32 col_idx[nnz] (sorted per row),
 *    fp64 vals[nnz]; the pattern is the union of element couplings.
 */
#ifndef SKTOPT_B200_H
#define SKTOPT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKTB_HEX8 0
#define SKTB_TET4 1

#define SKTB_KE_ELASTIC 0 /* fea/composer.py:88-98  lam tr(e)tr(e) + 2 mu e:e, E=1 */
#define SKTB_KE_LAPLACE 1 /* fea/composer.py:139-141 grad u . grad v, k=1         */
#define SKTB_KE_MASS 2    /* filters/helmholtz_filter_nodal.py:136-138  u v         */

const char *sktb_last_error(void);
int sktb_version(void);
/* number of kernels this library has launched in the calling process          */
int64_t sktb_launch_count(void);

/* ------------------------------------------------------------------ mesh --
 * Connectivity-derived structures shared by every operator on one mesh:
 * node graph (union of element couplings), node->element adjacency and the
 * per-(row node, column node) contributor lists ("scatter map") that make the
 * assembly a deterministic gather.
 * Replaces the index bookkeeping hidden inside skfem.asm / COO->CSR
 * (fea/composer.py:100,144) and the Python loops at
 * filters/helmholtz_filter_nodal.py:47-52, mesh/utils.py:210-212.            */
typedef struct sktb_mesh sktb_mesh;

int sktb_mesh_create(sktb_mesh **out, int elem_type, int64_t n_elem,
                     int64_t n_nodes,
                     const int32_t *conn_h,  /* [nen][n_elem] */
                     const double *coords_h, /* [3][n_nodes]  */
                     int device);
void sktb_mesh_destroy(sktb_mesh *m);
int64_t sktb_mesh_node_nnz(const sktb_mesh *m);
/* copies the node graph (CSR over nodes) to host buffers */
int sktb_mesh_node_graph_h(const sktb_mesh *m, int32_t *row_ptr_h,
                           int32_t *col_idx_h);
/* dof-level CSR pattern for dpn dofs per node, written to device buffers:
 * row_ptr[dpn*n_nodes+1], col_idx[dpn*dpn*node_nnz]                          */
int sktb_mesh_dof_pattern(const sktb_mesh *m, int dpn, int32_t *row_ptr,
                          int32_t *col_idx, void *stream);

/* Unit element matrices Ke0 (material coefficient = 1) for `n_class` geometry
 * classes; class c is represented by element class_rep_h[c].
 * out: [n_class][nde][nde], nde = nen*dpn (dpn = 3 for ELASTIC else 1).
 * Quadrature (X_h[3][nqp] on the reference element, W_h[nqp]) is the basis
 * quadrature of the reference (skfem Basis(intorder)).                         */
int sktb_unit_ke(const sktb_mesh *m, int kind, double nu, int nqp,
                 const double *X_h, const double *W_h, int64_t n_class,
                 const int32_t *class_rep_h, double *out, void *stream);

/* K2: vals = sum_e scale[e] * Ke0[class[e]] gathered into CSR order.
 * elem_class may be NULL (class == element), scale may be NULL (all ones).
 * If dir_mask (uint8 per dof) is given the Dirichlet rows/cols are written as
 * identity (skfem.enforce semantics, fea/solver_elastic.py:211).
 * Replaces composer.assemble_stiffness_matrix / assemble_conduction_matrix.   */
int sktb_assemble(const sktb_mesh *m, int dpn, const double *unit_ke,
                  const int32_t *elem_class, const double *scale,
                  const uint8_t *dir_mask, double *vals, void *stream);

/* K3 helpers (skfem.enforce): zero rows/cols of masked dofs, unit diagonal.   */
int sktb_csr_enforce(int64_t n_rows, const int32_t *row_ptr,
                     const int32_t *col_idx, double *vals,
                     const uint8_t *dir_mask, void *stream);
/* out[r] = 1 / K[r,r]  (Jacobi preconditioner, fea/solver_elastic.py:85-86)   */
int sktb_csr_inv_diag(int64_t n_rows, const int32_t *row_ptr,
                      const int32_t *col_idx, const double *vals, double *out,
                      void *stream);
/* K4: y = A x  (the SpMV inside scipy.sparse.linalg.cg,
 * fea/solver_elastic.py:89,101)                                               */
int sktb_spmv(int64_t n_rows, int dpn_hint, const int32_t *row_ptr,
              const int32_t *col_idx, const double *vals, const double *x,
              double *y, void *stream);

/* --------------------------------------------------------------- PCG ------
 * K4-K6: Jacobi-preconditioned conjugate gradients, all scalars device
 * resident, convergence test ||r||_2 <= rtol*||b||_2 (scipy cg, atol=0).
 * Replaces solve_u(... 'cg_jacobi' / 'cg_pyamg') fea/solver_elastic.py:61-143.
 * The workspace holds r, z, p, q and the reduction scratch.                   */
typedef struct sktb_pcg sktb_pcg;
int sktb_pcg_create(sktb_pcg **out, int64_t n_rows, int device);
void sktb_pcg_destroy(sktb_pcg *s);
/* x holds the initial guess on entry (use_x0 != 0) or is overwritten (x0 = 0).
 * Synchronises the stream.  info_h: [0]=iterations, [1]=converged(0/1);
 * relres_h: final ||r||/||b||.                                                */
int sktb_pcg_solve(sktb_pcg *s, int dpn_hint, const int32_t *row_ptr,
                   const int32_t *col_idx, const double *vals,
                   const double *inv_diag, const double *b, double *x,
                   int use_x0, double rtol, int maxiter, int check_every,
                   int32_t *info_h, double *relres_h, void *stream);

/* Node-block variant for 3 dofs per node: the values keep the CSR layout
 * (row 3n+i = 3*deg(n) contiguous entries, the node's three rows back to back)
 * but columns are read once per 3x3 block from the node graph
 * (node_ptr[n_nodes+1], node_col[]): 8 + 4/9 bytes per non-zero instead of 12. */
int sktb_spmv_bsr3(int64_t n_nodes, const int32_t *node_ptr,
                   const int32_t *node_col, const double *vals, const double *x,
                   double *y, void *stream);
/* bulk-async (cp.async.bulk + mbarrier) pipelined variant: tiles of 16 nodes
 * are streamed into a 3-stage shared-memory ring by one thread per CTA while
 * the warps consume the previous tile; needs max_deg <= 27 (hex8 graphs).
 * n_blocks = node_ptr[n_nodes].  The PCG uses it automatically when eligible. */
int sktb_spmv_bsr3_tma(int64_t n_nodes, int64_t n_blocks, int max_deg,
                       const int32_t *node_ptr, const int32_t *node_col,
                       const double *vals, const double *x, double *y,
                       void *stream);
/* out[3n+i] = 1 / A[3n+i,3n+i] from the node-block layout; node0 = global id
 * of the first (local) row node when the rows are a shard                      */
int sktb_bsr3_inv_diag(int64_t n_nodes, int64_t node0, const int32_t *node_ptr,
                       const int32_t *node_col, const double *vals, double *out,
                       void *stream);
int sktb_pcg_solve_bsr3(sktb_pcg *s, const int32_t *node_ptr,
                        const int32_t *node_col, int64_t n_blocks, int max_deg,
                        const double *vals,
                        const double *inv_diag, const double *b, double *x,
                        int use_x0, double rtol, int maxiter, int check_every,
                        int32_t *info_h, double *relres_h, void *stream);

/* ----------------------------------------------- multigrid preconditioner --
 * Replaces pyamg.smoothed_aggregation_solver(K).aspreconditioner()
 * (fea/solver_elastic.py:94-104, rebuilt every optimiser iteration there) for
 * tensor-product hexahedral grids: geometric hierarchy (cell counts halved per
 * level, trilinear prolongation), exact Galerkin coarse operators formed
 * element-wise, V(1,1) damped-Jacobi cycle, Dirichlet dofs masked per level.
 * All level arrays are caller-owned device buffers.                           */
typedef struct sktb_mg sktb_mg;
int sktb_mg_create(sktb_mg **out, int n_levels, int device);
void sktb_mg_destroy(sktb_mg *m);
int sktb_mg_set_params(sktb_mg *m, double omega, int nu_coarse);
int sktb_mg_set_level_omega(sktb_mg *m, int level, double omega);
/* operator of one level: node-block CSR (values in CSR layout, enforced),
 * inverse diagonal, optional per-dof Dirichlet mask                           */
int sktb_mg_set_level(sktb_mg *m, int level, int64_t n_nodes, int64_t n_blocks,
                      int max_deg, const int32_t *node_ptr,
                      const int32_t *node_col, const double *vals,
                      const double *inv_diag, const uint8_t *mask);
/* level 0 of a row-sharded operator: this rank owns the nodes starting at
 * node0 out of n_global (call before sktb_mg_set_level(m, 0, n_owned, ...)).
 * Coarser levels stay replicated; the restricted residual is all-reduced and
 * the level-0 smoother exchanges halos through the PCG workspace it runs in.  */
int sktb_mg_set_level0_range(sktb_mg *m, int64_t node0, int64_t n_global);
/* z-slab sharding of ANY level (tensor grids; SURVEY.md 8e "element-block
 * partition"): this rank owns the whole node planes [node0, node0 + n_owned) of
 * n_global; prev_rank / next_rank own the adjacent planes (-1: none).  The
 * level's operator arrays then hold the owned rows only (global columns), its
 * iterate is a full-length vector whose ghost planes are exchanged before every
 * product; restriction / prolongation between two sharded levels exchange one
 * plane of the residual / the coarse iterate.  plane_nodes = 0: replicated.
 * Call before sktb_mg_set_level / sktb_mg_set_level0_grid of that level.       */
int sktb_mg_set_level_slab(sktb_mg *m, int level, int64_t node0, int64_t n_global,
                           int64_t plane_nodes, int prev_rank, int next_rank);
/* y[owned rows] = A_level x_full (ghost planes refreshed first; dist = the PCG
 * workspace that carries the communicator, NULL on one GPU)                    */
int sktb_mg_level_apply(sktb_mg *m, int level, sktb_pcg *dist, double *x_full,
                        double *y_own, void *stream);
/* transfer between level (fine) and level+1 (coarse).  Nodes per axis (x,y,z)
 * of both grids (node = iy + npy*ix + npy*npx*iz); per-axis interpolation
 * tables on the device, concatenated [x|y|z]: fine index i takes coarse
 * c0[i], c1[i] with weights w0[i], w1[i]; transposed tables [3 slots][x|y|z]:
 * coarse index I gathers fine axT_f (or -1) with weight axT_w.                */
int sktb_mg_set_transfer(sktb_mg *m, int level, const int32_t *fine_np_h,
                         const int32_t *coarse_np_h, const int32_t *ax_c0,
                         const int32_t *ax_c1, const double *ax_w0,
                         const double *ax_w1, const int32_t *axT_f,
                         const double *axT_w);
/* single-precision copy of a level's values (same node-block layout, made with
 * sktb_f64_to_f32; device, caller-owned): the V-cycle's products on that level
 * stream half the bytes, accumulation stays fp64.  Call after sktb_mg_set_level
 * (which forgets the copy); NULL switches back to the fp64 values.                */
int sktb_mg_set_level_vals32(sktb_mg *m, int level, const float *vals32);
/* out[i] = (float) in[i]  (in 16-byte, out 8-byte aligned)                         */
int sktb_f64_to_f32(int64_t n, const double *in, float *out, void *stream);
/* y = A x with single-precision values in the layout of sktb_spmv_bsr3 (bulk-async
 * pipeline of sktb_spmv_bsr3_tma, fp64 x / y / accumulation)                       */
int sktb_spmv_bsr3_tma_f32(int64_t n_nodes, int64_t n_blocks, int max_deg,
                           const int32_t *node_ptr, const int32_t *node_col,
                           const float *vals, const double *x, double *y, void *stream);
/* dst (a second set of work vectors on the same level operators, e.g. one per
 * concurrently solved load case) borrows src's exact coarsest-level inverse        */
int sktb_mg_share_coarsest(sktb_mg *dst, const sktb_mg *src);
/* z = M^-1 r (one V-cycle)                                                    */
int sktb_mg_vcycle(sktb_mg *m, const double *r, double *z, void *stream);
/* Galerkin coarse element matrices: out[E] = sum_c Q^T K_child Q over the
 * children child[c][E] (-1 = none) of coarse element E, Q = Qtab[ptype[E]][c]
 * (8x8 trilinear weights child vertex <- parent vertex).  Children are either
 * per-element matrices fine_ke[e][24][24] or scale[e]*unit[cls[e]] (level 0). */
int sktb_elem_restrict(int64_t n_coarse, const int32_t *child,
                       const uint8_t *ptype, const double *Qtab,
                       const double *fine_ke, const double *unit,
                       const int32_t *cls, const double *scale, double *out,
                       void *stream);
/* the same for the coarse elements [e_lo, e_hi) only (slab-sharded set-up):
 * out holds them from e_lo on, fine_ke holds the fine elements from fine_base  */
int sktb_elem_restrict_range(int64_t n_coarse, int64_t e_lo, int64_t e_hi,
                             int64_t fine_base, const int32_t *child,
                             const uint8_t *ptype, const double *Qtab,
                             const double *fine_ke, const double *unit,
                             const int32_t *cls, const double *scale, double *out,
                             void *stream);
/* Algebraic Galerkin product A_c = P^T A P of a node-block (3 dofs per node)
 * operator whose nodes form a lattice (node = iy + npy*ix + npy*npx*iz) with ANY
 * geometry or element type (jittered / graded hexahedra, the Kuhn tetrahedra of
 * MeshTet.init_tensor): the set-up of pyamg.smoothed_aggregation_solver(K),
 * fea/solver_elastic.py:94-100, where the element-wise kernels above (uniform
 * hexahedra) do not apply.  P = the trilinear index-space interpolation of
 * sktb_mg_set_transfer (same tables); fine / coarse operators in the node-block
 * layout of sktb_spmv_bsr3 (coarse graph = the 27-point lattice graph, given by
 * the caller); rows / columns of fixed fine dofs (fine_mask, may be NULL) are
 * left out, fixed coarse dofs (coarse_mask, may be NULL) become identity rows.
 * One warp per coarse node, fixed summation order, no atomics.                  */
int sktb_galerkin_bsr3_lattice(const int32_t *fine_np_h, const int32_t *coarse_np_h,
                               const int32_t *ax_c0, const int32_t *ax_c1,
                               const double *ax_w0, const double *ax_w1,
                               const int32_t *axT_f, const double *axT_w,
                               const int32_t *fine_node_ptr, const int32_t *fine_node_col,
                               const double *fine_vals, const uint8_t *fine_mask,
                               const int32_t *coarse_node_ptr, const int32_t *coarse_node_col,
                               const uint8_t *coarse_mask, double *coarse_vals, void *stream);
/* ---- matrix-free operators for uniform hexahedral tensor grids --------------
 * Replace the assembled matrix inside the solvers where the reference hands
 * scipy/pyamg an assembled one (elasticity: fea/solver_elastic.py:94-104,
 * 189-260; Helmholtz filter: filters/helmholtz_filter_nodal.py solve calls):
 * y = A x with A = sum_e scale[e] Ae0 evaluated element-wise from the grid
 * structure (node = iy + npy (ix + npx iz), element = ey + ny (ex + nx ez)).
 * dpn = 3 (elasticity, Ae0 24x24) or 1 (scalar, Ae0 8x8; scale may be NULL = 1).
 * ke_cc_h: Ae0 with local vertices re-ordered by corner code cx + 2 cy + 4 cz
 * (host).  dmask (device, per node): bit i = dof i fixed, bit 3 = some node of
 * the 27-neighbourhood has a fixed dof.                                        */
typedef struct sktb_gridop sktb_gridop;
int sktb_gridop_create(sktb_gridop **out, int dpn, const int32_t *np_h,
                       const double *ke_cc_h, int device);
/* tile_h[3] <- node brick (x, y, z) one CTA of the tiled kernel owns           */
int sktb_gridop_tile_shape(const sktb_gridop *op, int32_t *tile_h);
void sktb_gridop_destroy(sktb_gridop *op);
int sktb_gridop_set_fields(sktb_gridop *op, const double *scale,
                           const uint8_t *dmask);
/* rows of the nodes [node0, node0 + n_nodes); x is full length, y local        */
int sktb_gridop_apply(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                      const double *x, double *y, void *stream);
int sktb_gridop_inv_diag(const sktb_gridop *op, int64_t node0, int64_t n_nodes,
                         double *out, void *stream);
/* PCG on the matrix-free operator; mg may be NULL (Jacobi)                     */
int sktb_pcg_solve_grid(sktb_pcg *s, sktb_mg *mg, const sktb_gridop *op,
                        const double *inv_diag, const double *b, double *x,
                        int use_x0, double rtol, int maxiter, int check_every,
                        int32_t *info_h, double *relres_h, void *stream);
int sktb_pcg_lambda_max_grid(sktb_pcg *s, const sktb_gridop *op,
                             const double *inv_diag, int iters, double *out_h,
                             void *stream);
/* nu pre- and nu post-smoothing sweeps on one level (default 1)                 */
int sktb_mg_set_level_sweeps(sktb_mg *m, int level, int nu);
/* Chebyshev polynomial smoother of degree nu on a coarse level (>= 1) instead of
 * nu damped-Jacobi sweeps: d_k = c1_h[k] d_{k-1} + c2_h[k] D^-1 r_k, x += d_k    */
int sktb_mg_set_level_cheby(sktb_mg *m, int level, int nu, const double *c1_h,
                            const double *c2_h);
/* fp32_level0 != 0: the two products with a matrix-free level-0 operator inside
 * the V-cycle are formed in single precision (vectors stay fp64)              */
int sktb_mg_set_precision(sktb_mg *m, int fp32_level0);
/* on: all levels with <= 1024 nodes run inside one cooperative kernel (grid-wide
 * barriers instead of ~14 launches per level); off (default): one launch per
 * operation on every level                                                    */
int sktb_mg_set_fused_tail(sktb_mg *m, int on);
/* exact coarsest-level solve: dense Gauss-Jordan inverse of the last level's
 * operator (<= 160 dofs; larger levels keep damped-Jacobi sweeps).  Call after
 * sktb_mg_set_level(last, ...) whenever its values changed.                    */
int sktb_mg_factor_coarsest(sktb_mg *m, void *stream);
/* level 0 of the multigrid hierarchy applied matrix-free                       */
int sktb_mg_set_level0_grid(sktb_mg *m, const sktb_gridop *op, int64_t n_nodes,
                            const double *inv_diag, const uint8_t *mask);
/* level 0 -> 1 fast path: children are scale[e]*Ke0[cls[e]], so out[E] =
 * sum_c scale[child] * T[(cls*8 + ptype[E])*8 + c] with the precomputed tables
 * T = Q_c^T Ke0[cls] Q_c (576 doubles each)                                    */
int sktb_elem_combine(int64_t n_coarse, const int32_t *child,
                      const uint8_t *ptype, const double *T, const int32_t *cls,
                      const double *scale, double *out, void *stream);
int sktb_elem_combine_range(int64_t n_coarse, int64_t e_lo, int64_t e_hi,
                            const int32_t *child, const uint8_t *ptype,
                            const double *T, const int32_t *cls, const double *scale,
                            double *out, void *stream);
/* Stress tensor sigma = 2 mu eps(u) + lam tr eps(u) I at every quadrature point
 * (fea/composer.py:444-494 stress_tensor_skfem); G = physical shape-function
 * gradients [class][q][a][3] from sktb_geom_tables; out[3][3][n_elem][nqp].      */
int sktb_element_stress(const sktb_mesh *m, int nqp, const int32_t *elem_class,
                        const double *G, const double *E_elem, double nu, const double *u,
                        double *out, void *stream);
/* Batched, strided, row-major fp64 product C[b] = A[b] . B[b] (optionally times
 * scale[b] elementwise, same layout as C): the six small dense products of the
 * direct (fast-diagonalisation) Helmholtz solve that replaces the sparse LU of
 * filters/helmholtz_filter_nodal.py:121-157 on tensor grids.                    */
int sktb_dgemm_batched(int M, int N, int K, const double *A, int lda, int64_t stride_a,
                       const double *B, int ldb, int64_t stride_b, double *C, int ldc,
                       int64_t stride_c, int batch, const double *scale, int64_t stride_s,
                       void *stream);
/* ---- scalar multigrid on tensor grids (heat conduction: the reference's sparse
 * LU of K, fea/solver_heat.py:191-192, becomes an MG-preconditioned PCG).  The
 * operators are kept in a 27-point stencil ("DIA") format, vals[k][node] with
 * k = 9 (dz+1) + 3 (dx+1) + (dy+1): level 0 is converted from the caller's
 * enforced CSR matrix (conduction + Robin terms), the coarse operators are the
 * algebraic Galerkin products P^T A P with trilinear P.  np_h: nodes per axis
 * (x, y, z) of every level, [n_levels][3]; the coarsest level (<= 160 nodes) is
 * solved exactly.  Transfer tables as in sktb_mg_set_transfer; mask = per-node
 * uint8 of fixed (Dirichlet) nodes per level (device, caller-owned, may be NULL). */
typedef struct sktb_smg sktb_smg;
int sktb_smg_create(sktb_smg **out, int n_levels, const int32_t *np_h, int device);
void sktb_smg_destroy(sktb_smg *m);
int sktb_smg_set_mask(sktb_smg *m, int level, const uint8_t *mask);
int sktb_smg_set_level_sweeps(sktb_smg *m, int level, int nu);
int sktb_smg_set_transfer(sktb_smg *m, int level, const int32_t *ax_c0,
                          const int32_t *ax_c1, const double *ax_w0,
                          const double *ax_w1, const int32_t *axT_f,
                          const double *axT_w);
int sktb_smg_setup_csr(sktb_smg *m, const int32_t *row_ptr, const int32_t *col_idx,
                       const double *vals, void *stream);
int sktb_smg_vcycle(sktb_smg *m, const double *r, double *z, void *stream);
/* y = A_level x and a copy of a level's 27 x n stencil values (tests)           */
int sktb_smg_apply(sktb_smg *m, int level, const double *x, double *y, void *stream);
int sktb_smg_level_values(sktb_smg *m, int level, double *out27n, void *stream);
int sktb_pcg_solve_smg(sktb_pcg *s, sktb_smg *smg, const double *b, double *x,
                       int use_x0, double rtol, int maxiter, int check_every,
                       int32_t *info_h, double *relres_h, void *stream);
/* PCG preconditioned by the V-cycle (level 0 may be row-sharded)               */
int sktb_pcg_solve_bsr3_mg(sktb_pcg *s, sktb_mg *mg, const int32_t *node_ptr,
                           const int32_t *node_col, int64_t n_blocks,
                           int max_deg, const double *vals,
                           const double *inv_diag, const double *b, double *x,
                           int use_x0, double rtol, int maxiter,
                           int check_every, int32_t *info_h, double *relres_h,
                           void *stream);

/* lambda_max(D^-1 A) by `iters` power iterations on the (possibly row-sharded)
 * node-block operator; used to place the multigrid smoother's damping.        */
int sktb_pcg_lambda_max_bsr3(sktb_pcg *s, const int32_t *node_ptr,
                             const int32_t *node_col, int64_t n_blocks,
                             int max_deg, const double *vals,
                             const double *inv_diag, int iters, double *out_h,
                             void *stream);

/* -------------------------------------------------- multi-GPU --
 * Row-sharded operator: rank r owns the rows of the contiguous node range
 * [node_begin, node_end) (its dofs [dpn*node_begin, dpn*node_end)); column
 * indices stay global.  The reference has no distributed code (PETSc objects
 * live on COMM_SELF, fea/solver_petsc.py:143,157); this is the B200 addition. */
int sktb_mesh_dof_pattern_rows(const sktb_mesh *m, int dpn, int64_t node_begin,
                               int64_t node_end, int32_t *row_ptr,
                               int32_t *col_idx, void *stream);
int sktb_assemble_rows(const sktb_mesh *m, int dpn, int64_t node_begin,
                       int64_t node_end, const double *unit_ke,
                       const int32_t *elem_class, const double *scale,
                       const uint8_t *dir_mask, double *vals, void *stream);
int sktb_csr_inv_diag_rows(int64_t n_rows, int64_t row0, const int32_t *row_ptr,
                           const int32_t *col_idx, const double *vals,
                           double *out, void *stream);
/* NCCL communicator (one process per GPU); id128_h = 128-byte ncclUniqueId
 * created on rank 0 by sktb_comm_unique_id and shipped by the host launcher.  */
typedef struct sktb_comm sktb_comm;
int sktb_comm_unique_id(void *id128_h);
int sktb_comm_create(sktb_comm **out, const void *id128_h, int rank, int world,
                     int device);
void sktb_comm_destroy(sktb_comm *c);
int sktb_comm_rank(const sktb_comm *c);
int sktb_comm_world(const sktb_comm *c);
int sktb_comm_allreduce_sum(sktb_comm *c, const double *src, double *dst,
                            int64_t count, void *stream);
/* Peer-memory halos (NVLink P2P through cudaIpc; NCCL transport, one GPU per
 * rank): every rank creates one arena of `bytes` and gets its 64-byte IPC
 * handle; the launcher all-gathers the handles (rank order) and every rank maps
 * its neighbours' arenas with sktb_comm_arena_open.  Full-length slab-sharded
 * vectors (PCG direction, multigrid iterates) are then placed in the arena at
 * the same offset on every rank and their ghost planes are pulled straight from
 * the neighbour's copy by one kernel per exchange (device-side flags), instead of
 * ncclSend / ncclRecv.  status: bytes used, exchanges done, error flag (a spin
 * on a neighbour timed out).                                                    */
int sktb_comm_arena_create(sktb_comm *c, int64_t bytes, void *handle64_h);
int sktb_comm_arena_open(sktb_comm *c, const void *handles_h);
int sktb_comm_arena_status(const sktb_comm *c, int64_t *used_h, int64_t *epoch_h,
                           int32_t *err_h);
/* in-place all-gather of contiguous slices buf[displs[r] .. +counts[r])        */
int sktb_comm_allgatherv(sktb_comm *c, double *buf, const int64_t *counts_h,
                         const int64_t *displs_h, void *stream);
/* PCG workspace for the rows [row0, row0+n_local) of an n_global system.  The
 * halo is described per peer: send_idx_h[send_off_h[i]..send_off_h[i+1]) are
 * the global indices of owned entries peer i needs, recv_idx_h[...] the global
 * indices received from it (ghost slots of the full-length direction vector).
 * sktb_pcg_solve then takes local row_ptr/vals/inv_diag/b/x and global col_idx;
 * dot products are all-reduced in the stream (2 all-reduces per iteration).    */
/* first host poll of the next solve after n iterations (then every check_every) */
int sktb_pcg_set_first_batch(sktb_pcg *s, int n);
/* z-slab halo of a row-sharded PCG on a tensor grid: whole planes of plane_dofs
 * entries exchanged with prev_rank / next_rank (-1: none) straight from / into
 * the full-length vectors, instead of the packed index lists                   */
int sktb_pcg_set_slab_halo(sktb_pcg *s, int64_t plane_dofs, int prev_rank,
                           int next_rank);
int sktb_pcg_create_dist(sktb_pcg **out, sktb_comm *comm, int64_t n_global,
                         int64_t row0, int64_t n_local, int n_peers,
                         const int32_t *peers_h, const int64_t *send_off_h,
                         const int32_t *send_idx_h, const int64_t *recv_off_h,
                         const int32_t *recv_idx_h, int device);

/* In-situ timing of the dominant kernel: with every_n > 0 the SpMV launch of
 * the first iteration of every every_n-th batch is bracketed by CUDA events on
 * the solver's stream; get_profile returns the accumulated milliseconds and
 * the number of samples since the last set_profile.                          */
int sktb_pcg_set_profile(sktb_pcg *s, int every_n);
int sktb_pcg_get_profile(const sktb_pcg *s, double *ms_sum_h, int64_t *count_h);

/* ------------------------------------------------------ element kernels ---*/
/* K1: E = Emin + (E0-Emin) rho^p (fea/composer.py:19-22); ramp != 0 gives
 * Emin + (E0-Emin) rho/(1+p(1-rho)) (fea/composer.py:25-39)                   */
int sktb_interpolate_modulus(int64_t n, const double *rho, double E0,
                             double Emin, double p, int ramp, double *out,
                             void *stream);
/* K7: U_e = 1/2 scale[e] u_e^T Ke0[class[e]] u_e
 * (fea/solver_elastic.py:470-537, fea/solver_heat.py:256-303)                 */
int sktb_element_energy(const sktb_mesh *m, int dpn, const double *unit_ke,
                        const int32_t *elem_class, const double *scale,
                        const double *u, double *out, void *stream);
/* K7 for 3-dof hexahedral meshes whose elements share ONE geometry class (tensor
 * grids): unit_ke_h = the 24x24 unit matrix on the HOST (passed as kernel
 * parameters); same result as sktb_element_energy up to summation order        */
int sktb_element_energy_hex_uniform(const sktb_mesh *m, const double *unit_ke_h,
                                    const double *scale, const double *u,
                                    double *out, void *stream);
/* out[e] = factor scale[e] u_e^T Ke0[class[e]] v_e : the elemental integrals of
 * grad T . grad lambda behind energy_multi_load for the heat_exchange and
 * averaged_temp objectives (fea/solver_heat.py:327-383,518-549; scale = NULL,
 * factor = 1 there)                                                           */
int sktb_element_bilinear(const sktb_mesh *m, int dpn, const double *unit_ke,
                          const int32_t *elem_class, const double *scale,
                          const double *u, const double *v, double factor,
                          double *out, void *stream);
/* K8: g = -2 U dE/drho / max(E,1e-12) * dH  (core/derivatives.py:42-68,
 * core/projection.py:80-118, common_density.py:1083-1097); dH may be NULL.    */
int sktb_dc_drho(int64_t n, const double *rho_proj, const double *energy,
                 double E0, double Emin, double p, int ramp, const double *dH,
                 double *out, void *stream);
/* K13a: out = H_beta(x), dH = dH/dx (core/projection.py:27-118); either output
 * may be NULL.                                                                 */
int sktb_heaviside(int64_t n, const double *x, double beta, double eta,
                   double *out, double *dH, void *stream);

/* K9: Helmholtz filter transfer operators
 * (filters/helmholtz_filter_nodal.py:26-56).
 * e2n: out[n] = sum_{e in n} w[e]*v(e) / wsum[n], v(e) = val[e] if
 *      design[e] (or design == NULL) else fixed_value; wsum from
 *      sktb_e2n_wsum (sum of w over the node's elements, 0 -> 1).
 * n2e: out[e] = mean of x over the element's nodes; clamp_max0 != 0 applies
 *      min(.,0) (helmholtz_filter_nodal.py:232).                               */
int sktb_e2n(const sktb_mesh *m, const double *w, const double *val,
             const uint8_t *design, double fixed_value, const double *wsum,
             double *out, void *stream);
int sktb_e2n_wsum(const sktb_mesh *m, const double *w, double *out,
                      void *stream);
int sktb_n2e_mean(const sktb_mesh *m, const double *x, int clamp_max0,
                  double *out, void *stream);

/* ------------------------------------------------- heat: Robin terms -------
 * K16/K17 (fea/solver_heat.py:552-625,723-745,928-980).
 * Quadrature tables per geometry class (what Basis.interpolate evaluates):
 * N[cls][q][a], physical gradients G[cls][q][a][3], dx[cls][q] = w_q |det J|. */
int sktb_geom_tables(const sktb_mesh *m, int nqp, const double *X_h,
                     const double *W_h, int64_t n_class,
                     const int32_t *class_rep_h, double *N_out, double *G_out,
                     double *dx_out, void *stream);
/* Mq[cls][q][a][b] = dx_q N_a N_b  (unit matrices of the virtual Robin form)  */
int sktb_unit_qp_mass(const sktb_mesh *m, int nqp, int64_t n_class,
                      const double *N_tab, const double *dx_tab, double *out,
                      void *stream);
/* s[q][e] = h rho_q^p (1-rho_q)^q |grad rho_q|, rho interpolated from nodal
 * values (get_robin_virtual, fea/solver_heat.py:552-572)                      */
int sktb_robin_virtual_scale(const sktb_mesh *m, int nqp,
                             const int32_t *elem_class, const double *N_tab,
                             const double *G_tab, const double *dx_tab,
                             const double *rho_node, double h, double p,
                             double q, double *out, void *stream);
/* scalar gather assembly with n_terms (scale[k][e], unit[cls][k][a][b]) terms */
int sktb_assemble_terms(const sktb_mesh *m, int n_terms, const double *unit,
                        const int32_t *elem_class, const double *scale,
                        double *vals, void *stream);
/* element-local vectors of _robin_compliance_explicit_grad_
 * (fea/solver_heat.py:575-597): out_local[a][e]                               */
int sktb_robin_explicit_local(const sktb_mesh *m, int nqp,
                              const int32_t *elem_class, const double *N_tab,
                              const double *G_tab, const double *dx_tab,
                              const double *rho_node, const double *T, double h,
                              double T_env, double p, double q,
                              double *out_local, void *stream);
/* heat_exchange objective on the interface measure |grad rho_n|
 * (fea/solver_heat.py:306-324,385-413,416-446): per element den_e = int |g|,
 * num_e = int -T_env h (T - T_env) |g| (num_out / T may be NULL) and the
 * element-local adjoint load local[a][e] = int -T_env heff |g| N_a with heff
 * interpolated from the nodal values h rho_n^p (1-rho_n)^q (local_out may be
 * NULL)                                                                        */
int sktb_heat_exchange_local(const sktb_mesh *m, int nqp,
                             const int32_t *elem_class, const double *N_tab,
                             const double *G_tab, const double *dx_tab,
                             const double *rho_node, const double *T,
                             double p, double q, double h, double T_env,
                             double *den_out, double *num_out,
                             double *local_out, void *stream);
/* out[n] = sum over the node's (element, local) slots of local[a][e]
 * (optionally divided by divisor[n]): load-vector style assembly              */
int sktb_local_to_nodes(const sktb_mesh *m, const double *local,
                        const double *divisor, double *out, void *stream);

/* ------------------------------------------------------ update kernels ----*/
/* K12: OC candidate (core/optimizers/oc.py:50-68).  rho_e, dC over design
 * elements; writes scaling_rate, rho_cand (design) and scatters rho_cand into
 * rho_full_cand[design_idx].                                                   */
int sktb_oc_candidate(int64_t n_design, const double *dC, const double *rho_e,
                      double lmid, double eps, double eta, double move_limit,
                      double rho_min, double rho_max, double sr_min,
                      double sr_max, const int32_t *design_idx,
                      double *scaling_rate, double *rho_cand,
                      double *rho_full_cand, void *stream);
/* K14: log-space MOC step (core/optimizers/logmoc.py:36-66), rho in/out.      */
int sktb_logmoc_update(int64_t n, double *rho, const double *dL, double eta,
                       double move_limit, double rho_min, double rho_max,
                       double clip, double *scaling_rate, double *clip_lower,
                       double *clip_upper, void *stream);
/* K13b: deterministic single-kernel reductions; result written to out_h after
 * synchronising the stream.  idx may be NULL (identity).
 *   wsum : sum_i a[idx[i]] * (w ? w[i] : 1)
 *   stats: out_h[0..3] = min, mean, max, std (population) of a[idx[i]]
 *   absmax: max_i |a[i]|                                                      */
int sktb_reduce_wsum_h(int64_t n, const double *a, const int32_t *idx,
                       const double *w, double *out_h, void *stream);
int sktb_reduce_stats_h(int64_t n, const double *a, const int32_t *idx,
                        double *out_h, void *stream);
int sktb_reduce_absmax_h(int64_t n, const double *a, double *out_h,
                         void *stream);
/* dot product (compliance F.u, fea/solver_elastic.py:236)                     */
int sktb_dot_h(int64_t n, const double *a, const double *b, double *out_h,
               void *stream);
/* K15: np.percentile(np.abs(a), q) with linear interpolation
 * (core/optimizers/oc.py:189, logmoc.py:167-174,220); work: >= n uint64 +
 * 8192 bytes of scratch.  Synchronises.                                        */
int sktb_abs_percentile_h(int64_t n, const double *a, double q, void *work,
                          double *out_h, void *stream);
/* gather / scatter by index (rho[design] <-> rho_design)                      */
int sktb_gather(int64_t n, const double *src, const int32_t *idx, double *dst,
                void *stream);
int sktb_scatter(int64_t n, const double *src, const int32_t *idx, double *dst,
                 void *stream);
/* y = a*x + b*y elementwise helpers used by the filter RHS / LogMOC dL         */
int sktb_axpby(int64_t n, double a, const double *x, double b, double *y,
               void *stream);
/* out = a*x + b*y + c (y may be NULL); out = a*x*y                             */
int sktb_affine(int64_t n, double a, const double *x, double b,
                const double *y, double c, double *out, void *stream);
int sktb_hadamard(int64_t n, double a, const double *x, const double *y,
                  double *out, void *stream);
/* ---- host-side helpers of the task construction (HOST pointers, multi-threaded
 * C++, no device work).  Each reproduces its NumPy counterpart
 * in sktopt/ bit for bit.
 * sktb_host_hex_volumes: get_elements_volume for hexahedra (fea/composer.py:191-248):
 *   sum of six |tetrahedron volumes| on the reference's local quadruples;
 *   t = connectivity [8][n_elem], p = coordinates [3][n_nodes].
 * sktb_host_lattice_facets: facets / t2f / f2t / f2lf (what skfem's Mesh provides and
 *   mesh/task_common.py:105-270 reads) of a hexahedral mesh whose elements are cells
 *   of a lattice numbered like init_tensor (npy = nodes along y, P = npy*npx), in two
 *   passes: facets == NULL validates, fills key [6][n_elem] and *n_facets (-1: not
 *   such a mesh); the second call fills facets [4][n], t2f [6][n_elem], f2t [2][n],
 *   f2lf [2][n] (lexicographic facet order, slot 0 = first occurrence).              */
int sktb_host_hex_volumes(int64_t n_elem, int64_t n_nodes, const int32_t *t, const double *p,
                          double *vol);
int sktb_host_lattice_facets(int64_t n_elem, int64_t n_nodes, const int32_t *t,
                             const int32_t *lf, int64_t npy, int64_t P, int64_t *key,
                             int64_t *n_facets, int32_t *facets, int32_t *t2f, int32_t *f2t,
                             int8_t *f2lf);
/* out = |x| (x != NULL), else out = value                                      */
int sktb_fill_abs(int64_t n, const double *x, double value, double *out, void *stream);
/* KKT residual (core/optimizers/oc.py:230-240, logmoc.py:227-236):
 * out_h[0] = max |g + coef*dv| over lo < rho < hi, out_h[1] = how many such.  */
int sktb_kkt_residual_h(int64_t n, const double *rho, const double *g,
                        const double *dv, double coef, double lo, double hi,
                        double *out_h, void *stream);
/* max_i |a[idx[i]] - b[idx[i]]|  (rho_change_max, common_density.py:1137-1141) */
int sktb_reduce_maxdiff_h(int64_t n, const double *a, const double *b,
                          const int32_t *idx, double *out_h, void *stream);
/* rhs for enforce with prescribed values: out = mask ? xD : b - t             */
int sktb_enforce_rhs(int64_t n, const double *b, const double *t,
                     const uint8_t *mask, const double *xD, double *out,
                     void *stream);

/* benchmark utility: overwrite a > L2-sized scratch buffer                     */
int sktb_flush_l2(void *scratch, int64_t bytes, void *stream);
/* benchmark utility: FP64 FMA throughput probe; out holds 148*8*256 doubles,
 * *flops_h receives the flops of the launch                                    */
int sktb_fp64_probe(int iters, double *out, int64_t *flops_h, void *stream);
/* same with the multiplier read from the constant bank (iters % 8 == 0)        */
int sktb_fp64_probe_const(int iters, double *out, int64_t *flops_h, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SKTOPT_B200_H */
