/* CPU ORACLE (test infrastructure, NOT product code): C / OpenMP restatement of
 * the heavy stages of the reference's hot path, so that the CPU baseline can run
 * BASELINE's full-size configurations on all host cores instead of being
 * extrapolated from a small mesh.  Only tests/, __graft_entry__ and bench.py's
 * CPU legs may build, load or call this file (see oracle/cport/__init__.py).
 *
 * What it restates (paths relative to /root/reference/scikit-topt/sktopt/):
 *   cport_assemble     skfem asm(BilinearForm) as used by
 *                      fea/composer.py:88-100 (stiffness) and :139-144
 *                      (conduction): K = sum_e E_e Ke0[class_e], CSR with sorted
 *                      indices, duplicates summed; here as a per-row gather over
 *                      the node->element adjacency (deterministic order:
 *                      elements ascending), with skfem.enforce
 *                      (fea/solver_elastic.py:211) folded in when a mask is given
 *   cport_spmv         scipy CSR mat-vec (the product inside scipy cg)
 *   cport_pcg_jacobi   scipy.sparse.linalg.cg(K, f, M=1/diag, rtol, atol=0,
 *                      x0=0, maxiter) as called at fea/solver_elastic.py:84-92:
 *                      stop when ||r||_2 <= rtol ||b||_2
 *   cport_energy       _strain_energy_density_.elemental
 *                      (fea/solver_elastic.py:470-537) = 1/2 E_e u_e^T Ke0 u_e
 *                      (the same quadrature is inside Ke0)
 * Everything is fp64; indices are int32 (columns) / int64 (row pointers).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int cport_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void cport_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* dof-level CSR pattern from the node graph: row dpn*n+i holds the columns
 * dpn*m+j of every neighbour m (ascending) -- the sorted-index pattern
 * scipy's COO->CSR conversion produces */
void cport_expand_pattern(int64_t n_nodes, int dpn, const int64_t *nptr,
                          const int32_t *ncol, int64_t *indptr, int32_t *indices) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < n_nodes; ++n) {
    const int64_t s = nptr[n], deg = nptr[n + 1] - s;
    for (int i = 0; i < dpn; ++i) {
      const int64_t row = dpn * n + i;
      const int64_t o = dpn * dpn * s + (int64_t)i * dpn * deg;
      indptr[row] = o;
      for (int64_t k = 0; k < deg; ++k)
        for (int j = 0; j < dpn; ++j)
          indices[o + dpn * k + j] = (int32_t)(dpn * ncol[s + k] + j);
    }
  }
  indptr[dpn * n_nodes] = dpn * dpn * nptr[n_nodes];
}

static inline int64_t find_col(const int32_t *ncol, int64_t s, int64_t deg, int32_t c) {
  int64_t lo = 0, hi = deg - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ncol[s + mid] < c) lo = mid + 1; else hi = mid;
  }
  return lo;
}

/* K = sum_e scale[e] * Ke[cls[e]]  (cls == NULL: one shared matrix, or one
 * matrix per element when per_elem != 0).  conn is (nen, n_elem) row-major as
 * in skfem's mesh.t.  n2e_*: node -> (element, local vertex) adjacency with
 * elements ascending.  mask (n_dof, may be NULL): Dirichlet dofs, rows and
 * columns replaced by identity (skfem.enforce). */
void cport_assemble(int64_t n_nodes, int64_t n_elem, int nen, int dpn,
                    const int32_t *conn, const int64_t *n2e_ptr,
                    const int32_t *n2e_elem, const uint8_t *n2e_loc,
                    const int64_t *nptr, const int32_t *ncol,
                    const double *Ke, const int32_t *cls, int per_elem,
                    const double *scale, const uint8_t *mask, double *data) {
  const int nde = nen * dpn;
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t n = 0; n < n_nodes; ++n) {
    const int64_t s = nptr[n], deg = nptr[n + 1] - s;
    double *blk = data + (int64_t)dpn * dpn * s; /* dpn rows of dpn*deg values */
    memset(blk, 0, sizeof(double) * dpn * dpn * deg);
    for (int64_t q = n2e_ptr[n]; q < n2e_ptr[n + 1]; ++q) {
      const int64_t e = n2e_elem[q];
      const int a = n2e_loc[q];
      const double sc = scale ? scale[e] : 1.0;
      const double *K = Ke + (per_elem ? e : (cls ? cls[e] : 0)) * (int64_t)nde * nde;
      for (int b = 0; b < nen; ++b) {
        const int32_t m = conn[(int64_t)b * n_elem + e];
        const int64_t k = find_col(ncol, s, deg, m);
        for (int i = 0; i < dpn; ++i)
          for (int j = 0; j < dpn; ++j)
            blk[(int64_t)i * dpn * deg + dpn * k + j] +=
                sc * K[(dpn * a + i) * nde + dpn * b + j];
      }
    }
    if (mask) {
      for (int i = 0; i < dpn; ++i) {
        const int64_t row = dpn * n + i;
        for (int64_t k = 0; k < deg; ++k)
          for (int j = 0; j < dpn; ++j) {
            const int64_t col = (int64_t)dpn * ncol[s + k] + j;
            if (mask[row] || mask[col])
              blk[(int64_t)i * dpn * deg + dpn * k + j] = (row == col) ? 1.0 : 0.0;
          }
      }
    }
  }
}

void cport_spmv(int64_t n, const int64_t *indptr, const int32_t *indices,
                const double *data, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n; ++r) {
    double acc = 0.0;
    for (int64_t k = indptr[r]; k < indptr[r + 1]; ++k) acc += data[k] * x[indices[k]];
    y[r] = acc;
  }
}

/* scipy cg with M = diag^-1, x0 = 0, atol = 0.  Returns the iteration count;
 * *relres = ||r|| / ||b|| at exit; converged iff *relres <= rtol. */
int64_t cport_pcg_jacobi(int64_t n, const int64_t *indptr, const int32_t *indices,
                         const double *data, const double *b, double *x,
                         double rtol, int64_t maxiter, double *relres) {
  double *r = (double *)malloc(sizeof(double) * n);
  double *z = (double *)malloc(sizeof(double) * n);
  double *p = (double *)malloc(sizeof(double) * n);
  double *q = (double *)malloc(sizeof(double) * n);
  double *minv = (double *)malloc(sizeof(double) * n);
  double bb = 0.0, rz = 0.0, rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz, rr)
  for (int64_t i = 0; i < n; ++i) {
    double d = 1.0;
    for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k)
      if (indices[k] == i) d = data[k];
    minv[i] = 1.0 / d;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = minv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
    rr += r[i] * r[i];
  }
  const double tol2 = rtol * rtol * bb;
  int64_t it = 0;
  while (rr > tol2 && it < maxiter) {
    double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < n; ++i) {
      double acc = 0.0;
      for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) acc += data[k] * p[indices[k]];
      q[i] = acc;
      pq += p[i] * acc;
    }
    const double alpha = rz / pq;
    double rz_new = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz_new, rr)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      const double ri = r[i] - alpha * q[i];
      const double zi = minv[i] * ri;
      r[i] = ri;
      z[i] = zi;
      rz_new += ri * zi;
      rr += ri * ri;
    }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
    ++it;
  }
  if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(r); free(z); free(p); free(q); free(minv);
  return it;
}

/* out[e] = 1/2 scale[e] u_e^T Ke[cls[e]] u_e */
void cport_energy(int64_t n_elem, int nen, int dpn, const int32_t *conn,
                  const double *Ke, const int32_t *cls, const double *scale,
                  const double *u, double *out) {
  const int nde = nen * dpn;
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < n_elem; ++e) {
    double ue[24];
    for (int a = 0; a < nen; ++a)
      for (int i = 0; i < dpn; ++i)
        ue[dpn * a + i] = u[(int64_t)dpn * conn[(int64_t)a * n_elem + e] + i];
    const double *K = Ke + (cls ? cls[e] : 0) * (int64_t)nde * nde;
    double acc = 0.0;
    for (int r = 0; r < nde; ++r) {
      double t = 0.0;
      for (int c = 0; c < nde; ++c) t += K[r * nde + c] * ue[c];
      acc += ue[r] * t;
    }
    out[e] = 0.5 * (scale ? scale[e] : 1.0) * acc;
  }
}
