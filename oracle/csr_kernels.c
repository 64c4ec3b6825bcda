/* CPU restatement of the reference's assembled-operator hot loop (oracle; test
 * infrastructure and the timed CPU baseline -- never linked into the product).
 *
 *  - csr_spmv:  y = A x on an assembled CSR matrix, rows blocked over threads --
 *    what hypre_ParCSRMatrixMatvec does per MPI rank behind
 *    EquationSystem::Mult (/root/reference/src/equation_systems/equation_system.cpp:200-204).
 *  - csr_pcg:   hypre_PCGSolve semantics (two_norm = 0) with diagonal scaling, the
 *    configuration of DefaultJacobiPCGSolver
 *    (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57).
 * FP64 values, int32 column indices, int64 row pointers.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int csr_num_threads(void) { return omp_get_max_threads(); }
void csr_set_threads(int n) { omp_set_num_threads(n); }

void csr_spmv(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    double s = 0.0;
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; k++) s += val[k] * x[col[k]];
    y[i] = s;
  }
}

static double dot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

/* returns iterations; x holds the initial guess on entry (zero it for iterative_mode=false) */
int csr_pcg(int64_t n, const int64_t *rowptr, const int32_t *col, const double *val, const double *dinv,
            const double *b, double *x, double rtol, double atol, int max_iter, double *final_rel, int *converged) {
  double *r = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n), *s = malloc(sizeof(double) * n);
  *converged = 0;
  double bi = 0.0;
#pragma omp parallel for reduction(+ : bi) schedule(static)
  for (int64_t i = 0; i < n; i++) bi += (dinv ? dinv[i] : 1.0) * b[i] * b[i];
  if (!(bi > 0.0)) {
    memcpy(x, b, sizeof(double) * n);
    *final_rel = 0.0; *converged = 1;
    free(r); free(p); free(s);
    return 0;
  }
  const double eps = fmax(rtol * rtol, atol * atol / bi);
  csr_spmv(n, rowptr, col, val, x, r);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) { r[i] = b[i] - r[i]; p[i] = (dinv ? dinv[i] : 1.0) * r[i]; }
  double gamma = dot(n, r, p);
  int it = 0;
  while (it + 1 <= max_iter) {
    it++;
    csr_spmv(n, rowptr, col, val, p, s);
    const double sdotp = dot(n, s, p);
    if (sdotp == 0.0) break;
    const double alpha = gamma / sdotp;
    double gn = 0.0;
#pragma omp parallel for reduction(+ : gn) schedule(static)
    for (int64_t i = 0; i < n; i++) {
      x[i] += alpha * p[i];
      r[i] -= alpha * s[i];
      s[i] = (dinv ? dinv[i] : 1.0) * r[i];
      gn += r[i] * s[i];
    }
    if (gn / bi < eps) { gamma = gn; *converged = 1; break; }
    if (fabs(gn) < 1e-292) { gamma = gn; break; }
    const double beta = gn / gamma;
    gamma = gn;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) p[i] = s[i] + beta * p[i];
  }
  *final_rel = sqrt(fabs(gamma) / bi);
  free(r); free(p); free(s);
  return it;
}
