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

/* ---- assembly ------------------------------------------------------------------------------------------------
 * What blf->Assemble() (/root/reference/src/equation_systems/equation_system.cpp:345) followed by the essential-row /
 * column elimination of FormLinearSystem (ibid. :132-190, diagonal policy ONE) produce, from element matrices:
 *   elem_dof [ne][nd]   signed MFEM encoding of FiniteElementSpace::GetElementDofs (k, or -1-k = flipped)
 *   mat_id  [ne]        which of the nmat local element matrices elmat [nmat][nd][nd] the element uses (elements of
 *                       a structured mesh share the matrix of their attribute)
 *   ess     [n]         1 = essential dof (row and column replaced by the identity), may be NULL
 * Row-wise: every row gathers the contributions of the elements touching it, sorts them by column and merges
 * duplicates (fixed order: the result does not depend on the thread count).  Pass 1 (col == NULL) only counts:
 * rowptr [n+1] is written; pass 2 fills col / val. */
typedef struct { int32_t c; double v; } entry_t;
static int cmp_entry(const void *a, const void *b) {
  const int32_t x = ((const entry_t *)a)->c, y = ((const entry_t *)b)->c;
  return (x > y) - (x < y);
}
int csr_assemble(int64_t n, int64_t ne, int nd, const int32_t *elem_dof, const int32_t *mat_id, const double *elmat,
                 const uint8_t *ess, int64_t *rowptr, int32_t *col, double *val) {
  /* dof -> (element, local index) adjacency */
  int64_t *aoff = calloc((size_t)n + 2, sizeof(int64_t));
  if (!aoff) return 1;
  for (int64_t i = 0; i < ne * nd; i++) { const int32_t d = elem_dof[i]; aoff[(d >= 0 ? d : -1 - d) + 2]++; }
  for (int64_t i = 0; i < n; i++) aoff[i + 2] += aoff[i + 1];
  int64_t *adj = malloc(sizeof(int64_t) * (size_t)(ne * nd));
  if (!adj) { free(aoff); return 1; }
  for (int64_t i = 0; i < ne * nd; i++) { const int32_t d = elem_dof[i]; adj[aoff[(d >= 0 ? d : -1 - d) + 1]++] = i; }
  int64_t maxadj = 0;
  for (int64_t i = 0; i < n; i++) if (aoff[i + 1] - aoff[i] > maxadj) maxadj = aoff[i + 1] - aoff[i];
  const int fill = col != NULL;
  if (!fill) rowptr[0] = 0;
  int fail = 0;
#pragma omp parallel
  {
    entry_t *buf = malloc(sizeof(entry_t) * (size_t)(maxadj * nd + 1));
    if (!buf) {
#pragma omp atomic write
      fail = 1;
    }
#pragma omp for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
      if (!buf) continue;
      int64_t m = 0;
      if (ess && ess[i]) { buf[m].c = (int32_t)i; buf[m++].v = 1.0; }
      else
        for (int64_t q = aoff[i]; q < aoff[i + 1]; q++) {
          const int64_t e = adj[q] / nd;
          const int li = (int)(adj[q] % nd);
          const double si = elem_dof[e * nd + li] < 0 ? -1.0 : 1.0;
          const double *M = elmat + ((int64_t)mat_id[e] * nd + li) * nd;
          for (int lj = 0; lj < nd; lj++) {
            const int32_t d = elem_dof[e * nd + lj], cj = d >= 0 ? d : -1 - d;
            if (ess && ess[cj]) continue;
            buf[m].c = cj;
            buf[m++].v = (d < 0 ? -si : si) * M[lj];
          }
        }
      qsort(buf, (size_t)m, sizeof(entry_t), cmp_entry);
      int64_t u = 0;
      for (int64_t k = 0; k < m; k++) {
        if (u > 0 && buf[u - 1].c == buf[k].c) buf[u - 1].v += buf[k].v;
        else buf[u++] = buf[k];
      }
      if (!fill) rowptr[i + 1] = u;
      else
        for (int64_t k = 0; k < u; k++) { col[rowptr[i] + k] = buf[k].c; val[rowptr[i] + k] = buf[k].v; }
    }
    free(buf);
  }
  if (!fill && !fail) for (int64_t i = 0; i < n; i++) rowptr[i + 1] += rowptr[i];
  free(adj); free(aoff);
  return fail;
}
