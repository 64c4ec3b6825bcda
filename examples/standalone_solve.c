/* Stand-alone use of libhcurl_b200 from plain C -- no MFEM, no Python:
 *   ExodusII file -> mesh (hcurl_host_exodus_*) -> ND space of order p (hcurl_host_ndspace_*) ->
 *   A = K(alpha) + M(beta) on the GPU (hcurl_space_create / hcurl_op_*) -> manufactured solve A x = A x_ref with
 *   Jacobi-PCG (hcurl_pcg_*), essential dofs on the given side sets (all boundary faces if none are given).
 * This is the chain the reference runs as  mfem::Mesh(file) -> ParFiniteElementSpace("ND_3D_P<p>") -> ParBilinearForm
 * with CurlCurlIntegrator + VectorFEMassIntegrator -> HyprePCG  (examples/team7/Main.cpp:152-155,
 * src/problem_builders/problem_builder_base.cpp:98-127,220-254).
 *
 *   cc -std=c99 -Iinclude examples/standalone_solve.c -Lhephaestus_b200/lib -lhcurl_b200 -lm -o standalone_solve
 *   LD_LIBRARY_PATH=hephaestus_b200/lib ./standalone_solve /path/to/team7.g 1 [side-set ids ...]
 * HEX27 files are used with their 8 corner nodes (straight-sided elements). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "hcurl_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    if ((call) != 0) {                                                       \
      fprintf(stderr, "%s failed: %s\n", #call, hcurl_last_error());         \
      return 1;                                                              \
    }                                                                        \
  } while (0)
#define CHECK_PTR(p)                                                         \
  do {                                                                       \
    if (!(p)) {                                                              \
      fprintf(stderr, "%s: %s\n", #p, hcurl_last_error());                   \
      return 1;                                                              \
    }                                                                        \
  } while (0)

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s mesh.g order [essential side-set ids ...]\n", argv[0]);
    return 2;
  }
  const int order = atoi(argv[2]);
  /* ---- mesh */
  hcurl_host_exodus *ex = hcurl_host_exodus_open(argv[1]);
  CHECK_PTR(ex);
  int64_t n_nodes, n_elem, n_side;
  int geom, npe;
  CHECK(hcurl_host_exodus_counts(ex, &n_nodes, &n_elem, &geom, &npe, &n_side));
  const int nc = geom == HCURL_GEOM_HEX ? 8 : 4, nfv = geom == HCURL_GEOM_HEX ? 4 : 3;
  double *xyz = malloc(sizeof(double) * 3 * n_nodes);
  int32_t *conn = malloc(sizeof(int32_t) * npe * n_elem), *attr = malloc(sizeof(int32_t) * n_elem);
  int32_t *bdr = malloc(sizeof(int32_t) * nfv * (n_side > 0 ? n_side : 1)), *battr = malloc(sizeof(int32_t) * (n_side > 0 ? n_side : 1));
  CHECK(hcurl_host_exodus_read(ex, xyz, conn, attr, n_side ? bdr : NULL, n_side ? battr : NULL));
  hcurl_host_exodus_close(ex);
  int32_t *corners = malloc(sizeof(int32_t) * nc * n_elem);
  double *elem_xyz = malloc(sizeof(double) * 3 * nc * n_elem);
  for (int64_t e = 0; e < n_elem; e++)
    for (int v = 0; v < nc; v++) {
      const int32_t g = conn[e * npe + v];
      corners[e * nc + v] = g;
      for (int d = 0; d < 3; d++) elem_xyz[(e * nc + v) * 3 + d] = xyz[(int64_t)g * 3 + d];
    }
  /* ---- ND space */
  hcurl_host_ndspace *nds = geom == HCURL_GEOM_HEX ? hcurl_host_ndspace_create_hex(order, n_elem, corners)
                                                    : hcurl_host_ndspace_create_tet(order, n_elem, corners);
  CHECK_PTR(nds);
  const int nd = hcurl_host_ndspace_nd(nds);
  const int64_t ndof = hcurl_host_ndspace_ndof(nds);
  int32_t *ldof = malloc(sizeof(int32_t) * nd * n_elem);
  CHECK(hcurl_host_ndspace_elem_dofs(nds, ldof));
  int32_t *face_base = NULL;
  if (geom == HCURL_GEOM_TET && order >= 2) {
    face_base = malloc(sizeof(int32_t) * 12 * n_elem);
    CHECK(hcurl_host_ndspace_elem_face_base(nds, face_base));
  }
  /* essential dofs: the requested side sets, or the whole boundary */
  int32_t *ess = malloc(sizeof(int32_t) * ndof);
  int64_t n_ess;
  if (argc > 3 && n_side > 0) {
    int64_t nsel = 0;
    int32_t *sel = malloc(sizeof(int32_t) * nfv * n_side);
    for (int64_t s = 0; s < n_side; s++)
      for (int a = 3; a < argc; a++)
        if (battr[s] == atoi(argv[a])) {
          for (int k = 0; k < nfv; k++) sel[nsel * nfv + k] = bdr[s * nfv + k];
          nsel++;
          break;
        }
    n_ess = hcurl_host_ndspace_face_dofs(nds, nsel, sel, ess);
    free(sel);
  } else {
    n_ess = hcurl_host_ndspace_boundary_dofs(nds, ess);
  }
  if (n_ess < 0) { fprintf(stderr, "essential dofs: %s\n", hcurl_last_error()); return 1; }
  hcurl_host_ndspace_destroy(nds);
  printf("%s: %lld %s, ND%d: %lld dofs (%lld essential)\n", argv[1], (long long)n_elem, geom == HCURL_GEOM_HEX ? "hexes" : "tets", order,
         (long long)ndof, (long long)n_ess);
  /* ---- operator on the GPU: alpha = 0.05 / attribute, beta = 1 */
  hcurl_ctx *ctx = hcurl_ctx_create(0, NULL, 0, 1);
  CHECK_PTR(ctx);
  hcurl_space *sp = hcurl_space_create(ctx, geom, 1, order, n_elem, elem_xyz, ldof, geom == HCURL_GEOM_TET ? corners : NULL, face_base,
                                       (int32_t)ndof, NULL, 0, 0, NULL, NULL, NULL, NULL, NULL, NULL);
  CHECK_PTR(sp);
  hcurl_op *op = hcurl_op_create(ctx, sp, 0);
  CHECK_PTR(op);
  double *alpha = malloc(sizeof(double) * n_elem), *beta = malloc(sizeof(double) * n_elem);
  for (int64_t e = 0; e < n_elem; e++) { alpha[e] = 0.05 / (attr[e] > 0 ? attr[e] : 1); beta[e] = 1.0; }
  CHECK(hcurl_op_set_coefficients(op, HCURL_COEF_PER_ELEM, alpha, beta, HCURL_HOST));
  CHECK(hcurl_op_set_essential_tdofs(op, ess, (int32_t)n_ess));
  /* ---- manufactured solve */
  double *xref = malloc(sizeof(double) * ndof), *b = malloc(sizeof(double) * ndof), *x = malloc(sizeof(double) * ndof);
  unsigned long long st = 12345;
  for (int64_t i = 0; i < ndof; i++) { st = st * 6364136223846793005ULL + 1442695040888963407ULL; xref[i] = (double)(st >> 11) / 9007199254740992.0 - 0.5; }
  CHECK(hcurl_op_apply(op, xref, b, HCURL_HOST));
  hcurl_pcg *pcg = hcurl_pcg_create(ctx, op, HCURL_PC_JACOBI);
  CHECK_PTR(pcg);
  CHECK(hcurl_pcg_set(pcg, 1e-10, 0.0, 5000, 0));
  int iters = 0, conv = 0;
  double rel = 0.0;
  CHECK(hcurl_pcg_solve(pcg, b, x, 1, HCURL_HOST, &iters, &rel, &conv));
  double num = 0.0, den = 0.0;
  for (int64_t i = 0; i < ndof; i++) { num += (x[i] - xref[i]) * (x[i] - xref[i]); den += xref[i] * xref[i]; }
  printf("Jacobi-PCG: %d iterations, relative residual %.3e, converged %d, solution error %.3e, loop %.2f ms\n", iters, rel, conv,
         sqrt(num / den), hcurl_pcg_last_loop_ms(pcg));
  hcurl_pcg_destroy(pcg);
  hcurl_op_destroy(op);
  hcurl_space_destroy(sp);
  hcurl_ctx_destroy(ctx);
  return conv && sqrt(num / den) < 1e-6 ? 0 : 1;
}
