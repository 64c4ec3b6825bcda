// Patch path of the affine hex operators (orders 1 and 2): host-side set-up and launch glue.
//
// Same operator as hex_apply.cu / hex_fast.cu (CurlCurlKernel + VectorFEMassKernel,
// /root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24, applied as
// EquationSystem::Mult, /root/reference/src/equation_systems/equation_system.cpp:200-204), organised
// around what ncu showed to bound element-wise kernels on sm_100a:
//
//  * L1tex lane rate of gathers / FP64 RED atomics  ->  one CTA works on PATCHES of FE adjacent elements
//    (patch.cpp); x is gathered once per unique patch unit into a shared-memory staging vector, the
//    element side pulls its inputs from it through a per-slot table, y is summed per patch unit from
//    a shared-memory result buffer; a unit is STORED by the patch that owns it and RED-added by the others
//    after the owner's store (per-patch flags), so nothing is cleared before an apply.
//  * memory latency  ->  the kernel is PERSISTENT and software-pipelined: the parts of a patch blob arrive by
//    TMA bulk copies (cp.async.bulk + mbarrier) one / two patches ahead, the x gather of the next patch runs as
//    cp.async copies behind the element operator of the current one.
//  * FP64 pipe (36 TFLOP/s measured; the quadrature-based order-2 apply needs 2916 DFMA / element)  ->
//    on AFFINE elements with per-element coefficients the element operator is applied in an
//    orthonormal Legendre (modal) basis (hex_patch_p2_elem.h, gen/hcurl_modal_p1.h).
// The pipeline program: patch_pipeline.h; the kernels and their launcher: hex_patch_kernels.cu.
// Measured state and the variants that were tried and rejected: profiles/README.md.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "internal.h"

namespace hcurl {

// ---- setup ---------------------------------------------------------------------------------------
// affine geometry factors in patch order: S = alpha J^T J/det (6), M = beta adj adj^T/det (6)
__global__ void patch_edata_kernel(const int64_t n_pos, const int FE, const int32_t *__restrict__ perm, const double *__restrict__ xyz,
                                   const double *__restrict__ alpha, const double *__restrict__ beta,
                                   double *__restrict__ edata) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pos) return;
  const int32_t e = perm[j];
  double *out = edata + (j / FE) * 12 * FE + j % FE;  // [patch][12][FE]
  if (e < 0) {
    for (int s = 0; s < 12; s++) out[s * FE] = 0.0;
    return;
  }
  const double *X = xyz + (int64_t)e * 24;  // tensor order
  double c0[3], c1[3], c2[3];
  for (int d = 0; d < 3; d++) { c0[d] = X[3 + d] - X[d]; c1[d] = X[6 + d] - X[d]; c2[d] = X[12 + d] - X[d]; }
  const double a0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
  const double a1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  const double a2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
  const double det = c0[0] * a0[0] + c0[1] * a0[1] + c0[2] * a0[2];
  const double al = alpha ? alpha[e] / det : 0.0, be = beta ? beta[e] / det : 0.0;
  const double *c[3] = {c0, c1, c2};
  const double *ad[3] = {a0, a1, a2};
  int s = 0;
  for (int i = 0; i < 3; i++)
    for (int jj = i; jj < 3; jj++, s++) {
      out[s * FE] = al * (c[i][0] * c[jj][0] + c[i][1] * c[jj][1] + c[i][2] * c[jj][2]);
      out[(6 + s) * FE] = be * (ad[i][0] * ad[jj][0] + ad[i][1] * ad[jj][1] + ad[i][2] * ad[jj][2]);
    }
}

// OR the essential flags of both doubles into the pair words of every patch blob (one CTA per patch)
__global__ void patch_flag_kernel(const uint32_t *__restrict__ blob_off, const double *__restrict__ flag,
                                  uint32_t *__restrict__ blob) {
  uint32_t *B = blob + (size_t)blob_off[blockIdx.x] * 4;
  const uint32_t U = B[0];
  for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
    uint32_t m = B[PatchHost::HDR + u];
    const double2 f = reinterpret_cast<const double2 *>(flag)[m & PATCH_IDX];
    if (f.x != 0.0) m |= PATCH_ESS0;
    if (f.y != 0.0) m |= PATCH_ESS1;
    B[PatchHost::HDR + u] = m;
    if (m & (PATCH_ESS0 | PATCH_ESS1)) B[20] = 1u;  // header: the patch has essential dofs (benign race: every writer stores 1)
  }
}

// single-dof units (order 1): one flag per unit
__global__ void patch_flag8_kernel(const uint32_t *__restrict__ blob_off, const double *__restrict__ flag, uint32_t *__restrict__ blob) {
  uint32_t *B = blob + (size_t)blob_off[blockIdx.x] * 4;
  const uint32_t U = B[0];
  for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
    const uint32_t m = B[PatchHost::HDR + u];
    if (flag[m & PATCH_IDX] != 0.0) { B[PatchHost::HDR + u] = m | PATCH_ESS0; B[20] = 1u; }
  }
}

// orthonormal shifted Legendre on [0,1]: L0 = 1, L1 = sqrt3 (2x-1), L2 = sqrt5 (6x^2-6x+1);
// T[m][j] = int_0^1 L_m l_j, by a 5-point Gauss rule (exact).  Host only; returns 1 when the tables do not have the
// symmetric structure the kernels hard-wire.
int compute_modal_tables(ModalTables &T) {
  double xq[5], wq[5], gl[2], wl[2], gll[3];
  gauss_legendre01(5, xq, wq);
  gauss_legendre01(2, gl, wl);
  gauss_lobatto01(3, gll);
  double Bo[5 * 2], Bc[5 * 3];
  lagrange_tables(2, gl, 5, xq, Bo, nullptr);
  lagrange_tables(3, gll, 5, xq, Bc, nullptr);
  double Tc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, To[4] = {0, 0, 0, 0};
  for (int q = 0; q < 5; q++) {
    const double t = 2.0 * xq[q] - 1.0;
    const double L[3] = {1.0, sqrt(3.0) * t, sqrt(5.0) * 0.5 * (3.0 * t * t - 1.0)};
    for (int m = 0; m < 3; m++) {
      for (int j = 0; j < 3; j++) Tc[m * 3 + j] += wq[q] * L[m] * Bc[q * 3 + j];
      if (m < 2) for (int i = 0; i < 2; i++) To[m * 2 + i] += wq[q] * L[m] * Bo[q * 2 + i];
    }
  }
  const double tol = 1e-14;
  if (fabs(To[0] - 0.5) > tol || fabs(To[1] - 0.5) > tol || fabs(To[2] + 0.5) > tol || fabs(To[3] - 0.5) > tol ||
      fabs(Tc[0] - Tc[2]) > tol || fabs(Tc[3] + Tc[5]) > tol || fabs(Tc[4]) > tol || fabs(Tc[6] - Tc[8]) > tol) {
    set_error("modal tables do not have the expected symmetric structure");
    return 1;
  }
  T.a0 = 0.5 * (Tc[0] + Tc[2]); T.a1 = Tc[1]; T.b = 0.5 * (Tc[5] - Tc[3]); T.c0 = 0.5 * (Tc[6] + Tc[8]); T.c1 = Tc[7];
  T.d1 = 2.0 * sqrt(3.0);
  T.d2 = 2.0 * sqrt(15.0);
  return 0;
}

}  // namespace hcurl
extern "C" int hcurl_host_modal_tables_p2(double *out7) {
  hcurl::ModalTables T;
  if (!out7 || hcurl::compute_modal_tables(T)) return 1;
  const double v[7] = {T.a0, T.a1, T.b, T.c0, T.c1, T.d1, T.d2};
  memcpy(out7, v, sizeof(v));
  return 0;
}
namespace hcurl {

// set-up uploads go through the context's stream (it is non-blocking: no implicit ordering with the legacy stream)
template <class T>
static int upload(hcurl_ctx *ctx, T *&dst, const std::vector<T> &src) {
  HC_CUDA(cudaMalloc(&dst, sizeof(T) * (src.size() ? src.size() : 1)));
  HC_CUDA(cudaMemcpyAsync(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static bool all_elements_affine(const hcurl_space *sp, const double *elem_xyz, std::vector<double> &cen) {
  const int64_t ne = sp->n_elem;
  cen.resize((size_t)ne * 3);
  bool affine = true;
  for (int64_t e = 0; e < ne; e++) {
    const double *X = elem_xyz + e * 24;  // MFEM corner order: 0 1 2 3 bottom face, 4 5 6 7 top
    double h2 = 0.0, dev2 = 0.0;
    for (int d = 0; d < 3; d++) {
      double c = 0.0;
      for (int v = 0; v < 8; v++) c += X[v * 3 + d];
      cen[(size_t)e * 3 + d] = 0.125 * c;
      const double ex = X[3 + d] - X[d], ey = X[9 + d] - X[d], ez = X[12 + d] - X[d];
      h2 = std::max(h2, std::max(ex * ex, std::max(ey * ey, ez * ez)));
      const double pred[8] = {X[d], X[d] + ex, X[d] + ex + ey, X[d] + ey, X[d] + ez, X[d] + ex + ez, X[d] + ex + ey + ez, X[d] + ey + ez};
      for (int v = 0; v < 8; v++) dev2 = std::max(dev2, (X[v * 3 + d] - pred[v]) * (X[v * 3 + d] - pred[v]));
    }
    if (dev2 > 1e-26 * h2) affine = false;  // same tolerance as fast_affine_kernel (hex_fast.cu)
  }
  return affine;
}

// The patch path serves AFFINE meshes (every element a parallelepiped: the modal element operators are exact there),
// orders 1 and 2.  Order 2: the dof side works on aligned PAIRS of doubles.  In the rotated slot layout the slots 2r,
// 2r+1 of an element are the two order-2 dofs along one line (open index 0, 1); MFEM numbers the dofs of an edge /
// face / interior consecutively, so they are the doubles 2q, 2q+1 of the extended vector -- in this order, or swapped
// when the element sees the line reversed -- each with its own sign.  Checked here for every element slot pair; a map
// that does not have this structure (or an odd-length vector) simply does not get the patch path.
int hex_patch_build(hcurl_space *sp, const double *elem_xyz, const int32_t *elem_ldof, const int32_t *l2e) {
  if (sp->geom != HCURL_GEOM_HEX || sp->geom_order != 1 || (sp->order != 1 && sp->order != 2)) return 0;
  const char *off = getenv("HCURL_NO_PATCH");
  if (off && off[0] == '1') return 0;
  const bool p2 = sp->order == 2;
  // Elements per patch and kernel.  Order 2: the warp-specialised kernel (patch_stream.h) with 128-element patches (one
  // stream per SM: four element, eight sum and four gather warps) when that fits into one CTA's shared memory, else
  // 32-element patches through the single-warp pipeline
  // (patch_pipeline.h; irregular meshes have longer unit lists).  Order 1: 128 (fastest of 64 / 128 / 256).
  // HCURL_PATCH_FE fixes the patch size, HCURL_PATCH_STREAM=0 the single-warp pipeline (measurements, tests).
  int FE = 128;
  bool fe_fixed = false, want_stream = p2;
  if (const char *s = getenv("HCURL_PATCH_STREAM")) want_stream = p2 && atoi(s) != 0;
  if (p2 && !want_stream) FE = 32;
  if (const char *s = getenv("HCURL_PATCH_FE")) { FE = atoi(s); fe_fixed = true; }
  if (p2 ? (FE != 32 && FE != 64 && !(FE == 128 && want_stream)) : (FE != 32 && FE != 64 && FE != 128 && FE != 256)) {
    set_error(p2 ? "HCURL_PATCH_FE must be 32, 64 or (stream kernel only) 128 for order 2" : "HCURL_PATCH_FE must be 32, 64, 128 or 256 for order 1");
    return 1;
  }
  const int64_t ne = sp->n_elem;
  if (p2 && (sp->n_ext() & 1)) return 0;
  std::vector<double> cen;
  if (!all_elements_affine(sp, elem_xyz, cen)) return 0;
  const int nu = p2 ? 27 : 12, unit = p2 ? 16 : 8;
  const int64_t n_units = p2 ? sp->n_ext() / 2 : sp->n_ext();
  std::vector<uint32_t> code((size_t)ne * nu);
  PatchSet &P = sp->patch;
  P = PatchSet();
  if (p2) {
    if (hex_p2_pair_codes(ne, elem_ldof, l2e, code.data())) return 0;  // not aligned pairs: no patch path for this space
    for (size_t i = 0; i < code.size() && !P.general_orient; i++) {
      const uint32_t f = code[i] >> 29;
      P.general_orient = f != 0u && f != 7u;
    }
    if (const char *s = getenv("HCURL_PATCH_GENERAL")) P.general_orient = P.general_orient || s[0] == '1';  // tests
  } else hex_slot_codes(1, ne, elem_ldof, l2e, code.data());
  // units that take part in the inter-rank exchange: ghosts, and owned units other ranks hold as ghosts
  std::vector<uint8_t> xch;
  if (sp->n_nbr > 0) {
    xch.assign((size_t)n_units, 0);
    const int per = p2 ? 2 : 1;
    for (int64_t u = sp->n_tdof / per; u < n_units; u++) xch[(size_t)u] = 1;
    for (int32_t d : sp->own_ext) xch[(size_t)(d / per)] = 1;
  }
  PatchHost H;
  const uint8_t *xchp = xch.empty() ? nullptr : xch.data();
  for (;;) {
    H = PatchHost();
    if (build_patches(nu, ne, FE, FE + 1, unit, n_units, cen.data(), code.data(), xchp, H)) {
      if (p2 && FE == 128 && !fe_fixed) { FE = 32; continue; }  // (e.g. the 16-bit slot offsets do not reach: smaller patches)
      return 1;
    }
    if (!(p2 && FE >= 64 && want_stream)) break;
    int optin = 0;
    HC_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, sp->ctx->device));
    const size_t need = (128 / FE) * patch_stream_smem_bytes(FE, H.max_d_words * 4, H.max_l_words * 4, H.max_t_words * 4, (H.max_dofs * 16u + 15u) & ~15u) + 2048;
    P.stream = sp->ctx->coop_ok && H.max_dofs <= 8191 && need <= (size_t)optin;
    if (P.stream || fe_fixed) break;
    FE = 32;  // the single-warp pipeline, fastest with 32-element patches (two streams of 64-element patches -- HCURL_PATCH_FE=64 -- measured slower than it in the three-role form)
  }
  if (p2 && FE == 128 && !P.stream) { set_error("HCURL_PATCH_FE=128 (order 2) needs the stream kernel, which does not fit this mesh's patches into shared memory"); return 1; }
  if (H.max_dofs > 8191) return 0;  // 13-bit staging positions
  P.FE = FE; P.unit = unit; P.stride = H.stride; P.n_patch = H.n_patch; P.n_boundary = H.n_boundary;
  P.blob_words = (int64_t)H.blob.size(); P.max_d_words = H.max_d_words; P.max_l_words = H.max_l_words; P.max_t_words = H.max_t_words;
  P.max_dofs = H.max_dofs;
  for (size_t i = 0; i < 27; i++) P.swz_row[i] = i < H.swz_row.size() ? H.swz_row[i] : 0;
  for (size_t i = 0; i < 16; i++) P.swz_grp[i] = i < H.swz_grp.size() ? H.swz_grp[i] : 0;
  if (patch_smem_bytes(unit, FE, H.max_d_words * 4, H.max_l_words * 4, H.max_t_words * 4, H.max_dofs * unit) > 227 * 1024) return 0;
  HC_CUDA(cudaSetDevice(sp->ctx->device));
  if (upload(sp->ctx, P.d_perm, H.perm) || upload(sp->ctx, P.d_blob_off, H.blob_off) || upload(sp->ctx, P.d_part_off, H.part_off) ||
      upload(sp->ctx, P.d_blob, H.blob))
    return 1;
  if (upload(sp->ctx, P.d_own_off, H.own_off) || upload(sp->ctx, P.d_own_idx, H.own_idx)) return 1;
  P.n_count = P.n_patch / 8 + 4;  // one arrival counter per batch of 8 pipeline iterations (a grid of one CTA at worst)
  HC_CUDA(cudaMalloc(&P.d_count, sizeof(uint32_t) * (size_t)P.n_count));
  HC_CUDA(cudaMemsetAsync(P.d_count, 0, sizeof(uint32_t) * (size_t)P.n_count, sp->ctx->stream));
  HC_CUDA(cudaStreamSynchronize(sp->ctx->stream));
  P.arrivals = 0;
  sp->has_patches = true;
  return 0;
}

void hex_patch_free(hcurl_space *sp) {
  PatchSet &P = sp->patch;
  cudaFree(P.d_perm); cudaFree(P.d_blob_off); cudaFree(P.d_blob); cudaFree(P.d_part_off); cudaFree(P.d_own_off); cudaFree(P.d_own_idx); cudaFree(P.d_count);
  P = PatchSet();
  sp->has_patches = false;
}

bool hex_patch_eligible(const hcurl_op *op) {
  const hcurl_space *sp = op->sp;
  // the modal operator is exact where the (p+1)-point rule is: that is the rule it stands in for
  // either term may be absent (statics: curl-curl only; mass solves: mass only): its geometry factors are zero
  if (!(sp->has_patches && op->single_rule && op->qc == sp->order + 1 && op->coef_layout == HCURL_COEF_PER_ELEM &&
        (op->has_alpha || op->has_beta)))
    return false;
  return !(op->kernel_sel == 1 || op->kernel_sel == 2);
}

int hex_patch_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale, bool accumulate, double *d_dot) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  PatchSet &P = sp->patch;
  const int64_t n_pos = P.n_patch * P.FE;
  if (!op->edataP_valid) {
    if (!op->d_edataP) HC_CUDA(cudaMalloc(&op->d_edataP, sizeof(double) * 12 * n_pos));
    patch_edata_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, ctx->stream>>>(n_pos, P.FE, P.d_perm, sp->d_xyz, op->has_alpha ? op->d_alpha : nullptr,
                                                                                op->has_beta ? op->d_beta : nullptr, op->d_edataP);
    HC_CUDA(cudaGetLastError());
    op->edataP_valid = true;
  }
  PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n_patch = P.n_patch; a.k_begin = 0; a.dot_add = 0; a.ess_identity = sp->n_nbr == 0 ? 1 : 0; a.blob_off = P.d_blob_off; a.part_off = P.d_part_off;
  a.blob = (constrained && op->d_blob_ess) ? op->d_blob_ess : P.d_blob;
  a.d_cap_bytes = P.max_d_words * 4; a.l_cap_bytes = P.max_l_words * 4; a.t_cap_bytes = P.max_t_words * 4;
  a.xs_cap_bytes = (P.max_dofs * (uint32_t)P.unit + 15u) & ~15u;
  a.edata = op->d_edataP; a.x = d_x; a.y = d_y; a.scale = scale;
  a.own_off = P.d_own_off; a.own_idx = P.d_own_idx; a.count = P.d_count;  // count_target: set by the launcher (grid size)
  memcpy(a.swz_row, P.swz_row, sizeof(a.swz_row)); memcpy(a.swz_grp, P.swz_grp, sizeof(a.swz_grp));
  a.dot_out = d_dot; a.dot_partial = ctx->d_partial; a.dot_counter = ctx->d_counter;
  if (P.unit == 16) {
    static const struct Cached { ModalTables T; int rc; Cached() { rc = compute_modal_tables(T); } } modal;  // host data: device-independent
    if (modal.rc) return 1;
    a.modal = modal.T;
  }
  if (hex_patch_kernel_launch(op, a, accumulate)) return 1;
  HC_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// essential flags on the patch unit lists (called from hcurl_op_set_essential_tdofs with the flag vector)
int hex_patch_flag_essential(hcurl_op *op, const double *d_flag) {
  hcurl_space *sp = op->sp;
  if (!sp->has_patches) return 0;
  PatchSet &P = sp->patch;
  if (!op->d_blob_ess) HC_CUDA(cudaMalloc(&op->d_blob_ess, sizeof(uint32_t) * P.blob_words));
  HC_CUDA(cudaMemcpyAsync(op->d_blob_ess, P.d_blob, sizeof(uint32_t) * P.blob_words, cudaMemcpyDeviceToDevice, op->ctx->stream));
  if (P.unit == 8) patch_flag8_kernel<<<(unsigned)P.n_patch, 128, 0, op->ctx->stream>>>(P.d_blob_off, d_flag, op->d_blob_ess);
  else patch_flag_kernel<<<(unsigned)P.n_patch, 128, 0, op->ctx->stream>>>(P.d_blob_off, d_flag, op->d_blob_ess);
  HC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hcurl
