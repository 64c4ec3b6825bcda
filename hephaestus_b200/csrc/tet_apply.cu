// Batched dense element-matrix apply for Nedelec tetrahedra.
//
// Replaces, for simplicial meshes (TEAM7 / TEAM4: ND_3D_P1, P2 on TETRA4,
// /root/reference/examples/team7/Main.cpp:154-163), the assembled CSR operator
// of CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24,
//  /root/reference/src/kernels/vector_fe_mass_kernel.cpp:21-24,
//  /root/reference/src/equation_systems/equation_system.cpp:200-204, :345).
//
// Setup kernel: per element packed-symmetric matrix
//   A_e = sum_{a<=b} [ beta det (J^-1 J^-T)_ab Mref^{ab} + alpha/det (J^T J)_ab Kref^{ab} ]
// (J constant on straight tets; Mref/Kref integrated exactly on the reference
// tet), with MFEM's triangle-face DofTransformation folded in (R^T A R).
// Apply kernel: one thread per element, matrix streamed entry-major
// ([npk][n_elem_pad], coalesced), x_e / y_e in registers, signed atomic scatter.
// HBM-bound: 8 nd(nd+1)/2 + 4 nd bytes per element.
#include "internal.h"

namespace hcurl {

__device__ __forceinline__ double tet_gather(const double *__restrict__ x, uint32_t m) {
  if (m & MAP_ESS) return 0.0;
  double v = __ldg(x + (m & MAP_IDX));
  return (m & MAP_NEG) ? -v : v;
}

template <int ND>
__global__ void __launch_bounds__(128)
tet_apply_kernel(const int64_t n_elem, const int64_t n_pad, const uint32_t *__restrict__ map,
                 const double *__restrict__ emat, const double *__restrict__ x, double *__restrict__ y,
                 const double scale) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double xl[ND], yl[ND];
  uint32_t mm[ND];
#pragma unroll
  for (int i = 0; i < ND; i++) {
    mm[i] = __ldg(map + i * n_pad + e);
    xl[i] = tet_gather(x, mm[i]);
    yl[i] = 0.0;
  }
  int k = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) {
#pragma unroll
    for (int j = i; j < ND; j++, k++) {
      const double a = __ldg(emat + (int64_t)k * n_pad + e);
      yl[i] += a * xl[j];
      if (j != i) yl[j] += a * xl[i];
    }
  }
#pragma unroll
  for (int i = 0; i < ND; i++) {
    if (mm[i] & MAP_ESS) continue;
    const double v = scale * yl[i];
    atomicAdd(y + (mm[i] & MAP_IDX), (mm[i] & MAP_NEG) ? -v : v);
  }
}

// Complex (block) apply [[Ar, -Ai], [Ai, Ar]] in ONE pass: xr_e, xi_e gathered once, the two packed element matrices
// streamed once, yr_e, yi_e scattered once (ComplexMaxwellOperator::Solve builds this block operator,
// /root/reference/src/formulations/ComplexMaxwell/complex_maxwell_formulation.cpp:181-217).  Against four real applies
// this halves the matrix traffic that bounds the tet kernel and reads x / writes y once instead of twice.
template <int ND>
__global__ void __launch_bounds__(128)
tet_apply_complex_kernel(const int64_t n_elem, const int64_t n_pad, const uint32_t *__restrict__ map,
                         const double *__restrict__ emat_re, const double *__restrict__ emat_im,
                         const double *__restrict__ xr, const double *__restrict__ xi, double *__restrict__ yr, double *__restrict__ yi) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double ar_[ND], ai_[ND], br_[ND], bi_[ND];  // x_e (re, im), y_e (re, im)
  uint32_t mm[ND];
#pragma unroll
  for (int i = 0; i < ND; i++) {
    mm[i] = __ldg(map + i * n_pad + e);
    ar_[i] = tet_gather(xr, mm[i]);
    ai_[i] = tet_gather(xi, mm[i]);
    br_[i] = 0.0;
    bi_[i] = 0.0;
  }
  int k = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) {
#pragma unroll
    for (int j = i; j < ND; j++, k++) {
      const double mr = __ldg(emat_re + (int64_t)k * n_pad + e), mi = __ldg(emat_im + (int64_t)k * n_pad + e);
      br_[i] += mr * ar_[j] - mi * ai_[j];
      bi_[i] += mi * ar_[j] + mr * ai_[j];
      if (j != i) {
        br_[j] += mr * ar_[i] - mi * ai_[i];
        bi_[j] += mi * ar_[i] + mr * ai_[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < ND; i++) {
    if (mm[i] & MAP_ESS) continue;
    const bool neg = mm[i] & MAP_NEG;
    atomicAdd(yr + (mm[i] & MAP_IDX), neg ? -br_[i] : br_[i]);
    atomicAdd(yi + (mm[i] & MAP_IDX), neg ? -bi_[i] : bi_[i]);
  }
}

__global__ void tet_apply_complex_kernel_generic(const int nd, const int64_t n_elem, const int64_t n_pad, const uint32_t *__restrict__ map,
                                                 const double *__restrict__ emat_re, const double *__restrict__ emat_im,
                                                 const double *__restrict__ xr, const double *__restrict__ xi,
                                                 double *__restrict__ yr, double *__restrict__ yi) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double ar_[45], ai_[45], br_[45], bi_[45];
  for (int i = 0; i < nd; i++) {
    const uint32_t m = __ldg(map + i * n_pad + e);
    ar_[i] = tet_gather(xr, m); ai_[i] = tet_gather(xi, m); br_[i] = 0.0; bi_[i] = 0.0;
  }
  int k = 0;
  for (int i = 0; i < nd; i++)
    for (int j = i; j < nd; j++, k++) {
      const double mr = __ldg(emat_re + (int64_t)k * n_pad + e), mi = __ldg(emat_im + (int64_t)k * n_pad + e);
      br_[i] += mr * ar_[j] - mi * ai_[j];
      bi_[i] += mi * ar_[j] + mr * ai_[j];
      if (j != i) { br_[j] += mr * ar_[i] - mi * ai_[i]; bi_[j] += mi * ar_[i] + mr * ai_[i]; }
    }
  for (int i = 0; i < nd; i++) {
    const uint32_t m = __ldg(map + i * n_pad + e);
    if (m & MAP_ESS) continue;
    const bool neg = m & MAP_NEG;
    atomicAdd(yr + (m & MAP_IDX), neg ? -br_[i] : br_[i]);
    atomicAdd(yi + (m & MAP_IDX), neg ? -bi_[i] : bi_[i]);
  }
}

// generic (any nd) fallback used for nd = 45: x_e, y_e in local memory
__global__ void tet_apply_kernel_generic(const int nd, const int64_t n_elem, const int64_t n_pad,
                                         const uint32_t *__restrict__ map, const double *__restrict__ emat,
                                         const double *__restrict__ x, double *__restrict__ y, const double scale) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  double xl[45], yl[45];
  for (int i = 0; i < nd; i++) { xl[i] = tet_gather(x, __ldg(map + i * n_pad + e)); yl[i] = 0.0; }
  int k = 0;
  for (int i = 0; i < nd; i++)
    for (int j = i; j < nd; j++, k++) {
      const double a = __ldg(emat + (int64_t)k * n_pad + e);
      yl[i] += a * xl[j];
      if (j != i) yl[j] += a * xl[i];
    }
  for (int i = 0; i < nd; i++) {
    const uint32_t m = __ldg(map + i * n_pad + e);
    if (m & MAP_ESS) continue;
    const double v = scale * yl[i];
    atomicAdd(y + (m & MAP_IDX), (m & MAP_NEG) ? -v : v);
  }
}

// ---- setup: element matrices -----------------------------------------------------
__device__ __forceinline__ int packed_index(int nd, int i, int j) {  // i <= j
  return i * nd - i * (i - 1) / 2 + (j - i);
}

// one thread per (packed entry, element); ref = [Mref(6*npk) | Kref(6*npk)]
// nqc / nqm > 0: coefficients per quadrature point of the caller's rules (alpha [n_elem][nqc], beta [n_elem][nqm]) with
// the per-point reference matrices refq_k / refq_m [nq][6][npk] instead of the exactly integrated ones
__global__ void tet_emat_kernel(const int p, const int nd, const int npk, const int nf, const int64_t n_elem,
                                const int64_t n_pad, const double *__restrict__ xyz,
                                const double *__restrict__ ref, const double *__restrict__ faceR,
                                const double *__restrict__ alpha, const double *__restrict__ beta,
                                double *__restrict__ emat, const int nqc, const int nqm,
                                const double *__restrict__ refq_k, const double *__restrict__ refq_m) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (e >= n_elem) return;
  // decode (i, j) from packed k
  int i = 0, rem = k;
  while (rem >= nd - i) { rem -= nd - i; i++; }
  const int j = i + rem;
  const double *X = xyz + e * 12;
  double J[3][3];
  for (int d = 0; d < 3; d++)
    for (int c = 0; c < 3; c++) J[d][c] = X[3 * (c + 1) + d] - X[d];
  const double a0[3] = {J[1][1] * J[2][2] - J[1][2] * J[2][1], J[0][2] * J[2][1] - J[0][1] * J[2][2], J[0][1] * J[1][2] - J[0][2] * J[1][1]};
  const double a1[3] = {J[1][2] * J[2][0] - J[1][0] * J[2][2], J[0][0] * J[2][2] - J[0][2] * J[2][0], J[0][2] * J[1][0] - J[0][0] * J[1][2]};
  const double a2[3] = {J[1][0] * J[2][1] - J[1][1] * J[2][0], J[0][1] * J[2][0] - J[0][0] * J[2][1], J[0][0] * J[1][1] - J[0][1] * J[1][0]};
  // rows of J^-1 = a_r / det
  const double det = J[0][0] * a0[0] + J[1][0] * a0[1] + J[2][0] * a0[2];
  const double *ar[3] = {a0, a1, a2};
  double gm[6], gk[6];
  const bool perq = nqc > 0 || nqm > 0;
  const double al = perq ? 1.0 : (alpha ? alpha[e] : 0.0), be = perq ? 1.0 : (beta ? beta[e] : 0.0);
  int s = 0;
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++, s++) {
      // (J^-1 J^-T)_ab = row_a . row_b / det^2 ; (J^T J)_ab = col_a . col_b
      gm[s] = be / det * (ar[a][0] * ar[b][0] + ar[a][1] * ar[b][1] + ar[a][2] * ar[b][2]);
      gk[s] = al / det * (J[0][a] * J[0][b] + J[1][a] * J[1][b] + J[2][a] * J[2][b]);
    }
  auto raw = [&](int r, int c) {
    if (r > c) { int t = r; r = c; c = t; }
    const int kk = packed_index(nd, r, c);
    double v = 0.0;
    if (!perq) {
      for (int q = 0; q < 6; q++) v += gm[q] * ref[q * npk + kk] + gk[q] * ref[(6 + q) * npk + kk];
      return v;
    }
    if (beta)
      for (int q = 0; q < nqm; q++) {
        double t = 0.0;
        for (int s2 = 0; s2 < 6; s2++) t += gm[s2] * refq_m[((int64_t)q * 6 + s2) * npk + kk];
        v += beta[e * nqm + q] * t;
      }
    if (alpha)
      for (int q = 0; q < nqc; q++) {
        double t = 0.0;
        for (int s2 = 0; s2 < 6; s2++) t += gk[s2] * refq_k[((int64_t)q * 6 + s2) * npk + kk];
        v += alpha[e * nqc + q] * t;
      }
    return v;
  };
  const int fbeg = 6 * p, fend = 6 * p + 4 * nf;
  const bool fi = nf > 0 && i >= fbeg && i < fend, fj = nf > 0 && j >= fbeg && j < fend;
  double v;
  if (!faceR || (!fi && !fj)) {
    v = raw(i, j);
  } else {
    // A'_ij = sum_{k in blk(i)} sum_{l in blk(j)} R_ki R_lj A_kl
    const int bi = fi ? (i - fbeg) / nf : -1, bj = fj ? (j - fbeg) / nf : -1;
    const double *Ri = fi ? faceR + ((e * 4 + bi) * nf * nf) : nullptr;
    const double *Rj = fj ? faceR + ((e * 4 + bj) * nf * nf) : nullptr;
    const int i0 = fi ? fbeg + bi * nf : i, j0 = fj ? fbeg + bj * nf : j;
    const int ni = fi ? nf : 1, nj = fj ? nf : 1;
    v = 0.0;
    for (int kk = 0; kk < ni; kk++) {
      const double rk = fi ? Ri[kk * nf + (i - i0)] : 1.0;
      if (rk == 0.0) continue;
      for (int ll = 0; ll < nj; ll++) {
        const double rl = fj ? Rj[ll * nf + (j - j0)] : 1.0;
        if (rl == 0.0) continue;
        v += rk * rl * raw(i0 + kk, j0 + ll);
      }
    }
  }
  emat[(int64_t)k * n_pad + e] = v;
}

__global__ void tet_diag_kernel(const int nd, const int64_t n_elem, const int64_t n_pad,
                                const uint32_t *__restrict__ map, const double *__restrict__ emat,
                                double *__restrict__ diag) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  for (int i = 0; i < nd; i++) {
    const uint32_t m = map[i * n_pad + e];
    if (m & MAP_ESS) continue;
    atomicAdd(diag + (m & MAP_IDX), emat[(int64_t)packed_index(nd, i, i) * n_pad + e]);
  }
}

static double *g_ref_dev[4] = {nullptr, nullptr, nullptr, nullptr};  // per order, per process (single device per rank)

int tet_build_emat(hcurl_op *op) {
  hcurl_space *sp = op->sp;
  const int p = sp->order, nd = sp->nd, npk = nd * (nd + 1) / 2;
  if (!g_ref_dev[p]) {
    TetRef r;
    build_tet_ref(p, r);
    std::vector<double> h(12 * npk);
    for (int i = 0; i < 6 * npk; i++) { h[i] = r.Mref[i]; h[6 * npk + i] = r.Kref[i]; }
    HC_CUDA(cudaMalloc(&g_ref_dev[p], sizeof(double) * h.size()));
    HC_CUDA(h2d_sync(op->ctx, g_ref_dev[p], h.data(), sizeof(double) * h.size()));
  }
  if (!op->d_emat) HC_CUDA(cudaMalloc(&op->d_emat, sizeof(double) * (size_t)npk * sp->n_elem_pad));
  dim3 grid((unsigned)((sp->n_elem + 127) / 128), npk);
  tet_emat_kernel<<<grid, 128, 0, op->ctx->stream>>>(p, nd, npk, sp->nf, sp->n_elem, sp->n_elem_pad, sp->d_xyz, g_ref_dev[p],
                                                    sp->d_faceR, op->has_alpha ? op->d_alpha : nullptr,
                                                    op->has_beta ? op->d_beta : nullptr, op->d_emat,
                                                    op->coef_layout == HCURL_COEF_PER_QPT ? op->tet_nqc : 0,
                                                    op->coef_layout == HCURL_COEF_PER_QPT ? op->tet_nqm : 0, op->d_refq_k, op->d_refq_m);
  HC_CUDA(cudaGetLastError());
  op->emat_valid = true;
  return 0;
}

// the caller's quadrature rules for PER_QPT coefficients (hcurl_op_set_tet_rules)
int tet_set_rules(hcurl_op *op, int nqc, const double *pc, const double *wc, int nqm, const double *pm, const double *wm) {
  const int p = op->sp->order;
  std::vector<double> h;
  cudaFree(op->d_refq_k); cudaFree(op->d_refq_m);
  op->d_refq_k = op->d_refq_m = nullptr;
  build_tet_ref_q(p, true, nqc, pc, wc, h);
  HC_CUDA(cudaMalloc(&op->d_refq_k, sizeof(double) * (h.size() ? h.size() : 1)));
  HC_CUDA(h2d_sync(op->ctx, op->d_refq_k, h.data(), sizeof(double) * h.size()));
  build_tet_ref_q(p, false, nqm, pm, wm, h);
  HC_CUDA(cudaMalloc(&op->d_refq_m, sizeof(double) * (h.size() ? h.size() : 1)));
  HC_CUDA(h2d_sync(op->ctx, op->d_refq_m, h.data(), sizeof(double) * h.size()));
  op->tet_nqc = nqc; op->tet_nqm = nqm;
  op->emat_valid = false;
  return 0;
}

int tet_apply_launch(hcurl_op *op, const uint32_t *d_map, const double *d_x, double *d_y, double scale) {
  hcurl_space *sp = op->sp;
  if (!op->emat_valid && tet_build_emat(op)) return 1;
  const unsigned grid = (unsigned)((sp->n_elem + 127) / 128);
  cudaStream_t st = op->ctx->stream;
  if (sp->nd == 6)
    tet_apply_kernel<6><<<grid, 128, 0, st>>>(sp->n_elem, sp->n_elem_pad, d_map, op->d_emat, d_x, d_y, scale);
  else if (sp->nd == 20)
    tet_apply_kernel<20><<<grid, 128, 0, st>>>(sp->n_elem, sp->n_elem_pad, d_map, op->d_emat, d_x, d_y, scale);
  else
    tet_apply_kernel_generic<<<grid, 128, 0, st>>>(sp->nd, sp->n_elem, sp->n_elem_pad, d_map, op->d_emat, d_x, d_y, scale);
  HC_CUDA(cudaGetLastError());
  op->ctx->launches++;
  return 0;
}

// yr += Ar xr - Ai xi, yi += Ai xr + Ar xi over the element loop (essential dofs of the map: no input, no output)
int tet_apply_complex_launch(hcurl_op *op_re, hcurl_op *op_im, const uint32_t *d_map, const double *d_xr, const double *d_xi,
                             double *d_yr, double *d_yi) {
  hcurl_space *sp = op_re->sp;
  if (!op_re->emat_valid && tet_build_emat(op_re)) return 1;
  if (!op_im->emat_valid && tet_build_emat(op_im)) return 1;
  const unsigned grid = (unsigned)((sp->n_elem + 127) / 128);
  cudaStream_t st = op_re->ctx->stream;
  if (sp->nd == 6)
    tet_apply_complex_kernel<6><<<grid, 128, 0, st>>>(sp->n_elem, sp->n_elem_pad, d_map, op_re->d_emat, op_im->d_emat, d_xr, d_xi, d_yr, d_yi);
  else if (sp->nd == 20)
    tet_apply_complex_kernel<20><<<grid, 128, 0, st>>>(sp->n_elem, sp->n_elem_pad, d_map, op_re->d_emat, op_im->d_emat, d_xr, d_xi, d_yr, d_yi);
  else
    tet_apply_complex_kernel_generic<<<grid, 128, 0, st>>>(sp->nd, sp->n_elem, sp->n_elem_pad, d_map, op_re->d_emat, op_im->d_emat, d_xr, d_xi, d_yr, d_yi);
  HC_CUDA(cudaGetLastError());
  op_re->ctx->launches++;
  return 0;
}

int tet_diag_launch(hcurl_op *op, double *d_diag) {
  hcurl_space *sp = op->sp;
  if (!op->emat_valid && tet_build_emat(op)) return 1;
  const uint32_t *map = op->d_map_ess ? op->d_map_ess : sp->d_map;
  tet_diag_kernel<<<(unsigned)((sp->n_elem + 127) / 128), 128, 0, op->ctx->stream>>>(sp->nd, sp->n_elem, sp->n_elem_pad, map,
                                                                                    op->d_emat, d_diag);
  HC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hcurl
