// Order-2 Nedelec hex patch apply, ONE THREAD PER ELEMENT -- an environment-parametrised per-thread program like
// hex_patch_p1_body.h / hex_patch_p2_body.h (same source as CUDA kernel and on host threads, tests/emu/).
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204), affine order-2 hexes, per-element coefficients.
//
// Why: ncu on the three-threads-per-element kernel (profiles/r1j_hex_patch_fe32_ncu.txt, r2a sweep) shows the L1tex data
// pipe as the loaded unit (73 %), 58 % of all cycles being shared-memory wavefronts, and half of those the exchange of
// modal coefficients between the three component threads of an element (27 STS.128 + 54 LDS.128 per element); giving
// an element more threads (hex_patch_p2h: 0.67 ms against 0.447) makes it worse.  Here one thread owns all three
// components of its element, so
//   * the exchange buffer and its barrier disappear (shared memory per element: pull 27 + result 27 + sum 27 16-byte
//     accesses instead of 162), a 32-element patch is one warp, and CTA barriers are warp-cheap;
//   * the modal curl R = curl_h u (36 coefficients) and W = K R are computed once per element instead of once per
//     component thread (about 970 instead of 1190 FP64 instructions per element);
//   * the element operator runs in place on 90 live doubles (A -> M A mode by mode, R -> W mode by mode), the fused
//     <x, A x> is A.(M A) + R.W, so nothing else has to stay alive: fits the 255-register budget without spills.
// Per patch: pull + forward transforms -> barrier (Xs free) -> issue the x gather of the next patch, operator, results
// to EY -> barrier -> per-pair sums, stores / REDs -> barrier.  Blob format, staging vector, result-buffer swizzle and
// the TMA / cp.async pipeline are those of hex_patch_p2_body.h.
//
// Shared memory:  [16 B: two mbarriers] [2 x blob_cap] [Xs: xs_cap] [EY: 27 x (FE+1) double2]
#pragma once
#include "hex_patch_p2_body.h"

namespace hcurl {

// component c keeps its modal coefficients as IA(n_c, n_{c+1}, n_{c+2}) of the global mode (n0, n1, n2)
HCURL_BODY constexpr int p2s_ia(const int c, const int n0, const int n1, const int n2) {
  return c == 0 ? P2_IA(n0, n1, n2) : (c == 1 ? P2_IA(n1, n2, n0) : P2_IA(n2, n0, n1));
}
// curl component k: mode shape [3 along k][2][2], index n_k + 3 (n_{k+1} + 2 n_{k+2})
HCURL_BODY constexpr int p2s_ir(const int k, const int n0, const int n1, const int n2) {
  return k == 0 ? n0 + 3 * (n1 + 2 * n2) : (k == 1 ? n1 + 3 * (n2 + 2 * n0) : n2 + 3 * (n0 + 2 * n1));
}

// In-place modal element operator of one element.  A[c][18]: modal coefficients of the three components on entry,
// modal test values Yh on exit.  Gk / Gm: scaled geometry factors, packed symmetric (00, 01, 02, 11, 12, 22).
// Returns x_e^T A_e x_e when DOT.
template <bool DOT>
HCURL_BODY double p2s_modal_op(const ModalTables &M, double (&A)[3][18], const double *Gk, const double *Gm) {
  const double d1 = M.d1, d2 = M.d2;
#define P2S_DD(n) ((n) == 1 ? d1 : d2)
#define P2S_SYM(i, j) ((i) <= (j) ? (i) * 3 - (i) * ((i)-1) / 2 + ((j) - (i)) : (j) * 3 - (j) * ((j)-1) / 2 + ((i) - (j)))
  double R[3][12];
  // R_k(a, b, c) = d_{k+1} u_{k+2} - d_{k+2} u_{k+1}, (a, b, c) along the axes (k, k+1, k+2)
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
#pragma unroll
    for (int cc = 0; cc < 2; cc++)
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int aa = 0; aa < 3; aa++)
          R[k][aa + 3 * (b + 2 * cc)] = fma(P2S_DD(b + 1), A[k2][P2_IA(cc, aa, b + 1)], -P2S_DD(cc + 1) * A[k1][P2_IA(b, cc + 1, aa)]);
  }
  double dk0 = 0.0, dk1 = 0.0;
  // W = K R mode by mode, in place: the eight modes with all indices < 2 carry all three curl components ...
#pragma unroll
  for (int n2 = 0; n2 < 2; n2++)
#pragma unroll
    for (int n1 = 0; n1 < 2; n1++)
#pragma unroll
      for (int n0 = 0; n0 < 2; n0++) {
        const double r0 = R[0][p2s_ir(0, n0, n1, n2)], r1 = R[1][p2s_ir(1, n0, n1, n2)], r2 = R[2][p2s_ir(2, n0, n1, n2)];
        const double w0 = fma(Gk[P2S_SYM(0, 2)], r2, fma(Gk[P2S_SYM(0, 1)], r1, Gk[P2S_SYM(0, 0)] * r0));
        const double w1 = fma(Gk[P2S_SYM(1, 2)], r2, fma(Gk[P2S_SYM(1, 1)], r1, Gk[P2S_SYM(0, 1)] * r0));
        const double w2 = fma(Gk[P2S_SYM(2, 2)], r2, fma(Gk[P2S_SYM(1, 2)], r1, Gk[P2S_SYM(0, 2)] * r0));
        if (DOT) { dk0 = fma(r0, w0, dk0); dk1 = fma(r1, w1, dk1); dk0 = fma(r2, w2, dk0); }
        R[0][p2s_ir(0, n0, n1, n2)] = w0;
        R[1][p2s_ir(1, n0, n1, n2)] = w1;
        R[2][p2s_ir(2, n0, n1, n2)] = w2;
      }
  // ... the modes with index 2 along axis k only component k
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int cc = 0; cc < 2; cc++)
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const double r = R[k][2 + 3 * (b + 2 * cc)], w = Gk[P2S_SYM(k, k)] * r;
        if (DOT) dk1 = fma(r, w, dk1);
        R[k][2 + 3 * (b + 2 * cc)] = w;
      }
  // A <- M A mode by mode, in place: component c is present in the global mode n when n_c < 2
#pragma unroll
  for (int n2 = 0; n2 < 3; n2++)
#pragma unroll
    for (int n1 = 0; n1 < 3; n1++)
#pragma unroll
      for (int n0 = 0; n0 < 3; n0++) {
        const bool h0 = n0 < 2, h1 = n1 < 2, h2 = n2 < 2;
        if (!(h0 || h1 || h2)) continue;
        const double a0 = h0 ? A[0][p2s_ia(0, n0, n1, n2)] : 0.0, a1 = h1 ? A[1][p2s_ia(1, n0, n1, n2)] : 0.0,
                     a2 = h2 ? A[2][p2s_ia(2, n0, n1, n2)] : 0.0;
        if (h0) {
          double v = Gm[P2S_SYM(0, 0)] * a0;
          if (h1) v = fma(Gm[P2S_SYM(0, 1)], a1, v);
          if (h2) v = fma(Gm[P2S_SYM(0, 2)], a2, v);
          if (DOT) dk0 = fma(a0, v, dk0);
          A[0][p2s_ia(0, n0, n1, n2)] = v;
        }
        if (h1) {
          double v = Gm[P2S_SYM(1, 1)] * a1;
          if (h0) v = fma(Gm[P2S_SYM(0, 1)], a0, v);
          if (h2) v = fma(Gm[P2S_SYM(1, 2)], a2, v);
          if (DOT) dk1 = fma(a1, v, dk1);
          A[1][p2s_ia(1, n0, n1, n2)] = v;
        }
        if (h2) {
          double v = Gm[P2S_SYM(2, 2)] * a2;
          if (h0) v = fma(Gm[P2S_SYM(0, 2)], a0, v);
          if (h1) v = fma(Gm[P2S_SYM(1, 2)], a1, v);
          if (DOT) dk0 = fma(a2, v, dk0);
          A[2][p2s_ia(2, n0, n1, n2)] = v;
        }
      }
  // Yh_c(m0, m1, m2) = (M A)_c + d_{c+2}^T W_{c+1} - d_{c+1}^T W_{c+2}   (frame (c, c+1, c+2))
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const int c1 = (c + 1) % 3, c2 = (c + 2) % 3;
#pragma unroll
    for (int m2 = 0; m2 < 3; m2++)
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 2; m0++) {
          double v = A[c][P2_IA(m0, m1, m2)];
          if (m2 >= 1) v = fma(P2S_DD(m2), R[c1][m1 + 3 * ((m2 - 1) + 2 * m0)], v);
          if (m1 >= 1) v = fma(-P2S_DD(m1), R[c2][m2 + 3 * (m0 + 2 * (m1 - 1))], v);
          A[c][P2_IA(m0, m1, m2)] = v;
        }
  }
#undef P2S_DD
#undef P2S_SYM
  return dk0 + dk1;
}

// the transforms of hex_patch_p2_body.h with the open-direction butterfly left to the caller: a[2r] = x0 + x1 (mode L0
// along the line), a[2r+1] = x1 - x0 (mode L1).  A line seen REVERSED by the element (MFEM: order swapped and both signs
// flipped, local = (-g1, -g0)) only changes the sign of a[2r] -- on the way in and, transposed, on the way out.
HCURL_BODY void p2s_fwd_sd(const ModalTables &M, const double *a, double *A) {
  double b[18];
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      p2_tc_fwd(M, a[P2_IA(m0, 0, i2)], a[P2_IA(m0, 1, i2)], a[P2_IA(m0, 2, i2)], b[P2_IA(m0, 0, i2)], b[P2_IA(m0, 1, i2)], b[P2_IA(m0, 2, i2)]);
#pragma unroll
  for (int r = 0; r < 6; r++) p2_tc_fwd(M, b[r], b[r + 6], b[r + 12], A[r], A[r + 6], A[r + 12]);
}
HCURL_BODY void p2s_bwd_sd(const ModalTables &M, const double *Yh, double *b) {
  double a[18];
#pragma unroll
  for (int r = 0; r < 6; r++) p2_tc_bwd(M, Yh[r], Yh[r + 6], Yh[r + 12], a[r], a[r + 6], a[r + 12]);
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      p2_tc_bwd(M, a[P2_IA(m0, 0, i2)], a[P2_IA(m0, 1, i2)], a[P2_IA(m0, 2, i2)], b[P2_IA(m0, 0, i2)], b[P2_IA(m0, 1, i2)], b[P2_IA(m0, 2, i2)]);
}

// GENERAL = false: every flagged element slot pair of the space is a reversed line (all three orientation bits set; the
// only case MFEM's maps produce, checked on the host when the patches are built) -- two instructions per pair instead of
// a dozen; GENERAL = true: independent swap / sign bits.
template <class Env, int FE, bool ACC, bool DOT, bool GENERAL>
HCURL_BODY void hex_patch_p2s_body(Env &env, const PatchArgs &a) {
  constexpr int NT = FE, ST = FE + 1, HDR = PatchHost::HDR;
  const ModalTables &M = env.template tables<ModalTables>();
  unsigned char *smem = env.smem();
  unsigned char *blobB = smem + 16;
  double2 *Xs = reinterpret_cast<double2 *>(blobB + 2 * (size_t)a.blob_cap_bytes);
  double2 *EY = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
  double2 *y2 = reinterpret_cast<double2 *>(a.y);
  const int tid = env.tid(), el = tid;
  double dacc = 0.0;
  const int64_t k0 = a.k_begin + env.block(), kstep = env.nblocks();
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);

  auto request = [&](int b, uint32_t o0, uint32_t o1) {
    env.bulk_request(b, blobB + (size_t)b * a.blob_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u);
  };
  // launch the x gather of the patch whose blob sits in buffer b
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    const uint32_t U = B[0];
    const uint32_t *dof = B + HDR;
#pragma unroll 4
    for (uint32_t u = tid; u < U; u += NT) {
      const uint32_t m = dof[u];
      if (!(m & (PATCH_ESS0 | PATCH_ESS1))) env.gather16(Xs + u, x2 + (m & PATCH_IDX));
      else {  // essential dofs enter as zeros (rare: synchronous)
        double2 v = make_double2(0.0, 0.0);
        if ((m & (PATCH_ESS0 | PATCH_ESS1)) != (PATCH_ESS0 | PATCH_ESS1)) {
          v = env.ldg(x2 + (m & PATCH_IDX));
          if (m & PATCH_ESS0) v.x = 0.0;
          if (m & PATCH_ESS1) v.y = 0.0;
        }
        Xs[u] = v;
      }
    }
    env.gather_commit();
  };
  if (tid == 0) {
    env.bulk_init(2);
    request(0, env.ldg(a.blob_off + k0), env.ldg(a.blob_off + k0 + 1));
    if (n_it > 1) request(1, env.ldg(a.blob_off + k0 + kstep), env.ldg(a.blob_off + k0 + kstep + 1));
  }
  env.sync();  // mbarriers initialised
  env.bulk_wait(0, 0);
  gather(0);
  env.gather_wait();
  env.sync();  // Xs of the first patch complete and visible

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    uint32_t q0 = 0, q1 = 0;  // blob extent of the patch requested when this one retires
    if (tid == 0 && it + 2 < n_it) { q0 = env.ldg(a.blob_off + k + 2 * kstep); q1 = env.ldg(a.blob_off + k + 2 * kstep + 1); }
    // affine geometry factors of my element (coalesced: [12][FE] per patch): issued now, consumed after the transforms
    const double *g = a.edata + (size_t)k * 12 * FE + el;
    double G[12];
#pragma unroll
    for (int s = 0; s < 12; s++) G[s] = env.ldg(g + s * FE);
    const bool ok = (uint32_t)el < B[13];
    const uint16_t *lidx = reinterpret_cast<const uint16_t *>(B + B[12]) + el;
    // ---- pull the 27 input pairs through the per-slot table, forward transforms
    double A[3][18];
    if (ok) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double X[18];
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const uint32_t l = lidx[(c * 9 + r) * FE];
          const double2 v = Xs[l & 0x1fffu];
          if (GENERAL) {
            double x0, x1;
            p2_pair_to_local(v, l, x0, x1);
            X[2 * r] = x0 + x1;
            X[2 * r + 1] = x1 - x0;
          } else {
            X[2 * r] = p2_flip(v.x + v.y, (l << 16) & 0x80000000u);
            X[2 * r + 1] = v.y - v.x;
          }
        }
        p2s_fwd_sd(M, X, A[c]);
      }
    }
    env.sync();  // Xs free
    // ---- start the x gather of the next patch: it lands behind the element operator (the long phase of this kernel)
    if (it + 1 < n_it) {
      env.bulk_wait(b ^ 1, ((it + 1) >> 1) & 1);
      gather(b ^ 1);
    }
    if (ok) {
      const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
#pragma unroll
      for (int s = 0; s < 12; s++) G[s] *= sc;
      const double d = p2s_modal_op<DOT>(M, A, G, G + 6);
      if (DOT) dacc += d;
    }
    if (ok) {
      const int elx = el ^ a.swz_grp[el >> 3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        double Y[18];
        p2s_bwd_sd(M, A[c], Y);
        double2 *Eown = EY + (c * 9) * ST;
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const uint32_t l = lidx[(c * 9 + r) * FE];  // orientation flags again (the table is still in the blob buffer)
          double2 v;
          if (GENERAL) v = p2_local_to_pair(Y[2 * r] - Y[2 * r + 1], Y[2 * r] + Y[2 * r + 1], l);
          else {
            const double b0 = p2_flip(Y[2 * r], (l << 16) & 0x80000000u);
            v = make_double2(b0 - Y[2 * r + 1], b0 + Y[2 * r + 1]);
          }
          Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = v;
        }
      }
    }
    env.sync();  // element results visible
    // ---- sum the slots of every patch pair; private pairs are plain 16-byte stores, shared ones RED
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + HDR;
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(B + B[8]);
      const uint32_t *e32 = B + B[9];
      const uint32_t *e64 = B + B[10];  // two words per pair
      const uint32_t *ovf = B + B[11];
      constexpr uint32_t SKIP = PATCH_ESS0 | PATCH_ESS1;
      auto e_at = [&](uint32_t off) { return *reinterpret_cast<const double2 *>(Eb + off); };
      auto emit = [&](const uint32_t m, const double2 v) {
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & (SKIP | PATCH_SHARED))) {
          if (ACC) { const double2 o = *dst; *dst = make_double2(o.x + v.x, o.y + v.y); }
          else *dst = v;
        } else if ((m & SKIP) != SKIP) {
          if (m & PATCH_SHARED) {
            if (!(m & PATCH_ESS0)) env.red_add(&dst->x, v.x);
            if (!(m & PATCH_ESS1)) env.red_add(&dst->y, v.y);
          } else {
            if (!(m & PATCH_ESS0)) dst->x = ACC ? dst->x + v.x : v.x;
            if (!(m & PATCH_ESS1)) dst->y = ACC ? dst->y + v.y : v.y;
          }
        }
      };
      constexpr int UN = 4;  // pairs in flight per thread: table reads, then result reads, then the stores
      for (uint32_t r0 = tid; r0 < n1; r0 += UN * NT) {
        uint32_t m[UN], o[UN];
        double2 v[UN];
#pragma unroll
        for (int i = 0; i < UN; i++) {
          const uint32_t r = r0 + i * NT;
          m[i] = r < n1 ? dof[r] : SKIP;
          o[i] = r < n1 ? e16[r] : 0u;
        }
#pragma unroll
        for (int i = 0; i < UN; i++) v[i] = e_at(o[i]);
#pragma unroll
        for (int i = 0; i < UN; i++) emit(m[i], v[i]);
      }
      dof += n1;
      for (uint32_t r0 = tid; r0 < n2; r0 += UN * NT) {
        uint32_t m[UN], o[UN];
        double2 v[UN];
#pragma unroll
        for (int i = 0; i < UN; i++) {
          const uint32_t r = r0 + i * NT;
          m[i] = r < n2 ? dof[r] : SKIP;
          o[i] = r < n2 ? e32[r] : 0u;
        }
#pragma unroll
        for (int i = 0; i < UN; i++) {
          const double2 p = e_at(o[i] & 0xffffu), q = e_at(o[i] >> 16);
          v[i] = make_double2(p.x + q.x, p.y + q.y);
        }
#pragma unroll
        for (int i = 0; i < UN; i++) emit(m[i], v[i]);
      }
      dof += n2;
#pragma unroll 2
      for (uint32_t r = tid; r < n4; r += NT) {
        const uint32_t m = dof[r];
        const uint32_t ox = e64[2 * r], oy = e64[2 * r + 1];
        const double2 p = e_at(ox & 0xffffu), q = e_at(ox >> 16), s = e_at(oy & 0xffffu);
        double2 v = make_double2((p.x + q.x) + s.x, (p.y + q.y) + s.y);
        if ((oy >> 16) != 0xffffu) { const double2 t = e_at(oy >> 16); v.x += t.x; v.y += t.y; }
        emit(m, v);
      }
      dof = B + HDR;
      for (uint32_t t = tid; t < nov; t += NT) {  // slots beyond four (unstructured meshes)
        const uint32_t e = ovf[t], m = dof[e >> 16];
        const double2 v = e_at(e & 0xffffu);
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & PATCH_ESS0)) env.red_add(&dst->x, v.x);
        if (!(m & PATCH_ESS1)) env.red_add(&dst->y, v.y);
      }
    }
    env.gather_wait();
    env.sync();  // blob buffer b and EY are free; Xs of the next patch complete and visible
    if (tid == 0 && it + 2 < n_it) request(b, q0, q1);
  }
  if (DOT) env.dot_finish(dacc, a, reinterpret_cast<double *>(EY));
}

}  // namespace hcurl
