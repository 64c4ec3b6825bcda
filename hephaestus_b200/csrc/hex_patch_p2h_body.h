// Order-2 patch apply with TWO threads per element component ("half" variant) -- an Env program like
// hex_patch_p2_body.h, whose design and notation it shares.  NOT YET RUN ON A GPU (written after the round-1 GPU budget
// was spent; verified on host threads, tests/test_patch_p2_emu.py).
//
// Why: the order-2 kernel is latency-bound at 12 warps per SM; registers (140) and shared memory (49 KB)
// allow no more CTAs.  The modal element operator of component c splits exactly in two by the open-direction mode
// m0 = h in {0, 1}: half h produces Yh(h, :, :) from its own Am(h, :, :) and the rows m0 in {h, h+1} of the other two
// components' coefficients.  So thread (c, e, h) carries 9 instead of 18 values of everything (12 + 12 instead of
// 18 + 18 exchanged coefficients), the CTA has 6 FE threads for the same shared memory, and the register need per
// thread drops accordingly -- twice the warps per SM if it fits 80 registers, 1.5 x at 96.
//   * the two halves of an element are ADJACENT LANES (lane = 2 e' + h): the pull reads of a pair and of the geometry
//     factors are broadcasts, the 8-byte stores of the two halves of a pair coalesce into the 16-byte layout of the
//     full-thread kernel (exchange buffer E2 and result buffer EY are laid out exactly as there, so the blob format
//     and the dof side are unchanged), and the final open-direction butterfly (Y = b0 -+ b1) is one shuffle per pair.
//   * extra cost: the pull is done by both halves (broadcast: no extra shared-memory wavefronts), 12 instead of 9
//     16-byte exchange reads per thread (72 instead of 54 per element), 9 shuffles per thread.
// Env: as for hex_patch_p2_body.h, plus  double shfl_xor(double v, int lane_mask)  (called by every thread of the warp).
#pragma once
#include "hex_patch_p2_body.h"

namespace hcurl {

template <class Env, int FE, bool ACC, bool DOT>
HCURL_BODY void hex_patch_p2h_body(Env &env, const PatchArgs &a) {
  constexpr int NT = 6 * FE, ST = FE + 1, HDR = PatchHost::HDR;
  const ModalTables &M = env.template tables<ModalTables>();
  unsigned char *smem = env.smem();
  unsigned char *blobB = smem + 16;
  double2 *Xs = reinterpret_cast<double2 *>(blobB + 2 * (size_t)a.blob_cap_bytes);
  double2 *E2 = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);  // modal exchange
  double2 *EY = E2 + 27 * ST;                                                                            // element results
  double *E2d = reinterpret_cast<double *>(E2), *EYd = reinterpret_cast<double *>(EY);
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
  double2 *y2 = reinterpret_cast<double2 *>(a.y);
  const int tid = env.tid();
  double dacc = 0.0;
  const int64_t k0 = a.k_begin + env.block(), kstep = env.nblocks();
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);
  // component c of patch element el, half h (open-direction mode m0 = h), in the rotated frame (c, c+1, c+2)
  const int c = tid / (2 * FE), idx = tid - c * 2 * FE, el = idx >> 1, h = idx & 1;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int oK11 = p2_sym6(c1, c1) * FE + el, oK10 = p2_sym6(c1, c) * FE + el, oK12 = p2_sym6(c1, c2) * FE + el,
            oK22 = p2_sym6(c2, c2) * FE + el, oK20 = p2_sym6(c2, c) * FE + el, oM00 = (6 + p2_sym6(c, c)) * FE + el,
            oM01 = (6 + p2_sym6(c, c1)) * FE + el, oM02 = (6 + p2_sym6(c, c2)) * FE + el;
  const double d1 = M.d1, d2 = M.d2;
  const double dh = h ? d2 : d1;  // DD(h + 1)

  auto request = [&](int b, uint32_t o0, uint32_t o1) {
    env.bulk_request(b, blobB + (size_t)b * a.blob_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u);
  };
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    const uint32_t U = B[0];
    const uint32_t *dof = B + HDR;
#pragma unroll 2
    for (uint32_t u = tid; u < U; u += NT) {
      const uint32_t m = dof[u];
      if (!(m & (PATCH_ESS0 | PATCH_ESS1))) env.gather16(Xs + u, x2 + (m & PATCH_IDX));
      else {
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
  env.sync();
  env.bulk_wait(0, 0);
  gather(0);
  env.gather_wait();
  env.sync();

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    uint32_t q0 = 0, q1 = 0;
    if (tid == 0 && it + 2 < n_it) { q0 = env.ldg(a.blob_off + k + 2 * kstep); q1 = env.ldg(a.blob_off + k + 2 * kstep + 1); }
    const double *g = a.edata + (size_t)k * 12 * FE;
    double K11 = env.ldg(g + oK11), K10 = env.ldg(g + oK10), K12 = env.ldg(g + oK12), K22 = env.ldg(g + oK22), K20 = env.ldg(g + oK20),
           M00 = env.ldg(g + oM00), M01 = env.ldg(g + oM01), M02 = env.ldg(g + oM02);
    const bool ok = (uint32_t)el < B[13];
    // ---- pull (both halves read the pairs: broadcast), open-direction butterfly for my mode, closed transforms
    double A9[9];  // Am(h, m1, m2) at r = m1 + 3 m2
    // per-slot table entries (staging position | swap / sign flags): read here for the pull and AGAIN at the result
    // store of the non-plain path instead of being kept in nine registers across the element operator
    const uint16_t *lidx = reinterpret_cast<const uint16_t *>(B + B[12]) + (c * 9) * FE + el;
    uint32_t fl = 0;
    if (ok) {
#pragma unroll
      for (int r = 0; r < 9; r++) fl |= lidx[r * FE];
    }
    // warp-uniform: no element of this warp sees a pair swapped or sign-flipped (consistently oriented meshes)
    const bool plain = env.warp_all((fl & 0xe000u) == 0u);
    if (ok) {
      double ah[9], bh[9];
      uint32_t lx[9];
#pragma unroll
      for (int r = 0; r < 9; r++) lx[r] = lidx[r * FE];
      if (plain) {
#pragma unroll
        for (int r = 0; r < 9; r++) { const double2 v = Xs[lx[r]]; ah[r] = h ? v.y - v.x : v.x + v.y; }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) {
          double x0, x1;
          p2_pair_to_local(Xs[lx[r] & 0x1fffu], lx[r], x0, x1);
          ah[r] = h ? x1 - x0 : x0 + x1;
        }
      }
#pragma unroll
      for (int i2 = 0; i2 < 3; i2++) p2_tc_fwd(M, ah[3 * i2], ah[3 * i2 + 1], ah[3 * i2 + 2], bh[3 * i2], bh[3 * i2 + 1], bh[3 * i2 + 2]);
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++) p2_tc_fwd(M, bh[m1], bh[m1 + 3], bh[m1 + 6], A9[m1], A9[m1 + 3], A9[m1 + 6]);
      double *Eown = E2d + ((size_t)(c * 9) * ST + el) * 2 + h;  // my half of the pair (m0 = 0, m0 = 1)
#pragma unroll
      for (int r = 0; r < 9; r++) Eown[(size_t)r * ST * 2] = A9[r];
    }
    env.sync();  // modal coefficients visible; Xs free
    double b9[9];  // backward-transformed results of my mode: nodal in the closed directions
    if (ok) {
      const double2 *E1 = E2 + (c1 * 9) * ST + el, *E2c = E2 + (c2 * 9) * ST + el;
      // rows m0 = h ("lo") and m0 = h + 1 ("hi") of the other components' coefficients, in my frame:
      //   Bm(m0, m1 in {0,1}, m2): pair at row m2 + 3 m0 of component c+1;  Cm(m0, m1, m2 in {0,1}): pair at row m0 + 3 m1 of c+2
      const double sc = 0.25 * a.scale;
      K11 *= sc; K10 *= sc; K12 *= sc; K22 *= sc; K20 *= sc; M00 *= sc; M01 *= sc; M02 *= sc;
      const double K21 = K12;
#define P2H_DD(n) ((n) == 1 ? d1 : d2)
#define P2H_A(m1, m2) A9[(m1) + 3 * (m2)]
      // ordered for short live ranges: the "lo" rows feed the mass terms and r0 and die; then one "hi" block at a time
      double Yh[9], r0[2][2], r1[3][2], r2[2][3];
      {
        double Bl[2][3], Cl[3][2];
#pragma unroll
        for (int m2 = 0; m2 < 3; m2++) { const double2 v = E1[(m2 + 3 * h) * ST]; Bl[0][m2] = v.x; Bl[1][m2] = v.y; }
#pragma unroll
        for (int m1 = 0; m1 < 3; m1++) { const double2 v = E2c[(h + 3 * m1) * ST]; Cl[m1][0] = v.x; Cl[m1][1] = v.y; }
#pragma unroll
        for (int n2 = 0; n2 < 3; n2++)
#pragma unroll
          for (int n1 = 0; n1 < 3; n1++) {
            double v = M00 * P2H_A(n1, n2);
            if (n1 < 2) v = fma(M01, Bl[n1][n2], v);
            if (n2 < 2) v = fma(M02, Cl[n1][n2], v);
            Yh[n1 + 3 * n2] = v;
          }
#pragma unroll
        for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
          for (int m2 = 0; m2 < 2; m2++) r0[m1][m2] = fma(P2H_DD(m1 + 1), Cl[m1 + 1][m2], -P2H_DD(m2 + 1) * Bl[m1][m2 + 1]);
      }
      {
        double Ch[3][2];
#pragma unroll
        for (int m1 = 0; m1 < 3; m1++) { const double2 w = E2c[(h + 1 + 3 * m1) * ST]; Ch[m1][0] = w.x; Ch[m1][1] = w.y; }
#pragma unroll
        for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
          for (int m2 = 0; m2 < 2; m2++) r1[m1][m2] = fma(P2H_DD(m2 + 1), P2H_A(m1, m2 + 1), -dh * Ch[m1][m2]);
      }
      {
        double Bh[2][3];
#pragma unroll
        for (int m2 = 0; m2 < 3; m2++) { const double2 w = E1[(m2 + 3 * (h + 1)) * ST]; Bh[0][m2] = w.x; Bh[1][m2] = w.y; }
#pragma unroll
        for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
          for (int m2 = 0; m2 < 3; m2++) r2[m1][m2] = fma(dh, Bh[m1][m2], -P2H_DD(m1 + 1) * P2H_A(m1 + 1, m2));
      }
      // Yh += D^T (K r): w1 = K11 r1 + K10 r0 + K12 r2, w2 = K22 r2 + K20 r0 + K21 r1 (same operation order as before)
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m2 = 0; m2 < 2; m2++) {
          double v = K11 * r1[m1][m2];
          if (m1 < 2) v = fma(K12, r2[m1][m2], fma(K10, r0[m1][m2], v));
          Yh[m1 + 3 * (m2 + 1)] = fma(P2H_DD(m2 + 1), v, Yh[m1 + 3 * (m2 + 1)]);
        }
#pragma unroll
      for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
        for (int m2 = 0; m2 < 3; m2++) {
          double v = K22 * r2[m1][m2];
          if (m2 < 2) v = fma(K21, r1[m1][m2], fma(K20, r0[m1][m2], v));
          Yh[(m1 + 1) + 3 * m2] = fma(-P2H_DD(m1 + 1), v, Yh[(m1 + 1) + 3 * m2]);
        }
#undef P2H_DD
#undef P2H_A
      if (DOT) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int t = 0; t < 8; t += 2) { s0 = fma(A9[t], Yh[t], s0); s1 = fma(A9[t + 1], Yh[t + 1], s1); }
        dacc += (s0 + s1) + A9[8] * Yh[8];
      }
      double th[9];
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++) p2_tc_bwd(M, Yh[m1], Yh[m1 + 3], Yh[m1 + 6], th[m1], th[m1 + 3], th[m1 + 6]);
#pragma unroll
      for (int i2 = 0; i2 < 3; i2++) p2_tc_bwd(M, th[3 * i2], th[3 * i2 + 1], th[3 * i2 + 2], b9[3 * i2], b9[3 * i2 + 1], b9[3 * i2 + 2]);
    } else {
#pragma unroll
      for (int r = 0; r < 9; r++) b9[r] = 0.0;
    }
    // ---- start the x gather of the next patch
    if (it + 1 < n_it) {
      env.bulk_wait(b ^ 1, ((it + 1) >> 1) & 1);
      gather(b ^ 1);
    }
    // ---- open-direction butterfly across the two halves (adjacent lanes), my half of every result pair to EY
    {
      double *Eown = EYd + (size_t)(c * 9) * ST * 2;
      const int elx = el ^ a.swz_grp[el >> 3];
#pragma unroll
      for (int r = 0; r < 9; r++) {
        const double pb = env.shfl_xor(b9[r], 1);                  // every thread of the warp takes part
        const double yl = h ? pb + b9[r] : b9[r] - pb;              // local dof h of the pair: Y = b0 -+ b1
        if (ok) {
          double *dst = Eown + ((size_t)r * ST + (elx ^ a.swz_row[c * 9 + r])) * 2;
          if (plain) dst[h] = yl;
          else {
            const uint32_t l = lidx[r * FE];
            const int slot = h ^ (int)((l >> 13) & 1u);             // pair swapped: local dof h sits in the other slot
            const uint32_t neg = h ? (l << 17) & 0x80000000u : (l << 16) & 0x80000000u;  // PAIR_NEG1 / PAIR_NEG0
            dst[slot] = p2_flip(yl, neg);
          }
        }
      }
    }
    env.sync();  // element results visible
    // ---- per-pair sums: identical to the full-thread kernel, with 6 FE threads
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + HDR;
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(B + B[8]);
      const uint32_t *e32 = B + B[9];
      const uint32_t *e64 = B + B[10];
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
      for (uint32_t r = tid; r < n1; r += NT) emit(dof[r], e_at(e16[r]));
      dof += n1;
      for (uint32_t r = tid; r < n2; r += NT) {
        const uint32_t o = e32[r];
        const double2 p = e_at(o & 0xffffu), q = e_at(o >> 16);
        emit(dof[r], make_double2(p.x + q.x, p.y + q.y));
      }
      dof += n2;
      for (uint32_t r = tid; r < n4; r += NT) {
        const uint32_t ox = e64[2 * r], oy = e64[2 * r + 1];
        const double2 p = e_at(ox & 0xffffu), q = e_at(ox >> 16), s = e_at(oy & 0xffffu);
        double2 v = make_double2((p.x + q.x) + s.x, (p.y + q.y) + s.y);
        if ((oy >> 16) != 0xffffu) { const double2 t = e_at(oy >> 16); v.x += t.x; v.y += t.y; }
        emit(dof[r], v);
      }
      dof = B + HDR;
      for (uint32_t t = tid; t < nov; t += NT) {
        const uint32_t e = ovf[t], m = dof[e >> 16];
        const double2 v = e_at(e & 0xffffu);
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & PATCH_ESS0)) env.red_add(&dst->x, v.x);
        if (!(m & PATCH_ESS1)) env.red_add(&dst->y, v.y);
      }
    }
    env.gather_wait();
    env.sync();
    if (tid == 0 && it + 2 < n_it) request(b, q0, q1);
  }
  if (DOT) env.dot_finish(dacc, a, reinterpret_cast<double *>(E2));
}

}  // namespace hcurl
