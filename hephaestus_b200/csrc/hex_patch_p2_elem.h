// Element side of the order-2 Nedelec hex patch kernel: ONE THREAD PER ELEMENT (the pipeline around it: patch_pipeline.h).
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204), affine order-2 hexes, per-element coefficients.
//
// The unit of the dof side is a PAIR: the two order-2 dofs along one edge / face line / interior line are adjacent in
// MFEM's numbering and always used together (swapped and / or sign-flipped by the element's orientation), so gathers,
// staging, tables, sums and stores are all 16-byte.  Element slots: component c has the 18 slots t = i0 + 2 (i1 + 3 i2),
// i0 the open index along axis c, (i1, i2) the closed indices along the axes c+1, c+2; slot pair r = i1 + 3 i2 of
// component c is row c*9 + r of the per-slot table and of the result buffer.
//
// The element operator is applied in an orthonormal Legendre (MODAL) basis, where the mass matrix is a restricted
// identity and the curl has one term per mode:  A_e = T^T [ GM (x) I + C^T (GK (x) I) C ] T  with GK = alpha J^T J / det,
// GM = beta adj adj^T / det -- exact on affine elements with per-element coefficients, as is the (p+1)-point Gauss rule
// it replaces, at about 960 FP64 instructions per element instead of 2916 DFMA.
//
// Why one thread per element (measured, profiles/r2*): with three component threads per element the L1tex data pipe was
// the loaded unit (73 %), half of its shared-memory wavefronts being the exchange of modal coefficients between the
// component threads; six threads per element were slower still (0.67 against 0.447 ms).  Here the exchange buffer and
// its barrier do not exist, a 32-element patch is one warp, the modal curl R (36 coefficients) and W = K R are computed
// once per element, and the operator runs IN PLACE on 90 live doubles (A -> M A and R -> W mode by mode; the fused
// <x, A x> is A.(M A) + R.W), which fits the 255-register budget without spills.
#pragma once
#include "patch_pipeline.h"

namespace hcurl {

#define P2_IA(m0, m1, m2) ((m0) + 2 * ((m1) + 3 * (m2)))  // modal / nodal index of a component: [2 open][3][3]

// 1-D closed transform of a line (x0, x1, x2) -> (m0, m1, m2)
HCURL_BODY void p2_tc_fwd(const ModalTables &M, const double x0, const double x1, const double x2, double &m0, double &m1, double &m2) {
  const double s = x0 + x2, d = x2 - x0;
  m0 = fma(M.a1, x1, M.a0 * s);
  m1 = M.b * d;
  m2 = fma(M.c1, x1, M.c0 * s);
}
// transpose: (y0, y1, y2) modal test values -> nodal
HCURL_BODY void p2_tc_bwd(const ModalTables &M, const double y0, const double y1, const double y2, double &x0, double &x1, double &x2) {
  const double p = fma(M.c0, y2, M.a0 * y0), q = M.b * y1;
  x0 = p - q;
  x2 = p + q;
  x1 = fma(M.c1, y2, M.a1 * y0);
}
// orientation of an element slot pair relative to its global pair, carried in bits 13..15 of a per-slot table entry
// (PAIR_SWAP / PAIR_NEG1 / PAIR_NEG0 of the slot code >> 16): local = (s0 * g[swap], s1 * g[!swap])
HCURL_BODY double p2_flip(double v, uint32_t sign_bit31) {
#ifdef __CUDA_ARCH__
  return __hiloint2double(__double2hiint(v) ^ (int)sign_bit31, __double2loint(v));
#else
  return sign_bit31 ? -v : v;
#endif
}
HCURL_BODY void p2_pair_to_local(const double2 g, const uint32_t l, double &x0, double &x1) {
  const bool sw = l & 0x2000u;
  x0 = p2_flip(sw ? g.y : g.x, (l << 16) & 0x80000000u);
  x1 = p2_flip(sw ? g.x : g.y, (l << 17) & 0x80000000u);
}
HCURL_BODY double2 p2_local_to_pair(const double y0, const double y1, const uint32_t l) {
  const bool sw = l & 0x2000u;
  const double a = p2_flip(y0, (l << 16) & 0x80000000u), b = p2_flip(y1, (l << 17) & 0x80000000u);
  return sw ? make_double2(b, a) : make_double2(a, b);
}

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

// GENERAL = false: every flagged element slot pair of the space is a REVERSED LINE (all three orientation bits set: the
// only case MFEM's maps produce; checked on the host when the patches are built) -- two instructions per pair instead
// of a dozen; GENERAL = true: independent swap / sign bits.
template <int FE, bool GENERAL>
struct P2Elem {
  typedef UnitPair Unit;
  static constexpr int ROWS = 27;
  double A[3][18];  // nodal (s, d) values -> modal coefficients -> modal test values
  double G[12];     // GK (6) | GM (6), packed symmetric
  uint32_t rev;     // !GENERAL: bit c*9 + r = slot pair (c, r) is a reversed line (kept from the pull for the result store)

  template <class Env> HCURL_BODY void geometry(Env &env, const PatchArgs &a, const int64_t k, const int el) {
    const double *g = a.edata + (size_t)k * 12 * FE + el;  // [12][FE] per patch: coalesced
#pragma unroll
    for (int s = 0; s < 12; s++) G[s] = env.ldg(g + s * FE);
  }
  template <class Env> HCURL_BODY void pull(Env &, const PatchArgs &a, const double2 *Xs, const uint16_t *lidx) {
    rev = 0u;
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
          rev |= ((l >> 15) & 1u) << (c * 9 + r);
        }
      }
      p2s_fwd_sd(a.modal, X, A[c]);
    }
  }
  template <bool DOT, class Env> HCURL_BODY double apply(Env &, const PatchArgs &a) {
    const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
#pragma unroll
    for (int s = 0; s < 12; s++) G[s] *= sc;
    return p2s_modal_op<DOT>(a.modal, A, G, G + 6);
  }
  template <class Env> HCURL_BODY void store(Env &, const PatchArgs &a, double2 *EY, const uint16_t *lidx, const int el) {
    constexpr int ST = FE + 1;
    const int elx = el ^ a.swz_grp[el >> 3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double Y[18];
      p2s_bwd_sd(a.modal, A[c], Y);
      double2 *Eown = EY + (c * 9) * ST;
#pragma unroll
      for (int r = 0; r < 9; r++) {
        double2 v;
        if (GENERAL) {
          const uint32_t l = lidx[(c * 9 + r) * FE];  // orientation flags again (the table is still in shared memory)
          v = p2_local_to_pair(Y[2 * r] - Y[2 * r + 1], Y[2 * r] + Y[2 * r + 1], l);
        } else {
          const double b0 = p2_flip(Y[2 * r], (rev << (31 - (c * 9 + r))) & 0x80000000u);
          v = make_double2(b0 - Y[2 * r + 1], b0 + Y[2 * r + 1]);
        }
        Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = v;
      }
    }
  }
};

}  // namespace hcurl
