// Order-2 Nedelec hex patch apply as an environment-parametrised per-thread PROGRAM (see hex_patch_p1_body.h for the
// idea): the algorithm of hex_patch_modal_p2_kernel / hex_patch_split_p2_kernel (hex_patch.cu, where the design is
// described) restated against `Env`, so that the SAME source runs as a CUDA kernel (hex_patch_p2e.cu) and on host
// threads (tests/emu/patch_p2_emu.cpp), where it is checked against the oracle and under ThreadSanitizer.  This is the
// base for the next optimisation steps of the headline kernel: a change is verified on the host before it costs GPU time.
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204).
//
// SPLIT = false: one blob per patch, double-buffered (2 mbarriers), rows of the exchange / result buffers padded (FE+1).
// SPLIT = true:  the blob arrives in three parts by lifetime -- header + pair list (double-buffered, mbarriers 0/1),
//                slot-offset lists (mbarrier 2) and per-slot table (mbarrier 3), single-buffered and requested one
//                iteration ahead -- and the rows are not padded: 44 KB instead of 49 KB of shared memory per CTA.
//
// Shared memory, SPLIT = false:  [16 B mbarriers] [2 x blob_cap] [Xs: xs_cap] [E2: 27 x ST double2] [EY: 27 x ST double2]
//                SPLIT = true:   [32 B mbarriers] [2 x d_cap] [lists: l_cap] [table: T_CAP] [Xs] [E2] [EY]
//
// Env: as for hex_patch_p1_body.h, plus  gather16(double2 *, const double2 *),  double2 ldg(const double2 *),
//      bool warp_all(bool)  (= __all_sync over the full warp)  and  const ModalTables &tables<ModalTables>().
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef HCURL_BODY
#ifdef __CUDACC__
#define HCURL_BODY __device__ __forceinline__
#else
#define HCURL_BODY inline
#endif
#endif

namespace hcurl {

HCURL_BODY int p2_sym6(int i, int j) {  // packed index of the symmetric pair (i,j)
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  return lo * 3 - lo * (lo - 1) / 2 + (hi - lo);
}

#define P2_IA(m0, m1, m2) ((m0) + 2 * ((m1) + 3 * (m2)))  // own component  [2][3][3]
#define P2_IB(m0, m1, m2) ((m0) + 3 * ((m1) + 2 * (m2)))  // component c+1  [3][2][3] (my frame)
#define P2_IC(m0, m1, m2) ((m0) + 3 * ((m1) + 3 * (m2)))  // component c+2  [3][3][2] (my frame)

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
// nodal -> (2x) modal on own component
HCURL_BODY void p2_modal_fwd(const ModalTables &M, const double *X, double *A) {
  double a[18], b[18];
#pragma unroll
  for (int r = 0; r < 9; r++) { a[2 * r] = X[2 * r] + X[2 * r + 1]; a[2 * r + 1] = X[2 * r + 1] - X[2 * r]; }
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      p2_tc_fwd(M, a[P2_IA(m0, 0, i2)], a[P2_IA(m0, 1, i2)], a[P2_IA(m0, 2, i2)], b[P2_IA(m0, 0, i2)], b[P2_IA(m0, 1, i2)], b[P2_IA(m0, 2, i2)]);
#pragma unroll
  for (int r = 0; r < 6; r++) p2_tc_fwd(M, b[r], b[r + 6], b[r + 12], A[r], A[r + 6], A[r + 12]);
}
// modal -> nodal test functions (transpose, again with the open transform scaled by 2)
HCURL_BODY void p2_modal_bwd(const ModalTables &M, const double *Yh, double *Y) {
  double a[18], b[18];
#pragma unroll
  for (int r = 0; r < 6; r++) p2_tc_bwd(M, Yh[r], Yh[r + 6], Yh[r + 12], a[r], a[r + 6], a[r + 12]);
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      p2_tc_bwd(M, a[P2_IA(m0, 0, i2)], a[P2_IA(m0, 1, i2)], a[P2_IA(m0, 2, i2)], b[P2_IA(m0, 0, i2)], b[P2_IA(m0, 1, i2)], b[P2_IA(m0, 2, i2)]);
#pragma unroll
  for (int r = 0; r < 9; r++) { Y[2 * r] = b[2 * r] - b[2 * r + 1]; Y[2 * r + 1] = b[2 * r] + b[2 * r + 1]; }
}

// modal element operator for component c: reads own (Am) and the other two components' modal coefficients
// (Bm = comp c+1, Cm = comp c+2, in my frame), scaled geometry factors
HCURL_BODY void p2_modal_op(const ModalTables &M, const double *Am, const double *Bm, const double *Cm, const double M00, const double M01,
                            const double M02, const double K11, const double K10, const double K12, const double K22, const double K20,
                            const double K21, double *Yh) {
  const double d1 = M.d1, d2 = M.d2;
#define P2_DD(n) ((n) == 1 ? d1 : d2)
  double r0[8], r1[12], r2[12];
#pragma unroll
  for (int m0 = 0; m0 < 2; m0++) {
#pragma unroll
    for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
        r0[m0 + 2 * (m1 + 2 * m2)] = fma(P2_DD(m1 + 1), Cm[P2_IC(m0, m1 + 1, m2)], -P2_DD(m2 + 1) * Bm[P2_IB(m0, m1, m2 + 1)]);
#pragma unroll
    for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
        r1[m0 + 2 * (m1 + 3 * m2)] = fma(P2_DD(m2 + 1), Am[P2_IA(m0, m1, m2 + 1)], -P2_DD(m0 + 1) * Cm[P2_IC(m0 + 1, m1, m2)]);
#pragma unroll
    for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 3; m2++)
        r2[m0 + 2 * (m1 + 2 * m2)] = fma(P2_DD(m0 + 1), Bm[P2_IB(m0 + 1, m1, m2)], -P2_DD(m1 + 1) * Am[P2_IA(m0, m1 + 1, m2)]);
  }
  double w1[12], w2[12];
#pragma unroll
  for (int m0 = 0; m0 < 2; m0++) {
#pragma unroll
    for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++) {
        double v = K11 * r1[m0 + 2 * (m1 + 3 * m2)];
        if (m1 < 2) v = fma(K12, r2[m0 + 2 * (m1 + 2 * m2)], fma(K10, r0[m0 + 2 * (m1 + 2 * m2)], v));
        w1[m0 + 2 * (m1 + 3 * m2)] = v;
      }
#pragma unroll
    for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 3; m2++) {
        double v = K22 * r2[m0 + 2 * (m1 + 2 * m2)];
        if (m2 < 2) v = fma(K21, r1[m0 + 2 * (m1 + 3 * m2)], fma(K20, r0[m0 + 2 * (m1 + 2 * m2)], v));
        w2[m0 + 2 * (m1 + 2 * m2)] = v;
      }
  }
#pragma unroll
  for (int n2 = 0; n2 < 3; n2++)
#pragma unroll
    for (int n1 = 0; n1 < 3; n1++)
#pragma unroll
      for (int n0 = 0; n0 < 2; n0++) {
        double v = M00 * Am[P2_IA(n0, n1, n2)];
        if (n1 < 2) v = fma(M01, Bm[P2_IB(n0, n1, n2)], v);
        if (n2 < 2) v = fma(M02, Cm[P2_IC(n0, n1, n2)], v);
        if (n2 >= 1) v = fma(P2_DD(n2), w1[n0 + 2 * (n1 + 3 * (n2 - 1))], v);
        if (n1 >= 1) v = fma(-P2_DD(n1), w2[n0 + 2 * ((n1 - 1) + 2 * n2)], v);
        Yh[P2_IA(n0, n1, n2)] = v;
      }
#undef P2_DD
}

// orientation of an element slot pair relative to its global pair, carried in bits 13..15 of a lidx entry
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

template <class Env, int FE, bool ACC, bool DOT, bool SPLIT>
HCURL_BODY void hex_patch_p2_body(Env &env, const PatchArgs &a) {
  constexpr int NT = 3 * FE, ST = SPLIT ? FE : FE + 1, HDR = PatchHost::HDR;
  constexpr uint32_t T_CAP = (27 * FE * 2 + 15) & ~15;
  const ModalTables &M = env.template tables<ModalTables>();
  unsigned char *smem = env.smem();
  const uint32_t d_cap = SPLIT ? a.d_cap_bytes : a.blob_cap_bytes;  // room of one double-buffered part
  unsigned char *blobB = smem + (SPLIT ? 32 : 16);
  unsigned char *listB = blobB + 2 * (size_t)d_cap;                  // SPLIT: e16 | e32 | e64 | ovf of the patch being summed
  unsigned char *tabB = listB + (SPLIT ? a.l_cap_bytes : 0u);        // SPLIT: lidx of the patch being pulled
  double2 *Xs = reinterpret_cast<double2 *>(tabB + (SPLIT ? T_CAP : 0u));
  double2 *E2 = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);  // modal exchange
  double2 *EY = E2 + 27 * ST;                                                                            // element results
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
  double2 *y2 = reinterpret_cast<double2 *>(a.y);
  const int tid = env.tid();
  double dacc = 0.0;  // DOT: my share of x^T A x, summed in modal space (x_e^T A_e x_e = Am^T Yh)
  const int64_t k0 = a.k_begin + env.block(), kstep = env.nblocks();
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);
  // component c of patch element el, in the rotated frame (c, c+1, c+2)
  const int c = tid / FE, el = tid - c * FE;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int oK11 = p2_sym6(c1, c1) * FE + el, oK10 = p2_sym6(c1, c) * FE + el, oK12 = p2_sym6(c1, c2) * FE + el,
            oK22 = p2_sym6(c2, c2) * FE + el, oK20 = p2_sym6(c2, c) * FE + el, oM00 = (6 + p2_sym6(c, c)) * FE + el,
            oM01 = (6 + p2_sym6(c, c1)) * FE + el, oM02 = (6 + p2_sym6(c, c2)) * FE + el;

  // thread 0: request the parts of patch k (offsets in 16-byte units; SPLIT: begin | lists | table | end)
  auto request = [&](int b, uint32_t o0, uint32_t o1) {  // whole blob, or header + pair list -> buffer b
    env.bulk_request(b, blobB + (size_t)b * d_cap, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u);
  };
  auto request_lists = [&](uint32_t oL, uint32_t oT) { env.bulk_request(2, listB, a.blob + (size_t)oL * 4, (oT - oL) * 16u); };
  auto request_table = [&](uint32_t oT, uint32_t o1) { env.bulk_request(3, tabB, a.blob + (size_t)oT * 4, (o1 - oT) * 16u); };
  auto part_end = [&](int64_t k) { return SPLIT ? env.ldg(a.part_off + 2 * k) : env.ldg(a.blob_off + k + 1); };
  // launch the x gather of the patch whose blob sits in buffer b
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * d_cap);
    const uint32_t U = B[0];
    const uint32_t *dof = B + HDR;
#pragma unroll 2
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
    env.bulk_init(SPLIT ? 4 : 2);
    request(0, env.ldg(a.blob_off + k0), part_end(k0));
    if (SPLIT) {
      request_lists(env.ldg(a.part_off + 2 * k0), env.ldg(a.part_off + 2 * k0 + 1));
      request_table(env.ldg(a.part_off + 2 * k0 + 1), env.ldg(a.blob_off + k0 + 1));
    }
    if (n_it > 1) request(1, env.ldg(a.blob_off + k0 + kstep), part_end(k0 + kstep));
  }
  env.sync();  // mbarriers initialised
  env.bulk_wait(0, 0);
  gather(0);
  env.gather_wait();
  env.sync();  // Xs of the first patch complete and visible

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * d_cap);
    uint32_t q0 = 0, q1 = 0;          // extent of (the double-buffered part of) patch it+2, requested when this one retires
    uint32_t nL = 0, nT = 0, nE = 0;  // SPLIT: lists | table | end of patch it+1 (requested after the sums / after the pull)
    if (tid == 0 && it + 2 < n_it) { q0 = env.ldg(a.blob_off + k + 2 * kstep); q1 = part_end(k + 2 * kstep); }
    if (SPLIT && tid == 0 && it + 1 < n_it) {
      nL = env.ldg(a.part_off + 2 * (k + kstep)); nT = env.ldg(a.part_off + 2 * (k + kstep) + 1); nE = env.ldg(a.blob_off + k + kstep + 1);
    }
    // affine geometry factors of my element: issued now, consumed after the forward transform
    const double *g = a.edata + (size_t)k * 12 * FE;
    double K11 = env.ldg(g + oK11), K10 = env.ldg(g + oK10), K12 = env.ldg(g + oK12), K22 = env.ldg(g + oK22), K20 = env.ldg(g + oK20),
           M00 = env.ldg(g + oM00), M01 = env.ldg(g + oM01), M02 = env.ldg(g + oM02);
    const bool ok = (uint32_t)el < B[13];
    // ---- pull the element inputs through the per-slot table, forward transform
    double Am[18];
    uint32_t fl = 0;
    // swap / sign flags of my nine slots, three bits each: all the result store needs of the per-slot table (which, SPLIT,
    // is gone by then) -- one register across the element operator instead of the nine table entries
    uint32_t fpk = 0;
    if (SPLIT) env.bulk_wait(3, it & 1);  // per-slot table of this patch
    const uint16_t *lidx = (SPLIT ? reinterpret_cast<const uint16_t *>(tabB) : reinterpret_cast<const uint16_t *>(B + B[12])) + (c * 9) * FE + el;
    if (ok) {
#pragma unroll
      for (int r = 0; r < 9; r++) fl |= lidx[r * FE];
    }
    // warp-uniform: no element of this warp sees a pair swapped or sign-flipped (consistently oriented meshes)
    const bool plain = env.warp_all((fl & 0xe000u) == 0u);
    if (ok) {
      double X[18];
      if (plain) {
#pragma unroll
        for (int r = 0; r < 9; r++) { const double2 v = Xs[lidx[r * FE]]; X[2 * r] = v.x; X[2 * r + 1] = v.y; }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) {
          const uint32_t l = lidx[r * FE];
          p2_pair_to_local(Xs[l & 0x1fffu], l, X[2 * r], X[2 * r + 1]);
          fpk |= ((l >> 13) & 7u) << (3 * r);
        }
      }
      p2_modal_fwd(M, X, Am);
      double2 *Eown = E2 + (c * 9) * ST + el;
#pragma unroll
      for (int r = 0; r < 9; r++) Eown[r * ST] = make_double2(Am[2 * r], Am[2 * r + 1]);
    }
    env.sync();  // modal coefficients visible; Xs (and, SPLIT, the per-slot table) free
    if (SPLIT && tid == 0 && it + 1 < n_it) request_table(nT, nE);
    double Y[18];
    if (ok) {
      const double2 *E1 = E2 + (c1 * 9) * ST + el, *E2c = E2 + (c2 * 9) * ST + el;
      double Bm[18], Cm[18];
      // owner frames: comp c+1 stores (m1, m2, m0), comp c+2 stores (m2, m0, m1); the fastest owner index is the pair
#pragma unroll
      for (int m2 = 0; m2 < 3; m2++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) {
          const double2 v = E1[(m2 + 3 * m0) * ST];
          Bm[P2_IB(m0, 0, m2)] = v.x;
          Bm[P2_IB(m0, 1, m2)] = v.y;
        }
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) {
          const double2 v = E2c[(m0 + 3 * m1) * ST];
          Cm[P2_IC(m0, m1, 0)] = v.x;
          Cm[P2_IC(m0, m1, 1)] = v.y;
        }
      const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
      K11 *= sc; K10 *= sc; K12 *= sc; K22 *= sc; K20 *= sc; M00 *= sc; M01 *= sc; M02 *= sc;
      double Yh[18];
      p2_modal_op(M, Am, Bm, Cm, M00, M01, M02, K11, K10, K12, K22, K20, K12, Yh);
      if (DOT) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int t = 0; t < 18; t += 2) { d0 = fma(Am[t], Yh[t], d0); d1 = fma(Am[t + 1], Yh[t + 1], d1); }
        dacc += d0 + d1;
      }
      p2_modal_bwd(M, Yh, Y);
    }
    // ---- start the x gather of the next patch: it lands while this patch is scattered
    if (it + 1 < n_it) {
      env.bulk_wait(b ^ 1, ((it + 1) >> 1) & 1);
      gather(b ^ 1);
    }
    if (ok) {
      double2 *Eown = EY + (c * 9) * ST;
      const int elx = el ^ a.swz_grp[el >> 3];
      if (plain) {
#pragma unroll
        for (int r = 0; r < 9; r++) Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = make_double2(Y[2 * r], Y[2 * r + 1]);
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = p2_local_to_pair(Y[2 * r], Y[2 * r + 1], ((fpk >> (3 * r)) & 7u) << 13);
      }
    }
    env.sync();  // element results visible; everyone has read the modal coefficients
    // ---- sum the slots of every patch pair; private pairs are plain 16-byte stores, shared ones RED
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + HDR;
      if (SPLIT) env.bulk_wait(2, it & 1);  // slot-offset lists of this patch
      // SPLIT: the lists start at word B[8] of the blob and sit at the start of listB
      const uint32_t *Lw = SPLIT ? reinterpret_cast<const uint32_t *>(listB) - B[8] : B;
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(Lw + B[8]);
      const uint32_t *e32 = Lw + B[9];
      const uint32_t *e64 = Lw + B[10];  // two words per pair
      const uint32_t *ovf = Lw + B[11];
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
      constexpr int UN = 2;  // pairs in flight per thread: table reads, then E reads, then the stores
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
    env.sync();  // buffer b (SPLIT: and the lists) and EY are free; Xs of the next patch complete and visible
    if (tid == 0) {
      if (SPLIT && it + 1 < n_it) request_lists(nL, nT);
      if (it + 2 < n_it) request(b, q0, q1);
    }
  }
  if (DOT) env.dot_finish(dacc, a, reinterpret_cast<double *>(E2));
}

}  // namespace hcurl
