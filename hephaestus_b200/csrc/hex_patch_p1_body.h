// Order-1 Nedelec hex patch apply -- the per-thread program of one CTA, written against an execution environment `Env`
// so that the SAME source runs as the CUDA kernel (hex_patch_p1.cu: TMA bulk copies + mbarriers, cp.async gathers,
// red.global) and, for tests without a GPU, on host threads (tests/emu/patch_p1_emu.cpp: memcpy, std::barrier,
// compare-and-swap adds; run under ThreadSanitizer it also checks the barrier placement).
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204), order-1 Nedelec hexes (12 edge dofs), affine
// elements with per-element coefficients.
//
// Same organisation as the order-2 patch kernel (hex_patch.cu), with what order 1 changes:
//   * a dof is ONE double (edges carry one dof), so the units of the dof side are 8 bytes (patch.cpp, unit_bytes = 8)
//     and the only orientation flag is the sign (lidx bit 15);
//   * one thread owns a whole element (12 inputs, 12 results, 12 geometry factors): no exchange buffer between
//     component threads, FE threads per CTA;
//   * the element operator is the generated sparse modal operator gen/hcurl_modal_p1.h (72 + 96 FMA).
// Per patch: pull -> barrier -> operator, issue the x gather of the next patch, results to the shared result buffer ->
// barrier -> per-dof sums, private dofs stored, shared ones RED-added -> barrier.  The blob of patch i+2 is requested
// when patch i retires.
//
// Shared memory:  [16 B: two mbarriers] [2 x blob_cap_bytes] [Xs: xs_cap_bytes] [EY: 12 x (FE+1) doubles]
//
// Env provides:
//   int tid(); int64_t block(), nblocks(); unsigned char *smem(); void sync();
//   void bulk_init(int n);                                               (thread 0, before the first sync: n mbarriers)
//   void bulk_request(int b, void *dst, const void *src, uint32_t bytes) (thread 0: expect-tx + bulk copy on mbarrier b)
//   void bulk_wait(int b, uint32_t parity);                              (every thread)
//   void gather8(double *dst, const double *src); void gather_commit(); void gather_wait();
//   void red_add(double *p, double v);
//   uint32_t ldg(const uint32_t *p); double ldg(const double *p);
//   void dot_finish(double my_sum, const PatchArgs &a, double *scratch)  (all threads; scratch = FE doubles of shared memory)
#pragma once
#include "gen/hcurl_modal_p1.h"

#ifndef HCURL_BODY
#ifdef __CUDACC__
#define HCURL_BODY __device__ __forceinline__
#else
#define HCURL_BODY inline
#endif
#endif

namespace hcurl {

HCURL_BODY double p1_flip(double v, bool neg) { return neg ? -v : v; }

template <class Env, int FE, bool ACC, bool DOT>
HCURL_BODY void hex_patch_p1_body(Env &env, const PatchArgs &a) {
  constexpr int NT = FE, ST = FE + 1, ND = 12, HDR = PatchHost::HDR;
  unsigned char *smem = env.smem();
  unsigned char *blobB = smem + 16;
  double *Xs = reinterpret_cast<double *>(blobB + 2 * (size_t)a.blob_cap_bytes);
  double *EY = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
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
    for (uint32_t u = tid; u < U; u += NT) {
      const uint32_t m = dof[u];
      if (!(m & PATCH_ESS0)) env.gather8(Xs + u, a.x + (m & PATCH_IDX));
      else Xs[u] = 0.0;  // essential dofs enter as zeros
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
    // affine geometry factors of my element (coalesced: [12][FE] per patch)
    const double *g = a.edata + (size_t)k * 12 * FE + el;
    double G[12];
#pragma unroll
    for (int s = 0; s < 12; s++) G[s] = env.ldg(g + s * FE);
    const bool ok = (uint32_t)el < B[13];
    // ---- pull the 12 element inputs through the per-slot table
    double X[ND];
    uint32_t neg = 0;  // bit s: slot s is oriented against its global dof
    if (ok) {
      const uint16_t *lidx = reinterpret_cast<const uint16_t *>(B + B[12]) + el;
#pragma unroll
      for (int s = 0; s < ND; s++) {
        const uint32_t l = lidx[s * FE];
        neg |= (l >> 15) << s;
        X[s] = p1_flip(Xs[l & 0x1fffu], l >> 15);
      }
    }
    env.sync();  // Xs free
    double Y[ND];
    if (ok) {
      hcurl_modal_p1_apply(X, G, G + 6, a.scale, Y);
      if (DOT) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int s = 0; s < ND; s += 2) { d0 += X[s] * Y[s]; d1 += X[s + 1] * Y[s + 1]; }
        dacc += d0 + d1;
      }
    }
    // ---- start the x gather of the next patch: it lands while this patch is scattered
    if (it + 1 < n_it) {
      env.bulk_wait(b ^ 1, ((it + 1) >> 1) & 1);
      gather(b ^ 1);
    }
    if (ok) {
      const int elx = el ^ a.swz_grp[el >> 4];
#pragma unroll
      for (int s = 0; s < ND; s++) EY[s * ST + (elx ^ a.swz_row[s])] = p1_flip(Y[s], (neg >> s) & 1u);
    }
    env.sync();  // element results visible
    // ---- sum the slots of every patch dof; private dofs are plain stores, shared ones RED
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + HDR;
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(B + B[8]);
      const uint32_t *e32 = B + B[9];
      const uint32_t *e64 = B + B[10];  // two words per unit
      const uint32_t *ovf = B + B[11];
      auto e_at = [&](uint32_t off) { return *reinterpret_cast<const double *>(Eb + off); };
      auto emit = [&](const uint32_t m, const double v) {
        if (m & PATCH_ESS0) return;  // essential rows are written by the caller (identity)
        double *dst = a.y + (m & PATCH_IDX);
        if (m & PATCH_SHARED) env.red_add(dst, v);
        else *dst = ACC ? *dst + v : v;
      };
      constexpr int UN = 2;  // dofs in flight per thread: table reads, then result reads, then the stores
      for (uint32_t r0 = tid; r0 < n1; r0 += UN * NT) {
        uint32_t m[UN], o[UN];
        double v[UN];
#pragma unroll
        for (int i = 0; i < UN; i++) {
          const uint32_t r = r0 + i * NT;
          m[i] = r < n1 ? dof[r] : PATCH_ESS0;
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
        double v[UN];
#pragma unroll
        for (int i = 0; i < UN; i++) {
          const uint32_t r = r0 + i * NT;
          m[i] = r < n2 ? dof[r] : PATCH_ESS0;
          o[i] = r < n2 ? e32[r] : 0u;
        }
#pragma unroll
        for (int i = 0; i < UN; i++) v[i] = e_at(o[i] & 0xffffu) + e_at(o[i] >> 16);
#pragma unroll
        for (int i = 0; i < UN; i++) emit(m[i], v[i]);
      }
      dof += n2;
      for (uint32_t r = tid; r < n4; r += NT) {
        const uint32_t m = dof[r];
        const uint32_t ox = e64[2 * r], oy = e64[2 * r + 1];
        double v = (e_at(ox & 0xffffu) + e_at(ox >> 16)) + e_at(oy & 0xffffu);
        if ((oy >> 16) != 0xffffu) v += e_at(oy >> 16);
        emit(m, v);
      }
      dof = B + HDR;
      for (uint32_t t = tid; t < nov; t += NT) {  // slots beyond four (edge valence > 4 inside one patch)
        const uint32_t e = ovf[t], m = dof[e >> 16];
        if (!(m & PATCH_ESS0)) env.red_add(a.y + (m & PATCH_IDX), e_at(e & 0xffffu));
      }
    }
    env.gather_wait();
    env.sync();  // blob buffer b and EY are free; Xs of the next patch complete and visible
    if (tid == 0 && it + 2 < n_it) request(b, q0, q1);
  }
  if (DOT) env.dot_finish(dacc, a, EY);
}

}  // namespace hcurl
