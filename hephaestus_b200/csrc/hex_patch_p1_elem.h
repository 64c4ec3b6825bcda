// Element side of the order-1 Nedelec hex patch kernel: one thread per element (the pipeline around it:
// patch_pipeline.h).
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204), order-1 Nedelec hexes (12 edge dofs), affine
// elements with per-element coefficients.
//
// A dof is ONE double (edges carry one dof): 8-byte units, the only orientation flag is the sign (table bit 15).  The
// element operator is the generated sparse modal operator gen/hcurl_modal_p1.h (scripts/gen_modal_op.py: 72 operator +
// 96 transform FMA).
#pragma once
#include "gen/hcurl_modal_p1.h"
#include "patch_pipeline.h"

namespace hcurl {

template <int FE>
struct P1Elem {
  typedef UnitSingle Unit;
  static constexpr int ROWS = 12;
  double X[12], G[12];
  uint32_t neg;  // bit s: slot s is oriented against its global dof

  template <class Env> HCURL_BODY void geometry(Env &env, const PatchArgs &a, const int64_t k, const int el) {
    const double *g = a.edata + (size_t)k * 12 * FE + el;  // [12][FE] per patch: coalesced
#pragma unroll
    for (int s = 0; s < 12; s++) G[s] = env.ldg(g + s * FE);
  }
  template <class Env> HCURL_BODY void pull(Env &, const PatchArgs &, const double *Xs, const uint16_t *lidx) {
    neg = 0;
#pragma unroll
    for (int s = 0; s < 12; s++) {
      const uint32_t l = lidx[s * FE];
      neg |= (l >> 15) << s;
      const double v = Xs[l & 0x1fffu];
      X[s] = (l >> 15) ? -v : v;
    }
  }
  template <bool DOT, class Env> HCURL_BODY double apply(Env &, const PatchArgs &a) {
    double Y[12];
    hcurl_modal_p1_apply(X, G, G + 6, a.scale, Y);
    double d0 = 0.0, d1 = 0.0;
    if (DOT) {
#pragma unroll
      for (int s = 0; s < 12; s += 2) { d0 += X[s] * Y[s]; d1 += X[s + 1] * Y[s + 1]; }
    }
#pragma unroll
    for (int s = 0; s < 12; s++) X[s] = Y[s];
    return d0 + d1;
  }
  template <class Env> HCURL_BODY void store(Env &, const PatchArgs &a, double *EY, const uint16_t *, const int el) {
    constexpr int ST = FE + 1;
    const int elx = el ^ a.swz_grp[el >> 4];
#pragma unroll
    for (int s = 0; s < 12; s++) EY[s * ST + (elx ^ a.swz_row[s])] = ((neg >> s) & 1u) ? -X[s] : X[s];
  }
};

}  // namespace hcurl
