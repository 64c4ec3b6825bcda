// Discrete gradient (H1 -> ND) and discrete curl (ND -> RT) on hexes, matrix-free: the per-thread program, shared by the
// CUDA kernel (interp.cu) and the host check (tests/emu/interp_emu.cpp).
//
// Reference: CurlAuxSolver builds mfem::ParDiscreteLinearOperator + CurlInterpolator and applies it to L-vectors every
// step (/root/reference/src/auxsolvers/curl_aux.cpp:13-30); ScalarPotentialSource::BuildGrad does the same with
// GradientInterpolator (/root/reference/src/sources/scalar_potential_source.cpp:89-97, applied at :130).
//
// MFEM's H1 / ND / RT hex bases are tensor products of Lagrange bases on the closed (Gauss-Lobatto, p+1) and open
// (Gauss-Legendre, p) points and their dof functionals are point values of the pulled-back field, so both
// interpolators are geometry-free and act along ONE direction with the same 1-D matrix
//     D[i][j] = l_j^closed'(x_i^open)         (p x (p+1)):
//   grad:  (G phi)_c (i_c open, others closed)  = sum_j D[i_c][j] phi(.. j ..)
//   curl:  (C u)_c   (i_c closed, i_a, i_b open) = sum_j D[i_a][j] u_b(.. a: j ..) - sum_j D[i_b][j] u_a(.. b: j ..),
//          (a, b) = (c+1, c+2) cyclic.
// Tets (order 1 only: the TEAM7 configuration, ND1 -> RT0 and H1 P1 -> ND1): MFEM's reference tet has the Whitney edge
// functions (dof = u(midpoint) . (x_b - x_a)) and RT0 face functions with dof = v(centroid) . nk, nk = 2 x area-weighted
// outward normal [MFEM-upstream: ND_TetrahedronElement::tk, RT_TetrahedronElement::nk = {1,1,1, -1,0,0, 0,-1,0, 0,0,-1}],
// so the gradient matrix is the edge-vertex incidence (+-1) and the curl matrix the face-edge incidence times 2.
// Local layouts, hexes (all x fastest, component-major): H1 (p+1)^3; ND component c: p points along c, p+1 along the others;
// RT component c: p+1 points along c, p along the others.  The element -> dof maps arrive in THIS order (the adaptor
// converts with MFEM's own TensorBasisElement::GetDofMap()), encoded as index | bit 31 sign flip; bit 30 of an OUTPUT
// entry = "another element writes this dof" (a dof shared by several elements is written once, by its first element:
// all of them compute the same value, so the result does not depend on the choice, and the apply is deterministic).
#pragma once
#include <stdint.h>

#ifndef HCURL_HD
#ifdef __CUDACC__
#define HCURL_HD __host__ __device__ __forceinline__
#else
#define HCURL_HD inline
#endif
#endif

namespace hcurl {

static constexpr uint32_t INTERP_IDX = 0x3fffffffu, INTERP_SKIP = 0x40000000u, INTERP_NEG = 0x80000000u;

struct InterpArgs {
  int kind;  // HCURL_INTERP_GRAD (0) or HCURL_INTERP_CURL (1)
  int geom;  // HCURL_GEOM_HEX (tensor path) or HCURL_GEOM_TET (order 1: the dense local matrix Mloc)
  int p;
  int64_t n_elem;
  int nd_in, nd_out;
  const uint32_t *in_map;   // [n_elem][nd_in]
  const uint32_t *out_map;  // [n_elem][nd_out]
  const double *x;
  double *y;
  double D[4 * 5];  // hexes: [i * (p+1) + j]
  double Mloc[24];  // tets, order 1: [t_out * nd_in + j] (gradient 6 x 4, curl 4 x 6), entries 0, +-1 (gradient), 0, +-2 (curl)
};

HCURL_HD double interp_in(const InterpArgs &a, const uint32_t *im, int t) {
  const uint32_t m = im[t];
  const double v = a.x[m & INTERP_IDX];
  return (m & INTERP_NEG) ? -v : v;
}

// tensor index of node (i[0], i[1], i[2]) of component c of a vector space whose component c has `nc_c` points along c
// and `no` points along the other two directions
HCURL_HD int interp_vec_index(int c, const int *i, int n_along, int n_other) {
  const int n0 = c == 0 ? n_along : n_other, n1 = c == 1 ? n_along : n_other;
  const int per_comp = n_along * n_other * n_other;
  return c * per_comp + i[0] + n0 * (i[1] + n1 * i[2]);
}

// one output dof of one element: gid = e * nd_out + t
HCURL_HD void interp_thread(const InterpArgs &a, int64_t gid) {
  const int p = a.p, p1 = p + 1;
  const int64_t e = gid / a.nd_out;
  const int t = (int)(gid - e * a.nd_out);
  const uint32_t om = a.out_map[gid];
  if (om & INTERP_SKIP) return;
  const uint32_t *im = a.in_map + e * a.nd_in;
  if (a.geom != 0) {  // tets: lowest order, MFEM's local edge / face order on both sides
    double v = 0.0;
    for (int j = 0; j < a.nd_in; j++) {
      const double m = a.Mloc[t * a.nd_in + j];
      if (m != 0.0) v += m * interp_in(a, im, j);
    }
    a.y[om & INTERP_IDX] = (om & INTERP_NEG) ? -v : v;
    return;
  }
  const int per_comp = a.nd_out / 3;
  const int c = t / per_comp;
  int r = t - c * per_comp;
  int i[3];
  double v = 0.0;
  if (a.kind == 0) {
    // ND component c: open along c
    const int n0 = c == 0 ? p : p1, n1 = c == 1 ? p : p1;
    i[0] = r % n0; r /= n0; i[1] = r % n1; i[2] = r / n1;
    const int ic = i[c];
    const int stride = c == 0 ? 1 : (c == 1 ? p1 : p1 * p1);
    const int base = i[0] + p1 * (i[1] + p1 * i[2]) - ic * stride;
    for (int j = 0; j < p1; j++) v += a.D[ic * p1 + j] * interp_in(a, im, base + j * stride);
  } else {
    // RT component c: closed along c, open along a = c+1 and b = c+2
    const int n0 = c == 0 ? p1 : p, n1 = c == 1 ? p1 : p;
    i[0] = r % n0; r /= n0; i[1] = r % n1; i[2] = r / n1;
    const int ca = (c + 1) % 3, cb = (c + 2) % 3;
    const int ia = i[ca], ib = i[cb];
    int q[3] = {i[0], i[1], i[2]};
    for (int j = 0; j < p1; j++) {  // d/da of u_b: u_b is open along b, closed along a and c
      q[ca] = j;
      v += a.D[ia * p1 + j] * interp_in(a, im, interp_vec_index(cb, q, p, p1));
    }
    q[ca] = ia;
    for (int j = 0; j < p1; j++) {  // d/db of u_a
      q[cb] = j;
      v -= a.D[ib * p1 + j] * interp_in(a, im, interp_vec_index(ca, q, p, p1));
    }
  }
  a.y[om & INTERP_IDX] = (om & INTERP_NEG) ? -v : v;
}

}  // namespace hcurl
