// Register-blocked fast path of the order-2 Nedelec hex apply (Q = 3 Gauss points / dim).
//
// Same operator as hex_apply.cu (CurlCurlKernel + VectorFEMassKernel,
// /root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24),
// restructured for the FP64 pipe of sm_100a: measured on B200 the generic kernel is bound by
// shared-memory traffic and barriers, and FP64 RED atomics to HBM cap any scatter-by-element
// kernel at ~86 GDOF/s (scripts/microbench.py).  Here
//
//   * a CTA is 3 warps x 32 elements.  In phases 1/3 warp c owns vector component c of its
//     lane's element (18 dofs <-> 27 quadrature values, all contractions in registers with the
//     1D tables as constant-bank operands); in phase 2 warp c owns quadrature plane qz = c
//     (collocation curl, pointwise D, in-plane transposed curl).
//   * the only shared-memory traffic is the exchange between those two decompositions
//     (81 + 45*3 doubles per element, lane-contiguous => conflict-free).
//   * geometry: AFFINE elements use 12 precomputed doubles per element
//     (alpha J^T J/det, beta adj adj^T/det -- constant over the element), TRILINEAR elements
//     recompute J at every quadrature point from the 8 vertices.
//   * maps and element data are stored dof-major ([54][n_pad]) so every global access of a
//     warp is coalesced.
#include "internal.h"

namespace hcurl {

struct FastTables {
  double Bo[3 * 2];  // [q][i] open basis (Gauss-Legendre nodes) at the quadrature points
  double Bc[3 * 3];  // [q][i] closed basis (Gauss-Lobatto nodes)
  double D[3 * 3];   // [q][m] collocation derivative on the quadrature points
  double w[3];
  double Ng[3 * 2];  // trilinear geometry basis at quadrature points
};
__constant__ FastTables cT;

// ---- register tensor contractions (everything unrolled; indices compile-time) ----------
// in: [N0][N1][N2] x-fastest.  Contract dimension DIM (size n) with B[3][n] -> size 3.
template <int N0, int N1, int N2, int DIM>
__device__ __forceinline__ void fwd(const double *in, double *out, const double *B) {
  constexpr int M0 = DIM == 0 ? 3 : N0, M1 = DIM == 1 ? 3 : N1, M2 = DIM == 2 ? 3 : N2;
  constexpr int n = DIM == 0 ? N0 : (DIM == 1 ? N1 : N2);
#pragma unroll
  for (int k = 0; k < M2; k++)
#pragma unroll
    for (int j = 0; j < M1; j++)
#pragma unroll
      for (int i = 0; i < M0; i++) {
        const int q = DIM == 0 ? i : (DIM == 1 ? j : k);
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < n; a++) {
          const int ii = DIM == 0 ? a : i, jj = DIM == 1 ? a : j, kk = DIM == 2 ? a : k;
          const double v = in[ii + N0 * (jj + N1 * kk)];
          s = (a == 0) ? B[q * n + a] * v : fma(B[q * n + a], v, s);
        }
        out[i + M0 * (j + M1 * k)] = s;
      }
}
// transposed: in has size 3 along DIM, out size n along DIM: out[a] = sum_q B[q][a] in[q]
template <int N0, int N1, int N2, int DIM>
__device__ __forceinline__ void bwd(const double *in, double *out, const double *B) {
  constexpr int M0 = DIM == 0 ? 3 : N0, M1 = DIM == 1 ? 3 : N1;
  constexpr int n = DIM == 0 ? N0 : (DIM == 1 ? N1 : N2);
#pragma unroll
  for (int k = 0; k < N2; k++)
#pragma unroll
    for (int j = 0; j < N1; j++)
#pragma unroll
      for (int i = 0; i < N0; i++) {
        const int a = DIM == 0 ? i : (DIM == 1 ? j : k);
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 3; q++) {
          const int ii = DIM == 0 ? q : i, jj = DIM == 1 ? q : j, kk = DIM == 2 ? q : k;
          const double v = in[ii + M0 * (jj + M1 * kk)];
          s = (q == 0) ? B[q * n + a] * v : fma(B[q * n + a], v, s);
        }
        out[i + N0 * (j + N1 * k)] = s;
      }
}

__device__ __forceinline__ double fast_gather(const double *__restrict__ x, uint32_t m) {
  if (m & MAP_ESS) return 0.0;
  const double v = __ldg(x + (m & MAP_IDX));
  return (m & MAP_NEG) ? -v : v;
}

// component C: dofs [n0][n1][n2] with n_C = 2 (open), others 3 (closed)
template <int C>
__device__ __forceinline__ void interp_fwd(const double *X, double *U) {
  constexpr int N0 = C == 0 ? 2 : 3, N1 = C == 1 ? 2 : 3, N2 = C == 2 ? 2 : 3;
  double a[3 * N1 * N2], b[9 * N2];
  fwd<N0, N1, N2, 0>(X, a, C == 0 ? cT.Bo : cT.Bc);
  fwd<3, N1, N2, 1>(a, b, C == 1 ? cT.Bo : cT.Bc);
  fwd<3, 3, N2, 2>(b, U, C == 2 ? cT.Bo : cT.Bc);
}
template <int C>
__device__ __forceinline__ void interp_bwd(const double *Z, double *Y) {
  constexpr int N0 = C == 0 ? 2 : 3, N1 = C == 1 ? 2 : 3, N2 = C == 2 ? 2 : 3;
  double a[3 * N1 * N2], b[9 * N2];
  bwd<3, 3, N2, 2>(Z, b, C == 2 ? cT.Bo : cT.Bc);
  bwd<3, N1, N2, 1>(b, a, C == 1 ? cT.Bo : cT.Bc);
  bwd<N0, N1, N2, 0>(a, Y, C == 0 ? cT.Bo : cT.Bc);
}

enum { GEOM_AFFINE = 0, GEOM_TRILINEAR = 1 };
static constexpr int FE = 32;                    // elements per CTA
static constexpr int SU = 81 * FE;               // doubles: U exchange
static constexpr int SZ = 135 * FE;              // doubles: Z~ (81) + W_x, W_y (54) exchange
static constexpr size_t FAST_SMEM = sizeof(double) * (SU + SZ);

template <int C>
__device__ __forceinline__ void phase1(const int64_t e, const int64_t n_pad, const int lane,
                                       const uint32_t *__restrict__ mapT, const double *__restrict__ x, double *su) {
  double X[18], U[27];
#pragma unroll
  for (int t = 0; t < 18; t++) X[t] = fast_gather(x, __ldg(mapT + (int64_t)(C * 18 + t) * n_pad + e));
  interp_fwd<C>(X, U);
#pragma unroll
  for (int q = 0; q < 27; q++) su[(C * 27 + q) * FE + lane] = U[q];
}

template <int C>
__device__ __forceinline__ void phase3(const int64_t e, const int64_t n_pad, const int lane,
                                       const uint32_t *__restrict__ mapT, const double *sz, double *__restrict__ y) {
  double Z[27], Y[18];
#pragma unroll
  for (int q = 0; q < 27; q++) Z[q] = sz[(C * 27 + q) * FE + lane];
  if (C < 2) {
    // Z_x += Dz^T W_y ; Z_y -= Dz^T W_x     (W_x at 81.., W_y at 108..)
    const int off = C == 0 ? 108 : 81;
#pragma unroll
    for (int r = 0; r < 9; r++) {
      double wv[3];
#pragma unroll
      for (int m = 0; m < 3; m++) wv[m] = sz[(off + r + 9 * m) * FE + lane];
#pragma unroll
      for (int a = 0; a < 3; a++) {
        double s = Z[r + 9 * a];
#pragma unroll
        for (int m = 0; m < 3; m++) s = fma(C == 0 ? cT.D[m * 3 + a] : -cT.D[m * 3 + a], wv[m], s);
        Z[r + 9 * a] = s;
      }
    }
  }
  interp_bwd<C>(Z, Y);
#pragma unroll
  for (int t = 0; t < 18; t++) {
    const uint32_t m = __ldg(mapT + (int64_t)(C * 18 + t) * n_pad + e);
    if (!(m & MAP_ESS)) atomicAdd(y + (m & MAP_IDX), (m & MAP_NEG) ? -Y[t] : Y[t]);
  }
}

// plane QZ of the element: collocation curl, pointwise D, in-plane transposed curl
template <int GEOM>
__device__ __forceinline__ void phase2(const int qz, const int64_t e, const int64_t n_pad, const int lane,
                                       const double *__restrict__ edata, const double *__restrict__ xyzT,
                                       const double *__restrict__ alpha, const double *__restrict__ beta,
                                       const double scale, const double *su, double *sz) {
  // plane values of U (all three components) and the z-lines of U_x, U_y
  double Ux[9], Uy[9], Uz[9], cx[9], cy[9], cz[9];
#pragma unroll
  for (int r = 0; r < 9; r++) {
    double lx[3], ly[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
      lx[m] = su[(0 * 27 + r + 9 * m) * FE + lane];
      ly[m] = su[(1 * 27 + r + 9 * m) * FE + lane];
    }
    // dz U_x, dz U_y at (r, qz); qz is warp-uniform (0..2) -> select the D row without divergence
    const double d0 = qz == 0 ? cT.D[0] : (qz == 1 ? cT.D[3] : cT.D[6]);
    const double d1 = qz == 0 ? cT.D[1] : (qz == 1 ? cT.D[4] : cT.D[7]);
    const double d2 = qz == 0 ? cT.D[2] : (qz == 1 ? cT.D[5] : cT.D[8]);
    const double dzUx = fma(d2, lx[2], fma(d1, lx[1], d0 * lx[0]));
    const double dzUy = fma(d2, ly[2], fma(d1, ly[1], d0 * ly[0]));
    Ux[r] = qz == 0 ? lx[0] : (qz == 1 ? lx[1] : lx[2]);
    Uy[r] = qz == 0 ? ly[0] : (qz == 1 ? ly[1] : ly[2]);
    Uz[r] = su[(2 * 27 + r + 9 * qz) * FE + lane];
    cx[r] = -dzUy;  // curl_x = dyUz - dzUy
    cy[r] = dzUx;   // curl_y = dzUx - dxUz
  }
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double sx = cx[i + 3 * j], sy = cy[i + 3 * j], szz = 0.0;
#pragma unroll
      for (int m = 0; m < 3; m++) {
        sx = fma(cT.D[j * 3 + m], Uz[i + 3 * m], sx);    // + dy U_z
        sy = fma(-cT.D[i * 3 + m], Uz[m + 3 * j], sy);   // - dx U_z
        szz = fma(cT.D[i * 3 + m], Uy[m + 3 * j], szz);  // + dx U_y
        szz = fma(-cT.D[j * 3 + m], Ux[i + 3 * m], szz); // - dy U_x
      }
      cx[i + 3 * j] = sx; cy[i + 3 * j] = sy; cz[i + 3 * j] = szz;
    }
  // pointwise: W = w a J^T J/det curl,  V = w b adj adj^T/det U   (written over cx.. and Ux..)
  const double wz = qz == 0 ? cT.w[0] : (qz == 1 ? cT.w[1] : cT.w[2]);
  if (GEOM == GEOM_AFFINE) {
    double S[6], M[6];
#pragma unroll
    for (int s = 0; s < 6; s++) {
      S[s] = scale * __ldg(edata + (int64_t)s * n_pad + e);
      M[s] = scale * __ldg(edata + (int64_t)(6 + s) * n_pad + e);
    }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int r = i + 3 * j;
        const double w = cT.w[i] * cT.w[j] * wz;
        const double a = w * cx[r], b = w * cy[r], c = w * cz[r];
        cx[r] = fma(S[2], c, fma(S[1], b, S[0] * a));
        cy[r] = fma(S[4], c, fma(S[3], b, S[1] * a));
        cz[r] = fma(S[5], c, fma(S[4], b, S[2] * a));
        const double u = w * Ux[r], v = w * Uy[r], t = w * Uz[r];
        Ux[r] = fma(M[2], t, fma(M[1], v, M[0] * u));
        Uy[r] = fma(M[4], t, fma(M[3], v, M[1] * u));
        Uz[r] = fma(M[5], t, fma(M[4], v, M[2] * u));
      }
  } else {
    // trilinear: vertices in tensor order v = a + 2b + 4c; J columns recomputed per point
    double Xv[8][3];
#pragma unroll
    for (int v = 0; v < 8; v++)
#pragma unroll
      for (int d = 0; d < 3; d++) Xv[v][d] = __ldg(xyzT + (int64_t)(v * 3 + d) * n_pad + e);
    const double al = scale * __ldg(alpha + e), be = scale * __ldg(beta + e);
    const double nz0 = qz == 0 ? cT.Ng[0] : (qz == 1 ? cT.Ng[2] : cT.Ng[4]);
    const double nz1 = qz == 0 ? cT.Ng[1] : (qz == 1 ? cT.Ng[3] : cT.Ng[5]);
    // collapse z: P[a][b] = nz0 X[a,b,0] + nz1 X[a,b,1];  Pz[a][b] = X[a,b,1] - X[a,b,0]
    double P[4][3], Pz[4][3];
#pragma unroll
    for (int ab = 0; ab < 4; ab++)
#pragma unroll
      for (int d = 0; d < 3; d++) {
        P[ab][d] = fma(nz1, Xv[ab + 4][d], nz0 * Xv[ab][d]);
        Pz[ab][d] = Xv[ab + 4][d] - Xv[ab][d];
      }
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const int r = i + 3 * j;
        const double nx0 = cT.Ng[i * 2], nx1 = cT.Ng[i * 2 + 1], ny0 = cT.Ng[j * 2], ny1 = cT.Ng[j * 2 + 1];
        double c0[3], c1[3], c2[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          c0[d] = fma(ny1, P[3][d] - P[2][d], ny0 * (P[1][d] - P[0][d]));                  // d/dxi
          c1[d] = fma(nx1, P[3][d] - P[1][d], nx0 * (P[2][d] - P[0][d]));                  // d/deta
          c2[d] = fma(ny1, fma(nx1, Pz[3][d], nx0 * Pz[2][d]), ny0 * fma(nx1, Pz[1][d], nx0 * Pz[0][d]));  // d/dzeta
        }
        const double a0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
        const double a1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
        const double a2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
        const double det = c0[0] * a0[0] + c0[1] * a0[1] + c0[2] * a0[2];
        const double w = cT.w[i] * cT.w[j] * wz / det;
        const double ca = al * w, cb = be * w;
        double t[3], s[3];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          t[d] = ca * (c0[d] * cx[r] + c1[d] * cy[r] + c2[d] * cz[r]);
          s[d] = cb * (a0[d] * Ux[r] + a1[d] * Uy[r] + a2[d] * Uz[r]);
        }
        cx[r] = c0[0] * t[0] + c0[1] * t[1] + c0[2] * t[2];
        cy[r] = c1[0] * t[0] + c1[1] * t[1] + c1[2] * t[2];
        cz[r] = c2[0] * t[0] + c2[1] * t[1] + c2[2] * t[2];
        Ux[r] = a0[0] * s[0] + a0[1] * s[1] + a0[2] * s[2];
        Uy[r] = a1[0] * s[0] + a1[1] * s[1] + a1[2] * s[2];
        Uz[r] = a2[0] * s[0] + a2[1] * s[1] + a2[2] * s[2];
      }
  }
  // in-plane transposed derivative:  Z~_x = V_x - Dy^T W_z ; Z~_y = V_y + Dx^T W_z ;
  //                                  Z~_z = V_z + Dy^T W_x - Dx^T W_y
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double zx = Ux[i + 3 * j], zy = Uy[i + 3 * j], zz = Uz[i + 3 * j];
#pragma unroll
      for (int m = 0; m < 3; m++) {
        zx = fma(-cT.D[m * 3 + j], cz[i + 3 * m], zx);
        zy = fma(cT.D[m * 3 + i], cz[m + 3 * j], zy);
        zz = fma(cT.D[m * 3 + j], cx[i + 3 * m], zz);
        zz = fma(-cT.D[m * 3 + i], cy[m + 3 * j], zz);
      }
      const int q = i + 3 * j + 9 * qz;
      sz[(0 * 27 + q) * FE + lane] = zx;
      sz[(1 * 27 + q) * FE + lane] = zy;
      sz[(2 * 27 + q) * FE + lane] = zz;
      sz[(81 + q) * FE + lane] = cx[i + 3 * j];   // W_x
      sz[(108 + q) * FE + lane] = cy[i + 3 * j];  // W_y
    }
}

template <int GEOM>
__global__ void __launch_bounds__(96, 4)
hex_fast_p2_kernel(const int64_t n_elem, const int64_t n_pad, const uint32_t *__restrict__ mapT,
                   const double *__restrict__ edata, const double *__restrict__ xyzT, const double *__restrict__ alpha,
                   const double *__restrict__ beta, const double *__restrict__ x, double *__restrict__ y,
                   const double scale) {
  extern __shared__ double fsm[];
  double *su = fsm, *sz = fsm + SU;
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * FE + lane;
  const bool ok = e < n_elem;
  if (ok) {
    if (c == 0) phase1<0>(e, n_pad, lane, mapT, x, su);
    else if (c == 1) phase1<1>(e, n_pad, lane, mapT, x, su);
    else phase1<2>(e, n_pad, lane, mapT, x, su);
  }
  __syncthreads();
  if (ok) phase2<GEOM>(c, e, n_pad, lane, edata, xyzT, alpha, beta, scale, su, sz);
  __syncthreads();
  if (ok) {
    if (c == 0) phase3<0>(e, n_pad, lane, mapT, sz, y);
    else if (c == 1) phase3<1>(e, n_pad, lane, mapT, sz, y);
    else phase3<2>(e, n_pad, lane, mapT, sz, y);
  }
}

// ---- setup kernels ------------------------------------------------------------------------
__global__ void fast_transpose_map_kernel(const int64_t n_elem, const int nd, const int64_t n_pad,
                                          const uint32_t *__restrict__ in, uint32_t *__restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * nd) return;
  out[(gid % nd) * n_pad + gid / nd] = in[gid];
}
__global__ void fast_transpose_xyz_kernel(const int64_t n_elem, const int64_t n_pad, const double *__restrict__ in,
                                          double *__restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * 24) return;
  out[(gid % 24) * n_pad + gid / 24] = in[gid];
}
// per element: affine test (max deviation of the 8th..: all vertices vs the parallelepiped spanned at vertex 0)
__global__ void fast_affine_kernel(const int64_t n_elem, const double *__restrict__ xyz, const double tol, int *n_bad) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const double *X = xyz + e * 24;  // tensor order
  double h2 = 0.0, dev2 = 0.0;
  double ex[3], ey[3], ez[3];
  for (int d = 0; d < 3; d++) {
    ex[d] = X[3 + d] - X[d]; ey[d] = X[6 + d] - X[d]; ez[d] = X[12 + d] - X[d];
    h2 = fmax(h2, fmax(ex[d] * ex[d], fmax(ey[d] * ey[d], ez[d] * ez[d])));
  }
  for (int v = 0; v < 8; v++)
    for (int d = 0; d < 3; d++) {
      const double pred = X[d] + (v & 1) * ex[d] + ((v >> 1) & 1) * ey[d] + ((v >> 2) & 1) * ez[d];
      const double dv = X[3 * v + d] - pred;
      dev2 = fmax(dev2, dv * dv);
    }
  if (dev2 > tol * tol * h2) atomicAdd(n_bad, 1);
}
// per element: S = alpha J^T J/det (6), M = beta adj adj^T/det (6), J from the parallelepiped edges
__global__ void fast_edata_kernel(const int64_t n_elem, const int64_t n_pad, const double *__restrict__ xyz,
                                  const double *__restrict__ alpha, const double *__restrict__ beta,
                                  double *__restrict__ edata) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  const double *X = xyz + e * 24;
  double c0[3], c1[3], c2[3];
  for (int d = 0; d < 3; d++) { c0[d] = X[3 + d] - X[d]; c1[d] = X[6 + d] - X[d]; c2[d] = X[12 + d] - X[d]; }
  const double a0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
  const double a1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  const double a2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
  const double det = c0[0] * a0[0] + c0[1] * a0[1] + c0[2] * a0[2];
  const double al = alpha ? alpha[e] / det : 0.0, be = beta ? beta[e] / det : 0.0;
  const double *c[3] = {c0, c1, c2};
  const double *a[3] = {a0, a1, a2};
  int s = 0;
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++, s++) {
      edata[(int64_t)s * n_pad + e] = al * (c[i][0] * c[j][0] + c[i][1] * c[j][1] + c[i][2] * c[j][2]);
      edata[(int64_t)(6 + s) * n_pad + e] = be * (a[i][0] * a[j][0] + a[i][1] * a[j][1] + a[i][2] * a[j][2]);
    }
}

// affine classification alone (no dof-major copies): used by hcurl_op_create for HCURL_QUAD_AUTO
int hex_all_affine(hcurl_space *sp, bool *all_affine) {
  if (sp->fast_ready) { *all_affine = sp->all_affine; return 0; }
  hcurl_ctx *ctx = sp->ctx;
  int *d_bad = nullptr, h_bad = 0;
  HC_CUDA(cudaMalloc(&d_bad, sizeof(int)));
  HC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  fast_affine_kernel<<<(unsigned)((sp->n_elem + 255) / 256), 256, 0, ctx->stream>>>(sp->n_elem, sp->d_xyz, 1e-13, d_bad);
  HC_CUDA(cudaGetLastError());
  HC_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_bad);
  *all_affine = (h_bad == 0);
  return 0;
}

static bool g_tables_set = false;

int hex_fast_prepare_space(hcurl_space *sp) {
  // called lazily: dof-major copies of the map / vertices and the affine classification
  hcurl_ctx *ctx = sp->ctx;
  if (sp->fast_ready) return 0;
  if (!g_tables_set) {
    HexTables T;
    build_hex_tables(2, 3, 1, T);
    FastTables F;
    for (int q = 0; q < 3; q++) {
      for (int i = 0; i < 2; i++) F.Bo[q * 2 + i] = T.Bo[q * MAXP1 + i];
      for (int i = 0; i < 3; i++) { F.Bc[q * 3 + i] = T.Bc[q * MAXP1 + i]; F.D[q * 3 + i] = T.Dq[q * MAXQ + i]; }
      F.w[q] = T.wq[q];
      F.Ng[q * 2] = T.Ng[q * 3]; F.Ng[q * 2 + 1] = T.Ng[q * 3 + 1];
    }
    HC_CUDA(cudaMemcpyToSymbol(cT, &F, sizeof(F)));
    g_tables_set = true;
  }
  const int64_t n = sp->n_elem, n_pad = sp->n_elem_pad;
  HC_CUDA(cudaMalloc(&sp->d_mapT, sizeof(uint32_t) * sp->nd * n_pad));
  HC_CUDA(cudaMemsetAsync(sp->d_mapT, 0, sizeof(uint32_t) * sp->nd * n_pad, ctx->stream));
  fast_transpose_map_kernel<<<(unsigned)((n * sp->nd + 255) / 256), 256, 0, ctx->stream>>>(n, sp->nd, n_pad, sp->d_map, sp->d_mapT);
  int *d_bad = nullptr, h_bad = 0;
  HC_CUDA(cudaMalloc(&d_bad, sizeof(int)));
  HC_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
  fast_affine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, sp->d_xyz, 1e-13, d_bad);
  HC_CUDA(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_bad);
  sp->all_affine = (h_bad == 0);
  if (!sp->all_affine) {
    HC_CUDA(cudaMalloc(&sp->d_xyzT, sizeof(double) * 24 * n_pad));
    HC_CUDA(cudaMemsetAsync(sp->d_xyzT, 0, sizeof(double) * 24 * n_pad, ctx->stream));
    fast_transpose_xyz_kernel<<<(unsigned)((n * 24 + 255) / 256), 256, 0, ctx->stream>>>(n, n_pad, sp->d_xyz, sp->d_xyzT);
  }
  HC_CUDA(cudaGetLastError());
  sp->fast_ready = true;
  return 0;
}

bool hex_fast_eligible(const hcurl_op *op) {
  const hcurl_space *sp = op->sp;
  return sp->geom == HCURL_GEOM_HEX && sp->order == 2 && sp->geom_order == 1 && op->single_rule && op->qc == 3 &&
         op->coef_layout == HCURL_COEF_PER_ELEM && op->has_alpha && op->has_beta && !op->force_generic && op->kernel_sel != 3;
}

int hex_fast_update_coefficients(hcurl_op *op) {
  hcurl_space *sp = op->sp;
  if (hex_fast_prepare_space(sp)) return 1;
  if (!sp->all_affine) return 0;
  if (!op->d_edata) HC_CUDA(cudaMalloc(&op->d_edata, sizeof(double) * 12 * sp->n_elem_pad));
  fast_edata_kernel<<<(unsigned)((sp->n_elem + 255) / 256), 256, 0, op->ctx->stream>>>(sp->n_elem, sp->n_elem_pad, sp->d_xyz,
                                                                                      op->d_alpha, op->d_beta, op->d_edata);
  HC_CUDA(cudaGetLastError());
  op->edata_valid = true;
  return 0;
}

int hex_fast_update_map(hcurl_op *op) {
  // dof-major copy of the essential-flagged map
  hcurl_space *sp = op->sp;
  if (hex_fast_prepare_space(sp)) return 1;
  if (!op->d_map_ess) return 0;
  const int64_t n = sp->n_elem, n_pad = sp->n_elem_pad;
  if (!op->d_mapT_ess) HC_CUDA(cudaMalloc(&op->d_mapT_ess, sizeof(uint32_t) * sp->nd * n_pad));
  HC_CUDA(cudaMemsetAsync(op->d_mapT_ess, 0, sizeof(uint32_t) * sp->nd * n_pad, op->ctx->stream));
  fast_transpose_map_kernel<<<(unsigned)((n * sp->nd + 255) / 256), 256, 0, op->ctx->stream>>>(n, sp->nd, n_pad, op->d_map_ess,
                                                                                             op->d_mapT_ess);
  HC_CUDA(cudaGetLastError());
  return 0;
}

int hex_fast_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale) {
  hcurl_space *sp = op->sp;
  if (hex_fast_prepare_space(sp)) return 1;
  if (sp->all_affine && !op->edata_valid && hex_fast_update_coefficients(op)) return 1;
  if (constrained && op->d_map_ess && !op->d_mapT_ess && hex_fast_update_map(op)) return 1;
  const uint32_t *mapT = (constrained && op->d_mapT_ess) ? op->d_mapT_ess : sp->d_mapT;
  const unsigned grid = (unsigned)((sp->n_elem + FE - 1) / FE);
  static bool attr = false;
  if (!attr) {
    HC_CUDA(cudaFuncSetAttribute(hex_fast_p2_kernel<GEOM_AFFINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM));
    HC_CUDA(cudaFuncSetAttribute(hex_fast_p2_kernel<GEOM_TRILINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM));
    attr = true;
  }
  if (sp->all_affine)
    hex_fast_p2_kernel<GEOM_AFFINE><<<grid, 96, FAST_SMEM, op->ctx->stream>>>(sp->n_elem, sp->n_elem_pad, mapT, op->d_edata, nullptr,
                                                                             nullptr, nullptr, d_x, d_y, scale);
  else
    hex_fast_p2_kernel<GEOM_TRILINEAR><<<grid, 96, FAST_SMEM, op->ctx->stream>>>(sp->n_elem, sp->n_elem_pad, mapT, nullptr, sp->d_xyzT,
                                                                                op->d_alpha, op->d_beta, d_x, d_y, scale);
  HC_CUDA(cudaGetLastError());
  op->ctx->launches++;
  return 0;
}

}  // namespace hcurl
