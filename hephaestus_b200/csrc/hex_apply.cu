// Sum-factorised matrix-free apply of  K(alpha) + M(beta)  on Nedelec hexahedra.
//
// Replaces the element-matrix assembly + CSR SpMV the reference runs for
// CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24,
//  /root/reference/src/kernels/vector_fe_mass_kernel.cpp:21-24,
//  /root/reference/src/equation_systems/equation_system.cpp:200-204, :345).
//
// Generic kernel ("G1"): one CTA handles E elements, Q*Q threads per element.
//   gather x_e (signed, essential dofs read as 0)
//   -> per component: contract x, y through shared memory, z into registers
//      (thread (qx,qy) owns the z-line of quadrature values)
//   -> reference curl by collocation derivative on the quadrature points
//   -> pointwise: J from the geometry nodes (recomputed, not stored),
//      W = a w/det J^T J curl,  V = b w/det adj adj^T u
//   -> transposed derivative, transposed contractions, signed atomic scatter.
// FP64 throughout.  Algorithmic bytes per element: 4 nd (map) + 8*3*ngn
// (geometry) + 16 (coefficients) + gather/scatter of x and y.
#include <cstdio>

#include "internal.h"

namespace hcurl {

template <int P, int Q, int G>
struct HexCfg {
  static constexpr int P1 = P + 1;
  static constexpr int NC = P * P1 * P1;
  static constexpr int ND = 3 * NC;
  static constexpr int G1 = G + 1;
  static constexpr int NG = G1 * G1 * G1;
  static constexpr int T1 = Q * P1 * P1;
  static constexpr int T2 = Q * Q * P1;
  static constexpr int Q3 = Q * Q * Q;
  static constexpr int ELEM = ND + 3 * T1 + 3 * T2 + 3 * Q3 + 3 * NG;  // doubles per element
  static constexpr int TAB = Q * P + 2 * Q * P1 + Q * Q + 2 * Q * G1 + Q;
  static constexpr int E = (128 / (Q * Q)) > 0 ? (128 / (Q * Q)) : 1;
  static constexpr int THREADS = E * Q * Q;
  static constexpr size_t SMEM = sizeof(double) * (TAB + (size_t)E * ELEM);
};

enum { TERM_CURL = 1, TERM_MASS = 2, TERM_BOTH = 3 };

__device__ __forceinline__ double gather_val(const double *__restrict__ x, uint32_t m) {
  if (m & MAP_ESS) return 0.0;
  double v = __ldg(x + (m & MAP_IDX));
  return (m & MAP_NEG) ? -v : v;
}

#ifdef HC_P
template <int P, int Q, int G, int TERMS, bool QPT>
__global__ void __launch_bounds__(HexCfg<P, Q, G>::THREADS)
hex_apply_kernel(const HexTables T, const int64_t n_elem, const uint32_t *__restrict__ map,
                 const double *__restrict__ xyz, const double *__restrict__ alpha,
                 const double *__restrict__ beta, const double *__restrict__ x, double *__restrict__ y,
                 const double scale) {
  using C = HexCfg<P, Q, G>;
  constexpr int P1 = C::P1, NC = C::NC, ND = C::ND, G1 = C::G1, NG = C::NG, T1 = C::T1, T2 = C::T2, Q3 = C::Q3;
  extern __shared__ double sm[];
  double *sBo = sm;               // [Q][P]
  double *sBc = sBo + Q * P;      // [Q][P1]
  double *sGc = sBc + Q * P1;     // [Q][P1] (unused here; kept for the diagonal kernel layout)
  double *sD = sGc + Q * P1;      // [Q][Q]
  double *sNg = sD + Q * Q;       // [Q][G1]
  double *sdNg = sNg + Q * G1;    // [Q][G1]
  double *sw = sdNg + Q * G1;     // [Q]
  double *selem = sw + Q;

  const int tx = threadIdx.x, ty = threadIdx.y;
  const int tid = ty * (Q * Q) + tx;
  const int qx = tx % Q, qy = tx / Q;
  const int64_t e0 = (int64_t)blockIdx.x * C::E;
  const int ne_here = (int)min((int64_t)C::E, n_elem - e0);

  for (int i = tid; i < Q * P; i += C::THREADS) sBo[i] = T.Bo[(i / P) * MAXP1 + (i % P)];
  for (int i = tid; i < Q * P1; i += C::THREADS) {
    sBc[i] = T.Bc[(i / P1) * MAXP1 + (i % P1)];
    sGc[i] = T.Gc[(i / P1) * MAXP1 + (i % P1)];
  }
  for (int i = tid; i < Q * Q; i += C::THREADS) sD[i] = T.Dq[(i / Q) * MAXQ + (i % Q)];
  for (int i = tid; i < Q * G1; i += C::THREADS) {
    sNg[i] = T.Ng[(i / G1) * 3 + (i % G1)];
    sdNg[i] = T.dNg[(i / G1) * 3 + (i % G1)];
  }
  if (tid < Q) sw[tid] = T.wq[tid];

  // CTA-wide coalesced gather of x_e and geometry
  for (int i = tid; i < ne_here * ND; i += C::THREADS) {
    uint32_t m = __ldg(map + e0 * ND + i);
    selem[(i / ND) * C::ELEM + (i % ND)] = gather_val(x, m);
  }
  for (int i = tid; i < ne_here * NG * 3; i += C::THREADS)
    selem[(i / (NG * 3)) * C::ELEM + ND + 3 * T1 + 3 * T2 + 3 * Q3 + (i % (NG * 3))] = __ldg(xyz + e0 * NG * 3 + i);
  __syncthreads();

  const bool active = ty < ne_here;
  double *xe = selem + ty * C::ELEM;
  double *t1 = xe + ND;
  double *t2 = t1 + 3 * T1;
  double *su = t2 + 3 * T2;
  double *sg = su + 3 * Q3;
  const int64_t e = e0 + ty;

  // ---- forward: contract x --------------------------------------------------
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n0 = (c == 0) ? P : P1, n1 = (c == 1) ? P : P1, n2 = (c == 2) ? P : P1;
      const double *B0 = (c == 0) ? sBo : sBc;
      if (tx < n1 * n2) {
        double in[P1];
#pragma unroll
        for (int i = 0; i < P1; i++)
          if (i < n0) in[i] = xe[c * NC + i + n0 * tx];
#pragma unroll
        for (int q = 0; q < Q; q++) {
          double s = 0.0;
#pragma unroll
          for (int i = 0; i < P1; i++)
            if (i < n0) s += B0[q * n0 + i] * in[i];
          t1[c * T1 + q + Q * tx] = s;
        }
      }
    }
  }
  __syncthreads();
  // ---- forward: contract y --------------------------------------------------
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n1 = (c == 1) ? P : P1, n2 = (c == 2) ? P : P1;
      const double *B1 = (c == 1) ? sBo : sBc;
      const int q = tx % Q, k = tx / Q;
      if (k < n2) {
        double in[P1];
#pragma unroll
        for (int j = 0; j < P1; j++)
          if (j < n1) in[j] = t1[c * T1 + q + Q * (j + n1 * k)];
#pragma unroll
        for (int b = 0; b < Q; b++) {
          double s = 0.0;
#pragma unroll
          for (int j = 0; j < P1; j++)
            if (j < n1) s += B1[b * n1 + j] * in[j];
          t2[c * T2 + q + Q * (b + Q * k)] = s;
        }
      }
    }
  }
  __syncthreads();
  // ---- forward: contract z into registers ---------------------------------------
  double U[3][Q];
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n2 = (c == 2) ? P : P1;
      const double *B2 = (c == 2) ? sBo : sBc;
      double in[P1];
#pragma unroll
      for (int k = 0; k < P1; k++)
        if (k < n2) in[k] = t2[c * T2 + qx + Q * (qy + Q * k)];
#pragma unroll
      for (int a = 0; a < Q; a++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < P1; k++)
          if (k < n2) s += B2[a * n2 + k] * in[k];
        U[c][a] = s;
      }
    }
    if (TERMS & TERM_CURL) {
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int a = 0; a < Q; a++) su[c * Q3 + qx + Q * (qy + Q * a)] = U[c][a];
    }
  }
  if (TERMS & TERM_CURL) __syncthreads();

  // ---- pointwise ---------------------------------------------------------------
  double V[3][Q], W[3][Q];
  if (active) {
    // geometry: in-plane contractions for this (qx,qy) line
    double Pg[G1][3], Px[G1][3], Py[G1][3];
#pragma unroll
    for (int c = 0; c < G1; c++) {
#pragma unroll
      for (int d = 0; d < 3; d++) { Pg[c][d] = 0.0; Px[c][d] = 0.0; Py[c][d] = 0.0; }
#pragma unroll
      for (int b = 0; b < G1; b++)
#pragma unroll
        for (int a = 0; a < G1; a++) {
          const double na = sNg[qx * G1 + a], nb = sNg[qy * G1 + b];
          const double da = sdNg[qx * G1 + a], db = sdNg[qy * G1 + b];
          const double *X = sg + 3 * (a + G1 * (b + G1 * c));
#pragma unroll
          for (int d = 0; d < 3; d++) {
            Pg[c][d] += na * nb * X[d];
            Px[c][d] += da * nb * X[d];
            Py[c][d] += na * db * X[d];
          }
        }
    }
    const double ae = (TERMS & TERM_CURL) ? (QPT ? 0.0 : scale * __ldg(alpha + e)) : 0.0;
    const double be = (TERMS & TERM_MASS) ? (QPT ? 0.0 : scale * __ldg(beta + e)) : 0.0;
#pragma unroll
    for (int a = 0; a < Q; a++) {
      double c0[3] = {0, 0, 0}, c1[3] = {0, 0, 0}, c2[3] = {0, 0, 0};
#pragma unroll
      for (int c = 0; c < G1; c++) {
        const double n = sNg[a * G1 + c], dn = sdNg[a * G1 + c];
#pragma unroll
        for (int d = 0; d < 3; d++) {
          c0[d] += n * Px[c][d];
          c1[d] += n * Py[c][d];
          c2[d] += dn * Pg[c][d];
        }
      }
      // adjugate rows: a0 = c1 x c2, a1 = c2 x c0, a2 = c0 x c1
      const double a0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
      const double det = c0[0] * a0[0] + c0[1] * a0[1] + c0[2] * a0[2];
      const double wq = sw[qx] * sw[qy] * sw[a] / det;
      if (TERMS & TERM_CURL) {
        double dyUx = 0, dxUy = 0, dyUz = 0, dxUz = 0, dzUx = 0, dzUy = 0;
#pragma unroll
        for (int m = 0; m < Q; m++) {
          const double dqy = sD[qy * Q + m], dqx = sD[qx * Q + m], dqz = sD[a * Q + m];
          dyUx += dqy * su[0 * Q3 + qx + Q * (m + Q * a)];
          dyUz += dqy * su[2 * Q3 + qx + Q * (m + Q * a)];
          dxUy += dqx * su[1 * Q3 + m + Q * (qy + Q * a)];
          dxUz += dqx * su[2 * Q3 + m + Q * (qy + Q * a)];
          dzUx += dqz * U[0][m];
          dzUy += dqz * U[1][m];
        }
        const double cx = dyUz - dzUy, cy = dzUx - dxUz, cz = dxUy - dyUx;
        const double ca = (QPT ? scale * __ldg(alpha + e * Q3 + qx + Q * (qy + Q * a)) : ae) * wq;
        double t[3];
#pragma unroll
        for (int d = 0; d < 3; d++) t[d] = ca * (c0[d] * cx + c1[d] * cy + c2[d] * cz);
        W[0][a] = c0[0] * t[0] + c0[1] * t[1] + c0[2] * t[2];
        W[1][a] = c1[0] * t[0] + c1[1] * t[1] + c1[2] * t[2];
        W[2][a] = c2[0] * t[0] + c2[1] * t[1] + c2[2] * t[2];
      }
      if (TERMS & TERM_MASS) {
        const double a1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
        const double a2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
        const double cb = (QPT ? scale * __ldg(beta + e * Q3 + qx + Q * (qy + Q * a)) : be) * wq;
        double s[3];
#pragma unroll
        for (int d = 0; d < 3; d++) s[d] = cb * (a0[d] * U[0][a] + a1[d] * U[1][a] + a2[d] * U[2][a]);
        V[0][a] = a0[0] * s[0] + a0[1] * s[1] + a0[2] * s[2];
        V[1][a] = a1[0] * s[0] + a1[1] * s[1] + a1[2] * s[2];
        V[2][a] = a2[0] * s[0] + a2[1] * s[1] + a2[2] * s[2];
      } else {
#pragma unroll
        for (int c = 0; c < 3; c++) V[c][a] = 0.0;
      }
    }
  }
  // ---- transposed collocation derivative ------------------------------------------
  if (TERMS & TERM_CURL) {
    __syncthreads();  // all reads of su (U) done
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; c++)
#pragma unroll
        for (int a = 0; a < Q; a++) su[c * Q3 + qx + Q * (qy + Q * a)] = W[c][a];
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int a = 0; a < Q; a++) {
        double zx = 0, zy = 0, zz = 0;
#pragma unroll
        for (int m = 0; m < Q; m++) {
          const double dz = sD[m * Q + a], dy = sD[m * Q + qy], dx = sD[m * Q + qx];
          // Z_x = +Dz^T W_y - Dy^T W_z ; Z_y = -Dz^T W_x + Dx^T W_z ; Z_z = +Dy^T W_x - Dx^T W_y
          zx += dz * W[1][m] - dy * su[2 * Q3 + qx + Q * (m + Q * a)];
          zy += dx * su[2 * Q3 + m + Q * (qy + Q * a)] - dz * W[0][m];
          zz += dy * su[0 * Q3 + qx + Q * (m + Q * a)] - dx * su[1 * Q3 + m + Q * (qy + Q * a)];
        }
        V[0][a] += zx; V[1][a] += zy; V[2][a] += zz;
      }
    }
  }
  // ---- backward: z ----------------------------------------------------------------
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n2 = (c == 2) ? P : P1;
      const double *B2 = (c == 2) ? sBo : sBc;
#pragma unroll
      for (int k = 0; k < P1; k++) {
        if (k < n2) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < Q; a++) s += B2[a * n2 + k] * V[c][a];
          t2[c * T2 + qx + Q * (qy + Q * k)] = s;
        }
      }
    }
  }
  __syncthreads();
  // ---- backward: y ----------------------------------------------------------------
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n1 = (c == 1) ? P : P1, n2 = (c == 2) ? P : P1;
      const double *B1 = (c == 1) ? sBo : sBc;
      const int q = tx % Q, k = tx / Q;
      if (k < n2) {
        double in[Q];
#pragma unroll
        for (int b = 0; b < Q; b++) in[b] = t2[c * T2 + q + Q * (b + Q * k)];
#pragma unroll
        for (int j = 0; j < P1; j++) {
          if (j < n1) {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < Q; b++) s += B1[b * n1 + j] * in[b];
            t1[c * T1 + q + Q * (j + n1 * k)] = s;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- backward: x, into xe ---------------------------------------------------------
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int n0 = (c == 0) ? P : P1, n1 = (c == 1) ? P : P1, n2 = (c == 2) ? P : P1;
      const double *B0 = (c == 0) ? sBo : sBc;
      if (tx < n1 * n2) {
        double in[Q];
#pragma unroll
        for (int q = 0; q < Q; q++) in[q] = t1[c * T1 + q + Q * tx];
#pragma unroll
        for (int i = 0; i < P1; i++) {
          if (i < n0) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < Q; q++) s += B0[q * n0 + i] * in[q];
            xe[c * NC + i + n0 * tx] = s;
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- CTA-wide signed scatter ----------------------------------------------------------
  for (int i = tid; i < ne_here * ND; i += C::THREADS) {
    const uint32_t m = __ldg(map + e0 * ND + i);
    if (m & MAP_ESS) continue;
    const double v = selem[(i / ND) * C::ELEM + (i % ND)];
    atomicAdd(y + (m & MAP_IDX), (m & MAP_NEG) ? -v : v);
  }
}

#endif  // HC_P

#ifdef HC_P
// ------------------------------------------------------------------ per-order launchers ----
// This translation unit is compiled once per order (-DHC_P=1..4) to keep build times short.
template <int P, int Q, int G, int TERMS>
static int launch_one(hcurl_op *op, const uint32_t *d_map, const double *d_x, double *d_y, double scale) {
  using C = HexCfg<P, Q, G>;
  hcurl_space *sp = op->sp;
  HexTables T;
  build_hex_tables(P, Q, G, T);
  const bool qpt = op->coef_layout == HCURL_COEF_PER_QPT;
  auto kern = qpt ? hex_apply_kernel<P, Q, G, TERMS, true> : hex_apply_kernel<P, Q, G, TERMS, false>;
  HC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
  const int64_t grid = (sp->n_elem + C::E - 1) / C::E;
  kern<<<(unsigned)grid, dim3(Q * Q, C::E), C::SMEM, op->ctx->stream>>>(T, sp->n_elem, d_map, sp->d_xyz, op->d_alpha,
                                                                          op->d_beta, d_x, d_y, scale);
  HC_CUDA(cudaGetLastError());
  op->ctx->launches++;
  return 0;
}

template <int P, int G>
static int launch_pq(hcurl_op *op, int q, int terms, const uint32_t *m, const double *x, double *y, double s) {
#define HC_CASE(QQ)                                                                 \
  if (q == QQ) {                                                                    \
    if (terms == TERM_CURL) return launch_one<P, QQ, G, TERM_CURL>(op, m, x, y, s); \
    if (terms == TERM_MASS) return launch_one<P, QQ, G, TERM_MASS>(op, m, x, y, s); \
    return launch_one<P, QQ, G, TERM_BOTH>(op, m, x, y, s);                         \
  }
  HC_CASE(P + 1)
  HC_CASE(P + 2)
  if (G == 2 && q == P + 3 && terms == TERM_MASS) return launch_one<P, P + 3, G, TERM_MASS>(op, m, x, y, s);
#undef HC_CASE
  set_error("hex apply: unsupported quadrature size q=%d for order %d", q, P);
  return 1;
}

#define HC_CAT2(a, b) a##b
#define HC_CAT(a, b) HC_CAT2(a, b)
int HC_CAT(hex_launch_p, HC_P)(hcurl_op *op, int q, int terms, const uint32_t *m, const double *x, double *y, double s) {
  if (op->sp->geom_order == 1) return launch_pq<HC_P, 1>(op, q, terms, m, x, y, s);
  return launch_pq<HC_P, 2>(op, q, terms, m, x, y, s);
}

#else  // !HC_P : dispatcher + diagonal
int hex_launch_p1(hcurl_op *, int, int, const uint32_t *, const double *, double *, double);
int hex_launch_p2(hcurl_op *, int, int, const uint32_t *, const double *, double *, double);
int hex_launch_p3(hcurl_op *, int, int, const uint32_t *, const double *, double *, double);
int hex_launch_p4(hcurl_op *, int, int, const uint32_t *, const double *, double *, double);

static int launch_term(hcurl_op *op, int q, int terms, const uint32_t *m, const double *x, double *y, double s) {
  switch (op->sp->order) {
    case 1: return hex_launch_p1(op, q, terms, m, x, y, s);
    case 2: return hex_launch_p2(op, q, terms, m, x, y, s);
    case 3: return hex_launch_p3(op, q, terms, m, x, y, s);
    case 4: return hex_launch_p4(op, q, terms, m, x, y, s);
  }
  set_error("hex apply: unsupported order %d", op->sp->order);
  return 1;
}

int hex_apply_launch(hcurl_op *op, const uint32_t *d_map, const double *d_x, double *d_y, double scale) {
  if (!op->has_alpha && !op->has_beta) return 0;
  if (op->single_rule || !(op->has_alpha && op->has_beta)) {
    const int terms = (op->has_alpha ? TERM_CURL : 0) | (op->has_beta ? TERM_MASS : 0);
    const int q = op->single_rule ? op->qc : (op->has_alpha ? op->qc : op->qm);
    return launch_term(op, q, terms, d_map, d_x, d_y, scale);
  }
  if (launch_term(op, op->qc, TERM_CURL, d_map, d_x, d_y, scale)) return 1;
  return launch_term(op, op->qm, TERM_MASS, d_map, d_x, d_y, scale);
}

// Diagonal of the element matrices, summed into diag (simple, setup-time kernel):
// one thread per (element, tensor dof).
template <int G>
__global__ void hex_diag_kernel(const HexTables T, int P, int Q, int terms, int qpt, const int64_t n_elem,
                                const uint32_t *__restrict__ map, const double *__restrict__ xyz,
                                const double *__restrict__ alpha, const double *__restrict__ beta,
                                double *__restrict__ diag) {
  constexpr int G1 = G + 1, NG = G1 * G1 * G1;
  const int P1 = P + 1, NC = P * P1 * P1, ND = 3 * NC;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * ND) return;
  const int64_t e = gid / ND;
  const int t = (int)(gid % ND);
  const uint32_t m = map[gid];
  if (m & MAP_ESS) return;
  const int c = t / NC;
  const int r = t % NC;
  const int n0 = (c == 0) ? P : P1, n1 = (c == 1) ? P : P1;
  int idx[3] = {r % n0, (r / n0) % n1, r / (n0 * n1)};
  const double *X = xyz + e * NG * 3;
  const int d1 = (c + 1) % 3, d2 = (c + 2) % 3;
  double acc = 0.0;
  for (int qz = 0; qz < Q; qz++)
    for (int qyy = 0; qyy < Q; qyy++)
      for (int qxx = 0; qxx < Q; qxx++) {
        const int qq[3] = {qxx, qyy, qz};
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int cc = 0; cc < G1; cc++)
          for (int b = 0; b < G1; b++)
            for (int a = 0; a < G1; a++) {
              const double na = T.Ng[qxx * 3 + a], nb = T.Ng[qyy * 3 + b], nc = T.Ng[qz * 3 + cc];
              const double da = T.dNg[qxx * 3 + a], db = T.dNg[qyy * 3 + b], dc = T.dNg[qz * 3 + cc];
              const double *Xn = X + 3 * (a + G1 * (b + G1 * cc));
              for (int d = 0; d < 3; d++) {
                J[d][0] += da * nb * nc * Xn[d];
                J[d][1] += na * db * nc * Xn[d];
                J[d][2] += na * nb * dc * Xn[d];
              }
            }
        const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        const double w = T.wq[qxx] * T.wq[qyy] * T.wq[qz];
        // 1D factors of this basis function
        const double fo = T.Bo[qq[c] * MAXP1 + idx[c]];
        const double f1 = T.Bc[qq[d1] * MAXP1 + idx[d1]], g1 = T.Gc[qq[d1] * MAXP1 + idx[d1]];
        const double f2 = T.Bc[qq[d2] * MAXP1 + idx[d2]], g2 = T.Gc[qq[d2] * MAXP1 + idx[d2]];
        const int64_t qi = e * (int64_t)Q * Q * Q + qxx + Q * (qyy + Q * qz);
        if (terms & TERM_CURL) {
          double cr[3] = {0, 0, 0};
          cr[d1] = fo * f1 * g2;
          cr[d2] = -fo * g1 * f2;
          double tt[3];
          for (int d = 0; d < 3; d++) tt[d] = J[d][0] * cr[0] + J[d][1] * cr[1] + J[d][2] * cr[2];
          const double a = qpt ? alpha[qi] : alpha[e];
          acc += a * w / det * (tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
        }
        if (terms & TERM_MASS) {
          // J^-T e_c * phi = (row c of adj^T)/det ... = adj[c][:]/det with adj rows a_c = c_{c+1} x c_{c+2}
          double col1[3] = {J[0][d1], J[1][d1], J[2][d1]}, col2[3] = {J[0][d2], J[1][d2], J[2][d2]};
          double ac[3] = {col1[1] * col2[2] - col1[2] * col2[1], col1[2] * col2[0] - col1[0] * col2[2],
                          col1[0] * col2[1] - col1[1] * col2[0]};
          const double phi = fo * f1 * f2;
          const double b = qpt ? beta[qi] : beta[e];
          acc += b * w / det * phi * phi * (ac[0] * ac[0] + ac[1] * ac[1] + ac[2] * ac[2]);
        }
      }
  atomicAdd(diag + (m & MAP_IDX), acc);
}

int hex_diag_launch(hcurl_op *op, double *d_diag) {
  hcurl_space *sp = op->sp;
  const uint32_t *map = op->d_map_ess ? op->d_map_ess : sp->d_map;
  const int qpt = op->coef_layout == HCURL_COEF_PER_QPT;
  const int64_t n = sp->n_elem * sp->nd;
  const int bs = 128;
  for (int pass = 0; pass < 2; pass++) {
    int terms, q;
    if (op->single_rule) {
      if (pass) break;
      terms = (op->has_alpha ? TERM_CURL : 0) | (op->has_beta ? TERM_MASS : 0);
      q = op->qc;
    } else {
      terms = pass ? TERM_MASS : TERM_CURL;
      q = pass ? op->qm : op->qc;
      if ((pass == 0 && !op->has_alpha) || (pass == 1 && !op->has_beta)) continue;
    }
    if (!terms) continue;
    HexTables T;
    build_hex_tables(sp->order, q, sp->geom_order, T);
    if (sp->geom_order == 1)
      hex_diag_kernel<1><<<(unsigned)((n + bs - 1) / bs), bs, 0, op->ctx->stream>>>(T, sp->order, q, terms, qpt, sp->n_elem, map,
                                                                                  sp->d_xyz, op->d_alpha, op->d_beta, d_diag);
    else
      hex_diag_kernel<2><<<(unsigned)((n + bs - 1) / bs), bs, 0, op->ctx->stream>>>(T, sp->order, q, terms, qpt, sp->n_elem, map,
                                                                                  sp->d_xyz, op->d_alpha, op->d_beta, d_diag);
    HC_CUDA(cudaGetLastError());
  }
  return 0;
}

#endif  // HC_P

}  // namespace hcurl
