// Patch-based order-2 Nedelec hex apply:  y (+)= scale * (K(alpha) + M(beta)) x.
//
// Same operator as hex_apply.cu / hex_fast.cu (CurlCurlKernel + VectorFEMassKernel,
// /root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24, applied as
// EquationSystem::Mult, /root/reference/src/equation_systems/equation_system.cpp:200-204), organised
// around what ncu showed to bound the element-wise kernels on sm_100a:
//
//  * L1tex lane rate of gathers / FP64 RED atomics  ->  one CTA per PATCH of FE adjacent elements
//    (patch.cpp); x is gathered once per unique patch dof into a shared-memory E-vector, y is
//    summed per patch dof from the E-vector, private dofs are plain stores, only patch-boundary
//    dofs use RED.
//  * FP64 pipe (36 TFLOP/s measured; the quadrature-based apply needs 2916 DFMA / element)  ->
//    on AFFINE elements with per-element coefficients the element operator is applied in an
//    orthonormal Legendre (modal) basis, where the mass matrix is a restricted identity and the
//    curl has one term per mode:  A_e = T^T [ GM (x) I + C^T (GK (x) I) C ] T  with
//    GK = alpha J^T J / det, GM = beta adj adj^T / det.  474 DFMA per thread, 1422 per element,
//    exact (the p+1 Gauss rule it replaces is exact on these elements too).
//  * instruction-cache misses of fully unrolled per-component code  ->  every thread works in the
//    rotated frame (c, c+1, c+2) of its component c, so all warps run the same instruction stream.
//
// Thread layout: 3*FE threads; thread (c, e) owns component c of element e of the patch.
// E-vector layout: E[(c*18 + t)*(FE+1) + e], t = i0 + 2*(i1 + 3*i2) with i0 the open index along axis c
// (stride FE+1: lanes of the per-dof phases that hit different slots of one element fall in different banks).
#include <cstdlib>

#include "internal.h"

namespace hcurl {

struct ModalTables {
  // closed basis (Gauss-Lobatto nodes 0, 1/2, 1) -> orthonormal Legendre L0, L1, L2:
  //   T = [[a0, a1, a0], [-b, 0, b], [c0, c1, c0]]   (symmetry of the nodes)
  // open basis (2 Gauss-Legendre nodes) -> L0, L1 is [[1/2, 1/2], [-1/2, 1/2]]; the kernel applies
  // [[1, 1], [-1, 1]] (adds only) and folds the two factors 1/2 (trial and test side) into the
  // geometry factors.
  double a0, a1, b, c0, c1;
  double d1, d2;  // L1' = d1 L0, L2' = d2 L1
};
__constant__ ModalTables cM;

int hex_rot_slot_to_tensor(int p, int slot) {
  const int p1 = p + 1, nc = p * p1 * p1;
  const int c = slot / nc, t = slot % nc;
  int ir[3] = {t % p, (t / p) % p1, t / (p * p1)};  // along axes c, c+1, c+2
  int idx[3];
  for (int k = 0; k < 3; k++) idx[(c + k) % 3] = ir[k];
  const int n0 = (c == 0) ? p : p1, n1 = (c == 1) ? p : p1;
  return c * nc + idx[0] + n0 * (idx[1] + n1 * idx[2]);
}

__device__ __forceinline__ int sym6(int i, int j) {  // packed index of the symmetric pair (i,j)
  const int lo = min(i, j), hi = max(i, j);
  return lo * 3 - lo * (lo - 1) / 2 + (hi - lo);
}

#define IA(m0, m1, m2) ((m0) + 2 * ((m1) + 3 * (m2)))  // own component  [2][3][3]
#define IB(m0, m1, m2) ((m0) + 3 * ((m1) + 2 * (m2)))  // component c+1  [3][2][3] (my frame)
#define IC(m0, m1, m2) ((m0) + 3 * ((m1) + 3 * (m2)))  // component c+2  [3][3][2] (my frame)

// 1-D closed transform of a line (x0, x1, x2) -> (m0, m1, m2), 7 FP64 instructions
__device__ __forceinline__ void tc_fwd(const double x0, const double x1, const double x2, double &m0, double &m1, double &m2) {
  const double s = x0 + x2, d = x2 - x0;
  m0 = fma(cM.a1, x1, cM.a0 * s);
  m1 = cM.b * d;
  m2 = fma(cM.c1, x1, cM.c0 * s);
}
// transpose: (y0, y1, y2) modal test values -> nodal
__device__ __forceinline__ void tc_bwd(const double y0, const double y1, const double y2, double &x0, double &x1, double &x2) {
  const double p = fma(cM.c0, y2, cM.a0 * y0), q = cM.b * y1;
  x0 = p - q;
  x2 = p + q;
  x1 = fma(cM.c1, y2, cM.a1 * y0);
}
// nodal -> (2x) modal on own component
__device__ __forceinline__ void modal_fwd(const double *X, double *A) {
  double a[18], b[18];
#pragma unroll
  for (int r = 0; r < 9; r++) { a[2 * r] = X[2 * r] + X[2 * r + 1]; a[2 * r + 1] = X[2 * r + 1] - X[2 * r]; }
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      tc_fwd(a[IA(m0, 0, i2)], a[IA(m0, 1, i2)], a[IA(m0, 2, i2)], b[IA(m0, 0, i2)], b[IA(m0, 1, i2)], b[IA(m0, 2, i2)]);
#pragma unroll
  for (int r = 0; r < 6; r++) tc_fwd(b[r], b[r + 6], b[r + 12], A[r], A[r + 6], A[r + 12]);
}
// modal -> nodal test functions (transpose, again with the open transform scaled by 2)
__device__ __forceinline__ void modal_bwd(const double *Yh, double *Y) {
  double a[18], b[18];
#pragma unroll
  for (int r = 0; r < 6; r++) tc_bwd(Yh[r], Yh[r + 6], Yh[r + 12], a[r], a[r + 6], a[r + 12]);
#pragma unroll
  for (int i2 = 0; i2 < 3; i2++)
#pragma unroll
    for (int m0 = 0; m0 < 2; m0++)
      tc_bwd(a[IA(m0, 0, i2)], a[IA(m0, 1, i2)], a[IA(m0, 2, i2)], b[IA(m0, 0, i2)], b[IA(m0, 1, i2)], b[IA(m0, 2, i2)]);
#pragma unroll
  for (int r = 0; r < 9; r++) { Y[2 * r] = b[2 * r] - b[2 * r + 1]; Y[2 * r + 1] = b[2 * r] + b[2 * r + 1]; }
}

// modal element operator for component a0 = c: reads own (Am) and the other two components' modal
// coefficients (Bm = comp c+1, Cm = comp c+2, in my frame), G = scaled geometry factors
__device__ __forceinline__ void modal_op(const double *Am, const double *Bm, const double *Cm, const double M00,
                                         const double M01, const double M02, const double K11, const double K10,
                                         const double K12, const double K22, const double K20, const double K21,
                                         double *Yh) {
  const double d1 = cM.d1, d2 = cM.d2;
#define DD(n) ((n) == 1 ? d1 : d2)
  double r0[8], r1[12], r2[12];
#pragma unroll
  for (int m0 = 0; m0 < 2; m0++) {
#pragma unroll
    for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
        r0[m0 + 2 * (m1 + 2 * m2)] = fma(DD(m1 + 1), Cm[IC(m0, m1 + 1, m2)], -DD(m2 + 1) * Bm[IB(m0, m1, m2 + 1)]);
#pragma unroll
    for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 2; m2++)
        r1[m0 + 2 * (m1 + 3 * m2)] = fma(DD(m2 + 1), Am[IA(m0, m1, m2 + 1)], -DD(m0 + 1) * Cm[IC(m0 + 1, m1, m2)]);
#pragma unroll
    for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
      for (int m2 = 0; m2 < 3; m2++)
        r2[m0 + 2 * (m1 + 2 * m2)] = fma(DD(m0 + 1), Bm[IB(m0 + 1, m1, m2)], -DD(m1 + 1) * Am[IA(m0, m1 + 1, m2)]);
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
        double v = M00 * Am[IA(n0, n1, n2)];
        if (n1 < 2) v = fma(M01, Bm[IB(n0, n1, n2)], v);
        if (n2 < 2) v = fma(M02, Cm[IC(n0, n1, n2)], v);
        if (n2 >= 1) v = fma(DD(n2), w1[n0 + 2 * (n1 + 3 * (n2 - 1))], v);
        if (n1 >= 1) v = fma(-DD(n1), w2[n0 + 2 * ((n1 - 1) + 2 * n2)], v);
        Yh[IA(n0, n1, n2)] = v;
      }
#undef DD
}

struct PatchArgs {
  int64_t n_pos;             // n_patch * FE
  const int32_t *perm;       // element at each patch position (-1 padding)
  const uint32_t *blob_off;  // [n_patch+1] in 16-byte units
  const uint32_t *blob;      // patch blobs (patch.cpp)
  const double *edata;       // [12][n_pos] affine geometry factors in patch order
  const double *x;
  double *y;
  double scale;
  uint32_t blob_cap_bytes;   // shared-memory room reserved for one blob (multiple of 16)
  unsigned long long *prof;  // null, or 8 counters: cycles of thread 0 per phase summed over CTAs (scripts/phase_profile.py)
};
#define PROF_MARK(i)                                                                         \
  if (a.prof && tid == 0) {                                                                  \
    const long long t_now = clock64();                                                       \
    atomicAdd(a.prof + (i), (unsigned long long)(t_now - t_prev));                           \
    t_prev = t_now;                                                                          \
  }

// sign flip carried in bit 15 of a slot entry
__device__ __forceinline__ double apply_sign(double v, uint32_t entry) {
  return __hiloint2double(__double2hiint(v) ^ (int)((entry << 16) & 0x80000000u), __double2loint(v));
}

// ---- bulk async copy (TMA engine, 1-D) of the patch blob into shared memory ----------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void blob_load_async(uint64_t *mbar, void *dst, const void *src, uint32_t bytes) {
  const uint32_t mb = smem_addr(mbar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(mb)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
  const uint32_t mb = smem_addr(mbar);
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(mb), "r"(parity)
                 : "memory");
  } while (!done);
}

template <int FE, int MINB, bool ACC>
__global__ void __launch_bounds__(3 * FE, MINB)
hex_patch_modal_p2_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = 3 * FE, ST = FE + 1;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);
  const uint32_t *B = reinterpret_cast<const uint32_t *>(smem_raw + 16);
  double *E = reinterpret_cast<double *>(smem_raw + 16 + a.blob_cap_bytes);  // [54][FE + 1]
  const int tid = threadIdx.x;
  const int64_t k = blockIdx.x;
  long long t_prev = a.prof ? clock64() : 0;
  if (tid == 0) {
    const uint32_t o0 = __ldg(a.blob_off + k), o1 = __ldg(a.blob_off + k + 1);
    blob_load_async(mbar, smem_raw + 16, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u);
  }
  // component c of patch element el; its geometry factors are fetched now, consumed after phase A
  const int c = tid / FE, el = tid - c * FE;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int64_t pos = k * FE + el;
  const bool ok = __ldg(a.perm + pos) >= 0;
  const double *g = a.edata + pos;
  const int64_t np = a.n_pos;
  double K11 = __ldg(g + sym6(c1, c1) * np), K10 = __ldg(g + sym6(c1, c) * np), K12 = __ldg(g + sym6(c1, c2) * np),
         K22 = __ldg(g + sym6(c2, c2) * np), K20 = __ldg(g + sym6(c2, c) * np);
  double M00 = __ldg(g + (6 + sym6(c, c)) * np), M01 = __ldg(g + (6 + sym6(c, c1)) * np), M02 = __ldg(g + (6 + sym6(c, c2)) * np);
  __syncthreads();  // mbarrier initialised
  mbar_wait(mbar, 0);
  PROF_MARK(0)  // blob arrival
  const uint32_t U = B[0], b1 = B[1], b2 = B[1] + B[2], nP1 = B[4], nP2 = B[5], nP4 = B[6], nov = B[7];
  const uint32_t *dof = B + PatchHost::HDR;
  const uint16_t *e16 = reinterpret_cast<const uint16_t *>(B + B[8]);
  const uint32_t *e32 = B + B[9];
  const uint64_t *e64 = reinterpret_cast<const uint64_t *>(B + B[10]);
  const uint32_t *ovf = B + B[11];
  // ---- A: gather x once per unique patch dof (all loads of a chunk in flight together), push the
  // value to the element slots the dof feeds
  {
    constexpr int UNR = 6;
    for (uint32_t u0 = tid; u0 < U; u0 += UNR * NT) {
      uint32_t m[UNR];
      double v[UNR];
#pragma unroll
      for (int i = 0; i < UNR; i++) m[i] = (u0 + i * NT < U) ? dof[u0 + i * NT] : MAP_ESS;
#pragma unroll
      for (int i = 0; i < UNR; i++) v[i] = (m[i] & MAP_ESS) ? 0.0 : __ldg(a.x + (m[i] & MAP_IDX));
#pragma unroll
      for (int i = 0; i < UNR; i++) {
        const uint32_t u = u0 + i * NT;
        if (u < b1) {
          const uint32_t sl = e16[u];
          E[sl & 0x7fffu] = apply_sign(v[i], sl);
        } else if (u < b2) {
          const uint32_t w = e32[u - b1];
          E[w & 0x7fffu] = apply_sign(v[i], w);
          E[(w >> 16) & 0x7fffu] = apply_sign(v[i], w >> 16);
        } else if (u < U) {
          const uint64_t w = e64[u - b2];
          const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
          E[lo & 0x7fffu] = apply_sign(v[i], lo);
          E[(lo >> 16) & 0x7fffu] = apply_sign(v[i], lo >> 16);
          E[hi & 0x7fffu] = apply_sign(v[i], hi);
          if ((hi >> 16) != 0xffffu) E[(hi >> 16) & 0x7fffu] = apply_sign(v[i], hi >> 16);
        }
      }
    }
    for (uint32_t t = tid; t < nov; t += NT) {  // slots beyond four (unstructured meshes)
      const uint32_t e = ovf[t], m = dof[e >> 16];
      const double v = (m & MAP_ESS) ? 0.0 : __ldg(a.x + (m & MAP_IDX));
      E[e & 0x7fffu] = apply_sign(v, e);
    }
  }
  __syncthreads();
  PROF_MARK(1)  // phase A: gather
  // ---- element operator
  double *Eown = E + (c * 18) * ST + el;
  double Am[18];
  if (ok) {
    double X[18];
#pragma unroll
    for (int t = 0; t < 18; t++) X[t] = Eown[t * ST];
    modal_fwd(X, Am);
#pragma unroll
    for (int t = 0; t < 18; t++) Eown[t * ST] = Am[t];
  }
  __syncthreads();
  double Y[18];
  if (ok) {
    const double *E1 = E + (c1 * 18) * ST + el, *E2 = E + (c2 * 18) * ST + el;
    double Bm[18], Cm[18];
    // owner frames: comp c+1 stores (m1, m2, m0), comp c+2 stores (m2, m0, m1)
#pragma unroll
    for (int m2 = 0; m2 < 3; m2++)
#pragma unroll
      for (int m1 = 0; m1 < 2; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) Bm[IB(m0, m1, m2)] = E1[(m1 + 2 * (m2 + 3 * m0)) * ST];
#pragma unroll
    for (int m2 = 0; m2 < 2; m2++)
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) Cm[IC(m0, m1, m2)] = E2[(m2 + 2 * (m0 + 3 * m1)) * ST];
    const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
    K11 *= sc; K10 *= sc; K12 *= sc; K22 *= sc; K20 *= sc; M00 *= sc; M01 *= sc; M02 *= sc;
    double Yh[18];
    modal_op(Am, Bm, Cm, M00, M01, M02, K11, K10, K12, K22, K20, K12, Yh);
    modal_bwd(Yh, Y);
  }
  __syncthreads();  // everyone has read the modal coefficients
  PROF_MARK(3)  // exchange + modal operator + backward transform
  if (ok) {
#pragma unroll
    for (int t = 0; t < 18; t++) Eown[t * ST] = Y[t];
  }
  __syncthreads();
  // ---- C: sum the slots of every patch dof; private dofs are plain stores, shared ones RED
#pragma unroll 2
  for (uint32_t u = tid; u < U; u += NT) {
    const uint32_t m = dof[u];
    if (m & MAP_ESS) continue;
    double v;
    bool priv;
    if (u < b1) {
      const uint32_t sl = e16[u];
      v = apply_sign(E[sl & 0x7fffu], sl);
      priv = u < nP1;
    } else if (u < b2) {
      const uint32_t w = e32[u - b1];
      v = apply_sign(E[w & 0x7fffu], w) + apply_sign(E[(w >> 16) & 0x7fffu], w >> 16);
      priv = u - b1 < nP2;
    } else {
      const uint64_t w = e64[u - b2];
      const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
      v = apply_sign(E[lo & 0x7fffu], lo) + apply_sign(E[(lo >> 16) & 0x7fffu], lo >> 16) + apply_sign(E[hi & 0x7fffu], hi);
      if ((hi >> 16) != 0xffffu) v += apply_sign(E[(hi >> 16) & 0x7fffu], hi >> 16);
      priv = u - b2 < nP4;
    }
    double *dst = a.y + (m & MAP_IDX);
    if (priv) *dst = ACC ? *dst + v : v;
    else atomicAdd(dst, v);
  }
  for (uint32_t t = tid; t < nov; t += NT) {
    const uint32_t e = ovf[t], m = dof[e >> 16];
    if (m & MAP_ESS) continue;
    atomicAdd(a.y + (m & MAP_IDX), apply_sign(E[e & 0x7fffu], e));
  }
  PROF_MARK(5)  // phase C: scatter (thread 0's share)
}

// ---- setup ---------------------------------------------------------------------------------------
// affine geometry factors in patch order: S = alpha J^T J/det (6), M = beta adj adj^T/det (6)
__global__ void patch_edata_kernel(const int64_t n_pos, const int32_t *__restrict__ perm, const double *__restrict__ xyz,
                                   const double *__restrict__ alpha, const double *__restrict__ beta,
                                   double *__restrict__ edata) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pos) return;
  const int32_t e = perm[j];
  if (e < 0) {
    for (int s = 0; s < 12; s++) edata[(int64_t)s * n_pos + j] = 0.0;
    return;
  }
  const double *X = xyz + (int64_t)e * 24;  // tensor order
  double c0[3], c1[3], c2[3];
  for (int d = 0; d < 3; d++) { c0[d] = X[3 + d] - X[d]; c1[d] = X[6 + d] - X[d]; c2[d] = X[12 + d] - X[d]; }
  const double a0[3] = {c1[1] * c2[2] - c1[2] * c2[1], c1[2] * c2[0] - c1[0] * c2[2], c1[0] * c2[1] - c1[1] * c2[0]};
  const double a1[3] = {c2[1] * c0[2] - c2[2] * c0[1], c2[2] * c0[0] - c2[0] * c0[2], c2[0] * c0[1] - c2[1] * c0[0]};
  const double a2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
  const double det = c0[0] * a0[0] + c0[1] * a0[1] + c0[2] * a0[2];
  const double al = alpha ? alpha[e] / det : 0.0, be = beta ? beta[e] / det : 0.0;
  const double *c[3] = {c0, c1, c2};
  const double *ad[3] = {a0, a1, a2};
  int s = 0;
  for (int i = 0; i < 3; i++)
    for (int jj = i; jj < 3; jj++, s++) {
      edata[(int64_t)s * n_pos + j] = al * (c[i][0] * c[jj][0] + c[i][1] * c[jj][1] + c[i][2] * c[jj][2]);
      edata[(int64_t)(6 + s) * n_pos + j] = be * (ad[i][0] * ad[jj][0] + ad[i][1] * ad[jj][1] + ad[i][2] * ad[jj][2]);
    }
}

// OR the essential flag into the dof words of every patch blob (one CTA per patch)
__global__ void patch_flag_kernel(const uint32_t *__restrict__ blob_off, const double *__restrict__ flag,
                                  uint32_t *__restrict__ blob) {
  uint32_t *B = blob + (size_t)blob_off[blockIdx.x] * 4;
  const uint32_t U = B[0];
  for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
    const uint32_t m = B[PatchHost::HDR + u];
    if (flag[m & MAP_IDX] != 0.0) B[PatchHost::HDR + u] = m | MAP_ESS;
  }
}

unsigned long long *g_patch_prof = nullptr;  // set by hcurl_debug_patch_profile (debug.cu)
static bool g_modal_set = false;
static int set_modal_tables() {
  if (g_modal_set) return 0;
  // orthonormal shifted Legendre on [0,1]: L0 = 1, L1 = sqrt3 (2x-1), L2 = sqrt5 (6x^2-6x+1);
  // T[m][j] = int_0^1 L_m l_j, by a 5-point Gauss rule (exact)
  double xq[5], wq[5], gl[2], wl[2], gll[3];
  gauss_legendre01(5, xq, wq);
  gauss_legendre01(2, gl, wl);
  gauss_lobatto01(3, gll);
  double Bo[5 * 2], Bc[5 * 3];
  lagrange_tables(2, gl, 5, xq, Bo, nullptr);
  lagrange_tables(3, gll, 5, xq, Bc, nullptr);
  double Tc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, To[4] = {0, 0, 0, 0};
  for (int q = 0; q < 5; q++) {
    const double t = 2.0 * xq[q] - 1.0;
    const double L[3] = {1.0, sqrt(3.0) * t, sqrt(5.0) * 0.5 * (3.0 * t * t - 1.0)};
    for (int m = 0; m < 3; m++) {
      for (int j = 0; j < 3; j++) Tc[m * 3 + j] += wq[q] * L[m] * Bc[q * 3 + j];
      if (m < 2) for (int i = 0; i < 2; i++) To[m * 2 + i] += wq[q] * L[m] * Bo[q * 2 + i];
    }
  }
  // the kernel hard-wires the structure of these tables; check it
  const double tol = 1e-14;
  if (fabs(To[0] - 0.5) > tol || fabs(To[1] - 0.5) > tol || fabs(To[2] + 0.5) > tol || fabs(To[3] - 0.5) > tol ||
      fabs(Tc[0] - Tc[2]) > tol || fabs(Tc[3] + Tc[5]) > tol || fabs(Tc[4]) > tol || fabs(Tc[6] - Tc[8]) > tol) {
    set_error("modal tables do not have the expected symmetric structure");
    return 1;
  }
  ModalTables T;
  T.a0 = 0.5 * (Tc[0] + Tc[2]); T.a1 = Tc[1]; T.b = 0.5 * (Tc[5] - Tc[3]); T.c0 = 0.5 * (Tc[6] + Tc[8]); T.c1 = Tc[7];
  T.d1 = 2.0 * sqrt(3.0);
  T.d2 = 2.0 * sqrt(15.0);
  HC_CUDA(cudaMemcpyToSymbol(cM, &T, sizeof(T)));
  g_modal_set = true;
  return 0;
}

template <class T>
static int upload(T *&dst, const std::vector<T> &src) {
  HC_CUDA(cudaMalloc(&dst, sizeof(T) * (src.size() ? src.size() : 1)));
  HC_CUDA(cudaMemcpy(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
  return 0;
}

int hex_patch_build(hcurl_space *sp, const double *elem_xyz, const int32_t *elem_ldof, const int32_t *l2e) {
  if (sp->geom != HCURL_GEOM_HEX || sp->order != 2 || sp->geom_order != 1) return 0;
  const char *off = getenv("HCURL_NO_PATCH");
  if (off && off[0] == '1') return 0;
  int FE = 64;
  if (const char *s = getenv("HCURL_PATCH_FE")) FE = atoi(s);
  if (FE != 32 && FE != 64 && FE != 128) { set_error("HCURL_PATCH_FE must be 32, 64 or 128"); return 1; }
  const int nd = sp->nd, p = sp->order;
  const int64_t ne = sp->n_elem;
  std::vector<int> loc, sgn;
  nd_hex_dof_map(p, loc, sgn);
  std::vector<int> tens(nd);
  for (int s = 0; s < nd; s++) tens[s] = hex_rot_slot_to_tensor(p, s);
  std::vector<uint32_t> code((size_t)ne * nd);
  std::vector<double> cen((size_t)ne * 3);
  for (int64_t e = 0; e < ne; e++) {
    for (int s = 0; s < nd; s++) {
      const int t = tens[s];
      const int32_t v = elem_ldof[e * nd + loc[t]];
      const int32_t kk = v >= 0 ? v : -1 - v;
      bool neg = v < 0;
      if (sgn[t] < 0) neg = !neg;
      const uint32_t ext = l2e ? (uint32_t)l2e[kk] : (uint32_t)kk;
      code[(size_t)e * nd + s] = ext | (neg ? MAP_NEG : 0u);
    }
    for (int d = 0; d < 3; d++) {
      double c = 0.0;
      for (int v = 0; v < 8; v++) c += elem_xyz[(e * 8 + v) * 3 + d];
      cen[(size_t)e * 3 + d] = 0.125 * c;
    }
  }
  PatchHost H;
  if (build_patches(nd, ne, FE, FE + 1, sp->n_ext(), cen.data(), code.data(), H)) return 1;
  HC_CUDA(cudaSetDevice(sp->ctx->device));
  PatchSet &P = sp->patch;
  P.FE = FE; P.stride = H.stride; P.n_patch = H.n_patch;
  P.blob_words = (int64_t)H.blob.size(); P.max_blob_words = H.max_blob_words;
  if ((size_t)H.max_blob_words * 4 + sizeof(double) * 54 * (FE + 1) + 16 > 200 * 1024) return 0;  // patch path unavailable
  if (upload(P.d_perm, H.perm) || upload(P.d_blob_off, H.blob_off) || upload(P.d_blob, H.blob)) return 1;
  sp->has_patches = true;
  return 0;
}

void hex_patch_free(hcurl_space *sp) {
  PatchSet &P = sp->patch;
  cudaFree(P.d_perm); cudaFree(P.d_blob_off); cudaFree(P.d_blob);
  cudaFree(sp->d_xyzP);
  P = PatchSet();
  sp->has_patches = false;
}

bool hex_patch_eligible(const hcurl_op *op) {
  const hcurl_space *sp = op->sp;
  if (!(sp->has_patches && op->single_rule && op->qc == 3 && op->coef_layout == HCURL_COEF_PER_ELEM && op->has_alpha &&
        op->has_beta))
    return false;
  if (op->kernel_sel == 1 || op->kernel_sel == 2) return false;
  // affine meshes only (modal operator); checked lazily by hex_fast_prepare_space
  if (!sp->fast_ready && hex_fast_prepare_space(const_cast<hcurl_space *>(sp))) return false;
  return sp->all_affine;
}

template <int FE, int MINB>
static int launch_fe(const PatchArgs &a, int64_t n_patch, bool acc, cudaStream_t st) {
  const size_t smem = 16 + a.blob_cap_bytes + sizeof(double) * 54 * (FE + 1);
  auto k0 = hex_patch_modal_p2_kernel<FE, MINB, false>;
  auto k1 = hex_patch_modal_p2_kernel<FE, MINB, true>;
  static size_t attr = 0;
  if (attr < smem) {
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    attr = smem;
  }
  if (acc) k1<<<(unsigned)n_patch, 3 * FE, smem, st>>>(a);
  else k0<<<(unsigned)n_patch, 3 * FE, smem, st>>>(a);
  return 0;
}

int hex_patch_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale, bool accumulate) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  PatchSet &P = sp->patch;
  if (set_modal_tables()) return 1;
  const int64_t n_pos = P.n_patch * P.FE;
  if (!op->edataP_valid) {
    if (!op->d_edataP) HC_CUDA(cudaMalloc(&op->d_edataP, sizeof(double) * 12 * n_pos));
    patch_edata_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, ctx->stream>>>(n_pos, P.d_perm, sp->d_xyz, op->d_alpha, op->d_beta,
                                                                                op->d_edataP);
    HC_CUDA(cudaGetLastError());
    op->edataP_valid = true;
  }
  PatchArgs a;
  a.n_pos = n_pos; a.perm = P.d_perm; a.blob_off = P.d_blob_off;
  a.blob = (constrained && op->d_blob_ess) ? op->d_blob_ess : P.d_blob;
  a.blob_cap_bytes = (P.max_blob_words * 4 + 15u) & ~15u;
  a.edata = op->d_edataP; a.x = d_x; a.y = d_y; a.scale = scale;
  a.prof = g_patch_prof;
  int rc = 1;
  if (P.FE == 32) rc = launch_fe<32, 8>(a, P.n_patch, accumulate, ctx->stream);
  else if (P.FE == 64) rc = launch_fe<64, 4>(a, P.n_patch, accumulate, ctx->stream);
  else if (P.FE == 128) rc = launch_fe<128, 1>(a, P.n_patch, accumulate, ctx->stream);
  if (rc) return rc;
  HC_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// essential flags on the patch dof list (called from hcurl_op_set_essential_tdofs with the flag vector)
int hex_patch_flag_essential(hcurl_op *op, const double *d_flag) {
  hcurl_space *sp = op->sp;
  if (!sp->has_patches) return 0;
  PatchSet &P = sp->patch;
  cudaFree(op->d_blob_ess); op->d_blob_ess = nullptr;
  HC_CUDA(cudaMalloc(&op->d_blob_ess, sizeof(uint32_t) * P.blob_words));
  HC_CUDA(cudaMemcpyAsync(op->d_blob_ess, P.d_blob, sizeof(uint32_t) * P.blob_words, cudaMemcpyDeviceToDevice, op->ctx->stream));
  patch_flag_kernel<<<(unsigned)P.n_patch, 128, 0, op->ctx->stream>>>(P.d_blob_off, d_flag, op->d_blob_ess);
  HC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hcurl
