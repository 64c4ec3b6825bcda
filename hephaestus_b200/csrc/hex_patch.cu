// Patch-based order-2 Nedelec hex apply:  y (+)= scale * (K(alpha) + M(beta)) x.
//
// Same operator as hex_apply.cu / hex_fast.cu (CurlCurlKernel + VectorFEMassKernel,
// /root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24, applied as
// EquationSystem::Mult, /root/reference/src/equation_systems/equation_system.cpp:200-204), organised
// around what ncu showed to bound the element-wise kernels on sm_100a:
//
//  * L1tex lane rate of gathers / FP64 RED atomics  ->  one CTA works on PATCHES of FE adjacent elements
//    (patch.cpp); x is gathered once per unique patch dof into a shared-memory staging vector, the
//    element side pulls its 54 inputs from it through a per-slot table, y is summed per patch dof from
//    a shared-memory E-vector; private dofs are plain stores, only patch-boundary dofs use RED.
//  * shared-memory wavefronts and table bytes of that dof side  ->  the unit of the dof side is a PAIR:
//    the two order-2 dofs along one edge / face line / interior line sit next to each other in the
//    global numbering and are always used together (possibly swapped and sign-flipped by the element's
//    orientation), so gathers, staging, tables, sums and private stores all work on aligned double2.
//  * memory latency (62 % of a CTA's life in the one-patch-per-CTA version, profiles/r1g)  ->  the
//    kernel is PERSISTENT and software-pipelined: the blob of a patch arrives by a TMA bulk copy
//    (cp.async.bulk + mbarrier) two patches ahead, the x gather of the next patch runs as 16-byte
//    cp.async copies while the current patch is computed and scattered; three CTA barriers per patch.
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
// Exchange / result buffers (double2): E2[(c*9 + r)*(FE+1) + e] = slots t = 2r, 2r+1 of component c,
// t = i0 + 2*(i1 + 3*i2) with i0 the open index along axis c: the element side moves 16-byte vectors,
// conflict-free; the result buffer EY is additionally XOR-swizzled inside aligned runs of 8 elements
// (PatchArgs::swz_*) for the dof side's reads.
// Measured state and the variants that were tried and rejected: profiles/README.md.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "patch_dev.cuh"

namespace hcurl {

__constant__ ModalTables cM;

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

#define PROF_MARK(i)                                                                         \
  if (PROF && tid == 0) {                                                                    \
    const long long t_now = clock64();                                                       \
    prof_acc[i] += (unsigned long long)(t_now - t_prev);                                     \
    t_prev = t_now;                                                                          \
  }

// Persistent, software-pipelined patch kernel.  One CTA walks patches k = blockIdx.x, += gridDim.x.
// While patch i is being computed, the blob of patch i+1 (TMA bulk copy, issued an iteration ago)
// is used to launch the x gather of patch i+1 as 16-byte cp.async copies into the staging vector Xs,
// and the blob of patch i+2 is requested when patch i retires.
//   Xs  [max pairs]       patch-local x pairs, position u = order of the blob's dof list
//   E   [27][FE+1]        element slot pairs (double2): modal coefficients exchanged between the component threads
//   EY  [27][FE+1]        element slot pairs: element results, summed per dof pair
// (two buffers so that a patch needs three CTA barriers, not five)
template <int FE, int MINB, bool ACC, bool PROF, bool DOT>
__global__ void __maxnreg__(168)
hex_patch_modal_p2_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = 3 * FE, ST = FE + 1;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);  // [2]
  unsigned char *blobB = smem_raw + 16;
  double2 *Xs = reinterpret_cast<double2 *>(blobB + 2 * (size_t)a.blob_cap_bytes);
  double2 *E2 = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);  // modal exchange
  double2 *EY = E2 + 27 * ST;                                                                            // element results
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
  double2 *y2 = reinterpret_cast<double2 *>(a.y);
  const int tid = threadIdx.x;
  long long t_prev = PROF ? clock64() : 0;
  unsigned long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
  double dacc = 0.0;  // DOT: my share of x^T A x, summed in modal space (x_e^T A_e x_e = Am^T Yh)
  const int64_t k0 = a.k_begin + blockIdx.x, kstep = gridDim.x;
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);
  // component c of patch element el, in the rotated frame (c, c+1, c+2)
  const int c = tid / FE, el = tid - c * FE;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int oK11 = sym6(c1, c1) * FE + el, oK10 = sym6(c1, c) * FE + el, oK12 = sym6(c1, c2) * FE + el, oK22 = sym6(c2, c2) * FE + el,
            oK20 = sym6(c2, c) * FE + el, oM00 = (6 + sym6(c, c)) * FE + el, oM01 = (6 + sym6(c, c1)) * FE + el,
            oM02 = (6 + sym6(c, c2)) * FE + el;

  // request the blob of patch k into buffer b (thread 0)
  auto request = [&](int b, uint32_t o0, uint32_t o1) {
    const uint32_t bytes = (o1 - o0) * 16u;
    mbar_expect_tx(mbar + b, bytes);
    bulk_load(mbar + b, blobB + (size_t)b * a.blob_cap_bytes, a.blob + (size_t)o0 * 4, bytes);
  };
  // launch the x gather of the patch whose blob sits in buffer b
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    const uint32_t U = B[0];
    const uint32_t *dof = B + PatchHost::HDR;
#pragma unroll 2
    for (uint32_t u = tid; u < U; u += NT) {
      const uint32_t m = dof[u];
      if (!(m & (PATCH_ESS0 | PATCH_ESS1))) gather16_async(Xs + u, x2 + (m & PATCH_IDX));
      else {  // essential dofs enter as zeros (rare: synchronous)
        double2 v = make_double2(0.0, 0.0);
        if ((m & (PATCH_ESS0 | PATCH_ESS1)) != (PATCH_ESS0 | PATCH_ESS1)) {
          v = __ldg(x2 + (m & PATCH_IDX));
          if (m & PATCH_ESS0) v.x = 0.0;
          if (m & PATCH_ESS1) v.y = 0.0;
        }
        Xs[u] = v;
      }
    }
    gather_commit();
  };
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    request(0, __ldg(a.blob_off + k0), __ldg(a.blob_off + k0 + 1));
    if (n_it > 1) request(1, __ldg(a.blob_off + k0 + kstep), __ldg(a.blob_off + k0 + kstep + 1));
  }
  __syncthreads();  // mbarriers initialised
  mbar_wait(mbar, 0);
  gather(0);
  gather_wait_all();
  __syncthreads();  // Xs of the first patch complete and visible
  PROF_MARK(0)  // prologue

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.blob_cap_bytes);
    uint32_t q0 = 0, q1 = 0;  // blob extent of the patch requested when this one retires
    if (tid == 0 && it + 2 < n_it) { q0 = __ldg(a.blob_off + k + 2 * kstep); q1 = __ldg(a.blob_off + k + 2 * kstep + 1); }
    // affine geometry factors of my element: issued now, consumed after the forward transform
    const double *g = a.edata + (size_t)k * 12 * FE;
    double K11 = __ldg(g + oK11), K10 = __ldg(g + oK10), K12 = __ldg(g + oK12), K22 = __ldg(g + oK22), K20 = __ldg(g + oK20),
           M00 = __ldg(g + oM00), M01 = __ldg(g + oM01), M02 = __ldg(g + oM02);
    const bool ok = (uint32_t)el < B[13];
    // ---- pull the element inputs through the per-slot table, forward transform
    double Am[18];
    uint32_t lx[9];
    uint32_t fl = 0;
    if (ok) {
      const uint16_t *lidx = reinterpret_cast<const uint16_t *>(B + B[12]) + (c * 9) * FE + el;
#pragma unroll
      for (int r = 0; r < 9; r++) { lx[r] = lidx[r * FE]; fl |= lx[r]; }
    }
    // warp-uniform: no element of this warp sees a pair swapped or sign-flipped (consistently oriented meshes)
    const bool plain = __all_sync(0xffffffffu, (fl & 0xe000u) == 0u);
    if (ok) {
      double X[18];
      if (plain) {
#pragma unroll
        for (int r = 0; r < 9; r++) { const double2 v = Xs[lx[r]]; X[2 * r] = v.x; X[2 * r + 1] = v.y; }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) pair_to_local(Xs[lx[r] & 0x1fffu], lx[r], X[2 * r], X[2 * r + 1]);
      }
      modal_fwd(X, Am);
      double2 *Eown = E2 + (c * 9) * ST + el;
#pragma unroll
      for (int r = 0; r < 9; r++) Eown[r * ST] = make_double2(Am[2 * r], Am[2 * r + 1]);
    }
    __syncthreads();  // modal coefficients visible; Xs free
    PROF_MARK(2)  // pull + forward transform
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
          Bm[IB(m0, 0, m2)] = v.x;
          Bm[IB(m0, 1, m2)] = v.y;
        }
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) {
          const double2 v = E2c[(m0 + 3 * m1) * ST];
          Cm[IC(m0, m1, 0)] = v.x;
          Cm[IC(m0, m1, 1)] = v.y;
        }
      const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
      K11 *= sc; K10 *= sc; K12 *= sc; K22 *= sc; K20 *= sc; M00 *= sc; M01 *= sc; M02 *= sc;
      double Yh[18];
      modal_op(Am, Bm, Cm, M00, M01, M02, K11, K10, K12, K22, K20, K12, Yh);
      if (DOT) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int t = 0; t < 18; t += 2) { d0 = fma(Am[t], Yh[t], d0); d1 = fma(Am[t + 1], Yh[t + 1], d1); }
        dacc += d0 + d1;
      }
      modal_bwd(Yh, Y);
    }
    // ---- start the x gather of the next patch: it lands while this patch is scattered
    if (it + 1 < n_it) {
      mbar_wait(mbar + (b ^ 1), ((it + 1) >> 1) & 1);
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
        for (int r = 0; r < 9; r++) Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = local_to_pair(Y[2 * r], Y[2 * r + 1], lx[r]);
      }
    }
    __syncthreads();  // element results visible; everyone has read the modal coefficients
    PROF_MARK(3)  // exchange + modal operator + backward transform + next gather issue + results to shared memory
    // ---- sum the slots of every patch pair; private pairs are plain 16-byte stores, shared ones RED
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + PatchHost::HDR;
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(B + B[8]);
      const uint32_t *e32 = B + B[9];
      const uint2 *e64 = reinterpret_cast<const uint2 *>(B + B[10]);
      const uint32_t *ovf = B + B[11];
      constexpr uint32_t SKIP = PATCH_ESS0 | PATCH_ESS1;
#define E_AT(off) (*reinterpret_cast<const double2 *>(Eb + (off)))
      auto emit = [&](const uint32_t m, const double2 v) {
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & (SKIP | PATCH_SHARED))) {
          if (ACC) { const double2 o = *dst; *dst = make_double2(o.x + v.x, o.y + v.y); }
          else *dst = v;
        } else if ((m & SKIP) != SKIP) {
          if (m & PATCH_SHARED) {
            if (!(m & PATCH_ESS0)) red_add(&dst->x, v.x);
            if (!(m & PATCH_ESS1)) red_add(&dst->y, v.y);
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
        for (int i = 0; i < UN; i++) v[i] = E_AT(o[i]);
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
          const double2 p = E_AT(o[i] & 0xffffu), q = E_AT(o[i] >> 16);
          v[i] = make_double2(p.x + q.x, p.y + q.y);
        }
#pragma unroll
        for (int i = 0; i < UN; i++) emit(m[i], v[i]);
      }
      dof += n2;
      for (uint32_t r = tid; r < n4; r += NT) {
        const uint32_t m = dof[r];
        const uint2 o = e64[r];
        const double2 p = E_AT(o.x & 0xffffu), q = E_AT(o.x >> 16), s = E_AT(o.y & 0xffffu);
        double2 v = make_double2((p.x + q.x) + s.x, (p.y + q.y) + s.y);
        if ((o.y >> 16) != 0xffffu) { const double2 t = E_AT(o.y >> 16); v.x += t.x; v.y += t.y; }
        emit(m, v);
      }
      dof = B + PatchHost::HDR;
      for (uint32_t t = tid; t < nov; t += NT) {  // slots beyond four (unstructured meshes)
        const uint32_t e = ovf[t], m = dof[e >> 16];
        const double2 v = E_AT(e & 0xffffu);
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & PATCH_ESS0)) red_add(&dst->x, v.x);
        if (!(m & PATCH_ESS1)) red_add(&dst->y, v.y);
      }
#undef E_AT
    }
    gather_wait_all();
    __syncthreads();  // blob buffer b and EY are free; Xs of the next patch complete and visible
    PROF_MARK(5)  // dof sums + stores + wait for the x gather
    if (tid == 0 && it + 2 < n_it) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      request(b, q0, q1);
    }
  }
  if (PROF && tid == 0)
    for (int i = 0; i < 6; i++) atomicAdd(a.prof + i, prof_acc[i]);
  if (DOT) {
    // deterministic for a fixed grid: fixed-order sums inside the CTA, per-CTA partials, the last CTA to arrive
    // adds them in index order (same protocol as block_finish() of the vector kernels, api.cu)
    double *red = reinterpret_cast<double *>(E2);  // all patches retired: E is free
    __shared__ bool last;
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((tid & 31) == 0) red[tid >> 5] = dacc;
    __syncthreads();
    if (tid == 0) {
      double v = 0.0;
      for (int w = 0; w < NT / 32; w++) v += red[w];
      a.dot_partial[blockIdx.x] = v;
      __threadfence();
      last = atomicInc(a.dot_counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && tid < 32) {
      __threadfence();
      double v = 0.0;
      for (int i = tid; i < (int)gridDim.x; i += 32) v += ((volatile double *)a.dot_partial)[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (tid == 0) *a.dot_out = a.dot_add ? *a.dot_out + v : v;
    }
  }
}

// Variant with the blob split by lifetime (HCURL_PATCH_SPLIT=1 when the space is created; NOT YET MEASURED -- written at
// the end of round 1 after the GPU budget was spent): the pair list (needed from the gather of a patch
// until its sums) stays double-buffered, the slot-offset lists (sums only) and the per-slot table (pull only) get one
// buffer each and are requested one iteration ahead; rows of the two 27-row buffers are not padded.  44 KB of shared
// memory and 128 registers: five 96-thread CTAs per SM instead of four.
template <int FE, bool ACC, bool DOT>
__global__ void __maxnreg__(128)
hex_patch_split_p2_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NT = 3 * FE, ST = FE;  // no row padding: the element side is conflict-free anyway, the swizzle serves the dof side
  constexpr uint32_t T_CAP = (27 * FE * 2 + 15) & ~15;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);  // [0], [1]: pair-list buffers; [2]: slot-offset lists; [3]: per-slot table
  unsigned char *blobB = smem_raw + 32;                     // 2 x (header + pair list)
  unsigned char *listB = blobB + 2 * (size_t)a.d_cap_bytes; // e16 | e32 | e64 | ovf of the patch being summed
  unsigned char *tabB = listB + a.l_cap_bytes;              // lidx of the patch being pulled
  double2 *Xs = reinterpret_cast<double2 *>(tabB + T_CAP);
  double2 *E2 = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);  // modal exchange
  double2 *EY = E2 + 27 * ST;                                                                            // element results
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const double2 *x2 = reinterpret_cast<const double2 *>(a.x);
  double2 *y2 = reinterpret_cast<double2 *>(a.y);
  const int tid = threadIdx.x;
  double dacc = 0.0;  // DOT: my share of x^T A x, summed in modal space (x_e^T A_e x_e = Am^T Yh)
  const int64_t k0 = a.k_begin + blockIdx.x, kstep = gridDim.x;
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);
  // component c of patch element el, in the rotated frame (c, c+1, c+2)
  const int c = tid / FE, el = tid - c * FE;
  const int c1 = c == 2 ? 0 : c + 1, c2 = c == 0 ? 2 : c - 1;
  const int oK11 = sym6(c1, c1) * FE + el, oK10 = sym6(c1, c) * FE + el, oK12 = sym6(c1, c2) * FE + el, oK22 = sym6(c2, c2) * FE + el,
            oK20 = sym6(c2, c) * FE + el, oM00 = (6 + sym6(c, c)) * FE + el, oM01 = (6 + sym6(c, c1)) * FE + el,
            oM02 = (6 + sym6(c, c2)) * FE + el;

  // the three parts of a patch blob (offsets in 16-byte units: begin | lists | table | end), thread 0
  auto request = [&](int b, uint32_t o0, uint32_t oL) {  // header + pair list -> buffer b
    const uint32_t bytes = (oL - o0) * 16u;
    mbar_expect_tx(mbar + b, bytes);
    bulk_load(mbar + b, blobB + (size_t)b * a.d_cap_bytes, a.blob + (size_t)o0 * 4, bytes);
  };
  auto request_lists = [&](uint32_t oL, uint32_t oT) {
    const uint32_t bytes = (oT - oL) * 16u;
    mbar_expect_tx(mbar + 2, bytes);
    if (bytes) bulk_load(mbar + 2, listB, a.blob + (size_t)oL * 4, bytes);
  };
  auto request_table = [&](uint32_t oT, uint32_t o1) {
    const uint32_t bytes = (o1 - oT) * 16u;
    mbar_expect_tx(mbar + 3, bytes);
    bulk_load(mbar + 3, tabB, a.blob + (size_t)oT * 4, bytes);
  };
  // launch the x gather of the patch whose blob sits in buffer b
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.d_cap_bytes);
    const uint32_t U = B[0];
    const uint32_t *dof = B + PatchHost::HDR;
#pragma unroll 2
    for (uint32_t u = tid; u < U; u += NT) {
      const uint32_t m = dof[u];
      if (!(m & (PATCH_ESS0 | PATCH_ESS1))) gather16_async(Xs + u, x2 + (m & PATCH_IDX));
      else {  // essential dofs enter as zeros (rare: synchronous)
        double2 v = make_double2(0.0, 0.0);
        if ((m & (PATCH_ESS0 | PATCH_ESS1)) != (PATCH_ESS0 | PATCH_ESS1)) {
          v = __ldg(x2 + (m & PATCH_IDX));
          if (m & PATCH_ESS0) v.x = 0.0;
          if (m & PATCH_ESS1) v.y = 0.0;
        }
        Xs[u] = v;
      }
    }
    gather_commit();
  };
  if (tid == 0) {
    for (int i = 0; i < 4; i++) mbar_init(mbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    request(0, __ldg(a.blob_off + k0), __ldg(a.part_off + 2 * k0));
    request_lists(__ldg(a.part_off + 2 * k0), __ldg(a.part_off + 2 * k0 + 1));
    request_table(__ldg(a.part_off + 2 * k0 + 1), __ldg(a.blob_off + k0 + 1));
    if (n_it > 1) request(1, __ldg(a.blob_off + k0 + kstep), __ldg(a.part_off + 2 * (k0 + kstep)));
  }
  __syncthreads();  // mbarriers initialised
  mbar_wait(mbar, 0);
  gather(0);
  gather_wait_all();
  __syncthreads();  // Xs of the first patch complete and visible

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(blobB + (size_t)b * a.d_cap_bytes);
    uint32_t q0 = 0, q1 = 0;  // pair-list extent of patch it+2, requested when this one retires
    uint32_t nL = 0, nT = 0, nE = 0;  // lists | table | end of patch it+1 (requested after the sums / after the pull)
    if (tid == 0 && it + 2 < n_it) { q0 = __ldg(a.blob_off + k + 2 * kstep); q1 = __ldg(a.part_off + 2 * (k + 2 * kstep)); }
    if (tid == 0 && it + 1 < n_it) {
      nL = __ldg(a.part_off + 2 * (k + kstep)); nT = __ldg(a.part_off + 2 * (k + kstep) + 1); nE = __ldg(a.blob_off + k + kstep + 1);
    }
    // affine geometry factors of my element: issued now, consumed after the forward transform
    const double *g = a.edata + (size_t)k * 12 * FE;
    double K11 = __ldg(g + oK11), K10 = __ldg(g + oK10), K12 = __ldg(g + oK12), K22 = __ldg(g + oK22), K20 = __ldg(g + oK20),
           M00 = __ldg(g + oM00), M01 = __ldg(g + oM01), M02 = __ldg(g + oM02);
    const bool ok = (uint32_t)el < B[13];
    // ---- pull the element inputs through the per-slot table, forward transform
    double Am[18];
    uint32_t lx[9];
    uint32_t fl = 0;
    mbar_wait(mbar + 3, it & 1);  // per-slot table of this patch
    if (ok) {
      const uint16_t *lidx = reinterpret_cast<const uint16_t *>(tabB) + (c * 9) * FE + el;
#pragma unroll
      for (int r = 0; r < 9; r++) { lx[r] = lidx[r * FE]; fl |= lx[r]; }
    }
    // warp-uniform: no element of this warp sees a pair swapped or sign-flipped (consistently oriented meshes)
    const bool plain = __all_sync(0xffffffffu, (fl & 0xe000u) == 0u);
    if (ok) {
      double X[18];
      if (plain) {
#pragma unroll
        for (int r = 0; r < 9; r++) { const double2 v = Xs[lx[r]]; X[2 * r] = v.x; X[2 * r + 1] = v.y; }
      } else {
#pragma unroll
        for (int r = 0; r < 9; r++) pair_to_local(Xs[lx[r] & 0x1fffu], lx[r], X[2 * r], X[2 * r + 1]);
      }
      modal_fwd(X, Am);
      double2 *Eown = E2 + (c * 9) * ST + el;
#pragma unroll
      for (int r = 0; r < 9; r++) Eown[r * ST] = make_double2(Am[2 * r], Am[2 * r + 1]);
    }
    __syncthreads();  // modal coefficients visible; Xs and the per-slot table free
    if (tid == 0 && it + 1 < n_it) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      request_table(nT, nE);
    }
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
          Bm[IB(m0, 0, m2)] = v.x;
          Bm[IB(m0, 1, m2)] = v.y;
        }
#pragma unroll
      for (int m1 = 0; m1 < 3; m1++)
#pragma unroll
        for (int m0 = 0; m0 < 3; m0++) {
          const double2 v = E2c[(m0 + 3 * m1) * ST];
          Cm[IC(m0, m1, 0)] = v.x;
          Cm[IC(m0, m1, 1)] = v.y;
        }
      const double sc = 0.25 * a.scale;  // 1/4: the open transforms are applied without their 1/2
      K11 *= sc; K10 *= sc; K12 *= sc; K22 *= sc; K20 *= sc; M00 *= sc; M01 *= sc; M02 *= sc;
      double Yh[18];
      modal_op(Am, Bm, Cm, M00, M01, M02, K11, K10, K12, K22, K20, K12, Yh);
      if (DOT) {
        double d0 = 0.0, d1 = 0.0;
#pragma unroll
        for (int t = 0; t < 18; t += 2) { d0 = fma(Am[t], Yh[t], d0); d1 = fma(Am[t + 1], Yh[t + 1], d1); }
        dacc += d0 + d1;
      }
      modal_bwd(Yh, Y);
    }
    // ---- start the x gather of the next patch: it lands while this patch is scattered
    if (it + 1 < n_it) {
      mbar_wait(mbar + (b ^ 1), ((it + 1) >> 1) & 1);
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
        for (int r = 0; r < 9; r++) Eown[r * ST + (elx ^ a.swz_row[c * 9 + r])] = local_to_pair(Y[2 * r], Y[2 * r + 1], lx[r]);
      }
    }
    __syncthreads();  // element results visible; everyone has read the modal coefficients
    // ---- sum the slots of every patch pair; private pairs are plain 16-byte stores, shared ones RED
    {
      const uint32_t n1 = B[1], n2 = B[2], n4 = B[3], nov = B[7];
      const uint32_t *dof = B + PatchHost::HDR;
      mbar_wait(mbar + 2, it & 1);  // slot-offset lists of this patch
      const uint32_t *Lw = reinterpret_cast<const uint32_t *>(listB) - B[8];  // the lists start at word B[8] of the blob
      const uint16_t *e16 = reinterpret_cast<const uint16_t *>(Lw + B[8]);
      const uint32_t *e32 = Lw + B[9];
      const uint2 *e64 = reinterpret_cast<const uint2 *>(Lw + B[10]);
      const uint32_t *ovf = Lw + B[11];
      constexpr uint32_t SKIP = PATCH_ESS0 | PATCH_ESS1;
#define E_AT(off) (*reinterpret_cast<const double2 *>(Eb + (off)))
      auto emit = [&](const uint32_t m, const double2 v) {
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & (SKIP | PATCH_SHARED))) {
          if (ACC) { const double2 o = *dst; *dst = make_double2(o.x + v.x, o.y + v.y); }
          else *dst = v;
        } else if ((m & SKIP) != SKIP) {
          if (m & PATCH_SHARED) {
            if (!(m & PATCH_ESS0)) red_add(&dst->x, v.x);
            if (!(m & PATCH_ESS1)) red_add(&dst->y, v.y);
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
        for (int i = 0; i < UN; i++) v[i] = E_AT(o[i]);
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
          const double2 p = E_AT(o[i] & 0xffffu), q = E_AT(o[i] >> 16);
          v[i] = make_double2(p.x + q.x, p.y + q.y);
        }
#pragma unroll
        for (int i = 0; i < UN; i++) emit(m[i], v[i]);
      }
      dof += n2;
      for (uint32_t r = tid; r < n4; r += NT) {
        const uint32_t m = dof[r];
        const uint2 o = e64[r];
        const double2 p = E_AT(o.x & 0xffffu), q = E_AT(o.x >> 16), s = E_AT(o.y & 0xffffu);
        double2 v = make_double2((p.x + q.x) + s.x, (p.y + q.y) + s.y);
        if ((o.y >> 16) != 0xffffu) { const double2 t = E_AT(o.y >> 16); v.x += t.x; v.y += t.y; }
        emit(m, v);
      }
      dof = B + PatchHost::HDR;
      for (uint32_t t = tid; t < nov; t += NT) {  // slots beyond four (unstructured meshes)
        const uint32_t e = ovf[t], m = dof[e >> 16];
        const double2 v = E_AT(e & 0xffffu);
        double2 *dst = y2 + (m & PATCH_IDX);
        if (!(m & PATCH_ESS0)) red_add(&dst->x, v.x);
        if (!(m & PATCH_ESS1)) red_add(&dst->y, v.y);
      }
#undef E_AT
    }
    gather_wait_all();
    __syncthreads();  // pair-list buffer b, the lists and EY are free; Xs of the next patch complete and visible
    if (tid == 0 && it + 1 < n_it) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      request_lists(nL, nT);
      if (it + 2 < n_it) request(b, q0, q1);
    }
  }
  if (DOT) {
    // deterministic for a fixed grid: fixed-order sums inside the CTA, per-CTA partials, the last CTA to arrive
    // adds them in index order (same protocol as block_finish() of the vector kernels, api.cu)
    double *red = reinterpret_cast<double *>(E2);  // all patches retired: E is free
    __shared__ bool last;
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((tid & 31) == 0) red[tid >> 5] = dacc;
    __syncthreads();
    if (tid == 0) {
      double v = 0.0;
      for (int w = 0; w < NT / 32; w++) v += red[w];
      a.dot_partial[blockIdx.x] = v;
      __threadfence();
      last = atomicInc(a.dot_counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && tid < 32) {
      __threadfence();
      double v = 0.0;
      for (int i = tid; i < (int)gridDim.x; i += 32) v += ((volatile double *)a.dot_partial)[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (tid == 0) *a.dot_out = a.dot_add ? *a.dot_out + v : v;
    }
  }
}

// ---- setup ---------------------------------------------------------------------------------------
// affine geometry factors in patch order: S = alpha J^T J/det (6), M = beta adj adj^T/det (6)
__global__ void patch_edata_kernel(const int64_t n_pos, const int FE, const int32_t *__restrict__ perm, const double *__restrict__ xyz,
                                   const double *__restrict__ alpha, const double *__restrict__ beta,
                                   double *__restrict__ edata) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pos) return;
  const int32_t e = perm[j];
  double *out = edata + (j / FE) * 12 * FE + j % FE;  // [patch][12][FE]
  if (e < 0) {
    for (int s = 0; s < 12; s++) out[s * FE] = 0.0;
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
      out[s * FE] = al * (c[i][0] * c[jj][0] + c[i][1] * c[jj][1] + c[i][2] * c[jj][2]);
      out[(6 + s) * FE] = be * (ad[i][0] * ad[jj][0] + ad[i][1] * ad[jj][1] + ad[i][2] * ad[jj][2]);
    }
}

// OR the essential flags of both doubles into the pair words of every patch blob (one CTA per patch)
__global__ void patch_flag_kernel(const uint32_t *__restrict__ blob_off, const double *__restrict__ flag,
                                  uint32_t *__restrict__ blob) {
  uint32_t *B = blob + (size_t)blob_off[blockIdx.x] * 4;
  const uint32_t U = B[0];
  for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
    uint32_t m = B[PatchHost::HDR + u];
    const double2 f = reinterpret_cast<const double2 *>(flag)[m & PATCH_IDX];
    if (f.x != 0.0) m |= PATCH_ESS0;
    if (f.y != 0.0) m |= PATCH_ESS1;
    B[PatchHost::HDR + u] = m;
  }
}

// single-dof units (order 1): one flag per unit
__global__ void patch_flag8_kernel(const uint32_t *__restrict__ blob_off, const double *__restrict__ flag, uint32_t *__restrict__ blob) {
  uint32_t *B = blob + (size_t)blob_off[blockIdx.x] * 4;
  const uint32_t U = B[0];
  for (uint32_t u = threadIdx.x; u < U; u += blockDim.x) {
    const uint32_t m = B[PatchHost::HDR + u];
    if (flag[m & PATCH_IDX] != 0.0) B[PatchHost::HDR + u] = m | PATCH_ESS0;
  }
}

unsigned long long *g_patch_prof = nullptr;  // set by hcurl_debug_patch_profile (debug.cu)
// orthonormal shifted Legendre on [0,1]: L0 = 1, L1 = sqrt3 (2x-1), L2 = sqrt5 (6x^2-6x+1);
// T[m][j] = int_0^1 L_m l_j, by a 5-point Gauss rule (exact).  Host only; returns 1 when the tables do not have the
// symmetric structure the kernels hard-wire.
int compute_modal_tables(ModalTables &T) {
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
  const double tol = 1e-14;
  if (fabs(To[0] - 0.5) > tol || fabs(To[1] - 0.5) > tol || fabs(To[2] + 0.5) > tol || fabs(To[3] - 0.5) > tol ||
      fabs(Tc[0] - Tc[2]) > tol || fabs(Tc[3] + Tc[5]) > tol || fabs(Tc[4]) > tol || fabs(Tc[6] - Tc[8]) > tol) {
    set_error("modal tables do not have the expected symmetric structure");
    return 1;
  }
  T.a0 = 0.5 * (Tc[0] + Tc[2]); T.a1 = Tc[1]; T.b = 0.5 * (Tc[5] - Tc[3]); T.c0 = 0.5 * (Tc[6] + Tc[8]); T.c1 = Tc[7];
  T.d1 = 2.0 * sqrt(3.0);
  T.d2 = 2.0 * sqrt(15.0);
  return 0;
}

}  // namespace hcurl
extern "C" int hcurl_host_modal_tables_p2(double *out7) {
  hcurl::ModalTables T;
  if (!out7 || hcurl::compute_modal_tables(T)) return 1;
  const double v[7] = {T.a0, T.a1, T.b, T.c0, T.c1, T.d1, T.d2};
  memcpy(out7, v, sizeof(v));
  return 0;
}
namespace hcurl {

static bool g_modal_set = false;
static int set_modal_tables() {
  if (g_modal_set) return 0;
  ModalTables T;
  if (compute_modal_tables(T)) return 1;
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

// dynamic shared memory of the patch kernel: 2 mbarriers | 2 blob buffers | Xs | E | EY
static size_t patch_smem_bytes(int FE, uint32_t max_blob_words, uint32_t max_units) {
  const size_t blob_cap = ((size_t)max_blob_words * 4 + 15u) & ~(size_t)15u;
  return 16 + 2 * blob_cap + (size_t)max_units * 16 + 2 * (size_t)27 * (FE + 1) * 16;
}

// The dof side of the patch kernel works on aligned PAIRS of doubles.  In the rotated slot layout the
// slots 2r, 2r+1 of an element are the two order-2 dofs along one line (open index 0, 1); MFEM numbers
// the dofs of an edge / face / interior consecutively, so they are the doubles 2q, 2q+1 of the
// extended vector -- in this order, or swapped when the element sees the line reversed -- each with its
// own sign.  Checked here for every element slot pair; a map that does not have this structure (or an
// odd-length vector) simply does not get the patch path.
// Order-1 spaces (hex_patch_p1.cu): units are single doubles, one thread per element.  Opt-in (HCURL_PATCH_P1=1) until
// the kernel's device environment has been run on a GPU; affine elements only (the check is done here, on the host).
static int hex_patch_build_p1(hcurl_space *sp, const double *elem_xyz, const int32_t *elem_ldof, const int32_t *l2e) {
  const char *on = getenv("HCURL_PATCH_P1");
  if (!(on && on[0] == '1')) return 0;
  int FE = 128;
  if (const char *s = getenv("HCURL_PATCH_FE")) FE = atoi(s);
  if (FE != 32 && FE != 64 && FE != 128 && FE != 256) { set_error("HCURL_PATCH_FE must be 32, 64, 128 or 256 for order 1"); return 1; }
  const int64_t ne = sp->n_elem;
  std::vector<double> cen((size_t)ne * 3);
  bool affine = true;
  for (int64_t e = 0; e < ne; e++) {
    const double *X = elem_xyz + e * 24;  // MFEM corner order: 0 1 2 3 bottom face, 4 5 6 7 top
    double h2 = 0.0, dev2 = 0.0;
    for (int d = 0; d < 3; d++) {
      double c = 0.0;
      for (int v = 0; v < 8; v++) c += X[v * 3 + d];
      cen[(size_t)e * 3 + d] = 0.125 * c;
      const double ex = X[3 + d] - X[d], ey = X[9 + d] - X[d], ez = X[12 + d] - X[d];
      h2 = std::max(h2, std::max(ex * ex, std::max(ey * ey, ez * ez)));
      const double pred[8] = {X[d], X[d] + ex, X[d] + ex + ey, X[d] + ey, X[d] + ez, X[d] + ex + ez, X[d] + ex + ey + ez, X[d] + ey + ez};
      for (int v = 0; v < 8; v++) dev2 = std::max(dev2, (X[v * 3 + d] - pred[v]) * (X[v * 3 + d] - pred[v]));
    }
    if (dev2 > 1e-26 * h2) { affine = false; break; }  // same tolerance as fast_affine_kernel (hex_fast.cu)
  }
  if (!affine) return 0;
  std::vector<uint32_t> code((size_t)ne * 12);
  hex_slot_codes(1, ne, elem_ldof, l2e, code.data());
  std::vector<uint8_t> xch;
  if (sp->n_nbr > 0) {
    xch.assign((size_t)sp->n_ext(), 0);
    for (int64_t u = sp->n_tdof; u < sp->n_ext(); u++) xch[(size_t)u] = 1;
    for (int32_t d : sp->own_ext) xch[(size_t)d] = 1;
  }
  PatchHost H;
  if (build_patches(12, ne, FE, FE + 1, 8, sp->n_ext(), cen.data(), code.data(), xch.empty() ? nullptr : xch.data(), H)) return 1;
  if (H.max_dofs > 8191) return 0;  // 13-bit staging positions
  HC_CUDA(cudaSetDevice(sp->ctx->device));
  PatchSet &P = sp->patch;
  P.FE = FE; P.unit = 8; P.affine = true; P.stride = H.stride; P.n_patch = H.n_patch; P.n_boundary = H.n_boundary;
  P.blob_words = (int64_t)H.blob.size(); P.max_blob_words = H.max_blob_words; P.max_dofs = H.max_dofs;
  for (size_t i = 0; i < 27; i++) P.swz_row[i] = i < H.swz_row.size() ? H.swz_row[i] : 0;
  for (size_t i = 0; i < 16; i++) P.swz_grp[i] = i < H.swz_grp.size() ? H.swz_grp[i] : 0;
  const uint32_t blob_cap = (H.max_blob_words * 4 + 15u) & ~15u, xs_cap = (H.max_dofs * 8u + 15u) & ~15u;
  if (patch_p1_smem_bytes(FE, blob_cap, xs_cap) > 227 * 1024) return 0;
  if (upload(P.d_perm, H.perm) || upload(P.d_blob_off, H.blob_off) || upload(P.d_blob, H.blob) || upload(P.d_shared, H.shared_lines)) return 1;
  P.n_shared = (int64_t)H.shared_lines.size();
  sp->has_patches = true;
  return 0;
}

int hex_patch_build(hcurl_space *sp, const double *elem_xyz, const int32_t *elem_ldof, const int32_t *l2e) {
  if (sp->geom == HCURL_GEOM_HEX && sp->order == 1 && sp->geom_order == 1) return hex_patch_build_p1(sp, elem_xyz, elem_ldof, l2e);
  if (sp->geom != HCURL_GEOM_HEX || sp->order != 2 || sp->geom_order != 1) return 0;
  const char *off = getenv("HCURL_NO_PATCH");
  if (off && off[0] == '1') return 0;
  int FE = 32;  // 32-element patches, 4 CTAs of 96 threads per SM: fastest of 32 / 64 / 128 (profiles/README.md)
  if (const char *s = getenv("HCURL_PATCH_FE")) FE = atoi(s);
  if (FE != 32 && FE != 64 && FE != 128) { set_error("HCURL_PATCH_FE must be 32, 64 or 128"); return 1; }
  const int np = sp->nd / 2;
  const int64_t ne = sp->n_elem;
  if (sp->n_ext() & 1) return 0;
  std::vector<uint32_t> code((size_t)ne * np);
  if (hex_p2_pair_codes(ne, elem_ldof, l2e, code.data())) return 0;  // not aligned pairs: no patch path for this space
  std::vector<double> cen((size_t)ne * 3);
  for (int64_t e = 0; e < ne; e++)
    for (int d = 0; d < 3; d++) {
      double c = 0.0;
      for (int v = 0; v < 8; v++) c += elem_xyz[(e * 8 + v) * 3 + d];
      cen[(size_t)e * 3 + d] = 0.125 * c;
    }
  // pairs that take part in the inter-rank exchange: ghosts, and owned pairs other ranks hold as ghosts
  std::vector<uint8_t> xch;
  if (sp->n_nbr > 0) {
    xch.assign((size_t)(sp->n_ext() / 2), 0);
    for (int64_t u = sp->n_tdof / 2; u < sp->n_ext() / 2; u++) xch[(size_t)u] = 1;
    for (int32_t d : sp->own_ext) xch[(size_t)(d / 2)] = 1;
  }
  const char *spl = getenv("HCURL_PATCH_SPLIT");
  const bool split = spl && spl[0] == '1' && FE == 32;  // experimental variant, see hex_patch_split_p2_kernel
  PatchHost H;
  if (build_patches(np, ne, FE, split ? FE : FE + 1, 16, sp->n_ext() / 2, cen.data(), code.data(), xch.empty() ? nullptr : xch.data(), H)) return 1;
  HC_CUDA(cudaSetDevice(sp->ctx->device));
  PatchSet &P = sp->patch;
  P.FE = FE; P.stride = H.stride; P.n_patch = H.n_patch; P.n_boundary = H.n_boundary;
  const char *envk = getenv("HCURL_PATCH_ENV");
  P.env_kernel = envk && envk[0] == '1' && FE == 32;  // the Env-program restatement of the kernel (hex_patch_p2e.cu), opt-in
  const char *halfk = getenv("HCURL_PATCH_HALF");
  P.half_kernel = halfk && (halfk[0] == '1' || halfk[0] == '2') && FE == 32 && !split;  // two threads per element component, opt-in
  P.half_regs = (halfk && halfk[0] == '2') ? 80 : 96;
  for (size_t i = 0; i < code.size() && !P.general_orient; i++) {
    const uint32_t f = code[i] >> 29;
    P.general_orient = f != 0u && f != 7u;
  }
  if (const char *s = getenv("HCURL_PATCH_GENERAL")) P.general_orient = P.general_orient || s[0] == '1';  // tests
  const char *sglk = getenv("HCURL_PATCH_SINGLE");
  P.single_kernel = sglk && sglk[0] == '1' && (FE == 32 || FE == 64) && !split;  // one thread per element, opt-in
  P.split = split; P.max_d_words = H.max_d_words; P.max_l_words = H.max_l_words;
  if (split && upload(P.d_part_off, H.part_off)) return 1;
  P.blob_words = (int64_t)H.blob.size(); P.max_blob_words = H.max_blob_words; P.max_dofs = H.max_dofs;
  for (int i = 0; i < 27; i++) P.swz_row[i] = H.swz_row[(size_t)i];
  for (size_t i = 0; i < 16; i++) P.swz_grp[i] = i < H.swz_grp.size() ? H.swz_grp[i] : 0;
  if (patch_smem_bytes(FE, H.max_blob_words, H.max_dofs) > 227 * 1024) return 0;  // patch path unavailable
  if (upload(P.d_perm, H.perm) || upload(P.d_blob_off, H.blob_off) || upload(P.d_blob, H.blob) || upload(P.d_shared, H.shared_lines)) return 1;
  P.n_shared = (int64_t)H.shared_lines.size();
  sp->has_patches = true;
  return 0;
}

void hex_patch_free(hcurl_space *sp) {
  PatchSet &P = sp->patch;
  cudaFree(P.d_perm); cudaFree(P.d_blob_off); cudaFree(P.d_blob); cudaFree(P.d_shared); cudaFree(P.d_part_off);
  cudaFree(sp->d_xyzP);
  P = PatchSet();
  sp->has_patches = false;
}

bool hex_patch_eligible(const hcurl_op *op) {
  const hcurl_space *sp = op->sp;
  // the modal operator is exact where the (p+1)-point rule is: that is the rule it stands in for
  // either term may be absent (statics: curl-curl only; mass solves: mass only): its geometry factors are zero
  if (!(sp->has_patches && op->single_rule && op->qc == sp->order + 1 && op->coef_layout == HCURL_COEF_PER_ELEM &&
        (op->has_alpha || op->has_beta)))
    return false;
  if (op->kernel_sel == 1 || op->kernel_sel == 2) return false;
  if (sp->patch.unit == 8) return sp->patch.affine;  // order 1: classified on the host when the patches were built
  // affine meshes only (modal operator); checked lazily by hex_fast_prepare_space
  if (!sp->fast_ready && hex_fast_prepare_space(const_cast<hcurl_space *>(sp))) return false;
  return sp->all_affine;
}

template <int FE, int MINB>
static int launch_fe(const PatchArgs &a, int64_t n_patch, int sm_count, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_modal_p2_kernel<FE, MINB, false, false, false>;
  auto k1 = hex_patch_modal_p2_kernel<FE, MINB, true, false, false>;
  auto kp = hex_patch_modal_p2_kernel<FE, MINB, false, true, false>;
  auto kd = hex_patch_modal_p2_kernel<FE, MINB, false, false, true>;
  static size_t attr = 0;
  static int per_sm = 0;
  if (attr < smem) {
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k0, 3 * FE, smem));
    if (per_sm < 1) { set_error("patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  // persistent CTAs: every SM holds per_sm of them, each walks patches blockIdx.x, += gridDim.x
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_patch) grid = n_patch;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    kd<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  } else if (a.prof) kp<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  else if (acc) k1<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  return 0;
}

// split-blob variant: 2 mbarrier pairs | 2 pair-list buffers | lists | table | Xs | E | EY (unpadded rows)
static size_t patch_split_smem_bytes(int FE, const PatchSet &P) {
  const size_t d_cap = ((size_t)P.max_d_words * 4 + 15u) & ~(size_t)15u, l_cap = ((size_t)P.max_l_words * 4 + 15u) & ~(size_t)15u;
  return 32 + 2 * d_cap + l_cap + (((size_t)27 * FE * 2 + 15) & ~(size_t)15) + (size_t)P.max_dofs * 16 + 2 * (size_t)27 * FE * 16;
}
template <int FE>
static int launch_split(const PatchArgs &a, int64_t n_run, int sm_count, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_split_p2_kernel<FE, false, false>;
  auto k1 = hex_patch_split_p2_kernel<FE, true, false>;
  auto kd = hex_patch_split_p2_kernel<FE, false, true>;
  static size_t attr = 0;
  static int per_sm = 0;
  if (attr < smem) {
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd, 3 * FE, smem));
    if (per_sm < 1) { set_error("split patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_run) grid = n_run;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    kd<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  } else if (acc) k1<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  return 0;
}

int hex_patch_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale, bool accumulate, double *d_dot,
                     int64_t k_begin, int64_t k_end, bool dot_add) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  PatchSet &P = sp->patch;
  if (P.unit == 16 && set_modal_tables()) return 1;
  const int64_t n_pos = P.n_patch * P.FE;
  if (!op->edataP_valid) {
    if (!op->d_edataP) HC_CUDA(cudaMalloc(&op->d_edataP, sizeof(double) * 12 * n_pos));
    patch_edata_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, ctx->stream>>>(n_pos, P.FE, P.d_perm, sp->d_xyz, op->has_alpha ? op->d_alpha : nullptr,
                                                                                op->has_beta ? op->d_beta : nullptr, op->d_edataP);
    HC_CUDA(cudaGetLastError());
    op->edataP_valid = true;
  }
  PatchArgs a;
  a.n_patch = k_end < 0 ? P.n_patch : k_end; a.k_begin = k_begin; a.dot_add = dot_add ? 1 : 0; a.blob_off = P.d_blob_off;
  if (a.n_patch <= a.k_begin) return 0;
  a.blob = (constrained && op->d_blob_ess) ? op->d_blob_ess : P.d_blob;
  a.blob_cap_bytes = (P.max_blob_words * 4 + 15u) & ~15u;
  a.xs_cap_bytes = P.unit == 8 ? ((P.max_dofs * 8u + 15u) & ~15u) : P.max_dofs * 16u;
  a.edata = op->d_edataP; a.x = d_x; a.y = d_y; a.scale = scale;
  a.prof = g_patch_prof;
  memcpy(a.swz_row, P.swz_row, sizeof(a.swz_row)); memcpy(a.swz_grp, P.swz_grp, sizeof(a.swz_grp));
  a.dot_out = d_dot; a.dot_partial = ctx->d_partial; a.dot_counter = ctx->d_counter;
  if (d_dot && ctx->n_partial < 8 * ctx->sm_count) { set_error("reduction scratch too small for the patch grid"); return 1; }
  if (a.prof && accumulate) a.prof = nullptr;
  if (P.unit == 8) {  // order 1
    a.prof = nullptr; a.part_off = nullptr; a.d_cap_bytes = a.l_cap_bytes = 0;
    if (hex_patch_p1_launch(op, a, accumulate)) return 1;
    HC_CUDA(cudaGetLastError());
    ctx->launches++;
    return 0;
  }
  a.part_off = P.d_part_off;
  a.d_cap_bytes = (P.max_d_words * 4 + 15u) & ~15u;
  a.l_cap_bytes = (P.max_l_words * 4 + 15u) & ~15u;
  const size_t smem = P.split ? patch_split_smem_bytes(P.FE, P) : patch_smem_bytes(P.FE, P.max_blob_words, P.max_dofs);
  if (P.env_kernel || P.half_kernel || P.single_kernel) {
    a.prof = nullptr;
    if (hex_patch_p2e_launch(op, a, accumulate)) return 1;
    HC_CUDA(cudaGetLastError());
    ctx->launches++;
    return 0;
  }
  int rc = 1;
  // 168 registers per thread: 2 CTAs of 192 threads per SM.  (A 3-CTA variant capped at 112 / 104 registers spilled
  // and ran 45 % slower, profiles/README.md.)
  const int64_t n_run = a.n_patch - a.k_begin;
  if (P.split) rc = launch_split<32>(a, n_run, ctx->sm_count, smem, accumulate, ctx->stream);
  else if (P.FE == 32) rc = launch_fe<32, 4>(a, n_run, ctx->sm_count, smem, accumulate, ctx->stream);
  else if (P.FE == 64) rc = launch_fe<64, 2>(a, n_run, ctx->sm_count, smem, accumulate, ctx->stream);
  else if (P.FE == 128) rc = launch_fe<128, 1>(a, n_run, ctx->sm_count, smem, accumulate, ctx->stream);
  if (rc) return rc;
  HC_CUDA(cudaGetLastError());
  ctx->launches++;
  return 0;
}

// eight lanes clear one 128-byte line (the last line of the vector may be partial)
__global__ void patch_zero_shared_kernel(const int64_t n_lines, const uint32_t *__restrict__ line, double2 *__restrict__ y2,
                                         const int64_t n_pairs) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_lines * 8) return;
  const int64_t q = (int64_t)line[t >> 3] * 8 + (t & 7);
  if (q < n_pairs) y2[q] = make_double2(0.0, 0.0);
}

// the same for single-dof units (order 1): sixteen lanes per 128-byte line
__global__ void patch_zero_shared8_kernel(const int64_t n_lines, const uint32_t *__restrict__ line, double *__restrict__ y, const int64_t n) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_lines * 16) return;
  const int64_t q = (int64_t)line[t >> 4] * 16 + (t & 15);
  if (q < n) y[q] = 0.0;
}

// Private pairs are stored, shared pairs are RED-added by the patch kernel: only the latter need to start from zero.
// Whole 128-byte lines holding a shared pair are cleared (full-sector writes; private pairs in them are stored later by
// the kernel anyway) instead of all of y.
int hex_patch_zero_shared(hcurl_op *op, double *d_y) {
  PatchSet &P = op->sp->patch;
  if (P.n_shared == 0) return 0;
  if (P.unit == 8) {
    const int64_t n = P.n_shared * 16;
    patch_zero_shared8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, op->ctx->stream>>>(P.n_shared, P.d_shared, d_y, op->sp->n_ext());
    HC_CUDA(cudaGetLastError());
    op->ctx->launches++;
    return 0;
  }
  const int64_t nt = P.n_shared * 8;
  patch_zero_shared_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, op->ctx->stream>>>(P.n_shared, P.d_shared, reinterpret_cast<double2 *>(d_y),
                                                                                     op->sp->n_ext() / 2);
  HC_CUDA(cudaGetLastError());
  op->ctx->launches++;
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
  if (P.unit == 8) patch_flag8_kernel<<<(unsigned)P.n_patch, 128, 0, op->ctx->stream>>>(P.d_blob_off, d_flag, op->d_blob_ess);
  else patch_flag_kernel<<<(unsigned)P.n_patch, 128, 0, op->ctx->stream>>>(P.d_blob_off, d_flag, op->d_blob_ess);
  HC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hcurl
