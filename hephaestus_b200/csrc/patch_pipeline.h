// The persistent patch pipeline shared by the order-1 and order-2 affine hex kernels, written as an
// environment-parametrised per-thread PROGRAM: the same source is the CUDA kernel (hex_patch_kernels.cu: TMA bulk
// copies + mbarriers, cp.async gathers, red.global, release / acquire flags) and, for tests without a GPU, runs on host
// threads (tests/emu/: memcpy, std::barrier, atomics; under ThreadSanitizer it also checks the barrier placement).
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204) with the essential rows / columns of
// FormLinearSystem (equation_system.cpp:132-190: y = x on essential dofs) written by the same kernel.
//
// One thread per element, one CTA per patch of FE elements (patch.cpp), every CTA walks patches block, += nblocks:
//   wait table(i) | pull x_e through the per-slot table, forward transform      [Elem::pull]
//   barrier (staging vector free) | start the x gather of patch i+1 (cp.async)
//   element operator                                                               [Elem::apply]
//   results to the shared result buffer EY                                         [Elem::store]
//   barrier | request table(i+1) | wait lists(i)
//   per-unit sums of the private units: plain stores; essential dofs get y = x here
//   per-unit sums of the shared units: red.global.add.f64
//   gather wait, barrier | request lists(i+1), header + unit list (i+2)
// Nothing is zero-filled before an apply and nothing is fixed up after it: one launch per apply.  A unit touched by
// several patches is owned by the lowest-numbered of them, and the CTA that will process the owner ZEROES it ahead of
// time: at iteration i it zeroes what its patch of iteration i + 2 ZB owns (the unit indices are loaded at the top of
// the iteration, the stores issued behind the pull).  Iterations are grouped in batches of ZB; at the first iteration of
// batch m a CTA ARRIVES for batch m+1 (fence + one RED on the batch's arrival counter: the zeros it wrote for that
// batch during the last ZB iterations are visible), and before its first REDs of batch m it WAITS until every CTA of
// the grid has arrived for batch m -- which happened about ZB iterations earlier: a single acquire load per batch.
// The zeros are still in L2 when the adds arrive.  A unit shared by two patches (every face unit) gets a
// bit-reproducible sum (0 + a + b); units shared by three or more (edges) receive their REDs in arrival order.
// Accumulating applies (y += ...) need no zeros and no watermark.
//
// Deadlock freedom: arrivals never wait for anything of the same or a later batch, and all CTAs are co-resident
// (cooperative launch).
//
// Shared memory:  [64 B: four mbarriers] [2 x d_cap: header + unit list] [l_cap: lists] [t_cap: per-slot table]
//                 [Xs: xs_cap] [EY: Elem::ROWS x (FE+1) units]
//
// Env provides:
//   int tid(); int64_t block(), nblocks(); unsigned char *smem(); void sync();
//   void bulk_init(int n);                                               (thread 0, before the first sync: n mbarriers)
//   void bulk_request(int b, void *dst, const void *src, uint32_t bytes) (thread 0: expect-tx + bulk copy on mbarrier b)
//   void bulk_wait(int b, uint32_t parity);                              (every thread, once per request)
//   void gather8(double *, const double *); void gather16(double2 *, const double2 *); void gather_commit(); void gather_wait();
//   void red_add(double *p, double v);
//   void arrive(uint32_t *counter)  (release: fence + add 1);  uint32_t counter_acquire(const uint32_t *);  void backoff(unsigned tries);
//   uint32_t ldg(const uint32_t *); double ldg(const double *); double2 ldg(const double2 *);
//   void dot_finish(double my_sum, const PatchArgs &a, double *scratch)  (all threads; scratch: shared memory)
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

// ---- units of the dof side -------------------------------------------------------------------------------------------
#ifndef HCURL_SUMS_UNROLL
#define HCURL_SUMS_UNROLL 1  // measured: 0.504 ms against 0.546 with 2 (instruction count and code size beat loads in flight)
#endif
#ifdef HCURL_PHASE_PROFILE  // diagnostic builds only (scripts/phase_profile.sh): clock64 deltas of thread 0 per pipeline phase
#define HCURL_MARK(k) env.mark(k)
#else
#define HCURL_MARK(k) ((void)0)
#endif
#define HCURL_PRAGMA_(x) _Pragma(#x)
#define HCURL_UNROLL(n) HCURL_PRAGMA_(unroll n)

struct UnitSingle {  // order 1: one double
  typedef double T;
  static constexpr uint32_t ESS = PATCH_ESS0;
  template <class Env, bool ACC> static HCURL_BODY void store_free(Env &, T *dst, const T v) { *dst = ACC ? *dst + v : v; }
  template <class Env> static HCURL_BODY void red_free(Env &env, T *dst, const T v) { env.red_add(dst, v); }
  static HCURL_BODY T zero() { return 0.0; }
  static HCURL_BODY T add(const T a, const T b) { return a + b; }
  template <class Env> static HCURL_BODY void gather(Env &env, T *dst, const T *src) { env.gather8(dst, src); }
  static HCURL_BODY T free_part(const T v, const uint32_t m) { return (m & PATCH_ESS0) ? 0.0 : v; }
  // the owner's store: free dofs get the sum, essential dofs y = x; returns x_ess^2 for the fused <x, y>
  // ident = false (several ranks: the caller writes the essential rows after the exchange): essential dofs get 0
  template <class Env, bool ACC> static HCURL_BODY double store(Env &env, T *dst, const T v, const uint32_t m, const T *xsrc, const bool ident) {
    if (!(m & PATCH_ESS0)) { *dst = ACC ? *dst + v : v; return 0.0; }
    const T xv = ident ? env.ldg(xsrc) : 0.0;
    *dst = xv;
    return xv * xv;
  }
  // shared unit with an essential flag: the owner writes y = x (nobody adds to an essential dof)
  template <class Env> static HCURL_BODY double red_ess(Env &env, T *dst, const T v, const uint32_t m, const T *xsrc, const bool ident) {
    if (!(m & PATCH_ESS0)) { env.red_add(dst, v); return 0.0; }
    if (!(m & PATCH_OWNED)) return 0.0;
    const T xv = ident ? env.ldg(xsrc) : 0.0;
    *dst = xv;
    return xv * xv;
  }
};
struct UnitPair {  // order 2: the two dofs along a line, an aligned pair of doubles
  typedef double2 T;
  static constexpr uint32_t ESS = PATCH_ESS0 | PATCH_ESS1;
  template <class Env, bool ACC> static HCURL_BODY void store_free(Env &, T *dst, const T v) {
    if (ACC) { const T o = *dst; *dst = make_double2(o.x + v.x, o.y + v.y); }
    else *dst = v;
  }
  template <class Env> static HCURL_BODY void red_free(Env &env, T *dst, const T v) {
    env.red_add(&dst->x, v.x);
    env.red_add(&dst->y, v.y);
  }
  static HCURL_BODY T zero() { return make_double2(0.0, 0.0); }
  static HCURL_BODY T add(const T a, const T b) { return make_double2(a.x + b.x, a.y + b.y); }
  template <class Env> static HCURL_BODY void gather(Env &env, T *dst, const T *src) { env.gather16(dst, src); }
  static HCURL_BODY T free_part(T v, const uint32_t m) {
    if (m & PATCH_ESS0) v.x = 0.0;
    if (m & PATCH_ESS1) v.y = 0.0;
    return v;
  }
  template <class Env, bool ACC> static HCURL_BODY double store(Env &env, T *dst, const T v, const uint32_t m, const T *xsrc, const bool ident) {
    if (!(m & ESS)) {
      if (ACC) { const T o = *dst; *dst = make_double2(o.x + v.x, o.y + v.y); }
      else *dst = v;
      return 0.0;
    }
    const T xv = ident ? env.ldg(xsrc) : make_double2(0.0, 0.0);  // rare: a pair with an essential dof
    T o = make_double2(0.0, 0.0);
    if (ACC) o = *dst;
    double e2 = 0.0;
    T r;
    if (m & PATCH_ESS0) { r.x = xv.x; e2 += xv.x * xv.x; } else r.x = o.x + v.x;
    if (m & PATCH_ESS1) { r.y = xv.y; e2 += xv.y * xv.y; } else r.y = o.y + v.y;
    *dst = r;
    return e2;
  }
  template <class Env> static HCURL_BODY double red_ess(Env &env, T *dst, const T v, const uint32_t m, const T *xsrc, const bool ident) {
    if (!(m & PATCH_ESS0)) env.red_add(&dst->x, v.x);
    if (!(m & PATCH_ESS1)) env.red_add(&dst->y, v.y);
    if (!(m & ESS) || !(m & PATCH_OWNED)) return 0.0;
    const T xv = ident ? env.ldg(xsrc) : make_double2(0.0, 0.0);
    double e2 = 0.0;
    if (m & PATCH_ESS0) { dst->x = xv.x; e2 += xv.x * xv.x; }
    if (m & PATCH_ESS1) { dst->y = xv.y; e2 += xv.y * xv.y; }
    return e2;
  }
};

// ---- per-unit sums of one phase (0: private units, stored; 1: shared units + overflow slots, RED-added) ----------------
// ESSP = false: no unit of the patch carries an essential flag (header word 20, set with the flags): no per-dof checks
template <class Env, class Unit, int NT, bool ACC, bool DOT, int phase, bool ESSP>
HCURL_BODY double patch_unit_sums(Env &env, const PatchArgs &a, const uint32_t *B, const uint32_t *Lw, const unsigned char *Eb,
                                  const int tid) {
  typedef typename Unit::T T;
  constexpr int HDR = PatchHost::HDR;
  const T *xu = reinterpret_cast<const T *>(a.x);
  T *yu = reinterpret_cast<T *>(a.y);
  const uint32_t n1 = B[4 + 3 * phase], n2 = B[5 + 3 * phase], n4 = B[6 + 3 * phase];
  const uint32_t *dof = B + HDR + (phase ? B[4] + B[5] + B[6] : 0u);
  const uint16_t *e16 = reinterpret_cast<const uint16_t *>(Lw + B[10 + 3 * phase]);
  const uint32_t *e32 = Lw + B[11 + 3 * phase];
  const uint32_t *e64 = Lw + B[12 + 3 * phase];  // two words per unit
  double dess = 0.0;
  auto e_at = [&](const uint32_t off) { return *reinterpret_cast<const T *>(Eb + off); };
  auto emit = [&](const uint32_t m, const T v) {
    const uint32_t q = m & PATCH_IDX;
    if (!ESSP) {
      if (phase == 0) Unit::template store_free<Env, ACC>(env, yu + q, v);
      else Unit::red_free(env, yu + q, v);
    } else {
      const double e2 = phase == 0 ? Unit::template store<Env, ACC>(env, yu + q, v, m, xu + q, a.ess_identity != 0)
                                   : Unit::red_ess(env, yu + q, v, m, xu + q, a.ess_identity != 0);
      if (DOT) dess += e2;
    }
  };
  HCURL_UNROLL(HCURL_SUMS_UNROLL)
  for (uint32_t r = tid; r < n1; r += NT) emit(dof[r], e_at(e16[r]));
  dof += n1;
  HCURL_UNROLL(HCURL_SUMS_UNROLL)
  for (uint32_t r = tid; r < n2; r += NT) {
    const uint32_t o = e32[r];
    emit(dof[r], Unit::add(e_at(o & 0xffffu), e_at(o >> 16)));
  }
  dof += n2;
  HCURL_UNROLL(HCURL_SUMS_UNROLL)
  for (uint32_t r = tid; r < n4; r += NT) {
    const uint32_t m = dof[r];
    const uint32_t ox = e64[2 * r], oy = e64[2 * r + 1];
    T v = Unit::add(Unit::add(e_at(ox & 0xffffu), e_at(ox >> 16)), e_at(oy & 0xffffu));
    if ((oy >> 16) != 0xffffu) v = Unit::add(v, e_at(oy >> 16));
    emit(m, v);
  }
  if (phase == 1) {  // slots beyond four (unstructured meshes)
    const uint32_t nov = B[2];
    const uint32_t *ovf = Lw + B[16];
    const uint32_t *dall = B + HDR;
    for (uint32_t t = tid; t < nov; t += NT) {
      const uint32_t e = ovf[t], m = dall[e >> 16];
      if (ESSP) Unit::red_ess(env, yu + (m & PATCH_IDX), e_at(e & 0xffffu), m & ~PATCH_OWNED, xu + (m & PATCH_IDX), false);
      else Unit::red_free(env, yu + (m & PATCH_IDX), e_at(e & 0xffffu));
    }
  }
  return dess;
}

// ---- the pipeline ----------------------------------------------------------------------------------------------------
// Elem: the element side.  typedef Unit; static constexpr int ROWS (unit rows of the result buffer);
//   void geometry(Env &, const PatchArgs &, int64_t patch, int el)              issue the loads of the geometry factors
//   void pull(Env &, const PatchArgs &, const Unit::T *Xs, const uint16_t *lidx_el)   inputs through the per-slot table (stride FE)
//   double apply<DOT>(Env &, const PatchArgs &)                                  element operator in place; returns x_e^T A_e x_e
//   void store(Env &, const PatchArgs &, Unit::T *EY, const uint16_t *lidx_el, int el)   results to the (swizzled) EY rows
template <class Env, class Elem, int FE, bool ACC, bool DOT>
HCURL_BODY void patch_pipeline(Env &env, const PatchArgs &a) {
  typedef typename Elem::Unit Unit;
  typedef typename Unit::T T;
  constexpr int NT = FE, HDR = PatchHost::HDR;
  unsigned char *smem = env.smem();
  unsigned char *bufD = smem + 64;                               // 2 x (header + unit list)
  unsigned char *bufL = bufD + 2 * (size_t)a.d_cap_bytes;        // lists of the patch being summed
  unsigned char *bufT = bufL + a.l_cap_bytes;                    // per-slot table of the patch being pulled
  T *Xs = reinterpret_cast<T *>(bufT + a.t_cap_bytes);
  T *EY = reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(Xs) + a.xs_cap_bytes);
  const unsigned char *Eb = reinterpret_cast<const unsigned char *>(EY);
  const T *xu = reinterpret_cast<const T *>(a.x);
  const int tid = env.tid(), el = tid;
  double dacc = 0.0;
  const int64_t k0 = a.k_begin + env.block(), kstep = env.nblocks();
  const int n_it = (int)((a.n_patch - k0 + kstep - 1) / kstep);

  // thread 0: request the parts of a patch (offsets in 16-byte units)
  auto request_d = [&](int b, uint32_t o0, uint32_t o1) {
    env.bulk_request(b, bufD + (size_t)b * a.d_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u);
  };
  auto request_l = [&](uint32_t o0, uint32_t o1) { env.bulk_request(2, bufL, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u); };
  auto request_t = [&](uint32_t o0, uint32_t o1) { env.bulk_request(3, bufT, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u); };
  // launch the x gather of the patch whose unit list sits in buffer b
  auto gather = [&](int b) {
    const uint32_t *B = reinterpret_cast<const uint32_t *>(bufD + (size_t)b * a.d_cap_bytes);
    const uint32_t U = B[0];
    const uint32_t *dof = B + HDR;
    if (B[20] == 0) {  // no essential dof in this patch
#pragma unroll 4
      for (uint32_t u = tid; u < U; u += NT) Unit::gather(env, Xs + u, xu + (dof[u] & PATCH_IDX));
    } else {
#pragma unroll 2
      for (uint32_t u = tid; u < U; u += NT) {
        const uint32_t m = dof[u];
        if (!(m & Unit::ESS)) Unit::gather(env, Xs + u, xu + (m & PATCH_IDX));
        else Xs[u] = (m & Unit::ESS) == Unit::ESS ? Unit::zero() : Unit::free_part(env.ldg(xu + (m & PATCH_IDX)), m);  // essential dofs enter as zeros (rare: synchronous)
      }
    }
    env.gather_commit();
  };
  // ---- zeroing ahead.  Batch m = my iterations [m ZB, (m+1) ZB); NB batches in the grid (CTAs with fewer iterations
  // still arrive for every batch, so that each counter receives exactly nblocks arrivals per launch)
#ifndef HCURL_ZERO_BATCH
#define HCURL_ZERO_BATCH 8  // measured on B200 (order 2): 0.454 ms with 8, 0.476 with 4
#endif
  constexpr int ZB = HCURL_ZERO_BATCH, ZU = 6;  // iterations per batch; zero indices a thread keeps in flight across the pull
  const int NB = (int)(((a.n_patch - a.k_begin + kstep - 1) / kstep + ZB - 1) / ZB);
  T *yu = reinterpret_cast<T *>(a.y);
  // prologue: the extents of the 2 ZB owned lists are loaded by 4 ZB threads at once and handed over through shared memory
  // (EY is still unused), so that the index loads of all patches are independent
  auto zero_prologue = [&]() {
    constexpr int NJ = 2 * ZB;
    uint32_t *ext = reinterpret_cast<uint32_t *>(EY);  // [2 NJ]
    for (int j = tid; j < 2 * NJ; j += NT) {
      const int jj = j >> 1;
      ext[j] = jj < n_it ? env.ldg(a.own_off + k0 + (int64_t)jj * kstep + (j & 1)) : 0u;
    }
    env.sync();
#pragma unroll 4
    for (int j = 0; j < NJ; j++) {
      const uint32_t u1 = ext[2 * j + 1];
      for (uint32_t u = ext[2 * j] + tid; u < u1; u += NT) yu[env.ldg(a.own_idx + u)] = Unit::zero();
    }
  };  // (EY is first written behind the barrier that follows)
  // extent of the owned list zeroed in the coming iteration, loaded an iteration ahead (the index loads depend on it)
  uint32_t zn0 = 0, zn1 = 0;
  if (!ACC && 2 * ZB < n_it) { zn0 = env.ldg(a.own_off + k0 + (int64_t)(2 * ZB) * kstep); zn1 = env.ldg(a.own_off + k0 + (int64_t)(2 * ZB) * kstep + 1); }
  int arrived = 0;  // batches [0, arrived) have been arrived for
  bool passed = ACC;  // the watermark of the current batch has been seen
  if (tid == 0) {
    env.bulk_init(4);
    const uint32_t o1 = env.ldg(a.part_off + 2 * k0), o2 = env.ldg(a.part_off + 2 * k0 + 1);
    request_d(0, env.ldg(a.blob_off + k0), o1);
    request_l(o1, o2);
    request_t(o2, env.ldg(a.blob_off + k0 + 1));
    if (n_it > 1) request_d(1, env.ldg(a.blob_off + k0 + kstep), env.ldg(a.part_off + 2 * (k0 + kstep)));
  }
  env.sync();  // mbarriers initialised
  env.bulk_wait(0, 0);
  gather(0);
  if (!ACC) zero_prologue();
  env.gather_wait();
  env.sync();  // Xs of the first patch complete and visible; the zeros of batches 0, 1 issued by every thread
  if (!ACC) {
    if (tid == 0) for (int m = 0; m < 2 && m < NB; m++) env.arrive(a.count + m);
    arrived = 2;
  }

  for (int it = 0; it < n_it; it++) {
    const int b = it & 1;
    const int64_t k = k0 + (int64_t)it * kstep;
    const uint32_t *B = reinterpret_cast<const uint32_t *>(bufD + (size_t)b * a.d_cap_bytes);
    const uint32_t *Lw = reinterpret_cast<const uint32_t *>(bufL);
    const uint16_t *lidx = reinterpret_cast<const uint16_t *>(bufT) + el;
    uint32_t nL = 0, nT = 0, nE = 0, qD = 0, qL = 0;  // thread 0: extents of the parts requested during this iteration
    // zeroing ahead: the owned shared units of my patch of iteration it + 2 ZB (indices loaded now, zeros stored after the pull)
    uint32_t zu[ZU];
    const uint32_t z0 = zn0 + tid, z1 = zn1;
    if (!ACC) {
#pragma unroll
      for (int i = 0; i < ZU; i++) zu[i] = z0 + i * NT < z1 ? env.ldg(a.own_idx + z0 + i * NT) : 0xffffffffu;
      zn0 = zn1 = 0;
      if (it + 1 + 2 * ZB < n_it) { const int64_t pz = k + (int64_t)(2 * ZB + 1) * kstep; zn0 = env.ldg(a.own_off + pz); zn1 = env.ldg(a.own_off + pz + 1); }
    }
    Elem e;
    HCURL_MARK(0);
    e.geometry(env, a, k, el);  // issued now, consumed by the operator
    const bool ok = (uint32_t)el < B[1];
    env.bulk_wait(3, it & 1);  // per-slot table of this patch
    HCURL_MARK(1);
    if (ok) e.pull(env, a, Xs, lidx);
    HCURL_MARK(2);
    env.sync();  // Xs free
    HCURL_MARK(3);
    // the release fence of an arrival waits for the thread's outstanding memory operations: fewer here than behind the
    // gather (there it cost 2 % of the kernel)
    if (!ACC) {
      if (it % ZB == 0) {  // first iteration of batch m: arrive for batch m+1 (its zeros were written during batch m-1)
        const int m = it / ZB;
        if (m >= 1 && m + 1 < NB) {
          if (tid == 0) env.arrive(a.count + m + 1);
          arrived = m + 2;
        }
        passed = false;
      }
    }
    // ---- the x gather of the next patch lands behind the element operator
    if (it + 1 < n_it) {
      env.bulk_wait(b ^ 1, ((it + 1) >> 1) & 1);
      HCURL_MARK(4);
      gather(b ^ 1);
    }
    HCURL_MARK(5);
    if (!ACC) {
      if (z0 < z1) {
#pragma unroll
        for (int i = 0; i < ZU; i++) if (zu[i] != 0xffffffffu) yu[zu[i]] = Unit::zero();
        for (uint32_t u = z0 + ZU * NT; u < z1; u += NT) yu[env.ldg(a.own_idx + u)] = Unit::zero();  // (patches owning more than ZU NT units)
      }
    }
    HCURL_MARK(6);
    if (ok) {
      const double d = e.template apply<DOT>(env, a);
      if (DOT) dacc += d;
    }
    HCURL_MARK(7);
    // thread 0: extents of the parts requested below, loaded HERE (the operator's registers are free again, the result
    // store hides the latency): at the top of the iteration these loads sat in front of the first use of zn0 and made it
    // wait a full global-memory latency per iteration (shared scoreboard; 4 % of the kernel in the ncu source view)
    if (tid == 0) {
      if (it + 1 < n_it) { nL = env.ldg(a.part_off + 2 * (k + kstep)); nT = env.ldg(a.part_off + 2 * (k + kstep) + 1); nE = env.ldg(a.blob_off + k + kstep + 1); }
      if (it + 2 < n_it) { qD = env.ldg(a.blob_off + k + 2 * kstep); qL = env.ldg(a.part_off + 2 * (k + 2 * kstep)); }
    }
    if (ok) e.store(env, a, EY, lidx, el);
    HCURL_MARK(8);
    env.sync();  // element results visible; the per-slot table is free
    if (tid == 0 && it + 1 < n_it) request_t(nT, nE);
    env.bulk_wait(2, it & 1);  // slot-offset lists of this patch
    HCURL_MARK(9);
    const bool essp = B[20] != 0;
    if (essp) {
      const double d = patch_unit_sums<Env, Unit, NT, ACC, DOT, 0, true>(env, a, B, Lw, Eb, tid);
      if (DOT) dacc += d;
    } else patch_unit_sums<Env, Unit, NT, ACC, false, 0, false>(env, a, B, Lw, Eb, tid);
    HCURL_MARK(10);
    if (B[7] + B[8] + B[9] + B[2]) {
      if (!passed) {  // every CTA has arrived for this batch: all zeros under the shared units of its patches are visible
        for (unsigned ns = 0; (int32_t)(env.counter_acquire(a.count + it / ZB) - a.count_target) < 0; ns++) env.backoff(ns);
        passed = true;
      }
      HCURL_MARK(11);
      if (essp) {
        const double d = patch_unit_sums<Env, Unit, NT, ACC, DOT, 1, true>(env, a, B, Lw, Eb, tid);
        if (DOT) dacc += d;
      } else patch_unit_sums<Env, Unit, NT, ACC, false, 1, false>(env, a, B, Lw, Eb, tid);
    }
    HCURL_MARK(12);
    env.gather_wait();
    HCURL_MARK(13);
    env.sync();  // buffer b, the lists and EY are free; Xs of the next patch complete and visible
    if (tid == 0) {
      if (it + 1 < n_it) request_l(nL, nT);
      if (it + 2 < n_it) request_d(b, qD, qL);
    }
    HCURL_MARK(14);
  }
  if (!ACC && tid == 0) for (int m = arrived; m < NB; m++) env.arrive(a.count + m);  // (zeros of these batches: none of mine left)
  if (DOT) env.dot_finish(dacc, a, reinterpret_cast<double *>(EY));
}

}  // namespace hcurl
