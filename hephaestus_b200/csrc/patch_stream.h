// The WARP-SPECIALISED form of the persistent patch pipeline (patch_pipeline.h) for the order-2 affine hex kernel: the
// same patch blobs, element side, per-unit sums, zero-ahead watermark and essential-row handling, but the element side,
// the x gather (with the sums of the shared units) and the sums of the private units of a patch run in DIFFERENT warps,
// connected by four one-bit flags per stream.
//
// Operator and reference: y (+)= scale * (K(alpha) + M(beta)) x, CurlCurlKernel + VectorFEMassKernel
// (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24, vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult
// (/root/reference/src/equation_systems/equation_system.cpp:200-204) with the essential rows / columns of
// FormLinearSystem (equation_system.cpp:132-190).
//
// Why (measured, profiles/README.md round 2): with one warp doing everything a pipeline iteration takes ~14 000 cycles
// of which the element side (pull, modal operator, result store: all the FP64) is ~4 700 and the dof side (x gather,
// zeroing ahead, per-unit sums, REDs) ~8 000, almost all of it exposed latency; at 255 registers only 7-8 such warps
// fit on an SM, so neither side hides the other.  Here a STREAM is FE/32 element warps (248 registers, one thread per
// element of an FE-element patch), ND/32 SUM warps and NG/32 GATHER warps (few registers: setmaxnreg moves the register
// file to the element warps), and an SM runs 128/FE streams -- one element warp per scheduler / FP64 pipe:
//
//   element warp, patch i:  wait T(i), XS_FULL[i%2] | pull | signal XS_FREE[i%2] | operator | wait EY_FREE (sums of patch i-1)
//                           | results to EY | signal EY_FULL, request T(i+2)
//   gather warps, patch i:  the gather of patch i (started an iteration ago) has landed: signal XS_FULL[i%2] | wait
//                           D(i+1), XS_FREE[(i+1)%2]: start the x gather of patch i+1 | wait L(i-1), watermark, EY_FULL(i-1):
//                           sums of the SHARED units of patch i-1 (REDs) | arrive on EY_FREE
//   sum warps, patch i:     arrive (batch) | wait D(i), L(i) | zero ahead | wait EY_FULL | sums of the PRIVATE units (stores) |
//                           arrive on EY_FREE; once the gather group has arrived too: request L(i+2), D(i+4)
//
// How the roles were cut (round 2, phase profiles under profiles/: clock64 per phase): with one dof group doing the gather AND
// the sums (two streams of 64-element patches) an iteration took 7 900 cycles per stream and the element warps waited 1 700
// of them; on ONE stream of 128-element patches the three groups depend on each other patch by patch and their balance
// decides -- the group whose waits were shortest gave work to the one that waited most: gather warps split off (0.407 ms
// per apply), shared sums moved to them (0.378), zeroing moved to the sum warps (0.373), slot lists double-buffered (0.366 =
// 71 % of the measured HBM copy bandwidth).
// The staging vector Xs is a DOUBLE-BUFFERED ring (one FULL / FREE flag pair per buffer: the usual producer / consumer
// protocol, so a one-bit phase parity per flag is enough); the gather warps run as far ahead as the ring allows (16-byte
// scattered copies need ~3 000 cycles to land under load) and the unit lists are requested four patches ahead (bulk
// copies: ~2 000).  The unit list D(i) is read by the gather warps (indices for the gather; counts, indices and flags for
// the shared sums) and by the sum warps; its buffer is requested again by the first sum thread once BOTH groups have
// arrived on EY_FREE(i) (an mbarrier with two arrivals per phase; the element warps wait for the same phase before they
// store patch i+1).  The slot lists L are double-buffered: L(i+2) is requested into the buffer of L(i) at the same moment, a
// whole iteration ahead of its use.  No flag can run two phases ahead of a
// waiter: EY_FULL(i+1) needs EY_FREE(i), which needs the arrivals of both groups that wait for EY_FULL(i) (the host emulation
// checks this for every wait).  Everything else (blob format, ownership and zeroing batches, deadlock freedom by
// co-residency, bit-reproducible two-patch sums) is as described in patch_pipeline.h and patch.cpp.
//
// Shared memory of a stream:  [128 B: mbarriers D0 D1 D2 D3 L0 T0 T1 L1 | 6 flags] [4 x d_cap] [2 x l_cap] [2 x t_cap] [2 x Xs] [EY]
//
// Env (per THREAD of a stream; tid() in [0, FE): the element warps, [FE, FE + ND): the sum group, [FE + ND, FE + ND + NG):
// the gather group), in addition to the interface listed in patch_pipeline.h (block() / nblocks() count STREAMS, smem() is
// the stream's):
//   void role(int r)                              register re-allocation of the calling warp group (device: setmaxnreg); 0 element, 1 sums, 2 gather
//   void group_sync()  void ggroup_sync()         barrier of the sum group / of the gather group
//   void elem_sync()                              barrier of the element warps of the stream (one warp when FE = 32)
//   void bulk_request_w(int b, void *dst, const void *src, uint32_t bytes, int nwait)   as bulk_request; nwait threads will wait
//   void flag_init(int f, int arrivals)  (thread 0, with bulk_init; arrivals = signals per phase)
//   void flag_signal(int f)                       (one thread, after a warp / group barrier)
//   void flag_wait(int f, uint32_t n)             until phase n of the flag is complete (n = 0, 1, ...: every phase waited for in turn)
#pragma once
#include "patch_pipeline.h"

#ifndef HCURL_STREAM_GATHER_UNROLL
#define HCURL_STREAM_GATHER_UNROLL 4
#endif
#ifdef HCURL_PHASE_PROFILE
#define HCURL_SMARK(k) env.smark(k)
#else
#define HCURL_SMARK(k) ((void)0)
#endif

namespace hcurl {

template <class Env, class Elem, int FE, int ND, int NG, bool ACC, bool DOT>
HCURL_BODY void patch_stream(Env &env, const PatchArgs &a) {
  typedef typename Elem::Unit Unit;
  typedef typename Unit::T T;
  constexpr int HDR = PatchHost::HDR;
  enum { MB_D0 = 0, ND_BUF = 4, MB_L0 = 4, MB_T0 = 5, MB_L1 = 7, N_MB = 8 };  // D0 D1 D2 D3 | L0 | T0 T1 | L1
  enum { XS_FULL = 0, XS_FREE = 2, EY_FULL = 4, EY_FREE = 5, N_FLAG = 6 };  // XS_*: + buffer index
  unsigned char *smem = env.smem();
  unsigned char *bufD = smem + 128;
  unsigned char *bufL = bufD + ND_BUF * (size_t)a.d_cap_bytes;
  unsigned char *bufT = bufL + 2 * (size_t)a.l_cap_bytes;  // (the slot lists are double-buffered: those of patch i + 1 are requested while patch i is summed)
  unsigned char *bufX = bufT + 2 * (size_t)a.t_cap_bytes;
  unsigned char *bufE = bufX + 2 * (size_t)a.xs_cap_bytes;
  const int tid = env.tid();
  const int role = tid < FE ? 0 : (tid < FE + ND ? 1 : 2);
  const int64_t k0 = a.k_begin + env.block(), kstep = env.nblocks();
  const int64_t n_it64 = (a.n_patch - k0 + kstep - 1) / kstep;
  const int n_it = n_it64 > 0 ? (int)n_it64 : 0;
  double dacc = 0.0;

  if (tid == 0) {
    env.bulk_init(N_MB);
    for (int f = 0; f < N_FLAG; f++) env.flag_init(f, f == EY_FREE ? 2 : 1);  // EY_FREE: one arrival of the sum group, one of the gather group
  }
  env.sync();  // mbarriers and flags initialised
  env.role(role);

  if (role == 0) {
    // ================================================================ element warps
    const int el = tid;
    auto request_t = [&](int b, int64_t k) {  // per-slot table of patch k into table buffer b
      const uint32_t o0 = env.ldg(a.part_off + 2 * k + 1), o1 = env.ldg(a.blob_off + k + 1);
      env.bulk_request_w(MB_T0 + b, bufT + (size_t)b * a.t_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u, FE);
    };
    if (tid == 0) {
      if (n_it > 0) request_t(0, k0);
      if (n_it > 1) request_t(1, k0 + kstep);
    }
    for (int it = 0; it < n_it; it++) {
      const int b = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);  // how often buffer b has been used before
      const int64_t k = k0 + (int64_t)it * kstep;
      const uint16_t *lidx = reinterpret_cast<const uint16_t *>(bufT + (size_t)b * a.t_cap_bytes) + el;
      const T *Xs = reinterpret_cast<const T *>(bufX + (size_t)b * a.xs_cap_bytes);
      T *EY = reinterpret_cast<T *>(bufE);
      Elem e;
      HCURL_SMARK(0);
      e.geometry(env, a, k, el);  // issued now, consumed by the operator
      env.bulk_wait(MB_T0 + b, use & 1u);
      HCURL_SMARK(1);
      env.flag_wait(XS_FULL + b, use);
      HCURL_SMARK(2);
      e.pull(env, a, Xs, lidx);   // (padding elements of a patch: table entries 0, geometry factors 0 -- harmless, never summed)
      env.elem_sync();
      if (tid == 0) env.flag_signal(XS_FREE + b);
      HCURL_SMARK(3);
      const double d = e.template apply<DOT>(env, a);
      if (DOT) dacc += d;
      HCURL_SMARK(4);
      uint32_t t0 = 0, t1 = 0;  // thread 0: extents of the table requested below (the result store hides the latency)
      if (tid == 0 && it + 2 < n_it) { t0 = env.ldg(a.part_off + 2 * (k + 2 * kstep) + 1); t1 = env.ldg(a.blob_off + k + 2 * kstep + 1); }
      if (it >= 1) env.flag_wait(EY_FREE, (uint32_t)(it - 1));  // the sums of patch it - 1
      HCURL_SMARK(5);
      e.store(env, a, EY, lidx, el);
      env.elem_sync();  // results written, the table buffer is free
      HCURL_SMARK(6);
      if (tid == 0) {
        env.flag_signal(EY_FULL);
        if (it + 2 < n_it) env.bulk_request_w(MB_T0 + b, bufT + (size_t)b * a.t_cap_bytes, a.blob + (size_t)t0 * 4, (t1 - t0) * 16u, FE);
      }
      HCURL_SMARK(7);
    }
  } else if (role == 1) {
    // ================================================================ sum group
    const int dt = tid - FE;
    auto request_d = [&](int b, uint32_t o0, uint32_t o1) {  // (waited for by the gather group and by the sum group)
      env.bulk_request_w(MB_D0 + b, bufD + (size_t)b * a.d_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u, ND + NG);
    };
    auto request_l = [&](int j, uint32_t o0, uint32_t o1) {  // the slot lists of patch j
      env.bulk_request_w((j & 1) ? MB_L1 : MB_L0, bufL + (size_t)(j & 1) * a.l_cap_bytes, a.blob + (size_t)o0 * 4, (o1 - o0) * 16u, ND + NG);
    };
    if (dt == 0 && n_it > 0) {
      const uint32_t o1 = env.ldg(a.part_off + 2 * k0), o2 = env.ldg(a.part_off + 2 * k0 + 1);
      request_d(0, env.ldg(a.blob_off + k0), o1);
      request_l(0, o1, o2);
      if (n_it > 1) request_l(1, env.ldg(a.part_off + 2 * (k0 + kstep)), env.ldg(a.part_off + 2 * (k0 + kstep) + 1));
      for (int j = 1; j < ND_BUF && j < n_it; j++) request_d(j, env.ldg(a.blob_off + k0 + j * kstep), env.ldg(a.part_off + 2 * (k0 + j * kstep)));
    }
    // zeroing ahead (patch_pipeline.h): batches of ZB iterations, one arrival per stream and batch.  This group does it: it
    // waits for the element warps anyway while the gather group is busy all the time
    constexpr int ZB = HCURL_ZERO_BATCH, ZU = (6 * FE + ND - 1) / ND;
    const int NB = (int)(((a.n_patch - a.k_begin + kstep - 1) / kstep + ZB - 1) / ZB);
    T *yu = reinterpret_cast<T *>(a.y);
    uint32_t zn0 = 0, zn1 = 0;
    if (!ACC && 2 * ZB < n_it) { zn0 = env.ldg(a.own_off + k0 + (int64_t)(2 * ZB) * kstep); zn1 = env.ldg(a.own_off + k0 + (int64_t)(2 * ZB) * kstep + 1); }
    int arrived = 0;
    if (!ACC) {  // prologue: zeros under the shared units my patches of iterations [0, 2 ZB) own -- ND / (2 ZB) threads per patch, so
      // that the extent and index loads of all patches are independent
      constexpr int NJ = 2 * ZB, TP = ND / NJ;
      static_assert(ND % NJ == 0, "threads per patch of the zeroing prologue");
      const int j = dt / TP;
      if (j < n_it) {
        const int64_t p = k0 + (int64_t)j * kstep;
        const uint32_t u1 = env.ldg(a.own_off + p + 1);
#pragma unroll 4
        for (uint32_t u = env.ldg(a.own_off + p) + (uint32_t)(dt % TP); u < u1; u += TP) yu[env.ldg(a.own_idx + u)] = Unit::zero();
      }
      env.group_sync();  // the zeros of batches 0, 1 issued by every thread of the group
      if (dt == 0) for (int m = 0; m < 2 && m < NB; m++) env.arrive(a.count + m);
      arrived = 2;
    }
    for (int it = 0; it < n_it; it++) {
      const int64_t k = k0 + (int64_t)it * kstep;
      const int bd = it & (ND_BUF - 1);
      const uint32_t *B = reinterpret_cast<const uint32_t *>(bufD + (size_t)bd * a.d_cap_bytes);
      const uint32_t *Lw = reinterpret_cast<const uint32_t *>(bufL + (size_t)(it & 1) * a.l_cap_bytes);
      const unsigned char *Eb = bufE;
      // zeroing ahead: the owned shared units of my patch of iteration it + 2 ZB
      uint32_t zu[ZU];
      const uint32_t z0 = zn0 + dt, z1 = zn1;
      if (!ACC) {
#pragma unroll
        for (int i = 0; i < ZU; i++) zu[i] = z0 + i * ND < z1 ? env.ldg(a.own_idx + z0 + i * ND) : 0xffffffffu;
        zn0 = zn1 = 0;
        if (it + 1 + 2 * ZB < n_it) { const int64_t pz = k + (int64_t)(2 * ZB + 1) * kstep; zn0 = env.ldg(a.own_off + pz); zn1 = env.ldg(a.own_off + pz + 1); }
      }
      if (!ACC && it % ZB == 0) {  // first iteration of batch m: arrive for batch m+1 (its zeros were written during batch m-1,
        const int m = it / ZB;     // each iteration of which ended with a barrier of the group)
        if (m >= 1 && m + 1 < NB) {
          if (dt == 0) env.arrive(a.count + m + 1);
          arrived = m + 2;
        }
      }
      HCURL_SMARK(8);
      // thread 0 of the group: extents of the parts requested at the end of this iteration
      uint32_t nL = 0, nT = 0, qD = 0, qL = 0;
      if (dt == 0) {
        if (it + 2 < n_it) { nL = env.ldg(a.part_off + 2 * (k + 2 * kstep)); nT = env.ldg(a.part_off + 2 * (k + 2 * kstep) + 1); }
        if (it + ND_BUF < n_it) { qD = env.ldg(a.blob_off + k + ND_BUF * kstep); qL = env.ldg(a.part_off + 2 * (k + ND_BUF * kstep)); }
      }
      HCURL_SMARK(14);
      // ---- sums of the PRIVATE units of this patch (plain stores); the gather group adds the shared ones
      env.bulk_wait(MB_D0 + bd, (uint32_t)(it / ND_BUF) & 1u);
      env.bulk_wait((it & 1) ? MB_L1 : MB_L0, (uint32_t)(it >> 1) & 1u);
      HCURL_SMARK(15);
      if (!ACC && z0 < z1) {
#pragma unroll
        for (int i = 0; i < ZU; i++) if (zu[i] != 0xffffffffu) yu[zu[i]] = Unit::zero();
        for (uint32_t u = z0 + ZU * ND; u < z1; u += ND) yu[env.ldg(a.own_idx + u)] = Unit::zero();  // (patches owning more than ZU ND units)
      }
      HCURL_SMARK(13);
      env.flag_wait(EY_FULL, (uint32_t)it);
      HCURL_SMARK(16);
      if (B[20] != 0) {
        const double d = patch_unit_sums<Env, Unit, ND, ACC, DOT, 0, true>(env, a, B, Lw, Eb, dt);
        if (DOT) dacc += d;
      } else patch_unit_sums<Env, Unit, ND, ACC, false, 0, false>(env, a, B, Lw, Eb, dt);
      HCURL_SMARK(17);
      env.group_sync();
      if (dt == 0) {
        env.flag_signal(EY_FREE);
        // unit buffer bd and the slot lists of this patch are free once the gather group has arrived as well
        env.flag_wait(EY_FREE, (uint32_t)it);
        if (it + 2 < n_it) request_l(it + 2, nL, nT);  // (into the buffer of the slot lists of this patch: a whole iteration ahead of their use)
        if (it + ND_BUF < n_it) request_d(bd, qD, qL);
      }
      HCURL_SMARK(20);
    }
    if (!ACC && dt == 0) for (int m = arrived; m < NB; m++) env.arrive(a.count + m);  // (zeros of these batches: none of mine left)
  } else {
    // ================================================================ gather group
    const int gt = tid - FE - ND;
    const T *xu = reinterpret_cast<const T *>(a.x);
    T *yu = reinterpret_cast<T *>(a.y);
    auto gather = [&](int bd, int bx) {  // the x gather of the patch whose unit list sits in unit buffer bd, into staging buffer bx
      const uint32_t *B = reinterpret_cast<const uint32_t *>(bufD + (size_t)bd * a.d_cap_bytes);
      T *Xs = reinterpret_cast<T *>(bufX + (size_t)bx * a.xs_cap_bytes);
      const uint32_t U = B[0];
      const uint32_t *dof = B + HDR;
      if (B[20] == 0) {  // no essential dof in this patch
HCURL_UNROLL(HCURL_STREAM_GATHER_UNROLL)
        for (uint32_t u = gt; u < U; u += NG) Unit::gather(env, Xs + u, xu + (dof[u] & PATCH_IDX));
      } else {
#pragma unroll 2
        for (uint32_t u = gt; u < U; u += NG) {
          const uint32_t m = dof[u];
          if (!(m & Unit::ESS)) Unit::gather(env, Xs + u, xu + (m & PATCH_IDX));
          else Xs[u] = (m & Unit::ESS) == Unit::ESS ? Unit::zero() : Unit::free_part(env.ldg(xu + (m & PATCH_IDX)), m);  // essential dofs enter as zeros
        }
      }
      env.gather_commit();
    };
    constexpr int ZB = HCURL_ZERO_BATCH;
    bool passed = ACC;
    // sums of the SHARED units of patch j (RED-adds), one patch behind the gather: the element warps store patch j while
    // the gather of patch j + 2 is issued
    auto shared_sums = [&](const int j) {
      const uint32_t *B = reinterpret_cast<const uint32_t *>(bufD + (size_t)(j & (ND_BUF - 1)) * a.d_cap_bytes);  // (waited for when patch j was gathered)
      const uint32_t *Lw = reinterpret_cast<const uint32_t *>(bufL + (size_t)(j & 1) * a.l_cap_bytes);
      env.bulk_wait((j & 1) ? MB_L1 : MB_L0, (uint32_t)(j >> 1) & 1u);
      if (!ACC && j % ZB == 0) passed = false;
      const bool shared = B[7] + B[8] + B[9] + B[2] != 0;
      if (!passed && shared) {  // every stream has arrived for this batch: all zeros under the shared units of its patches are visible
        for (unsigned ns = 0; (int32_t)(env.counter_acquire(a.count + j / ZB) - a.count_target) < 0; ns++) env.backoff(ns);
        passed = true;
      }
      HCURL_SMARK(18);
      env.flag_wait(EY_FULL, (uint32_t)j);  // (also orders this group's EY_FREE arrivals: the element warps have seen EY_FREE(j - 1))
      HCURL_SMARK(23);
      if (shared) {
        if (B[20] != 0) {
          const double d = patch_unit_sums<Env, Unit, NG, ACC, DOT, 1, true>(env, a, B, Lw, bufE, gt);
          if (DOT) dacc += d;
        } else patch_unit_sums<Env, Unit, NG, ACC, false, 1, false>(env, a, B, Lw, bufE, gt);
      }
      HCURL_SMARK(19);
      env.ggroup_sync();
      if (gt == 0) env.flag_signal(EY_FREE);
    };
    if (n_it > 0) {
      env.bulk_wait(MB_D0, 0);
      gather(0, 0);
    }
    for (int it = 0; it < n_it; it++) {
      HCURL_SMARK(9);
      // the gather of patch it (started an iteration ago) has landed: hand it to the element warps at once
      env.gather_wait();
      HCURL_SMARK(21);
      env.ggroup_sync();  // ... for every thread of the group
      if (gt == 0) env.flag_signal(XS_FULL + (it & 1));
      HCURL_SMARK(22);
      // ---- the x gather of patch it + 1 into the other staging buffer, once the element warps have pulled patch it - 1 from it
      if (it + 1 < n_it) {
        const int j = it + 1;
        env.bulk_wait(MB_D0 + (j & (ND_BUF - 1)), (uint32_t)(j / ND_BUF) & 1u);
        HCURL_SMARK(10);
        if (j >= 2) env.flag_wait(XS_FREE + (j & 1), (uint32_t)((j - 2) >> 1));
        HCURL_SMARK(11);
        gather(j & (ND_BUF - 1), j & 1);
      }
      HCURL_SMARK(12);
      if (it >= 1) shared_sums(it - 1);
    }
    if (n_it >= 1) shared_sums(n_it - 1);
  }
  if (DOT) env.dot_finish(dacc, a, reinterpret_cast<double *>(bufE));
}

}  // namespace hcurl
