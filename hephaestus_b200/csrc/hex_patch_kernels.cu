// The patch kernels: device environment of the pipeline program (patch_pipeline.h) with the order-1 and order-2
// element sides (hex_patch_p1_elem.h, hex_patch_p2_elem.h), and their launcher.
//
// CurlCurlKernel + VectorFEMassKernel (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24,
// vector_fe_mass_kernel.cpp:21-24) applied as EquationSystem::Mult (/root/reference/src/equation_systems/equation_system.cpp:200-204).
//
// Device environment: cp.async.bulk (TMA) + mbarrier for the blob parts, cp.async (LDGSTS) for the x gather,
// red.global.add.f64 for the shared units, red.release.gpu / ld.acquire.gpu on the arrival counters of the zeroing batches.
// Launch: persistent CTAs, one (order 2, 32 elements: a single warp) per patch in flight, grid = resident CTAs per SM x
// SMs, COOPERATIVE so that the watermark waits cannot dead-lock on a CTA that is not resident.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "patch_dev.cuh"
#include "hex_patch_p1_elem.h"
#include "hex_patch_p2_elem.h"
#include "patch_stream.h"

namespace hcurl {

#ifdef HCURL_PHASE_PROFILE
__device__ unsigned long long g_phase_cycles[32];  // summed over CTAs (thread 0 of each); [15] = iterations (stream program: [0..7] element warp, [8..23] dof group, [31] iterations)
#endif

struct PatchDeviceEnv {
  unsigned char *smem_;
#ifdef HCURL_PHASE_PROFILE
  long long *prof_;  // shared: [0..15] accumulated cycles by phase, [16] last time stamp
  __device__ __forceinline__ void mark(int k) const {
    if (threadIdx.x == 0) { const long long t = clock64(); prof_[k] += t - prof_[16]; prof_[16] = t; if (k == 14) prof_[15]++; }
  }
  __device__ __forceinline__ void prof_begin() const { if (threadIdx.x == 0) { for (int i = 0; i < 16; i++) prof_[i] = 0; prof_[16] = clock64(); } }
  __device__ __forceinline__ void prof_end() const { if (threadIdx.x == 0) for (int i = 0; i < 16; i++) atomicAdd(g_phase_cycles + i, (unsigned long long)prof_[i]); }
#endif
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int64_t block() const { return blockIdx.x; }
  __device__ __forceinline__ int64_t nblocks() const { return gridDim.x; }
  __device__ __forceinline__ unsigned char *smem() const { return smem_; }
  __device__ __forceinline__ uint64_t *mbar() const { return reinterpret_cast<uint64_t *>(smem_); }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void bulk_init(int n) const {
    for (int i = 0; i < n; i++) mbar_init(mbar() + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void bulk_request(int b, void *dst, const void *src, uint32_t bytes) const {
    // generic-proxy accesses of the buffer (by all threads, ordered before this point by a CTA barrier) come before
    // the async-proxy write of the bulk copy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(mbar() + b, bytes);
    if (bytes) bulk_load(mbar() + b, dst, src, bytes);
  }
  __device__ __forceinline__ void bulk_wait(int b, uint32_t parity) const { mbar_wait(mbar() + b, parity); }
  __device__ __forceinline__ void gather8(double *dst, const double *src) const { gather8_async(dst, src); }
  __device__ __forceinline__ void gather16(double2 *dst, const double2 *src) const { gather16_async(dst, src); }
  __device__ __forceinline__ void gather_commit() const { hcurl::gather_commit(); }
  __device__ __forceinline__ void gather_wait() const { gather_wait_all(); }
  __device__ __forceinline__ void red_add(double *p, double v) const { hcurl::red_add(p, v); }
  // arrival on a batch counter.  Release: the zeros written by all threads of the CTA (ordered before this point by CTA
  // barriers) are visible to whoever acquires the counter value
  __device__ __forceinline__ void arrive(uint32_t *p) const {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
  }
  __device__ __forceinline__ uint32_t counter_acquire(const uint32_t *p) const {
    uint32_t f;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(p) : "memory");
    return f;
  }
  __device__ __forceinline__ void backoff(unsigned tries) const { __nanosleep(tries < 4 ? 200u : 1000u); }  // waiting CTAs must not tax the others
  __device__ __forceinline__ uint32_t ldg(const uint32_t *p) const { return __ldg(p); }
  __device__ __forceinline__ double ldg(const double *p) const { return __ldg(p); }
  __device__ __forceinline__ double2 ldg(const double2 *p) const { return __ldg(p); }
  // deterministic for a fixed grid: fixed-order sums inside the CTA, per-CTA partials, the last CTA to arrive adds them
  // in index order (same protocol as block_finish() of the vector kernels, api.cu)
  __device__ __forceinline__ void dot_finish(double dacc, const PatchArgs &a, double *red) const {
    __shared__ bool last;
    const int t = threadIdx.x, nw = (blockDim.x + 31) / 32;
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((t & 31) == 0) red[t >> 5] = dacc;
    __syncthreads();
    if (t == 0) {
      double v = 0.0;
      for (int w = 0; w < nw; w++) v += red[w];
      a.dot_partial[blockIdx.x] = v;
      __threadfence();
      last = atomicInc(a.dot_counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && t < 32) {
      __threadfence();
      double v = 0.0;
      for (int i = t; i < (int)gridDim.x; i += 32) v += ((volatile double *)a.dot_partial)[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (t == 0) *a.dot_out = a.dot_add ? *a.dot_out + v : v;
    }
  }
};

template <int FE, bool ACC, bool DOT, bool GENERAL>
__global__ void __launch_bounds__(FE, 1) hex_patch_p2_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_patch[];
#ifdef HCURL_PHASE_PROFILE
  __shared__ long long prof[17];
  PatchDeviceEnv env{smem_patch, prof};
  env.prof_begin();
  patch_pipeline<PatchDeviceEnv, P2Elem<FE, GENERAL>, FE, ACC, DOT>(env, a);
  env.prof_end();
#else
  PatchDeviceEnv env{smem_patch};
  patch_pipeline<PatchDeviceEnv, P2Elem<FE, GENERAL>, FE, ACC, DOT>(env, a);
#endif
}

template <int FE, bool ACC, bool DOT>
__global__ void __launch_bounds__(FE) hex_patch_p1_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_patch[];
#ifdef HCURL_PHASE_PROFILE
  __shared__ long long prof[17];
  PatchDeviceEnv env{smem_patch, prof};
#else
  PatchDeviceEnv env{smem_patch};
#endif
  patch_pipeline<PatchDeviceEnv, P1Elem<FE>, FE, ACC, DOT>(env, a);
}

// ---- warp-specialised order-2 kernels (patch_stream.h): one CTA per SM ---------------------------------------------------
// 64-element patches: two streams of two element warps + four sum warps + two gather warps (16 warps).  Warps 0-3 are the
// element warps (warp group 0, one per scheduler / FP64 pipe), warps 4-11 the sum warps (warp groups 1, 2), warps 12-15 the
// gather warps (warp group 3): setmaxnreg moves the register file -- the kernel is launched at <= 128 registers per thread,
// the element warp group then takes 248, the sum warp groups drop to 72 and the gather warp group to 56.  (The same program
// with 32-element patches -- four streams -- measured 4 % slower in its two-role form: profiles/README.md; it still runs in
// the host emulation, tests/test_patch_p2_emu.py.)
#ifndef HCURL_STREAM_REG_SUM
#define HCURL_STREAM_REG_SUM 72
#endif
#ifndef HCURL_STREAM_REG_GATHER
#define HCURL_STREAM_REG_GATHER 88
#endif
template <int FE_, int ND_, int NG_, int NSTREAM_>
struct StreamDeviceEnv : PatchDeviceEnv {
  static constexpr int FE = FE_, ND = ND_, NG = NG_, NSTREAM = NSTREAM_, NTHREADS = NSTREAM * (FE + ND + NG);
  static constexpr int REG_ELEMENT = 248, REG_SUM = HCURL_STREAM_REG_SUM, REG_GATHER = HCURL_STREAM_REG_GATHER;
  static constexpr int FLAG0 = 8;  // mbarriers 0..7: bulk copies (D0 D1 D2 D3 L0 T0 T1 L1); 8..13: flags
  static_assert(NSTREAM * FE == 128, "the element warps are exactly warp group 0");
  static_assert((NSTREAM * ND) % 128 == 0 && (NSTREAM * NG) % 128 == 0, "setmaxnreg acts on whole warp groups");
  // setmaxnreg re-distributes the registers the CTA was LAUNCHED with (at most 65536 / NTHREADS per thread, in steps of 8):
  // an increase beyond that pool waits forever
  static_assert(128 * REG_ELEMENT + NSTREAM * ND * REG_SUM + NSTREAM * NG * REG_GATHER <= ((65536 / NTHREADS) & ~7) * NTHREADS, "register pool of the CTA");
  int tid_, s_;
#ifdef HCURL_PHASE_PROFILE
  long long *sprof_;  // shared, per stream: [0..31] cycles by phase, [32] / [33] / [34] last time stamps of the element warp / the sum group / the gather group
  __device__ __forceinline__ void smark(int k) const {
    if (tid_ == 0 || tid_ == FE || tid_ == FE + ND) { long long *last = sprof_ + 32 + (tid_ != 0) + (tid_ == FE + ND); const long long t = clock64(); sprof_[k] += t - *last; *last = t; }
  }
#endif
  __device__ __forceinline__ int tid() const { return tid_; }
  __device__ __forceinline__ int64_t block() const { return (int64_t)blockIdx.x * NSTREAM + s_; }
  __device__ __forceinline__ int64_t nblocks() const { return (int64_t)gridDim.x * NSTREAM; }
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(1 + s_), "n"(FE + ND + NG) : "memory"); }
  __device__ __forceinline__ void group_sync() const { asm volatile("bar.sync %0, %1;" ::"r"(1 + NSTREAM + s_), "n"(ND) : "memory"); }
  __device__ __forceinline__ void ggroup_sync() const { asm volatile("bar.sync %0, %1;" ::"r"(1 + 3 * NSTREAM + s_), "n"(NG) : "memory"); }
  __device__ __forceinline__ void elem_sync() const {
    if (FE == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + 2 * NSTREAM + s_), "n"(FE) : "memory");
  }
  __device__ __forceinline__ void role(int r) const {
    if (r == 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_ELEMENT));
    else if (r == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_SUM));
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_GATHER));
  }
  __device__ __forceinline__ void bulk_request_w(int b, void *dst, const void *src, uint32_t bytes, int) const { bulk_request(b, dst, src, bytes); }
  // flags: mbarriers of the stream with one arrival per phase.  A flag is never signalled twice before its waiter has
  // passed (ring protocol, patch_stream.h), so the one-bit phase parity identifies the n-th signal
  __device__ __forceinline__ void flag_init(int f, int arrivals) const { mbar_init(mbar() + FLAG0 + f, (uint32_t)arrivals); }
  __device__ __forceinline__ void flag_signal(int f) const {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(mbar() + FLAG0 + f)) : "memory");
  }
  __device__ __forceinline__ void flag_wait(int f, uint32_t n) const { mbar_wait(mbar() + FLAG0 + f, n & 1u); }
  // per-CTA partial sum over all warps, then the last-CTA protocol of PatchDeviceEnv::dot_finish
  __device__ __forceinline__ void dot_finish(double dacc, const PatchArgs &a, double *) const {
    __shared__ double red[NTHREADS / 32];
    __shared__ bool last;
    const int t = threadIdx.x;
    for (int o = 16; o > 0; o >>= 1) dacc += __shfl_down_sync(0xffffffffu, dacc, o);
    if ((t & 31) == 0) red[t >> 5] = dacc;
    __syncthreads();
    if (t == 0) {
      double v = 0.0;
      for (int w = 0; w < NTHREADS / 32; w++) v += red[w];
      a.dot_partial[blockIdx.x] = v;
      __threadfence();
      last = atomicInc(a.dot_counter, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && t < 32) {
      __threadfence();
      double v = 0.0;
      for (int i = t; i < (int)gridDim.x; i += 32) v += ((volatile double *)a.dot_partial)[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (t == 0) *a.dot_out = a.dot_add ? *a.dot_out + v : v;
    }
  }
};
typedef StreamDeviceEnv<128, 256, 128, 1> StreamEnv128;  // one stream: 4 element + 8 sum + 4 gather warps on 128-element patches (the default)
typedef StreamDeviceEnv<64, 128, 64, 2> StreamEnv64;     // two streams of 2 + 4 + 2 warps on 64-element patches (when the unit lists of 128 elements do not fit)

// shared memory of ONE stream (a CTA holds 128 / FE of them)
size_t patch_stream_smem_bytes(int FE, uint32_t d_cap, uint32_t l_cap, uint32_t t_cap, uint32_t xs_cap) {
  const size_t one = 128 + 4 * (size_t)d_cap + 2 * (size_t)l_cap + 2 * (size_t)t_cap + 2 * (size_t)xs_cap + (size_t)27 * (FE + 1) * 16;
  return (one + 127) & ~(size_t)127;
}

template <class SEnv, bool ACC, bool DOT, bool GENERAL>
__global__ void __launch_bounds__(SEnv::NTHREADS, 1) hex_patch_p2_stream_kernel(const PatchArgs a, const uint32_t stream_bytes) {
  extern __shared__ __align__(128) unsigned char smem_patch[];
  constexpr int EW = SEnv::FE / 32, NEW = SEnv::NSTREAM * EW;  // element warps per stream / per CTA
  constexpr int NSW = SEnv::NSTREAM * SEnv::ND / 32;            // sum warps per CTA
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SEnv env;
  if (warp < NEW) { env.s_ = warp / EW; env.tid_ = (warp % EW) * 32 + lane; }
  else if (warp < NEW + NSW) { env.s_ = (warp - NEW) % SEnv::NSTREAM; env.tid_ = SEnv::FE + ((warp - NEW) / SEnv::NSTREAM) * 32 + lane; }
  else { env.s_ = (warp - NEW - NSW) % SEnv::NSTREAM; env.tid_ = SEnv::FE + SEnv::ND + ((warp - NEW - NSW) / SEnv::NSTREAM) * 32 + lane; }
  env.smem_ = smem_patch + (size_t)env.s_ * stream_bytes;
#ifdef HCURL_PHASE_PROFILE
  __shared__ long long sprof[4][35];
  env.sprof_ = sprof[env.s_];
  if (env.tid_ == 0) { for (int i = 0; i < 32; i++) sprof[env.s_][i] = 0; sprof[env.s_][32] = sprof[env.s_][33] = sprof[env.s_][34] = clock64(); }
  __syncthreads();
#endif
  patch_stream<SEnv, P2Elem<SEnv::FE, GENERAL>, SEnv::FE, SEnv::ND, SEnv::NG, ACC, DOT>(env, a);
#ifdef HCURL_PHASE_PROFILE
  __syncthreads();
  if (threadIdx.x < 32) for (int s = 0; s < SEnv::NSTREAM; s++) atomicAdd(g_phase_cycles + threadIdx.x, (unsigned long long)sprof[s][threadIdx.x]);
#endif
}

// dynamic shared memory: 4 mbarriers | 2 x (header + unit list) | lists | per-slot table | Xs | EY
size_t patch_smem_bytes(int unit, int FE, uint32_t d_cap, uint32_t l_cap, uint32_t t_cap, uint32_t xs_cap) {
  const int rows = unit == 16 ? 27 : 12;
  return 64 + 2 * (size_t)d_cap + l_cap + t_cap + xs_cap + (size_t)rows * (FE + 1) * unit;
}

// one launch state per (context, kernel variant): the opt-in to large dynamic shared memory and the occupancy are
// per device, and a context is bound to one device
template <class K>
static int launch_variant(hcurl_ctx *ctx, const PatchSet *pset, int var, K k0, K k1, K kd, int nt, const PatchArgs &a, size_t smem, bool acc) {
  if (var < 0 || var >= hcurl_ctx::N_KVAR) { set_error("patch kernel variant out of range"); return 1; }
  if (ctx->kattr_smem[var] < smem || ctx->kattr_per_sm[var] < 1) {
    for (K k : {k0, k1, kd}) {
      HC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    int per_sm = 0;
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd, nt, smem));
    if (per_sm < 1) { set_error("patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    ctx->kattr_smem[var] = smem;
    ctx->kattr_per_sm[var] = per_sm;
  }
  // persistent CTAs: every SM holds per_sm of them, each walks patches blockIdx.x, += gridDim.x
  int64_t grid = (int64_t)ctx->kattr_per_sm[var] * ctx->sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= ctx->kattr_per_sm[var]) grid = (int64_t)v * ctx->sm_count; }
  const int64_t n_run = a.n_patch - a.k_begin;
  if (grid > n_run) grid = n_run;
  K k = k0;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > ctx->n_partial) grid = ctx->n_partial;  // one partial sum per CTA
    k = kd;
  } else if (acc) k = k1;
  PatchArgs args = a;
  if (!acc) {  // every arrival counter receives `grid` arrivals in this launch
    PatchSet &P = const_cast<PatchSet &>(*pset);
    P.arrivals += (uint32_t)grid;
    args.count_target = P.arrivals;
    const int64_t nb = ((n_run + grid - 1) / grid + 7) / 8;
    if (nb > P.n_count) { set_error("patch kernel: arrival counters too few for this grid"); return 1; }
  }
  void *params[] = {&args};
  // cooperative: all CTAs of the grid are resident together, which the flag waits of the pipeline rely on
  HC_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)grid), dim3((unsigned)nt), params, smem, ctx->stream));
  return 0;
}

template <int FE, bool GENERAL>
static int launch_p2(hcurl_ctx *ctx, const PatchSet *ps, int var, const PatchArgs &a, size_t smem, bool acc) {
  return launch_variant(ctx, ps, var, hex_patch_p2_kernel<FE, false, false, GENERAL>, hex_patch_p2_kernel<FE, true, false, GENERAL>,
                        hex_patch_p2_kernel<FE, false, true, GENERAL>, FE, a, smem, acc);
}
template <int FE>
static int launch_p1(hcurl_ctx *ctx, const PatchSet *ps, int var, const PatchArgs &a, size_t smem, bool acc) {
  return launch_variant(ctx, ps, var, hex_patch_p1_kernel<FE, false, false>, hex_patch_p1_kernel<FE, true, false>,
                        hex_patch_p1_kernel<FE, false, true>, FE, a, smem, acc);
}

// warp-specialised order-2 kernels: one CTA per SM, cooperative (the watermark waits need co-residency)
template <class SEnv, bool GENERAL>
static int launch_stream(hcurl_ctx *ctx, const PatchSet *pset, int var, const PatchArgs &a, bool acc) {
  typedef void (*K)(const PatchArgs, const uint32_t);
  const K k0 = hex_patch_p2_stream_kernel<SEnv, false, false, GENERAL>, k1 = hex_patch_p2_stream_kernel<SEnv, true, false, GENERAL>,
          kd = hex_patch_p2_stream_kernel<SEnv, false, true, GENERAL>;
  const size_t one = patch_stream_smem_bytes(SEnv::FE, a.d_cap_bytes, a.l_cap_bytes, a.t_cap_bytes, a.xs_cap_bytes);
  const size_t smem = one * SEnv::NSTREAM;
  if (ctx->kattr_smem[var] < smem || ctx->kattr_per_sm[var] < 1) {
    for (K k : {k0, k1, kd}) {
      HC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    int per_sm = 0;
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd, SEnv::NTHREADS, smem));
    if (per_sm < 1) { set_error("stream patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    ctx->kattr_smem[var] = smem;
    ctx->kattr_per_sm[var] = 1;
  }
  const int64_t n_run = a.n_patch - a.k_begin;
  int64_t grid = ctx->sm_count;
  if (grid * SEnv::NSTREAM > n_run) grid = (n_run + SEnv::NSTREAM - 1) / SEnv::NSTREAM;
  K k = k0;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > ctx->n_partial) grid = ctx->n_partial;  // one partial sum per CTA
    k = kd;
  } else if (acc) k = k1;
  PatchArgs args = a;
  if (!acc) {  // every arrival counter receives one arrival per STREAM of this launch
    PatchSet &P = const_cast<PatchSet &>(*pset);
    const int64_t streams = grid * SEnv::NSTREAM;
    P.arrivals += (uint32_t)streams;
    args.count_target = P.arrivals;
    const int64_t nb = ((n_run + streams - 1) / streams + HCURL_ZERO_BATCH - 1) / HCURL_ZERO_BATCH;
    if (nb > P.n_count) { set_error("patch kernel: arrival counters too few for this grid"); return 1; }
  }
  uint32_t sb = (uint32_t)one;
  void *params[] = {&args, &sb};
  HC_CUDA(cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)grid), dim3(SEnv::NTHREADS), params, smem, ctx->stream));
  return 0;
}

int hex_patch_kernel_launch(hcurl_op *op, const PatchArgs &a, bool accumulate) {
  hcurl_ctx *ctx = op->ctx;
  const PatchSet &P = op->sp->patch;
  if (!ctx->coop_ok) { set_error("the device does not support cooperative launches (needed by the patch kernels)"); return 1; }
  const size_t smem = patch_smem_bytes(P.unit, P.FE, a.d_cap_bytes, a.l_cap_bytes, a.t_cap_bytes, a.xs_cap_bytes);
  if (P.unit == 16) {
    if (P.stream && P.FE == 128) return P.general_orient ? launch_stream<StreamEnv128, true>(ctx, &P, 12, a, accumulate) : launch_stream<StreamEnv128, false>(ctx, &P, 13, a, accumulate);
    if (P.stream) return P.general_orient ? launch_stream<StreamEnv64, true>(ctx, &P, 10, a, accumulate) : launch_stream<StreamEnv64, false>(ctx, &P, 11, a, accumulate);
    if (P.FE == 32) return P.general_orient ? launch_p2<32, true>(ctx, &P, 0, a, smem, accumulate) : launch_p2<32, false>(ctx, &P, 1, a, smem, accumulate);
    if (P.FE == 64) return P.general_orient ? launch_p2<64, true>(ctx, &P, 2, a, smem, accumulate) : launch_p2<64, false>(ctx, &P, 3, a, smem, accumulate);
    set_error("order-2 patch kernel: unsupported patch size %d", P.FE);
    return 1;
  }
  switch (P.FE) {
    case 32: return launch_p1<32>(ctx, &P, 4, a, smem, accumulate);
    case 64: return launch_p1<64>(ctx, &P, 5, a, smem, accumulate);
    case 128: return launch_p1<128>(ctx, &P, 6, a, smem, accumulate);
    case 256: return launch_p1<256>(ctx, &P, 7, a, smem, accumulate);
  }
  set_error("order-1 patch kernel: unsupported patch size %d", P.FE);
  return 1;
}

}  // namespace hcurl

#ifdef HCURL_PHASE_PROFILE
// diagnostic builds only: cycles per pipeline phase summed over the CTAs since the last reset
extern "C" int hcurl_debug_phase_cycles(unsigned long long *out32, int reset) {
  if (out32 && cudaMemcpyFromSymbol(out32, hcurl::g_phase_cycles, 32 * sizeof(unsigned long long)) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[32] = {0}; if (cudaMemcpyToSymbol(hcurl::g_phase_cycles, z, sizeof z) != cudaSuccess) return 1; }
  return 0;
}
#endif
