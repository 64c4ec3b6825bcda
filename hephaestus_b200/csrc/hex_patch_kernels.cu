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

namespace hcurl {

struct PatchDeviceEnv {
  unsigned char *smem_;
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
  PatchDeviceEnv env{smem_patch};
  patch_pipeline<PatchDeviceEnv, P2Elem<FE, GENERAL>, FE, ACC, DOT>(env, a);
}

template <int FE, bool ACC, bool DOT>
__global__ void __launch_bounds__(FE) hex_patch_p1_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_patch[];
  PatchDeviceEnv env{smem_patch};
  patch_pipeline<PatchDeviceEnv, P1Elem<FE>, FE, ACC, DOT>(env, a);
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

int hex_patch_kernel_launch(hcurl_op *op, const PatchArgs &a, bool accumulate) {
  hcurl_ctx *ctx = op->ctx;
  const PatchSet &P = op->sp->patch;
  if (!ctx->coop_ok) { set_error("the device does not support cooperative launches (needed by the patch kernels)"); return 1; }
  const size_t smem = patch_smem_bytes(P.unit, P.FE, a.d_cap_bytes, a.l_cap_bytes, a.t_cap_bytes, a.xs_cap_bytes);
  if (P.unit == 16) {
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
