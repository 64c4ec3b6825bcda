// Order-1 Nedelec hex patch kernel: the CUDA execution environment of hex_patch_p1_body.h and its launcher.
//
// STATUS: written after the round's GPU budget was spent.  The program itself (hex_patch_p1_body.h) is verified on host
// threads against the oracle (tests/test_patch_p1_emu.py: apply, essential dofs, accumulation, fused dot product,
// unstructured valence-6 meshes, ThreadSanitizer-clean); what has NOT run yet is this file's device environment.
// Therefore opt-in: HCURL_PATCH_P1=1 when the space is created; the GPU parity tests for it are opt-in too
// (HCURL_TEST_PATCH_P1=1, tests/test_gpu_apply.py).  Without the switch order 1 keeps using hex_apply.cu.
//
// One thread per element, FE = 128 elements per patch by default (HCURL_PATCH_FE = 32 / 64 / 128 / 256): about
// 8.5 KB of blob (x 2 buffers), 4.6 KB of staging vector and 12.4 KB of result buffer per CTA -> 6 CTAs of 128 threads
// per SM on the structured cube.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "patch_dev.cuh"
#include "hex_patch_p1_body.h"

namespace hcurl {

struct P1DeviceEnv {
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
    bulk_load(mbar() + b, dst, src, bytes);
  }
  __device__ __forceinline__ void bulk_wait(int b, uint32_t parity) const { mbar_wait(mbar() + b, parity); }
  __device__ __forceinline__ void gather8(double *dst, const double *src) const { gather8_async(dst, src); }
  __device__ __forceinline__ void gather_commit() const { hcurl::gather_commit(); }
  __device__ __forceinline__ void gather_wait() const { gather_wait_all(); }
  __device__ __forceinline__ void red_add(double *p, double v) const { hcurl::red_add(p, v); }
  __device__ __forceinline__ uint32_t ldg(const uint32_t *p) const { return __ldg(p); }
  __device__ __forceinline__ double ldg(const double *p) const { return __ldg(p); }
  // deterministic for a fixed grid: fixed-order sums inside the CTA, per-CTA partials, the last CTA to arrive adds them
  // in index order (same protocol as the order-2 kernel and block_finish() of the vector kernels, api.cu)
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

template <int FE, bool ACC, bool DOT>
__global__ void __launch_bounds__(FE) hex_patch_modal_p1_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_p1[];
  P1DeviceEnv env{smem_p1};
  hex_patch_p1_body<P1DeviceEnv, FE, ACC, DOT>(env, a);
}

// dynamic shared memory: 2 mbarriers | 2 blob buffers | Xs | EY
size_t patch_p1_smem_bytes(int FE, uint32_t blob_cap_bytes, uint32_t xs_cap_bytes) {
  return 16 + 2 * (size_t)blob_cap_bytes + xs_cap_bytes + (size_t)12 * (FE + 1) * 8;
}

template <int FE>
static int launch_p1(const PatchArgs &a, int64_t n_run, int sm_count, int n_partial, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_modal_p1_kernel<FE, false, false>;
  auto k1 = hex_patch_modal_p1_kernel<FE, true, false>;
  auto kd = hex_patch_modal_p1_kernel<FE, false, true>;
  static size_t attr = 0;
  static int per_sm = 0;
  if (attr < smem) {
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd, FE, smem));
    if (per_sm < 1) { set_error("order-1 patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  // persistent CTAs: every SM holds per_sm of them, each walks patches blockIdx.x, += gridDim.x
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_run) grid = n_run;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > n_partial) grid = n_partial;  // one partial sum per CTA
    kd<<<(unsigned)grid, FE, smem, st>>>(a);
  } else if (acc) k1<<<(unsigned)grid, FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, FE, smem, st>>>(a);
  return 0;
}

int hex_patch_p1_launch(hcurl_op *op, const PatchArgs &a, bool accumulate) {
  hcurl_ctx *ctx = op->ctx;
  const PatchSet &P = op->sp->patch;
  const size_t smem = patch_p1_smem_bytes(P.FE, a.blob_cap_bytes, a.xs_cap_bytes);
  const int64_t n_run = a.n_patch - a.k_begin;
  if (a.dot_out && ctx->n_partial < 1) { set_error("no reduction scratch for the fused dot product"); return 1; }
  switch (P.FE) {
    case 32: return launch_p1<32>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
    case 64: return launch_p1<64>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
    case 128: return launch_p1<128>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
    case 256: return launch_p1<256>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
  }
  set_error("order-1 patch kernel: unsupported patch size %d", P.FE);
  return 1;
}

}  // namespace hcurl
