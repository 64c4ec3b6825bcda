// Order-2 patch kernel from the Env program hex_patch_p2_body.h: device environment and launcher.
//
// STATUS: opt-in (HCURL_PATCH_ENV=1 at space creation, FE = 32; with HCURL_PATCH_SPLIT=1 the split-blob variant).  The
// program is verified on host threads against the oracle (tests/test_patch_p2_emu.py) and is a statement-by-statement
// restatement of hex_patch.cu's kernels, which stay the measured default until this one has run on a GPU.
#include <cstdlib>
#include <cstring>

#include "internal.h"
#include "patch_dev.cuh"
#include "hex_patch_p2_body.h"
#include "hex_patch_p2h_body.h"
#include "hex_patch_p2s_body.h"

namespace hcurl {

__constant__ ModalTables cM2;

struct P2DeviceEnv {
  unsigned char *smem_;
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int64_t block() const { return blockIdx.x; }
  __device__ __forceinline__ int64_t nblocks() const { return gridDim.x; }
  __device__ __forceinline__ unsigned char *smem() const { return smem_; }
  __device__ __forceinline__ uint64_t *mbar() const { return reinterpret_cast<uint64_t *>(smem_); }
  template <class T> __device__ __forceinline__ const T &tables() const { return cM2; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
  __device__ __forceinline__ void bulk_init(int n) const {
    for (int i = 0; i < n; i++) mbar_init(mbar() + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void bulk_request(int b, void *dst, const void *src, uint32_t bytes) const {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(mbar() + b, bytes);
    if (bytes) bulk_load(mbar() + b, dst, src, bytes);
  }
  __device__ __forceinline__ void bulk_wait(int b, uint32_t parity) const { mbar_wait(mbar() + b, parity); }
  __device__ __forceinline__ void gather16(double2 *dst, const double2 *src) const { gather16_async(dst, src); }
  __device__ __forceinline__ void gather_commit() const { hcurl::gather_commit(); }
  __device__ __forceinline__ void gather_wait() const { gather_wait_all(); }
  __device__ __forceinline__ void red_add(double *p, double v) const { hcurl::red_add(p, v); }
  __device__ __forceinline__ bool warp_all(bool pred) const { return __all_sync(0xffffffffu, pred); }
  __device__ __forceinline__ double shfl_xor(double v, int lane_mask) const { return __shfl_xor_sync(0xffffffffu, v, lane_mask); }
  __device__ __forceinline__ uint32_t ldg(const uint32_t *p) const { return __ldg(p); }
  __device__ __forceinline__ double ldg(const double *p) const { return __ldg(p); }
  __device__ __forceinline__ double2 ldg(const double2 *p) const { return __ldg(p); }
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

template <int FE, bool ACC, bool DOT, bool SPLIT>
__global__ void __maxnreg__(SPLIT ? 128 : 168) hex_patch_p2e_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_p2e[];
  P2DeviceEnv env{smem_p2e};
  hex_patch_p2_body<P2DeviceEnv, FE, ACC, DOT, SPLIT>(env, a);
}

// two threads per element component (hex_patch_p2h_body.h): 6 FE threads per CTA, same shared memory as the plain kernel
// REGS: 96 (no spills; 3 CTAs = 18 warps per SM) or 80 (32-48 bytes of stack; 4 CTAs = 24 warps per SM); natural need 116
template <int FE, bool ACC, bool DOT, int REGS>
__global__ void __maxnreg__(REGS) hex_patch_p2h_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_p2h[];
  P2DeviceEnv env{smem_p2h};
  hex_patch_p2h_body<P2DeviceEnv, FE, ACC, DOT>(env, a);
}

// one thread per element (hex_patch_p2s_body.h): FE threads per CTA, no exchange buffer
template <int FE, bool ACC, bool DOT, bool GENERAL>
__global__ void __launch_bounds__(FE, 1) hex_patch_p2s_kernel(const PatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_p2s[];
  P2DeviceEnv env{smem_p2s};
  hex_patch_p2s_body<P2DeviceEnv, FE, ACC, DOT, GENERAL>(env, a);
}

template <int FE, bool GENERAL>
static int launch_p2s(const PatchArgs &a, int64_t n_run, int sm_count, int n_partial, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_p2s_kernel<FE, false, false, GENERAL>;
  auto k1 = hex_patch_p2s_kernel<FE, true, false, GENERAL>;
  auto kd = hex_patch_p2s_kernel<FE, false, true, GENERAL>;
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
    if (per_sm < 1) { set_error("patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_run) grid = n_run;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > n_partial) grid = n_partial;
    kd<<<(unsigned)grid, FE, smem, st>>>(a);
  } else if (acc) k1<<<(unsigned)grid, FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, FE, smem, st>>>(a);
  return 0;
}

template <int FE, int REGS>
static int launch_p2h(const PatchArgs &a, int64_t n_run, int sm_count, int n_partial, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_p2h_kernel<FE, false, false, REGS>;
  auto k1 = hex_patch_p2h_kernel<FE, true, false, REGS>;
  auto kd = hex_patch_p2h_kernel<FE, false, true, REGS>;
  static size_t attr = 0;
  static int per_sm = 0;
  if (attr < smem) {
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HC_CUDA(cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    HC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kd, 6 * FE, smem));
    if (per_sm < 1) { set_error("patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_run) grid = n_run;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > n_partial) grid = n_partial;
    kd<<<(unsigned)grid, 6 * FE, smem, st>>>(a);
  } else if (acc) k1<<<(unsigned)grid, 6 * FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, 6 * FE, smem, st>>>(a);
  return 0;
}

template <int FE, bool SPLIT>
static int launch_p2e(const PatchArgs &a, int64_t n_run, int sm_count, int n_partial, size_t smem, bool acc, cudaStream_t st) {
  auto k0 = hex_patch_p2e_kernel<FE, false, false, SPLIT>;
  auto k1 = hex_patch_p2e_kernel<FE, true, false, SPLIT>;
  auto kd = hex_patch_p2e_kernel<FE, false, true, SPLIT>;
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
    if (per_sm < 1) { set_error("patch kernel does not fit on an SM (%zu bytes of shared memory)", smem); return 1; }
    attr = smem;
  }
  int64_t grid = (int64_t)per_sm * sm_count;
  if (const char *s = getenv("HCURL_PATCH_CTAS_PER_SM")) { const int v = atoi(s); if (v >= 1 && v <= per_sm) grid = (int64_t)v * sm_count; }
  if (grid > n_run) grid = n_run;
  if (a.dot_out) {
    if (acc) { set_error("patch kernel: fused dot product is not available with accumulation"); return 1; }
    if (grid > n_partial) grid = n_partial;  // one partial sum per CTA
    kd<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  } else if (acc) k1<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  else k0<<<(unsigned)grid, 3 * FE, smem, st>>>(a);
  return 0;
}

static bool g_modal2_set = false;

int hex_patch_p2e_launch(hcurl_op *op, const PatchArgs &a, bool accumulate) {
  hcurl_ctx *ctx = op->ctx;
  const PatchSet &P = op->sp->patch;
  if (P.FE != 32 && !(P.single_kernel && P.FE == 64)) { set_error("Env patch kernel: FE must be 32"); return 1; }
  if (!g_modal2_set) {
    ModalTables T;
    if (compute_modal_tables(T)) return 1;
    HC_CUDA(cudaMemcpyToSymbol(cM2, &T, sizeof(T)));
    g_modal2_set = true;
  }
  constexpr int FE = 32;
  constexpr size_t T_CAP = (27 * FE * 2 + 15) & ~15;
  const size_t smem = P.split ? 32 + 2 * (size_t)a.d_cap_bytes + a.l_cap_bytes + T_CAP + a.xs_cap_bytes + 2 * (size_t)27 * FE * 16
                              : 16 + 2 * (size_t)a.blob_cap_bytes + a.xs_cap_bytes + 2 * (size_t)27 * (FE + 1) * 16;
  const int64_t n_run = a.n_patch - a.k_begin;
  if (P.single_kernel) {
    if (P.split) { set_error("the one-thread-per-element kernel uses the plain blob layout (unset HCURL_PATCH_SPLIT)"); return 1; }
    const size_t sm1 = 16 + 2 * (size_t)a.blob_cap_bytes + a.xs_cap_bytes + (size_t)27 * (P.FE + 1) * 16;
    if (P.general_orient) {
      if (P.FE == 64) return launch_p2s<64, true>(a, n_run, ctx->sm_count, ctx->n_partial, sm1, accumulate, ctx->stream);
      return launch_p2s<32, true>(a, n_run, ctx->sm_count, ctx->n_partial, sm1, accumulate, ctx->stream);
    }
    if (P.FE == 64) return launch_p2s<64, false>(a, n_run, ctx->sm_count, ctx->n_partial, sm1, accumulate, ctx->stream);
    return launch_p2s<32, false>(a, n_run, ctx->sm_count, ctx->n_partial, sm1, accumulate, ctx->stream);
  }
  if (P.half_kernel) {
    if (P.split) { set_error("the two-threads-per-component kernel uses the plain blob layout (unset HCURL_PATCH_SPLIT)"); return 1; }
    if (P.half_regs == 80) return launch_p2h<FE, 80>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
    return launch_p2h<FE, 96>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
  }
  if (P.split) return launch_p2e<FE, true>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
  return launch_p2e<FE, false>(a, n_run, ctx->sm_count, ctx->n_partial, smem, accumulate, ctx->stream);
}

}  // namespace hcurl
