// Hardware micro-benchmarks behind design decisions: not part of the
// public ABI, used by scripts/microbench.py only.
//  - gather/scatter floor of the order-2 hex apply: map + x gather + FP64 RED scatter
//  - FP64 DMMA (mma.sync.m8n8k4.f64) throughput alone and co-issued with DFMA
#include "internal.h"

namespace hcurl {

__global__ void transpose_map_kernel(const int64_t n_elem, const int nd, const int64_t n_pad,
                                     const uint32_t *__restrict__ in, uint32_t *__restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * nd) return;
  const int64_t e = gid / nd;
  const int t = (int)(gid % nd);
  out[(int64_t)t * n_pad + e] = in[gid];
}

// 3 warps per 32 elements: warp c handles the 18 dofs of component c (order 2)
template <int MODE>
__global__ void __launch_bounds__(96) gs_kernel(const int64_t n_elem, const int64_t n_pad, const uint32_t *__restrict__ mapT,
                                                const double *__restrict__ x, double *__restrict__ y, double *sink) {
  const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int64_t e = (int64_t)blockIdx.x * 32 + lane;
  if (e >= n_elem) return;
  double v[18];
  uint32_t m[18];
#pragma unroll
  for (int t = 0; t < 18; t++) {
    m[t] = __ldg(mapT + (int64_t)(c * 18 + t) * n_pad + e);
    v[t] = __ldg(x + (m[t] & MAP_IDX));
  }
  if (MODE == 0) {
    double s = 0;
#pragma unroll
    for (int t = 0; t < 18; t++) s += v[t];
    if (s == 1.2345e-300) *sink = s;
  } else {
#pragma unroll
    for (int t = 0; t < 18; t++) {
      if (MODE == 2 && (m[t] & MAP_ESS)) y[m[t] & MAP_IDX] = v[t];  // "exclusive" dofs: plain store
      else atomicAdd(y + (m[t] & MAP_IDX), v[t]);
    }
  }
}

__global__ void dmma_kernel(double *out, int iters, int mode) {
  // mode 0: DMMA only; 1: DFMA only; 2: even warps DMMA, odd warps DFMA
  const int warp = threadIdx.x >> 5;
  double acc0 = 0, acc1 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  const bool do_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  if (do_mma) {
    for (int i = 0; i < iters; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc0), "+d"(acc1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2), "+d"(c3) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc1), "+d"(c0) : "d"(b), "d"(a));
    }
  } else {
    double d0 = a, d1 = b, d2 = a + 1, d3 = b + 1, d4 = a + 2, d5 = b + 2, d6 = a + 3, d7 = b + 3;
    for (int i = 0; i < iters; i++) {
      d0 = d0 * a + b; d1 = d1 * a + b; d2 = d2 * a + b; d3 = d3 * a + b;
      d4 = d4 * a + b; d5 = d5 * a + b; d6 = d6 * a + b; d7 = d7 * a + b;
    }
    acc0 = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1 + c0 + c1 + c2 + c3;
}

}  // namespace hcurl

using namespace hcurl;

extern "C" {

// returns ms per launch (best of reps); mode 0 gather only, 1 gather+RED scatter, 2 flagged dofs as plain stores
double hcurl_debug_gather_scatter_ms(hcurl_op *op, int mode, int reps) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  if (sp->geom != HCURL_GEOM_HEX || sp->order != 2) return -1.0;
  cudaSetDevice(ctx->device);
  const int nd = sp->nd;
  const int64_t n_pad = sp->n_elem_pad;
  uint32_t *mapT = nullptr;
  double *x = nullptr, *y = nullptr, *sink = nullptr;
  cudaMalloc(&mapT, sizeof(uint32_t) * nd * n_pad);
  cudaMalloc(&x, sizeof(double) * sp->n_ext());
  cudaMalloc(&y, sizeof(double) * sp->n_ext());
  cudaMalloc(&sink, 8);
  cudaMemset(x, 0, sizeof(double) * sp->n_ext());
  const uint32_t *src = op->d_map_ess ? op->d_map_ess : sp->d_map;
  transpose_map_kernel<<<(unsigned)((sp->n_elem * nd + 255) / 256), 256, 0, ctx->stream>>>(sp->n_elem, nd, n_pad, src, mapT);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  const unsigned grid = (unsigned)((sp->n_elem + 31) / 32);
  for (int r = 0; r < reps + 1; r++) {
    cudaMemsetAsync(y, 0, sizeof(double) * sp->n_ext(), ctx->stream);
    cudaEventRecord(e0, ctx->stream);
    if (mode == 0) gs_kernel<0><<<grid, 96, 0, ctx->stream>>>(sp->n_elem, n_pad, mapT, x, y, sink);
    else if (mode == 1) gs_kernel<1><<<grid, 96, 0, ctx->stream>>>(sp->n_elem, n_pad, mapT, x, y, sink);
    else gs_kernel<2><<<grid, 96, 0, ctx->stream>>>(sp->n_elem, n_pad, mapT, x, y, sink);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(mapT); cudaFree(x); cudaFree(y); cudaFree(sink);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

// TFLOP/s; mode 0 DMMA only, 1 DFMA only, 2 both co-issued (returns the sum)
double hcurl_debug_dmma_tflops(hcurl_ctx *ctx, int mode) {
  cudaSetDevice(ctx->device);
  const int blocks = ctx->sm_count * 4, threads = 512, iters = 1 << 13;
  double *out = nullptr;
  cudaMalloc(&out, sizeof(double) * blocks * threads);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0, ctx->stream);
    dmma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, mode);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double nwarps = (double)blocks * threads / 32;
  // DMMA m8n8k4: 8*8*4*2 = 512 flop per warp instruction, 4 per iteration; DFMA: 2 flop * 8 per thread per iteration
  const double f_mma = 512.0 * 4 * iters, f_fma = 2.0 * 8 * 32 * iters;
  double flops;
  if (mode == 0) flops = nwarps * f_mma;
  else if (mode == 1) flops = nwarps * f_fma;
  else flops = nwarps / 2 * (f_mma + f_fma);
  return flops / (best * 1e-3) / 1e12;
}

}  // extern "C"
