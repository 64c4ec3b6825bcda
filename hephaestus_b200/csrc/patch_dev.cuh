// Device-side primitives shared by the patch kernels (hex_patch.cu, hex_patch_p1.cu): TMA bulk copies counted on an
// mbarrier, asynchronous 8- / 16-byte gathers (LDGSTS), fire-and-forget FP64 RED, sign flips.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hcurl {

// orientation of an element slot pair relative to its global pair, carried in bits 13..15 of a lidx entry
// (PAIR_SWAP / PAIR_NEG1 / PAIR_NEG0 of the slot code >> 16): local = (s0 * g[swap], s1 * g[!swap])
__device__ __forceinline__ double flip(double v, uint32_t sign_bit31) {
  return __hiloint2double(__double2hiint(v) ^ (int)sign_bit31, __double2loint(v));
}
__device__ __forceinline__ void pair_to_local(const double2 g, const uint32_t l, double &x0, double &x1) {
  const bool sw = l & 0x2000u;
  x0 = flip(sw ? g.y : g.x, (l << 16) & 0x80000000u);
  x1 = flip(sw ? g.x : g.y, (l << 17) & 0x80000000u);
}
__device__ __forceinline__ double2 local_to_pair(const double y0, const double y1, const uint32_t l) {
  const bool sw = l & 0x2000u;
  const double a = flip(y0, (l << 16) & 0x80000000u), b = flip(y1, (l << 17) & 0x80000000u);
  return sw ? make_double2(b, a) : make_double2(a, b);
}

// ---- async copies -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(mbar)), "r"(bytes) : "memory");
}
// 1-D bulk copy through the TMA engine, completion counted on an mbarrier
__device__ __forceinline__ void bulk_load(uint64_t *mbar, void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(mbar))
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
// 16-byte asynchronous gather global -> shared (LDGSTS, L1 bypassed), no register staging.  (The L1-allocating
// .ca form was measured 7-9 % slower per apply although ncu counts fewer shared-memory wavefronts for it; staging
// through registers -- LDG.128 where the gather is issued, STS.128 at the end of the iteration -- 57 % slower.)
__device__ __forceinline__ void gather16_async(double2 *dst, const double2 *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}
// fire-and-forget FP64 add (REDG).  Spelled in PTX: once a kernel variant contains fences (the fused dot product's
// last-CTA reduction) nvcc turns atomicAdd() with an unused result into returning ATOMG, which made that variant
// 37 % slower.
__device__ __forceinline__ void red_add(double *p, double v) {
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void gather_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gather_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// 8-byte asynchronous gather (order-1 dofs are single doubles; cp.async.cg exists for 16 bytes only)
__device__ __forceinline__ void gather8_async(double *dst, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(dst)), "l"(src) : "memory");
}

}  // namespace hcurl
