// Host execution environment for hephaestus_b200/csrc/hex_patch_p1_body.h (TEST INFRASTRUCTURE, not part of the product
// library): the per-thread program of the order-1 patch kernel runs unchanged on host threads -- one std::thread per
// CUDA thread of a CTA, std::barrier for __syncthreads, memcpy for the TMA bulk copy and the cp.async gathers,
// compare-and-swap for red.global.add.f64 -- so its indexing, pipelining order and barrier placement can be checked
// against the oracle without a GPU (tests/test_patch_p1_emu.py; build with -fsanitize=thread to check for races between
// barriers).  CTAs run one after the other; each walks patches block, += nblocks like the persistent kernel.
//
// Deliberately pessimistic where the hardware is asynchronous: the destination of an asynchronous copy is poisoned when
// the copy is issued (it may land at once) and filled at its WAIT (blob_wait / gather_wait: it may land that late), so a
// thread that still reads the old content after the request, or the new one before waiting, computes garbage.
#include <atomic>
#include <barrier>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"
#include "hex_patch_p1_body.h"

namespace {

struct Pending { void *dst = nullptr; const void *src = nullptr; uint32_t bytes = 0; bool armed = false, landed = false; };

struct CtaShared {
  std::vector<unsigned char> smem;
  std::barrier<> bar;
  std::mutex mu[2];
  Pending blob[2];     // guarded by mu[b]
  uint32_t phase[2];   // completed requests per buffer (the mbarrier's phase), guarded by mu[b]
  int waiters[2];      // guarded by mu[b]
  double dot_sum;
  std::vector<double> dot_part;
  CtaShared(size_t bytes, int nt) : smem(bytes, 0xAB), bar(nt), phase{0, 0}, waiters{0, 0}, dot_sum(0.0), dot_part((size_t)nt, 0.0) {}
};

struct HostEnv {
  CtaShared *cta;
  int tid_;
  int64_t block_, nblocks_;
  int nt_;
  std::vector<std::pair<double *, const double *>> pend;  // my issued, not yet waited-for gathers
  int *error;
  int tid() const { return tid_; }
  int64_t block() const { return block_; }
  int64_t nblocks() const { return nblocks_; }
  unsigned char *smem() const { return cta->smem.data(); }
  void sync() { cta->bar.arrive_and_wait(); }
  void blob_init() {}
  void blob_request(int b, void *dst, const void *src, uint32_t bytes) {
    std::lock_guard<std::mutex> lk(cta->mu[b]);
    if (cta->blob[b].armed) *error |= 1;  // request into a buffer whose previous copy has not been consumed by everyone
    if (bytes % 16 != 0 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) *error |= 2;  // cp.async.bulk alignment rules
    memset(dst, 0xCD, bytes);  // the copy may start at once: whoever still reads the old content gets garbage
    cta->blob[b].dst = dst; cta->blob[b].src = src; cta->blob[b].bytes = bytes; cta->blob[b].armed = true; cta->blob[b].landed = false;
  }
  void blob_wait(int b, uint32_t parity) {
    std::lock_guard<std::mutex> lk(cta->mu[b]);
    Pending &q = cta->blob[b];
    if (!q.armed) { *error |= 4; return; }
    if ((cta->phase[b] & 1u) != parity) *error |= 8;   // waiting on the wrong phase of the mbarrier
    if (!q.landed) { memcpy(q.dst, q.src, q.bytes); q.landed = true; }  // ... and it completes as late as allowed
    if (++cta->waiters[b] == nt_) {  // every thread of the CTA waits exactly once per request
      cta->waiters[b] = 0;
      cta->phase[b]++;
      q.armed = false;
    }
  }
  void gather8(double *dst, const double *src) {
    *dst = std::numeric_limits<double>::quiet_NaN();  // in flight: readers that do not wait see garbage
    pend.emplace_back(dst, src);
  }
  void gather_commit() {}
  void gather_wait() {
    for (auto &p : pend) *p.first = *p.second;
    pend.clear();
  }
  void red_add(double *p, double v) {
    std::atomic_ref<double> r(*p);
    double old = r.load(std::memory_order_relaxed);
    while (!r.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
  }
  uint32_t ldg(const uint32_t *p) const { return *p; }
  double ldg(const double *p) const { return *p; }
  void dot_finish(double dacc, const hcurl::PatchArgs &, double *) {
    cta->dot_part[(size_t)tid_] = dacc;
    sync();
    if (tid_ == 0) {
      double v = 0.0;
      for (double d : cta->dot_part) v += d;
      cta->dot_sum = v;
    }
  }
};

template <int FE, bool ACC, bool DOT>
int run(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes) {
  int error = 0;
  double dot = 0.0;
  for (int64_t blk = 0; blk < grid; blk++) {
    CtaShared cta(smem_bytes, FE);
    std::vector<int> err((size_t)FE, 0);
    std::vector<std::thread> th;
    for (int t = 0; t < FE; t++)
      th.emplace_back([&, t]() {
        HostEnv env{&cta, t, blk, grid, FE, {}, &err[(size_t)t]};
        hcurl::hex_patch_p1_body<HostEnv, FE, ACC, DOT>(env, a);
        if (!env.pend.empty()) err[(size_t)t] |= 16;  // gathers never waited for
      });
    for (auto &t : th) t.join();
    for (int e : err) error |= e;
    if (cta.blob[0].armed || cta.blob[1].armed) error |= 32;  // a requested blob was never consumed
    dot += cta.dot_sum;
  }
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return error;
}

}  // namespace

extern "C" int emu_patch_p1_apply(int FE, int accumulate, int64_t n_patch, int64_t k_begin, int64_t grid, const uint32_t *blob_off,
                                  const uint32_t *blob, const double *edata, const double *x, double *y, double scale,
                                  uint32_t blob_cap_bytes, uint32_t xs_cap_bytes, const uint8_t *swz_row, const uint8_t *swz_grp,
                                  int n_grp, double *dot_out) {
  hcurl::PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n_patch = n_patch; a.k_begin = k_begin; a.blob_off = blob_off; a.blob = blob; a.edata = edata; a.x = x; a.y = y; a.scale = scale;
  a.blob_cap_bytes = blob_cap_bytes; a.xs_cap_bytes = xs_cap_bytes; a.dot_out = dot_out;
  memcpy(a.swz_row, swz_row, 12);
  memcpy(a.swz_grp, swz_grp, (size_t)n_grp);
  const size_t smem = 16 + 2 * (size_t)blob_cap_bytes + xs_cap_bytes + (size_t)12 * (FE + 1) * 8;
  if (grid > n_patch - k_begin) grid = n_patch - k_begin;
  if (grid < 1) return 0;
#define RUN(fe)                                                           \
  case fe:                                                                \
    if (dot_out) return accumulate ? -1 : run<fe, false, true>(a, grid, smem); \
    return accumulate ? run<fe, true, false>(a, grid, smem) : run<fe, false, false>(a, grid, smem);
  switch (FE) {
    RUN(32)
    RUN(64)
    RUN(128)
  }
  return -2;
}
