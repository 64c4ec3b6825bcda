// Host execution environment for kernel programs written against an `Env` (hephaestus_b200/csrc/patch_pipeline.h,
// interp_body.h) -- TEST INFRASTRUCTURE, not part of the product library.  One std::thread per CUDA thread of a
// CTA, std::barrier for __syncthreads, memcpy for TMA bulk copies and cp.async gathers, compare-and-swap for
// red.global.add.f64, release / acquire atomics for the arrival counters, a per-warp barrier for __all_sync.  All CTAs of
// the grid run CONCURRENTLY (the patch pipeline waits on flags of other CTAs, like the cooperative launch on the
// device); each walks patches block, += nblocks like the persistent kernels.
//
// Deliberately pessimistic where the hardware is asynchronous: the destination of an asynchronous copy is poisoned when
// the copy is issued (it may land at once) and filled at its WAIT (bulk_wait / gather_wait: it may land that late), so a
// thread that still reads the old content after the request, or the new one before waiting, computes garbage.  The
// mbarrier protocol is checked: one request per phase, every thread of the CTA waits exactly once per request and on the
// right parity, bulk-copy alignment rules.  Error flags are returned by run_grid().
#pragma once
#include <atomic>
#include <barrier>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace emu {

static constexpr int MAX_BAR = 8, MAX_FLAG = 8;
enum { ERR_REARM = 1, ERR_ALIGN = 2, ERR_WAIT_UNARMED = 4, ERR_PARITY = 8, ERR_GATHER_LEFT = 16, ERR_BULK_LEFT = 32, ERR_FLAG_AHEAD = 64 };

// one bulk-copy request on an mbarrier: issued by one thread, waited for (once) by nwait threads, each of which counts its own
// waits per mbarrier -- waiters of different warp groups may be a whole request apart (patch_stream.h: the gather group
// waits for the unit list of patch i + 4 while the sum group has not yet looked at the one of patch i)
struct Pending { void *dst = nullptr; const void *src = nullptr; uint32_t bytes = 0; bool active = false, landed = false; int nwait = 0, waited = 0; };

struct WarpShared {
  std::barrier<> bar;
  std::atomic<int> flag[3];
  double xch[3][32];  // shuffle scratch, rotating like the vote flags
  explicit WarpShared(int n) : bar(n) { for (auto &f : flag) f.store(1); }
};

struct CtaShared {
  std::vector<unsigned char> smem;
  std::barrier<> bar;
  std::mutex mu[MAX_BAR];
  Pending bulk[MAX_BAR][2];            // requests r and r - 1 of the mbarrier (index r & 1), guarded by mu[b]
  std::atomic<uint32_t> requests[MAX_BAR];  // requests issued so far per mbarrier
  std::vector<std::unique_ptr<WarpShared>> warp;
  std::unique_ptr<std::barrier<>> group, ggroup, elem;  // the sum group / the gather group / the element warps of a stream (patch_stream.h): threads [n_elem, n_elem + n_sum) / [n_elem + n_sum, nt) / [0, n_elem)
  int n_elem = 32;
  std::atomic<uint32_t> flag[MAX_FLAG];   // signals so far
  uint32_t flag_arrivals[MAX_FLAG] = {1, 1, 1, 1, 1, 1, 1, 1};  // signals per phase (mbarrier arrival count), set by flag_init
  double dot_sum = 0.0;
  std::vector<double> dot_part;
  void set_stream(int ne, int nsum, int nt) {
    n_elem = ne;
    elem.reset(new std::barrier<>(ne));
    group.reset(new std::barrier<>(nsum));
    ggroup.reset(new std::barrier<>(nt - ne - nsum));
  }
  CtaShared(size_t bytes, int nt) : smem(bytes + 16, 0xAB), bar(nt), dot_part((size_t)nt, 0.0) {
    for (int b = 0; b < MAX_BAR; b++) requests[b].store(0);
    for (auto &f : flag) f.store(0);

    for (int w = 0; w * 32 < nt; w++) warp.emplace_back(new WarpShared(std::min(32, nt - w * 32)));
  }
  unsigned char *base() {  // 16-byte aligned like dynamic shared memory
    return reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem.data()) + 15) & ~(uintptr_t)15);
  }
};

struct HostEnv {
  CtaShared *cta;
  int tid_;
  int64_t block_, nblocks_;
  int nt_;
  int *error;
  const void *tables_;
  struct Copy { void *dst; const void *src; int bytes; unsigned grp; };
  std::vector<Copy> pend;  // my issued, not yet waited-for gathers
  unsigned vote_n = 0, shfl_n = 0;
  bool wait_may_precede_request = false;  // set by role(): the stream program
  unsigned commits = 0;  // committed gather groups
  uint32_t my_waits[MAX_BAR] = {0, 0, 0, 0, 0, 0, 0, 0};  // my waits so far per mbarrier

  int tid() const { return tid_; }
  int64_t block() const { return block_; }
  int64_t nblocks() const { return nblocks_; }
  unsigned char *smem() const { return cta->base(); }
  void sync() { cta->bar.arrive_and_wait(); }
  template <class T> const T &tables() const { return *static_cast<const T *>(tables_); }
  void bulk_init(int) {}
  void bulk_request(int b, void *dst, const void *src, uint32_t bytes) { bulk_request_w(b, dst, src, bytes, nt_); }
  void bulk_request_w(int b, void *dst, const void *src, uint32_t bytes, int nwait) {  // nwait threads will wait for this request
    std::lock_guard<std::mutex> lk(cta->mu[b]);
    const uint32_t r = cta->requests[b].load(std::memory_order_relaxed);
    Pending &q = cta->bulk[b][r & 1u], &prev = cta->bulk[b][(r + 1u) & 1u];
    // request r while a thread has not yet waited for request r - 2: the one-bit phase parity of its wait could not tell them apart
    if (q.active && q.waited < q.nwait) *error |= ERR_REARM;
    if (bytes % 16 != 0 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) *error |= ERR_ALIGN;  // cp.async.bulk rules
    if (prev.active && !prev.landed) { if (prev.bytes) memcpy(prev.dst, prev.src, prev.bytes); prev.landed = true; }
    if (bytes) memset(dst, 0xCD, bytes);  // the copy may start at once: whoever still reads the old content gets garbage
    q.dst = dst; q.src = src; q.bytes = bytes; q.active = true; q.landed = false; q.nwait = nwait; q.waited = 0;
    cta->requests[b].store(r + 1u, std::memory_order_release);
  }
  void bulk_wait(int b, uint32_t parity) {
    // a wait may come before its request (warp-specialised streams: the requesting thread runs on its own); on the
    // device the mbarrier phase is simply incomplete until then.  A wait whose request never comes hangs here as it would there.
    const uint32_t r = my_waits[b]++;  // the request this thread is waiting for
    for (;;) {
      if (cta->requests[b].load(std::memory_order_acquire) > r) break;
      if (!wait_may_precede_request) { *error |= ERR_WAIT_UNARMED; return; }
      std::this_thread::yield();
    }
    std::lock_guard<std::mutex> lk(cta->mu[b]);
    if (cta->requests[b].load(std::memory_order_relaxed) > r + 1u) { *error |= ERR_REARM; return; }  // my request's buffer has been requested again already
    Pending &q = cta->bulk[b][r & 1u];
    if ((r & 1u) != parity) *error |= ERR_PARITY;                     // waiting on the wrong phase of the mbarrier
    if (!q.landed) { if (q.bytes) memcpy(q.dst, q.src, q.bytes); q.landed = true; }  // ... and it completes as late as allowed
    q.waited++;  // every thread concerned waits exactly once per request
  }
  void gather8(double *dst, const double *src) {
    *dst = std::numeric_limits<double>::quiet_NaN();  // in flight: readers that do not wait see garbage
    pend.push_back({dst, src, 8, commits});
  }
  void gather16(double2 *dst, const double2 *src) {
    dst->x = dst->y = std::numeric_limits<double>::quiet_NaN();
    pend.push_back({dst, src, 16, commits});
  }
  void gather_commit() { commits++; }
  void gather_wait() {
    for (auto &p : pend) memcpy(p.dst, p.src, (size_t)p.bytes);
    pend.clear();
  }
  void red_add(double *p, double v) {
    std::atomic_ref<double> r(*p);
    double old = r.load(std::memory_order_relaxed);
    while (!r.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
  }
  bool warp_all(bool pred) {  // __all_sync(0xffffffff, pred)
    WarpShared &W = *cta->warp[(size_t)(tid_ >> 5)];
    const unsigned n = vote_n++;
    if ((tid_ & 31) == 0) W.flag[(n + 1) % 3].store(1);  // slot of the next vote; readers of the previous one use another
    if (!pred) W.flag[n % 3].store(0);
    W.bar.arrive_and_wait();
    return W.flag[n % 3].load() != 0;
  }
  double shfl_xor(double v, int lane_mask) {  // __shfl_xor_sync(0xffffffff, v, lane_mask): every thread of the warp calls it
    WarpShared &W = *cta->warp[(size_t)(tid_ >> 5)];
    const unsigned n = shfl_n++ % 3;
    W.xch[n][tid_ & 31] = v;
    W.bar.arrive_and_wait();
    return W.xch[n][(tid_ & 31) ^ lane_mask];
  }
  // ---- warp-specialised streams (patch_stream.h)
  void role(int) { wait_may_precede_request = true; }
  void group_sync() { cta->group->arrive_and_wait(); }
  void ggroup_sync() { cta->ggroup->arrive_and_wait(); }
  void warp_sync() { cta->warp[(size_t)(tid_ >> 5)]->bar.arrive_and_wait(); }
  void elem_sync() { cta->elem->arrive_and_wait(); }
  void flag_init(int f, int arrivals) { cta->flag[f].store(0); cta->flag_arrivals[f] = (uint32_t)arrivals; }
  void flag_signal(int f) { cta->flag[f].fetch_add(1u, std::memory_order_release); }
  void flag_wait(int f, uint32_t n) {  // until phase n of the flag (its arrivals n * per .. (n + 1) * per - 1) is complete
    const uint32_t per = cta->flag_arrivals[f];
    for (;;) {
      const uint32_t c = cta->flag[f].load(std::memory_order_acquire);
      if (c >= (n + 2) * per) *error |= ERR_FLAG_AHEAD;  // the device's one-bit phase parity could not tell this phase from the one after next
      if (c >= (n + 1) * per) return;
      std::this_thread::yield();
    }
  }
  void arrive(uint32_t *p) { std::atomic_ref<uint32_t>(*p).fetch_add(1u, std::memory_order_release); }
  uint32_t counter_acquire(const uint32_t *p) { return std::atomic_ref<uint32_t>(*const_cast<uint32_t *>(p)).load(std::memory_order_acquire); }
  void backoff(unsigned) { std::this_thread::yield(); }
  uint32_t ldg(const uint32_t *p) const { return *p; }
  double ldg(const double *p) const { return *p; }
  double2 ldg(const double2 *p) const { return *p; }
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

// run `grid` CTAs of `nt` threads, all at the same time; body(env) is the per-thread program
template <class Body>
int run_grid(int64_t grid, int nt, size_t smem_bytes, const void *tables, double *dot_total, Body body, int stream_elem_threads = 0, int stream_sum_threads = 0) {
  std::vector<std::unique_ptr<CtaShared>> ctas;
  for (int64_t blk = 0; blk < grid; blk++) {
    ctas.emplace_back(new CtaShared(smem_bytes, nt));
    if (stream_elem_threads) ctas.back()->set_stream(stream_elem_threads, stream_sum_threads, nt);
  }
  std::vector<int> err((size_t)(grid * nt), 0);
  std::vector<std::thread> th;
  for (int64_t blk = 0; blk < grid; blk++)
    for (int t = 0; t < nt; t++)
      th.emplace_back([&, blk, t]() {
        HostEnv env{ctas[(size_t)blk].get(), t, blk, grid, nt, &err[(size_t)(blk * nt + t)], tables, {}, 0, 0};
        body(env);
        if (!env.pend.empty()) err[(size_t)(blk * nt + t)] |= ERR_GATHER_LEFT;  // gathers never waited for
      });
  for (auto &t : th) t.join();
  int error = 0;
  double dot = 0.0;
  for (int e : err) error |= e;
  for (auto &c : ctas) {
    for (int b = 0; b < MAX_BAR; b++)
      for (int h = 0; h < 2; h++) if (c->bulk[b][h].active && c->bulk[b][h].waited != c->bulk[b][h].nwait) error |= ERR_BULK_LEFT;  // a requested copy was not consumed by everyone concerned
    dot += c->dot_sum;
  }
  if (dot_total) *dot_total = dot;
  return error;
}

}  // namespace emu
