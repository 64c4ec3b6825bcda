// hephaestus_b200/csrc/hex_patch_p1_body.h on the host environment of host_env.h (TEST INFRASTRUCTURE, not part of the
// product library): the per-thread program of the order-1 patch kernel runs unchanged on host threads, so its indexing,
// pipelining order and barrier placement can be checked against the oracle without a GPU (tests/test_patch_p1_emu.py;
// build with -fsanitize=thread to check for races between barriers).
#include "host_env.h"
#include "hex_patch_p1_body.h"

namespace {
template <int FE, bool ACC, bool DOT>
int run(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes) {
  double dot = 0.0;
  const int err = emu::run_grid(grid, FE, smem_bytes, nullptr, &dot, [&](emu::HostEnv &env) { hcurl::hex_patch_p1_body<emu::HostEnv, FE, ACC, DOT>(env, a); });
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
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
