// hephaestus_b200/csrc/patch_pipeline.h with the order-1 / order-2 element sides on the host environment of
// host_env.h -- TEST INFRASTRUCTURE, see tests/test_patch_p1_emu.py, tests/test_patch_p2_emu.py.
#include "host_env.h"
#include "hex_patch_p1_elem.h"
#include "hex_patch_p2_elem.h"
#include "patch_stream.h"

namespace {
template <class Elem, int FE, bool ACC, bool DOT>
int run(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes) {
  double dot = 0.0;
  const int err = emu::run_grid(grid, FE, smem_bytes, nullptr, &dot,
                                [&](emu::HostEnv &env) { hcurl::patch_pipeline<emu::HostEnv, Elem, FE, ACC, DOT>(env, a); });
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
}
template <class Elem, int FE>
int run_mode(const hcurl::PatchArgs &a, int64_t grid, size_t smem, bool accumulate) {
  if (a.dot_out) return accumulate ? -1 : run<Elem, FE, false, true>(a, grid, smem);
  return accumulate ? run<Elem, FE, true, false>(a, grid, smem) : run<Elem, FE, false, false>(a, grid, smem);
}
// warp-specialised stream program (patch_stream.h): an emulated "CTA" is one stream of FE element + ND sum + NG gather threads
template <int FE, int ND, bool GENERAL, bool ACC, bool DOT>
int run_stream(const hcurl::PatchArgs &a, int64_t streams, size_t smem_bytes) {
  double dot = 0.0;
  constexpr int NG = FE;
  const int err = emu::run_grid(streams, FE + ND + NG, smem_bytes, nullptr, &dot, [&](emu::HostEnv &env) {
    hcurl::patch_stream<emu::HostEnv, hcurl::P2Elem<FE, GENERAL>, FE, ND, NG, ACC, DOT>(env, a);
  }, FE, ND);
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
}
template <int FE, int ND, bool GENERAL>
int run_stream_mode(const hcurl::PatchArgs &a, int64_t streams, size_t smem, bool accumulate) {
  if (a.dot_out) return accumulate ? -1 : run_stream<FE, ND, GENERAL, false, true>(a, streams, smem);
  return accumulate ? run_stream<FE, ND, GENERAL, true, false>(a, streams, smem) : run_stream<FE, ND, GENERAL, false, false>(a, streams, smem);
}
}  // namespace

// order 1 (fe = 32 / 64 / 128) or 2 (fe = 32 / 64, stream program also 128; general = orientation handling with independent swap / sign bits)
extern "C" int emu_patch_apply(int order, int fe, int general, int accumulate, int64_t n_patch, int64_t k_begin, int64_t grid,
                               const uint32_t *blob_off, const uint32_t *part_off, const uint32_t *blob, const double *edata,
                               const double *x, double *y, double scale, uint32_t d_cap_bytes, uint32_t l_cap_bytes,
                               uint32_t t_cap_bytes, uint32_t xs_cap_bytes, const uint8_t *swz_row, const uint8_t *swz_grp, int n_grp,
                               const double *tables7, const uint32_t *own_off, const uint32_t *own_idx, uint32_t *count, uint32_t count_target,
                               int ess_identity, double *dot_out) {
  using namespace hcurl;
  PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n_patch = n_patch; a.k_begin = k_begin; a.blob_off = blob_off; a.part_off = part_off; a.blob = blob; a.edata = edata; a.x = x; a.y = y;
  a.scale = scale; a.d_cap_bytes = d_cap_bytes; a.l_cap_bytes = l_cap_bytes; a.t_cap_bytes = t_cap_bytes; a.xs_cap_bytes = xs_cap_bytes;
  a.own_off = own_off; a.own_idx = own_idx; a.count = count; a.count_target = count_target; a.ess_identity = ess_identity; a.dot_out = dot_out;
  memcpy(a.swz_row, swz_row, order == 2 ? 27 : 12);
  memcpy(a.swz_grp, swz_grp, (size_t)n_grp);
  if (tables7) { a.modal.a0 = tables7[0]; a.modal.a1 = tables7[1]; a.modal.b = tables7[2]; a.modal.c0 = tables7[3]; a.modal.c1 = tables7[4]; a.modal.d1 = tables7[5]; a.modal.d2 = tables7[6]; }
  const int unit = order == 2 ? 16 : 8, rows = order == 2 ? 27 : 12;
  const size_t smem = 64 + 2 * (size_t)d_cap_bytes + l_cap_bytes + t_cap_bytes + xs_cap_bytes + (size_t)rows * (fe + 1) * unit;
  const bool acc = accumulate != 0;
  if (order == 2 && fe < 0) {  // the stream program; here `grid` counts streams and may exceed the number of patches (idle streams)
    const int FE = -fe;
    const size_t ssm = 128 + 4 * (size_t)d_cap_bytes + 2 * (size_t)l_cap_bytes + 2 * (size_t)t_cap_bytes + 2 * (size_t)xs_cap_bytes + (size_t)27 * (FE + 1) * 16;
    if (grid < 1) return 0;
    if (FE == 32) return general ? run_stream_mode<32, 96, true>(a, grid, ssm, acc) : run_stream_mode<32, 96, false>(a, grid, ssm, acc);
    if (FE == 64) return general ? run_stream_mode<64, 128, true>(a, grid, ssm, acc) : run_stream_mode<64, 128, false>(a, grid, ssm, acc);
    if (FE == 128) return general ? run_stream_mode<128, 256, true>(a, grid, ssm, acc) : run_stream_mode<128, 256, false>(a, grid, ssm, acc);
    return -2;
  }
  if (grid > n_patch - k_begin) grid = n_patch - k_begin;
  if (grid < 1) return 0;
  if (order == 2) {
    if (fe == 32) return general ? run_mode<P2Elem<32, true>, 32>(a, grid, smem, acc) : run_mode<P2Elem<32, false>, 32>(a, grid, smem, acc);
    if (fe == 64) return general ? run_mode<P2Elem<64, true>, 64>(a, grid, smem, acc) : run_mode<P2Elem<64, false>, 64>(a, grid, smem, acc);
    return -2;
  }
  if (fe == 32) return run_mode<P1Elem<32>, 32>(a, grid, smem, acc);
  if (fe == 64) return run_mode<P1Elem<64>, 64>(a, grid, smem, acc);
  if (fe == 128) return run_mode<P1Elem<128>, 128>(a, grid, smem, acc);
  return -2;
}
