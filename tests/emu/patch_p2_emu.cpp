// hephaestus_b200/csrc/hex_patch_p2_body.h (the order-2 patch kernel as an Env program, plain and split-blob variants)
// on the host environment of host_env.h -- TEST INFRASTRUCTURE, see tests/test_patch_p2_emu.py.
#include "host_env.h"
#include "hex_patch_p2_body.h"
#include "hex_patch_p2h_body.h"
#include "hex_patch_p2s_body.h"

namespace {
template <int FE, bool ACC, bool DOT, bool SPLIT>
int run(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes, const hcurl::ModalTables &T) {
  double dot = 0.0;
  const int err = emu::run_grid(grid, 3 * FE, smem_bytes, &T, &dot,
                                [&](emu::HostEnv &env) { hcurl::hex_patch_p2_body<emu::HostEnv, FE, ACC, DOT, SPLIT>(env, a); });
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
}
template <int FE, bool ACC, bool DOT>
int run_half(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes, const hcurl::ModalTables &T) {
  double dot = 0.0;
  const int err = emu::run_grid(grid, 6 * FE, smem_bytes, &T, &dot,
                                [&](emu::HostEnv &env) { hcurl::hex_patch_p2h_body<emu::HostEnv, FE, ACC, DOT>(env, a); });
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
}
template <int FE, bool ACC, bool DOT, bool GENERAL>
int run_single(const hcurl::PatchArgs &a, int64_t grid, size_t smem_bytes, const hcurl::ModalTables &T) {
  double dot = 0.0;
  const int err = emu::run_grid(grid, FE, smem_bytes, &T, &dot,
                                [&](emu::HostEnv &env) { hcurl::hex_patch_p2s_body<emu::HostEnv, FE, ACC, DOT, GENERAL>(env, a); });
  if (DOT && a.dot_out) *a.dot_out = a.dot_add ? *a.dot_out + dot : dot;
  return err;
}
template <int FE, bool GENERAL>
int run_single_fe(const hcurl::PatchArgs &a, int64_t grid, const hcurl::ModalTables &T, bool accumulate) {
  const size_t smem = 16 + 2 * (size_t)a.blob_cap_bytes + a.xs_cap_bytes + (size_t)27 * (FE + 1) * 16;
  if (a.dot_out) return accumulate ? -1 : run_single<FE, false, true, GENERAL>(a, grid, smem, T);
  return accumulate ? run_single<FE, true, false, GENERAL>(a, grid, smem, T) : run_single<FE, false, false, GENERAL>(a, grid, smem, T);
}
}  // namespace

extern "C" int emu_patch_p2_apply(int split, int accumulate, int64_t n_patch, int64_t k_begin, int64_t grid, const uint32_t *blob_off,
                                  const uint32_t *part_off, const uint32_t *blob, const double *edata, const double *x, double *y,
                                  double scale, uint32_t blob_cap_bytes, uint32_t xs_cap_bytes, uint32_t d_cap_bytes, uint32_t l_cap_bytes,
                                  const uint8_t *swz_row, const uint8_t *swz_grp, int n_grp, const double *tables7, double *dot_out) {
  constexpr int FE = 32;
  hcurl::PatchArgs a;
  memset(&a, 0, sizeof(a));
  a.n_patch = n_patch; a.k_begin = k_begin; a.blob_off = blob_off; a.part_off = part_off; a.blob = blob; a.edata = edata; a.x = x; a.y = y;
  a.scale = scale; a.blob_cap_bytes = blob_cap_bytes; a.xs_cap_bytes = xs_cap_bytes; a.d_cap_bytes = d_cap_bytes; a.l_cap_bytes = l_cap_bytes;
  a.dot_out = dot_out;
  memcpy(a.swz_row, swz_row, 27);
  memcpy(a.swz_grp, swz_grp, (size_t)n_grp);
  hcurl::ModalTables T;
  T.a0 = tables7[0]; T.a1 = tables7[1]; T.b = tables7[2]; T.c0 = tables7[3]; T.c1 = tables7[4]; T.d1 = tables7[5]; T.d2 = tables7[6];
  constexpr size_t T_CAP = (27 * FE * 2 + 15) & ~15;
  const size_t smem = split == 1 ? 32 + 2 * (size_t)d_cap_bytes + l_cap_bytes + T_CAP + xs_cap_bytes + 2 * (size_t)27 * FE * 16
                            : 16 + 2 * (size_t)blob_cap_bytes + xs_cap_bytes + 2 * (size_t)27 * (FE + 1) * 16;
  if (grid > n_patch - k_begin) grid = n_patch - k_begin;
  if (grid < 1) return 0;
  // one thread per element (hex_patch_p2s_body.h): reversed-line orientation flags only (3, 4) / general flags (5)
  if (split == 3) return run_single_fe<32, false>(a, grid, T, accumulate != 0);
  if (split == 4) return run_single_fe<64, false>(a, grid, T, accumulate != 0);
  if (split == 5) return run_single_fe<32, true>(a, grid, T, accumulate != 0);
  if (split == 2) {  // two threads per element component (hex_patch_p2h_body.h), plain blob layout
    if (dot_out) return accumulate ? -1 : run_half<FE, false, true>(a, grid, smem, T);
    return accumulate ? run_half<FE, true, false>(a, grid, smem, T) : run_half<FE, false, false>(a, grid, smem, T);
  }
  if (dot_out) {
    if (accumulate) return -1;
    return split ? run<FE, false, true, true>(a, grid, smem, T) : run<FE, false, true, false>(a, grid, smem, T);
  }
  if (accumulate) return split ? run<FE, true, false, true>(a, grid, smem, T) : run<FE, true, false, false>(a, grid, smem, T);
  return split ? run<FE, false, false, true>(a, grid, smem, T) : run<FE, false, false, false>(a, grid, smem, T);
}
