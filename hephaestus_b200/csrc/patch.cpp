// Patch decomposition of the element -> dof map (host side, setup time only).
//
// The reference resolves shared dofs by assembling a global CSR matrix
// (blf->Assemble() + FormLinearSystem, /root/reference/src/equation_systems/equation_system.cpp:345, :132-190).
// The matrix-free apply instead gathers x_e and scatter-adds y_e through the signed
// element -> dof map (FiniteElementSpace::GetElementDofs).  Doing that element by element
// costs nd gathers + nd FP64 RED atomics per element, and on sm_100a the L1tex lane rate of
// REDs (~1.1 cycles per lane per SM, measured: scripts/microbench.py) bounds the whole apply.
//
// So elements are grouped into PATCHES of FE spatially adjacent elements (Morton order of the
// element centroids), one CTA per patch, and the gather/scatter is done PER UNIQUE PATCH DOF:
//   * every dof of the patch is read from x once (sorted by global index -> sector-coalesced)
//     and pushed to the element slots of a shared-memory E-vector;
//   * after the element kernels, every patch dof sums its slots; dofs whose elements all lie
//     in this patch ("private") are written with plain stores, only patch-boundary dofs use
//     RED atomics.
// This file builds, for any mesh, the per-patch lists
//   dof[u]   : global (extended-vector) index | MAP_NEG never set here | flags
//   start[u] : first entry of u in slots[] (relative to the patch), U+1 entries
//   slots[s] : E-vector slot (bits 0..14) | sign flip (bit 15), grouped by dof
// with private dofs first, each class sorted by global index.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>

#include "internal.h"

namespace hcurl {

static inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffffULL;
  v = (v | v << 32) & 0x1f00000000ffffULL;
  v = (v | v << 16) & 0x1f0000ff0000ffULL;
  v = (v | v << 8) & 0x100f00f00f00f00fULL;
  v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
  v = (v | v << 2) & 0x1249249249249249ULL;
  return v;
}

template <class F>
static void parallel_for(int64_t n, F f) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 32) nt = 32;
  if ((int64_t)nt > n) nt = (unsigned)(n > 0 ? n : 1);
  std::atomic<int64_t> next(0);
  const int64_t chunk = std::max<int64_t>(1, n / (nt * 16));
  auto work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(chunk);
      if (b >= n) break;
      const int64_t e = std::min(n, b + chunk);
      for (int64_t i = b; i < e; i++) f(i);
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; t++) th.emplace_back(work);
  work();
  for (auto &t : th) t.join();
}

// centroid[e*3+d]; slot_dof[e*nd + s] = extended dof index of element slot s, slot_neg = sign flip
int build_patches(int nd, int64_t n_elem, int FE, int64_t n_ext, const double *centroid, const uint32_t *slot_code,
                  PatchHost &out) {
  if ((int64_t)FE * nd > 32767) { set_error("patch too large for the 15-bit slot encoding"); return 1; }
  const int64_t n_patch = (n_elem + FE - 1) / FE;
  out.FE = FE; out.nd = nd; out.n_patch = n_patch;
  // ---- Morton order of the centroids
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int64_t e = 0; e < n_elem; e++)
    for (int d = 0; d < 3; d++) {
      lo[d] = std::min(lo[d], centroid[e * 3 + d]);
      hi[d] = std::max(hi[d], centroid[e * 3 + d]);
    }
  // one common cell size so that Morton blocks are cubes in physical space
  double span = 0.0;
  for (int d = 0; d < 3; d++) span = std::max(span, hi[d] - lo[d]);
  if (!(span > 0.0)) span = 1.0;
  // resolution: about 2 * cbrt(n_elem) cells per axis, rounded up to a power of two, so that a
  // structured grid with power-of-two extents maps one element per cell pair exactly
  int bits = 1;
  while ((1LL << bits) < (int64_t)std::ceil(std::cbrt((double)n_elem)) && bits < 20) bits++;
  bits = std::min(20, bits + 1);
  const double inv = (double)(1LL << bits) / (span * (1.0 + 1e-9));
  std::vector<std::pair<uint64_t, int32_t>> key((size_t)n_elem);
  parallel_for(n_elem, [&](int64_t e) {
    uint64_t c[3];
    for (int d = 0; d < 3; d++) c[d] = (uint64_t)std::max(0.0, (centroid[e * 3 + d] - lo[d]) * inv);
    key[(size_t)e] = {spread21(c[0]) | spread21(c[1]) << 1 | spread21(c[2]) << 2, (int32_t)e};
  });
  std::sort(key.begin(), key.end());
  out.perm.assign((size_t)n_patch * FE, -1);
  for (int64_t i = 0; i < n_elem; i++) out.perm[(size_t)i] = key[(size_t)i].second;
  key.clear(); key.shrink_to_fit();

  // ---- per patch: entries sorted by dof, unique list
  struct Ent { uint32_t dof; uint16_t slot; };
  const int per = FE * nd;
  std::vector<std::vector<uint32_t>> udof((size_t)n_patch);     // unique dofs (ascending)
  std::vector<std::vector<uint16_t>> ustart((size_t)n_patch);   // U+1
  std::vector<std::vector<uint16_t>> uslots((size_t)n_patch);   // grouped by dof
  const int comp = nd / 3;  // E-vector index = (slot)*FE + element position
  (void)comp;
  parallel_for(n_patch, [&](int64_t k) {
    std::vector<Ent> ent;
    ent.reserve(per);
    for (int j = 0; j < FE; j++) {
      const int32_t e = out.perm[(size_t)k * FE + j];
      if (e < 0) continue;
      for (int s = 0; s < nd; s++) {
        const uint32_t code = slot_code[(size_t)e * nd + s];
        const uint16_t sl = (uint16_t)(s * FE + j) | ((code & MAP_NEG) ? 0x8000u : 0u);
        ent.push_back({code & MAP_IDX, sl});
      }
    }
    std::sort(ent.begin(), ent.end(), [](const Ent &a, const Ent &b) { return a.dof < b.dof || (a.dof == b.dof && a.slot < b.slot); });
    auto &ud = udof[(size_t)k];
    auto &us = ustart[(size_t)k];
    auto &sl = uslots[(size_t)k];
    sl.resize(ent.size());
    for (size_t i = 0; i < ent.size(); i++) {
      if (i == 0 || ent[i].dof != ent[i - 1].dof) { ud.push_back(ent[i].dof); us.push_back((uint16_t)i); }
      sl[i] = ent[i].slot;
    }
    us.push_back((uint16_t)ent.size());
  });
  // ---- which dofs are private to one patch
  std::vector<int32_t> owner((size_t)n_ext, -1);
  std::vector<uint8_t> multi((size_t)n_ext, 0);
  for (int64_t k = 0; k < n_patch; k++)
    for (uint32_t d : udof[(size_t)k]) {
      if (owner[d] < 0) owner[d] = (int32_t)k;
      else if (owner[d] != (int32_t)k) multi[d] = 1;
    }
  owner.clear(); owner.shrink_to_fit();
  // ---- offsets, final arrays (private first, each class ascending)
  out.off.assign((size_t)n_patch + 1, 0);
  for (int64_t k = 0; k < n_patch; k++) out.off[(size_t)k + 1] = out.off[(size_t)k] + (uint32_t)udof[(size_t)k].size();
  const size_t n_pd = out.off[(size_t)n_patch];
  out.npriv.assign((size_t)n_patch, 0);
  out.dof.assign(n_pd, 0);
  out.start.assign(n_pd + (size_t)n_patch, 0);
  out.slots.assign((size_t)n_patch * per, 0);
  parallel_for(n_patch, [&](int64_t k) {
    const auto &ud = udof[(size_t)k];
    const auto &us = ustart[(size_t)k];
    const auto &sl = uslots[(size_t)k];
    uint32_t *pd = out.dof.data() + out.off[(size_t)k];
    uint16_t *ps = out.start.data() + out.off[(size_t)k] + k;
    uint16_t *pl = out.slots.data() + (size_t)k * per;
    uint32_t u = 0;
    uint16_t pos = 0;
    for (int pass = 0; pass < 2; pass++) {
      for (size_t i = 0; i < ud.size(); i++) {
        if ((multi[ud[i]] != 0) != (pass == 1)) continue;
        pd[u] = ud[i];
        ps[u] = pos;
        for (uint16_t s = us[i]; s < us[i + 1]; s++) pl[pos++] = sl[s];
        u++;
      }
      if (pass == 0) out.npriv[(size_t)k] = u;
    }
    ps[u] = pos;
  });
  return 0;
}

}  // namespace hcurl

// ---- host-only C entry points (declared in include/hcurl_b200.h): the patch builder can be
// exercised without a GPU (tests/test_host_patches.py).
struct hcurl_host_patches { hcurl::PatchHost h; };

extern "C" {
hcurl_host_patches *hcurl_host_patches_create(int nd, int64_t n_elem, int elems_per_patch, int64_t n_dof,
                                              const double *centroid, const uint32_t *slot_code) {
  if (nd <= 0 || n_elem <= 0 || !centroid || !slot_code) { hcurl::set_error("hcurl_host_patches_create: bad arguments"); return nullptr; }
  for (int64_t i = 0; i < n_elem * nd; i++)
    if ((int64_t)(slot_code[i] & hcurl::MAP_IDX) >= n_dof) { hcurl::set_error("hcurl_host_patches_create: dof index out of range"); return nullptr; }
  hcurl_host_patches *p = new hcurl_host_patches;
  if (hcurl::build_patches(nd, n_elem, elems_per_patch, n_dof, centroid, slot_code, p->h)) { delete p; return nullptr; }
  return p;
}
void hcurl_host_patches_destroy(hcurl_host_patches *p) { delete p; }
int64_t hcurl_host_patches_count(const hcurl_host_patches *p) { return p->h.n_patch; }
int64_t hcurl_host_patches_ndof_entries(const hcurl_host_patches *p) { return (int64_t)p->h.dof.size(); }
int hcurl_host_patches_copy(const hcurl_host_patches *p, int32_t *perm, uint32_t *off, uint32_t *npriv, uint32_t *dof,
                            uint16_t *start, uint16_t *slots) {
  const hcurl::PatchHost &h = p->h;
  if (perm) memcpy(perm, h.perm.data(), sizeof(int32_t) * h.perm.size());
  if (off) memcpy(off, h.off.data(), sizeof(uint32_t) * h.off.size());
  if (npriv) memcpy(npriv, h.npriv.data(), sizeof(uint32_t) * h.npriv.size());
  if (dof) memcpy(dof, h.dof.data(), sizeof(uint32_t) * h.dof.size());
  if (start) memcpy(start, h.start.data(), sizeof(uint16_t) * h.start.size());
  if (slots) memcpy(slots, h.slots.data(), sizeof(uint16_t) * h.slots.size());
  return 0;
}
}
