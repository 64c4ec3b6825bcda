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
// This file builds, for any mesh, one contiguous BLOB per patch (so the kernel can stage it in
// shared memory with a single cp.async.bulk).  The builder works on abstract UNITS of `unit_bytes`
// (8: a dof is one double; 16: a dof is an aligned PAIR of doubles -- the two order-2 Nedelec dofs
// along one edge / face line / interior line, which always travel together, see hex_patch.cu).
// The units of a patch are ordered by slot count -- 1 | 2 | 3..4 element slots -- each class in
// ascending global index up to a bank-spreading permutation inside aligned 128-byte runs (see
// below), so that consecutive lanes touch consecutive addresses.  Position u in this order is also
// the unit's place in the patch-local x staging vector Xs.
//
// E-vector layout: slot s of patch position j sits at unit index s*stride + j (byte offset times
// unit_bytes).  Orientation flags (sign flips, pair swap) are applied on the element side (lidx bits
// 13..15, for the gather and for the scatter), so the dof side only sums.
//
// Ownership instead of a zero-filling pass: a unit touched by several patches is OWNED by the lowest-numbered of them.
// The CTA that will process the owner ZEROES it a batch of pipeline iterations ahead (own_off / own_idx: the owned
// shared units of every patch) and the kernel keeps a split-phase watermark over these batches; every patch touching
// the unit RED-adds its sum once the watermark has passed (patch_pipeline.h) -- so y never has to be cleared before an
// apply and the zeros are still in L2 when the adds arrive.
// The units of a patch are ordered private first (plain stores), shared second (REDs), each group by slot count.
//
// Blob layout in 32-bit words: three PARTS by lifetime in the kernel's pipeline, each starting on a 16-byte boundary,
// located by blob_off[k] (16-byte units) and part_off[2k], part_off[2k+1]:
//  part D  [blob_off[k], part_off[2k])      header (HDR = 32 words) + unit words; double-buffered by the kernels
//  part L  [part_off[2k], part_off[2k+1])   slot-offset lists, overflow list, dependency list (needed while summing)
//  part T  [part_off[2k+1], blob_off[k+1])  per-slot table (needed while the element side pulls / writes results)
// Header:
//   [0] U  number of patch units   [1] number of real (non-padding) elements   [2] n_ovf   [3] 0
//   [4..6]  n1a n2a n4a  private units (stored), with 1 / 2 / 3..4 element slots
//   [7..9]  n1b n2b n4b  shared units (RED-added); those with PATCH_OWNED are zeroed ahead by this patch's CTA
//   [10..15] word offsets, relative to the start of part L, of the lists e16a e32a e64a e16b e32b e64b
//   [16] offset of ovf (part L)   [17] 0   [18] words of part D   [19] words of part L
//   [20] 1 when a unit of the patch carries an essential flag (set together with the flags, hex_patch.cu)   [21..31] 0
// Part D:  dof[U]  global unit index (bits 0..27) | PATCH_OWNED (bit 28) | PATCH_SHARED (bit 31) when other patches also
//          touch the unit; bits 29, 30 are the essential flags the operator ORs in (hex_patch.cu).  Order: the six
//          classes above.
// Part L:  e16[n1], e32[n2], e64[n4] per class: per unit 1 / 2 / 4 entries of 16 bits: BYTE offset of the slot in the
//          E-vector; 0xffff = unused (a 3-slot unit in a 64-bit section)
//          ovf[n_ovf]  (u << 16 | byte offset) for the slots beyond four (edge valence > 4 inside one patch, only on
//          unstructured meshes); such units are classed shared (zeroed ahead, every slot sum RED-added).
// Part T:  lidx[nd][FE] 16 bits per element slot: Xs position u of the slot's unit (bits 0..12) | the slot
//          code's flag bits 29..31 (bits 13..15); the element side PULLS its inputs through this table.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
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
int build_patches(int nd, int64_t n_elem, int FE, int stride, int unit_bytes, int64_t n_ext, const double *centroid,
                  const uint32_t *slot_code, const uint8_t *exchange_units, PatchHost &out) {
  if (stride < FE) { set_error("slot stride must be >= elements per patch"); return 1; }
  if (unit_bytes != 8 && unit_bytes != 16) { set_error("unit size must be 8 or 16 bytes"); return 1; }
  if ((int64_t)stride * nd * unit_bytes > 65534 || (int64_t)FE * nd > 8191) { set_error("patch too large for the 16-bit slot encoding"); return 1; }
  if (n_ext > (int64_t)PATCH_IDX + 1) { set_error("too many units for the patch unit words"); return 1; }
  const int64_t n_patch = (n_elem + FE - 1) / FE;
  out.FE = FE; out.stride = stride; out.nd = nd; out.n_patch = n_patch;
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
  // ---- patches that touch a unit taking part in the inter-rank exchange (ghosts, owned units other ranks hold as
  // ghosts) come first, so that they can run between the two halves of the exchange while the rest overlaps with it
  out.n_boundary = 0;
  if (exchange_units) {
    std::vector<uint8_t> bdr((size_t)n_patch, 0);
    parallel_for(n_patch, [&](int64_t k) {
      for (int j = 0; j < FE && !bdr[(size_t)k]; j++) {
        const int32_t e = out.perm[(size_t)k * FE + j];
        if (e < 0) continue;
        for (int s = 0; s < nd; s++)
          if (exchange_units[slot_code[(size_t)e * nd + s] & SLOT_IDX]) { bdr[(size_t)k] = 1; break; }
      }
    });
    std::vector<int32_t> perm2((size_t)n_patch * FE, -1);
    int64_t q = 0;
    // the padded patch (the last one in Morton order) keeps its place at the end of its group
    for (int pass = 1; pass >= 0; pass--)
      for (int64_t k = 0; k < n_patch; k++)
        if (bdr[(size_t)k] == pass) {
          std::copy(out.perm.begin() + k * FE, out.perm.begin() + (k + 1) * FE, perm2.begin() + q * FE);
          q++;
          if (pass == 1) out.n_boundary++;
        }
    out.perm.swap(perm2);
  }

  // ---- per patch: entries sorted by dof, unique list
  struct Ent { uint32_t dof; uint16_t slot; uint16_t sj; };  // slot = E byte offset, sj = s*FE + j | flags << 13
  const int per = FE * nd;
  std::vector<std::vector<uint32_t>> udof((size_t)n_patch);     // unique dofs (ascending)
  std::vector<std::vector<uint16_t>> ustart((size_t)n_patch);   // U+1
  std::vector<std::vector<uint16_t>> usj((size_t)n_patch);      // s*FE + j | flags, same grouping
  parallel_for(n_patch, [&](int64_t k) {
    std::vector<Ent> ent;
    ent.reserve(per);
    for (int j = 0; j < FE; j++) {
      const int32_t e = out.perm[(size_t)k * FE + j];
      if (e < 0) continue;
      for (int s = 0; s < nd; s++) {
        const uint32_t code = slot_code[(size_t)e * nd + s];
        const uint16_t sl = (uint16_t)((s * stride + j) * unit_bytes);
        ent.push_back({code & SLOT_IDX, sl, (uint16_t)((uint32_t)(s * FE + j) | ((code >> 29) << 13))});
      }
    }
    std::sort(ent.begin(), ent.end(), [](const Ent &a, const Ent &b) { return a.dof < b.dof || (a.dof == b.dof && a.slot < b.slot); });
    auto &ud = udof[(size_t)k];
    auto &us = ustart[(size_t)k];
    auto &sj = usj[(size_t)k];
    sj.resize(ent.size());
    for (size_t i = 0; i < ent.size(); i++) {
      if (i == 0 || ent[i].dof != ent[i - 1].dof) { ud.push_back(ent[i].dof); us.push_back((uint16_t)i); }
      sj[i] = ent[i].sj;
    }
    us.push_back((uint16_t)ent.size());
  });
  // ---- owner of every unit = the lowest-numbered patch touching it; multi = touched by more than one patch
  std::vector<int32_t> owner((size_t)n_ext, -1);
  std::vector<uint8_t> multi((size_t)n_ext, 0);
  for (int64_t k = 0; k < n_patch; k++) {
    const auto &ud = udof[(size_t)k];
    const auto &us = ustart[(size_t)k];
    for (size_t i = 0; i < ud.size(); i++) {
      const uint32_t d = ud[i];
      if (owner[d] < 0) owner[d] = (int32_t)k;
      else if (owner[d] != (int32_t)k) multi[d] = 1;
      if (us[i + 1] - us[i] > 4) multi[d] = 1;  // overflow slots are RED-added: the unit must be pre-filled
    }
  }
  // class of unit i of patch k: (shared ? 3 : 0) + width
  auto width_of = [](int nslot) { return nslot <= 1 ? 0 : (nslot == 2 ? 1 : 2); };
  auto class_of = [&](int64_t k, size_t i) {
    const auto &us = ustart[(size_t)k];
    return (multi[udof[(size_t)k][i]] ? 3 : 0) + width_of(us[i + 1] - us[i]);
  };
  auto align4 = [](uint32_t w) { return (w + 3u) & ~3u; };  // words -> 16-byte boundary
  constexpr uint32_t HDR = PatchHost::HDR;
  const uint32_t lidx_words = align4((uint32_t)(per + 1) / 2);
  // the owned shared units of every patch, ascending (zeroed ahead of the sums by the CTA that processes the patch)
  out.own_off.assign((size_t)n_patch + 1, 0);
  for (int64_t k = 0; k < n_patch; k++) {
    uint32_t n = 0;
    for (uint32_t d : udof[(size_t)k]) n += multi[d] && owner[d] == (int32_t)k;
    out.own_off[(size_t)k + 1] = out.own_off[(size_t)k] + n;
  }
  out.own_idx.assign((size_t)out.own_off[(size_t)n_patch], 0);
  parallel_for(n_patch, [&](int64_t k) {
    uint32_t *o = out.own_idx.data() + out.own_off[(size_t)k];
    for (uint32_t d : udof[(size_t)k]) if (multi[d] && owner[d] == (int32_t)k) *o++ = d;
  });
  out.blob_off.assign((size_t)n_patch + 1, 0);
  out.max_blob_words = 0;
  out.max_dofs = 0;
  out.max_d_words = out.max_l_words = 0;
  out.max_t_words = lidx_words;
  out.part_off.assign((size_t)n_patch * 2, 0);
  struct Sizes { uint32_t n[6], nov, wd, wl; };
  auto sizes_of = [&](int64_t k) {
    const auto &ud = udof[(size_t)k];
    const auto &us = ustart[(size_t)k];
    Sizes z = {{0, 0, 0, 0, 0, 0}, 0, 0, 0};
    for (size_t i = 0; i < ud.size(); i++) {
      z.n[class_of(k, i)]++;
      if (us[i + 1] - us[i] > 4) z.nov += us[i + 1] - us[i] - 4;
    }
    z.wd = HDR + align4((uint32_t)ud.size());
    z.wl = 0;
    for (int ph = 0; ph < 2; ph++) z.wl += align4((z.n[3 * ph] + 1) / 2) + align4(z.n[3 * ph + 1]) + align4(2 * z.n[3 * ph + 2]);
    z.wl += align4(z.nov);
    return z;
  };
  for (int64_t k = 0; k < n_patch; k++) {
    const Sizes z = sizes_of(k);
    const uint32_t w = z.wd + z.wl + lidx_words;
    out.blob_off[(size_t)k + 1] = out.blob_off[(size_t)k] + w / 4;
    out.part_off[(size_t)k * 2] = out.blob_off[(size_t)k] + z.wd / 4;
    out.part_off[(size_t)k * 2 + 1] = out.blob_off[(size_t)k] + (z.wd + z.wl) / 4;
    out.max_d_words = std::max(out.max_d_words, z.wd);
    out.max_l_words = std::max(out.max_l_words, z.wl);
    out.max_blob_words = std::max(out.max_blob_words, w);
    out.max_dofs = std::max(out.max_dofs, (uint32_t)udof[(size_t)k].size());
  }
  // position -> index into ud, for patch k: the six classes in order, each in ascending global index (consecutive lanes
  // gather / store consecutive addresses) up to a bank-spreading permutation: the position of a unit is also its place in
  // the staging vector Xs, so its shared-memory bank on the element side (the pull through lidx) is position mod NB, NB =
  // 128 / unit_bytes lanes being served per wavefront.  Sorted order makes the pull 4.4-way bank conflicted on structured
  // meshes (strided dof numbering), so inside every aligned run of NB positions the units are PERMUTED to spread, for each
  // (slot, NB consecutive elements) access group, the units it reads over distinct banks.  A warp of the dof side still
  // touches the same 32 units, i.e. the same global sectors.
  auto make_order = [&](int64_t k, std::vector<uint32_t> &order, uint32_t *cls_end) {
    const auto &ud = udof[(size_t)k];
    const auto &us = ustart[(size_t)k];
    const auto &sj = usj[(size_t)k];
    order.clear();
    order.reserve(ud.size());
    for (int w = 0; w < 6; w++) {
      for (size_t i = 0; i < ud.size(); i++)
        if (class_of(k, i) == w) order.push_back((uint32_t)i);
      cls_end[w] = (uint32_t)order.size();
    }
    const int NB = 128 / unit_bytes, HG = (FE + NB - 1) / NB;
    std::vector<uint8_t> cnt((size_t)nd * HG * 16, 0);
    auto groups_of = [&](uint32_t i, int *g) {  // distinct access groups reading dof i
      int n = 0;
      for (int q = us[i]; q < us[i + 1] && n < 8; q++) {
        const int v = sj[q] & 0x1fff, gid = (v / FE) * HG + (v % FE) / NB;
        bool dup = false;
        for (int t = 0; t < n; t++) dup |= g[t] == gid;
        if (!dup) g[n++] = gid;
      }
      return n;
    };
    auto add = [&](uint32_t i, int bank, int v) {
      int g[8];
      const int n = groups_of(i, g);
      for (int t = 0; t < n; t++) cnt[(size_t)g[t] * 16 + bank] = (uint8_t)(cnt[(size_t)g[t] * 16 + bank] + v);
    };
    for (int pass = 0; pass < 2; pass++) {
      uint32_t p0 = 0;
      for (int w = 0; w < 6; w++)
        while (p0 < cls_end[w]) {
          const uint32_t p1 = std::min(cls_end[w], (p0 / NB + 1) * NB);
          const int m = (int)(p1 - p0);
          uint32_t ds[16];
          for (int t = 0; t < m; t++) ds[t] = order[p0 + t];
          if (pass > 0) for (int t = 0; t < m; t++) add(ds[t], (int)((p0 + t) % NB), -1);
          bool used[16] = {false};
          // dofs read by more groups first
          std::stable_sort(ds, ds + m, [&](uint32_t a, uint32_t b) { return us[a + 1] - us[a] > us[b + 1] - us[b]; });
          for (int t = 0; t < m; t++) {
            int g[8];
            const int n = groups_of(ds[t], g);
            int best = -1, best_cost = 1 << 30;
            for (int q = 0; q < m; q++) {
              if (used[q]) continue;
              const int bank = (int)((p0 + q) % NB);
              int c = 0;
              for (int r = 0; r < n; r++) c += 2 * cnt[(size_t)g[r] * 16 + bank] + 1;
              if (c < best_cost) { best_cost = c; best = q; }
            }
            used[best] = true;
            order[p0 + best] = ds[t];
            add(ds[t], (int)((p0 + best) % NB), +1);
          }
          p0 = p1;
        }
    }
  };
  // ---- swizzle of the E-vector positions (see PatchHost::swz_row): coordinate descent on the bank conflicts of the
  // dof side's slot reads -- lanes = NB consecutive units of a class, one read per slot rank -- over a sample
  // of patches.  Structured meshes number dofs in strides that make neighbouring units hit the same bank (2.8
  // wavefronts per read measured); one global swizzle removes about half of the excess.
  {
    const int NB = 128 / unit_bytes, NG = (FE + NB - 1) / NB;
    out.swz_row.assign((size_t)nd, 0);
    out.swz_grp.assign((size_t)NG, 0);
    struct Acc { uint16_t s, j; };
    std::vector<std::vector<Acc>> wins;  // one entry per (window, slot rank): the slots read together
    const int64_t n_sample = std::min<int64_t>(n_patch, 48);
    std::vector<uint32_t> order;
    for (int64_t q = 0; q < n_sample; q++) {
      const int64_t k = q * n_patch / n_sample;
      const auto &us = ustart[(size_t)k];
      const auto &sj = usj[(size_t)k];
      uint32_t cls_end[6];
      make_order(k, order, cls_end);
      for (int w = 0; w < 6; w++)
        for (uint32_t p0 = w ? cls_end[w - 1] : 0; p0 < cls_end[w]; p0 += NB)
          for (int kk = 0; kk < (w % 3 == 0 ? 1 : (w % 3 == 1 ? 2 : 4)); kk++) {
            std::vector<Acc> acc;
            for (uint32_t ul = p0; ul < std::min(cls_end[w], p0 + NB); ul++) {
              const uint32_t i = order[ul];
              if (us[i] + kk < us[i + 1]) { const int v = sj[us[i] + kk] & 0x1fff; acc.push_back({(uint16_t)(v / FE), (uint16_t)(v % FE)}); }
            }
            if (acc.size() > 1) wins.push_back(acc);
          }
    }
    auto cost = [&]() {
      int64_t c = 0;
      for (const auto &acc : wins) {
        int cnt[16] = {0}, mx = 0;
        for (const Acc &e : acc) {
          const int pos = e.j ^ out.swz_row[e.s] ^ out.swz_grp[e.j / NB];
          mx = std::max(mx, ++cnt[(e.s * stride + pos) % NB]);
        }
        c += mx;
      }
      return c;
    };
    int64_t best = cost();
    for (int sweep = 0; sweep < 3; sweep++)
      for (int var = 0; var < nd + NG; var++) {
        uint8_t &x = var < nd ? out.swz_row[(size_t)var] : out.swz_grp[(size_t)(var - nd)];
        uint8_t keep = x;
        for (int v = 0; v < NB; v++) {
          x = (uint8_t)v;
          const int64_t c = cost();
          if (c < best) { best = c; keep = (uint8_t)v; }
        }
        x = keep;
      }
  }
  out.blob.assign((size_t)out.blob_off[(size_t)n_patch] * 4, 0);
  parallel_for(n_patch, [&](int64_t k) {
    const auto &ud = udof[(size_t)k];
    const auto &us = ustart[(size_t)k];
    const auto &sj = usj[(size_t)k];
    uint32_t *B = out.blob.data() + (size_t)out.blob_off[(size_t)k] * 4;
    const Sizes z = sizes_of(k);
    const uint32_t U = (uint32_t)ud.size();
    uint32_t oL[6], o = 0;  // list offsets relative to part L
    for (int w = 0; w < 6; w++) {
      oL[w] = o;
      o += w % 3 == 0 ? align4((z.n[w] + 1) / 2) : (w % 3 == 1 ? align4(z.n[w]) : align4(2 * z.n[w]));
    }
    const uint32_t oov = o;
    uint32_t n_real = 0;
    for (int j = 0; j < FE; j++) n_real += out.perm[(size_t)k * FE + j] >= 0;
    B[0] = U; B[1] = n_real; B[2] = z.nov;
    for (int w = 0; w < 6; w++) { B[4 + w] = z.n[w]; B[10 + w] = oL[w]; }
    B[16] = oov; B[18] = z.wd; B[19] = z.wl;
    uint32_t *pd = B + HDR;
    uint32_t *Lp = B + z.wd;
    uint32_t *po = Lp + oov;
    uint16_t *plx = (uint16_t *)(Lp + z.wl);
    std::vector<uint32_t> order;  // position -> index into ud
    uint32_t cls_end[6];
    make_order(k, order, cls_end);
    for (uint32_t ul = 0; ul < U; ul++) {
      const uint32_t i = order[ul];
      const int ns = us[i + 1] - us[i], w = class_of(k, i);
      const uint32_t r = ul - (w ? cls_end[w - 1] : 0);  // index inside the class
      pd[ul] = ud[i] | (multi[ud[i]] ? PATCH_SHARED | (owner[ud[i]] == (int32_t)k ? PATCH_OWNED : 0u) : 0u);
      for (int j = 0; j < ns; j++) plx[sj[us[i] + j] & 0x1fffu] = (uint16_t)(ul | (sj[us[i] + j] & 0xe000u));
      auto off = [&](int q) {  // byte offset of the slot in the (swizzled) E-vector
        const int v = sj[q] & 0x1fff, ss = v / FE, jj = v % FE;
        return (uint32_t)((ss * stride + (jj ^ out.swz_row[(size_t)ss] ^ out.swz_grp[(size_t)(jj / (128 / unit_bytes))])) * unit_bytes);
      };
      if (w % 3 == 0) ((uint16_t *)(Lp + oL[w]))[r] = (uint16_t)off(us[i]);
      else if (w % 3 == 1) (Lp + oL[w])[r] = off(us[i]) | (off(us[i] + 1) << 16);
      else {
        uint64_t v = ~0ULL;
        for (int j = 0; j < ns; j++) {
          if (j < 4) v = (v & ~(0xffffULL << (16 * j))) | ((uint64_t)off(us[i] + j) << (16 * j));
          else *po++ = (ul << 16) | off(us[i] + j);
        }
        (Lp + oL[w])[2 * r] = (uint32_t)v;
        (Lp + oL[w])[2 * r + 1] = (uint32_t)(v >> 32);
      }
    }
  });
  return 0;
}

int hex_rot_slot_to_tensor(int p, int slot) {
  const int p1 = p + 1, nc = p * p1 * p1;
  const int c = slot / nc, t = slot % nc;
  int ir[3] = {t % p, (t / p) % p1, t / (p * p1)};  // along axes c, c+1, c+2
  int idx[3];
  for (int k = 0; k < 3; k++) idx[(c + k) % 3] = ir[k];
  const int n0 = (c == 0) ? p : p1, n1 = (c == 1) ? p : p1;
  return c * nc + idx[0] + n0 * (idx[1] + n1 * idx[2]);
}

// Pair codes of an order-2 hex space (the units of the patch kernel, hex_patch.cu): in the rotated slot layout the
// slots 2r, 2r+1 of an element are the two dofs along one line; they must be the doubles 2q, 2q+1 of the (extended)
// vector, in this order or swapped, each with its own sign.  code[e][r] = q | PAIR_SWAP | PAIR_NEG0 | PAIR_NEG1.
// Returns 1 (and leaves code partly written) when some slot pair does not have this structure.
int hex_p2_pair_codes(int64_t n_elem, const int32_t *elem_ldof, const int32_t *l2e, uint32_t *code) {
  const int p = 2, nd = 54, np = 27;
  std::vector<int> loc, sgn;
  nd_hex_dof_map(p, loc, sgn);
  int tens[54];
  for (int s = 0; s < nd; s++) tens[s] = hex_rot_slot_to_tensor(p, s);
  std::atomic<int> bad(0);
  parallel_for(n_elem, [&](int64_t e) {
    for (int r = 0; r < np; r++) {
      uint32_t ext[2];
      bool neg[2];
      for (int h = 0; h < 2; h++) {
        const int t = tens[2 * r + h];
        const int32_t v = elem_ldof[e * nd + loc[t]];
        const int32_t kk = v >= 0 ? v : -1 - v;
        neg[h] = (v < 0) != (sgn[t] < 0);
        ext[h] = l2e ? (uint32_t)l2e[kk] : (uint32_t)kk;
      }
      const bool swap = ext[0] > ext[1];
      const uint32_t lo = swap ? ext[1] : ext[0], hi = swap ? ext[0] : ext[1];
      if ((lo & 1u) || hi != lo + 1) { bad.store(1); return; }
      code[(size_t)e * np + r] = (lo >> 1) | (swap ? PAIR_SWAP : 0u) | (neg[0] ? PAIR_NEG0 : 0u) | (neg[1] ? PAIR_NEG1 : 0u);
    }
  });
  return bad.load();
}

// Single-dof slot codes in tensor order (the 8-byte units of the order-1 patch kernel, hex_patch_p1.cu).
void hex_slot_codes(int p, int64_t n_elem, const int32_t *elem_ldof, const int32_t *l2e, uint32_t *code) {
  const int nd = nd_hex_ndof(p);
  std::vector<int> loc, sgn;
  nd_hex_dof_map(p, loc, sgn);
  parallel_for(n_elem, [&](int64_t e) {
    for (int t = 0; t < nd; t++) {
      const int32_t v = elem_ldof[e * nd + loc[t]];
      const int32_t kk = v >= 0 ? v : -1 - v;
      const bool neg = (v < 0) != (sgn[t] < 0);
      code[(size_t)e * nd + t] = (l2e ? (uint32_t)l2e[kk] : (uint32_t)kk) | (neg ? 0x80000000u : 0u);
    }
  });
}

}  // namespace hcurl

// ---- host-only C entry points (declared in include/hcurl_b200.h): the patch builder can be
// exercised without a GPU (tests/test_host_patches.py).
struct hcurl_host_patches { hcurl::PatchHost h; };

extern "C" {
hcurl_host_patches *hcurl_host_patches_create(int nd, int64_t n_elem, int elems_per_patch, int slot_stride, int unit_bytes,
                                              int64_t n_dof, const double *centroid, const uint32_t *slot_code,
                                              const uint8_t *exchange_units) {
  if (nd <= 0 || n_elem <= 0 || !centroid || !slot_code) { hcurl::set_error("hcurl_host_patches_create: bad arguments"); return nullptr; }
  for (int64_t i = 0; i < n_elem * nd; i++)
    if ((int64_t)(slot_code[i] & hcurl::SLOT_IDX) >= n_dof) { hcurl::set_error("hcurl_host_patches_create: dof index out of range"); return nullptr; }
  hcurl_host_patches *p = new hcurl_host_patches;
  if (hcurl::build_patches(nd, n_elem, elems_per_patch, slot_stride, unit_bytes, n_dof, centroid, slot_code, exchange_units, p->h)) { delete p; return nullptr; }
  return p;
}
void hcurl_host_patches_destroy(hcurl_host_patches *p) { delete p; }
/* order-2 hex space: the 27 dof pairs of every element as the patch kernel uses them (see hex_p2_pair_codes above);
 * returns 0, or 1 when the map does not consist of aligned pairs (the operator then uses the element-wise kernels) */
int hcurl_host_hex_p2_pair_codes(int64_t n_elem, const int32_t *elem_ldof, uint32_t *codes) {
  if (n_elem <= 0 || !elem_ldof || !codes) { hcurl::set_error("hcurl_host_hex_p2_pair_codes: bad arguments"); return 1; }
  if (hcurl::hex_p2_pair_codes(n_elem, elem_ldof, nullptr, codes)) { hcurl::set_error("the element -> dof map does not consist of aligned dof pairs"); return 1; }
  return 0;
}
int hcurl_host_hex_slot_codes(int order, int64_t n_elem, const int32_t *elem_ldof, uint32_t *codes) {
  if (order < 1 || order > 4 || n_elem <= 0 || !elem_ldof || !codes) { hcurl::set_error("hcurl_host_hex_slot_codes: bad arguments"); return 1; }
  hcurl::hex_slot_codes(order, n_elem, elem_ldof, nullptr, codes);
  return 0;
}
int64_t hcurl_host_patches_count(const hcurl_host_patches *p) { return p->h.n_patch; }
int64_t hcurl_host_patches_boundary_count(const hcurl_host_patches *p) { return p->h.n_boundary; }
/* total blob size in 32-bit words; largest single patch blob in words */
int64_t hcurl_host_patches_blob_words(const hcurl_host_patches *p) { return (int64_t)p->h.blob.size(); }
int64_t hcurl_host_patches_max_blob_words(const hcurl_host_patches *p) { return (int64_t)p->h.max_blob_words; }
int64_t hcurl_host_patches_max_dofs(const hcurl_host_patches *p) { return (int64_t)p->h.max_dofs; }
int hcurl_host_patches_swizzle(const hcurl_host_patches *p, uint8_t *row, uint8_t *grp) {
  memcpy(row, p->h.swz_row.data(), p->h.swz_row.size());
  memcpy(grp, p->h.swz_grp.data(), p->h.swz_grp.size());
  return 0;
}
/* the blob of patch k by lifetime: part_off [2*n_patch] = begin of the lists and of the per-slot table in 16-byte
 * units; largest header + unit list / set of lists / table in 32-bit words */
int hcurl_host_patches_parts(const hcurl_host_patches *p, uint32_t *part_off, int64_t *max_d_words, int64_t *max_l_words, int64_t *max_t_words) {
  const hcurl::PatchHost &h = p->h;
  if (part_off) memcpy(part_off, h.part_off.data(), sizeof(uint32_t) * h.part_off.size());
  if (max_d_words) *max_d_words = h.max_d_words;
  if (max_l_words) *max_l_words = h.max_l_words;
  if (max_t_words) *max_t_words = h.max_t_words;
  return 0;
}
/* the shared units every patch owns (it is the lowest-numbered patch touching them): own_off [n_patch+1], own_idx
 * [own_off[n_patch]]; a null own_idx only returns the offsets */
int64_t hcurl_host_patches_owned(const hcurl_host_patches *p, uint32_t *own_off, uint32_t *own_idx) {
  const hcurl::PatchHost &h = p->h;
  if (own_off) memcpy(own_off, h.own_off.data(), sizeof(uint32_t) * h.own_off.size());
  if (own_idx) memcpy(own_idx, h.own_idx.data(), sizeof(uint32_t) * h.own_idx.size());
  return (int64_t)h.own_idx.size();
}
int hcurl_host_patches_copy(const hcurl_host_patches *p, int32_t *perm, uint32_t *blob_off, uint32_t *blob) {
  const hcurl::PatchHost &h = p->h;
  if (perm) memcpy(perm, h.perm.data(), sizeof(int32_t) * h.perm.size());
  if (blob_off) memcpy(blob_off, h.blob_off.data(), sizeof(uint32_t) * h.blob_off.size());
  if (blob) memcpy(blob, h.blob.data(), sizeof(uint32_t) * h.blob.size());
  return 0;
}
}
