// Host-side rank-local space for ANY mesh under a host-supplied element partition (no GPU needed).
//
// What mfem::ParMesh(comm, mesh, partitioning) + ParFiniteElementSpace produce from METIS output
// (/root/reference/examples/team7/Main.cpp:155): the rank's elements, its local (L-vector) dof numbering, true-dof
// ownership (lowest rank holding a dof owns it, MFEM's convention) and the shared-dof lists hcurl_space_create() takes.
// Built from the SERIAL space's signed element -> dof map, so every rank keeps the serial orientation and face-dof
// ordering of a shared dof (ghost signs are all +1) and the two lists of a neighbour pair are matched by ascending
// serial dof number.  Same result as hephaestus_b200/partition.py: GeneralProblem (tests/test_partition.py).
#include <algorithm>
#include <cstring>
#include <vector>

#include "internal.h"

struct hcurl_host_partition {
  int nd = 0;
  std::vector<int64_t> elems;        // serial element ids of this rank
  std::vector<int32_t> l2g;          // local dof -> serial dof (ascending)
  std::vector<int32_t> elem_ldof;    // [n_local][nd] signed local dofs
  std::vector<int32_t> ldof_ltdof;   // -1 = not owned
  int32_t n_tdof = 0;
  std::vector<int32_t> nbr_rank, own_off, own_ldof, ghost_off, ghost_ldof;
};

extern "C" {

hcurl_host_partition *hcurl_host_partition_create(int64_t n_elem, int nd, int64_t n_dof, const int32_t *elem_dofs, const int32_t *elem_rank,
                                                  int rank) {
  using hcurl::set_error;
  if (n_elem <= 0 || nd <= 0 || n_dof <= 0 || !elem_dofs || !elem_rank || rank < 0) { set_error("hcurl_host_partition_create: bad arguments"); return nullptr; }
  int world = 0;
  for (int64_t e = 0; e < n_elem; e++) {
    if (elem_rank[e] < 0) { set_error("hcurl_host_partition_create: negative rank for element %lld", (long long)e); return nullptr; }
    world = std::max(world, elem_rank[e] + 1);
  }
  auto dof_of = [&](int64_t i) { const int32_t v = elem_dofs[i]; return v >= 0 ? v : -1 - v; };
  for (int64_t i = 0; i < n_elem * nd; i++)
    if (dof_of(i) >= n_dof) { set_error("hcurl_host_partition_create: dof index out of range"); return nullptr; }
  // owner = lowest rank holding the dof; which dofs this rank holds
  std::vector<int32_t> owner((size_t)n_dof, world);
  std::vector<uint8_t> mine((size_t)n_dof, 0);
  for (int64_t e = 0; e < n_elem; e++)
    for (int s = 0; s < nd; s++) {
      const int32_t d = dof_of(e * nd + s);
      owner[(size_t)d] = std::min(owner[(size_t)d], elem_rank[e]);
      if (elem_rank[e] == rank) mine[(size_t)d] = 1;
    }
  hcurl_host_partition *P = new hcurl_host_partition;
  P->nd = nd;
  std::vector<int32_t> g2l((size_t)n_dof, -1);
  for (int64_t d = 0; d < n_dof; d++)
    if (mine[(size_t)d]) { g2l[(size_t)d] = (int32_t)P->l2g.size(); P->l2g.push_back((int32_t)d); }
  for (int64_t e = 0; e < n_elem; e++) {
    if (elem_rank[e] != rank) continue;
    P->elems.push_back(e);
    for (int s = 0; s < nd; s++) {
      const int32_t l = g2l[(size_t)dof_of(e * nd + s)];
      P->elem_ldof.push_back(elem_dofs[e * nd + s] >= 0 ? l : -1 - l);
    }
  }
  P->ldof_ltdof.assign(P->l2g.size(), -1);
  for (size_t l = 0; l < P->l2g.size(); l++)
    if (owner[(size_t)P->l2g[l]] == rank) P->ldof_ltdof[l] = P->n_tdof++;
  // per other rank q: the dofs both of us hold (ascending serial number = ascending local number on both sides)
  std::vector<std::vector<int32_t>> common((size_t)world);
  {
    std::vector<int32_t> stamp((size_t)n_dof, -1);  // last rank that recorded the dof, to list it once per rank
    for (int64_t e = 0; e < n_elem; e++) {
      const int q = elem_rank[e];
      if (q == rank) continue;
      for (int s = 0; s < nd; s++) {
        const int32_t d = dof_of(e * nd + s);
        if (mine[(size_t)d] && stamp[(size_t)d] != q) {
          // a dof may be visited again for q after another rank touched it: de-duplicate after sorting
          stamp[(size_t)d] = q;
          common[(size_t)q].push_back(d);
        }
      }
    }
  }
  P->own_off.push_back(0);
  P->ghost_off.push_back(0);
  for (int q = 0; q < world; q++) {
    auto &c = common[(size_t)q];
    if (c.empty()) continue;
    std::sort(c.begin(), c.end());
    c.erase(std::unique(c.begin(), c.end()), c.end());
    size_t n_own = 0, n_ghost = 0;
    for (int32_t d : c) {
      if (owner[(size_t)d] == rank) { P->own_ldof.push_back(g2l[(size_t)d]); n_own++; }
      else if (owner[(size_t)d] == q) { P->ghost_ldof.push_back(g2l[(size_t)d]); n_ghost++; }
    }
    if (n_own == 0 && n_ghost == 0) continue;  // shared only through dofs a third rank owns
    P->nbr_rank.push_back(q);
    P->own_off.push_back((int32_t)P->own_ldof.size());
    P->ghost_off.push_back((int32_t)P->ghost_ldof.size());
  }
  // every dof this rank does not own arrives from exactly one neighbour: its owner
  if ((int64_t)P->ghost_ldof.size() != (int64_t)P->l2g.size() - P->n_tdof) {
    set_error("hcurl_host_partition_create: inconsistent sharing (%zu ghosts listed, %lld expected)", P->ghost_ldof.size(),
              (long long)((int64_t)P->l2g.size() - P->n_tdof));
    delete P;
    return nullptr;
  }
  return P;
}

void hcurl_host_partition_destroy(hcurl_host_partition *p) { delete p; }

int hcurl_host_partition_sizes(const hcurl_host_partition *p, int64_t *n_elem_local, int32_t *n_ldof, int32_t *n_tdof, int *n_nbr,
                               int32_t *n_own, int32_t *n_ghost) {
  if (!p) { hcurl::set_error("hcurl_host_partition_sizes: null handle"); return 1; }
  if (n_elem_local) *n_elem_local = (int64_t)p->elems.size();
  if (n_ldof) *n_ldof = (int32_t)p->l2g.size();
  if (n_tdof) *n_tdof = p->n_tdof;
  if (n_nbr) *n_nbr = (int)p->nbr_rank.size();
  if (n_own) *n_own = (int32_t)p->own_ldof.size();
  if (n_ghost) *n_ghost = (int32_t)p->ghost_ldof.size();
  return 0;
}

int hcurl_host_partition_copy(const hcurl_host_partition *p, int64_t *elems, int32_t *l2g, int32_t *elem_ldof, int32_t *ldof_ltdof,
                              int32_t *nbr_rank, int32_t *own_off, int32_t *own_ldof, int32_t *ghost_off, int32_t *ghost_ldof) {
  if (!p) { hcurl::set_error("hcurl_host_partition_copy: null handle"); return 1; }
  auto cp = [](auto *dst, const auto &v) { if (dst && !v.empty()) memcpy(dst, v.data(), sizeof(v[0]) * v.size()); };
  cp(elems, p->elems); cp(l2g, p->l2g); cp(elem_ldof, p->elem_ldof); cp(ldof_ltdof, p->ldof_ltdof); cp(nbr_rank, p->nbr_rank);
  cp(own_off, p->own_off); cp(own_ldof, p->own_ldof); cp(ghost_off, p->ghost_off); cp(ghost_ldof, p->ghost_ldof);
  return 0;
}
}
