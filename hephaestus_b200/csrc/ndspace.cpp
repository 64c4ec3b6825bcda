// Host-side Nedelec (ND) space on a hexahedral or tetrahedral mesh: global edge / face enumeration and the signed
// element -> dof map in MFEM's conventions (setup time only, no GPU needed).
//
// In the reference this is what mfem::ParFiniteElementSpace provides once
// ProblemBuilder::AddFESpace("ND_3D_P<p>") has run (/root/reference/src/problem_builders/problem_builder_base.cpp:98-127):
// FiniteElementSpace::GetElementDofs returns, per element, the global dof of every local basis function with the
// orientation encoded as k (aligned) or -1-k (flipped).  A stand-alone caller (benchmarks, tests, a plug-in that is not
// handed MFEM's arrays) gets the same map from here; SURVEY.md 8(f4).  Conventions (restated in
// hephaestus_b200/fespace.py and oracle/space.py, against which tests/test_ndspace.py checks this file bit for bit):
//   * global edges / faces are numbered by first appearance in element order, local entity order;
//   * dofs: p per edge (numbered along the edge from its lower to its higher vertex id), then 2p(p-1) per face (in the
//     frame of the face's first occurrence: its "base" vertex order), then 3p(p-1)^2 per element;
//   * an element that sees an edge from its higher vertex reverses the edge's dofs and flips their signs; a face seen
//     in another orientation (8 of them) gets a signed permutation of the base face dofs.
#include <algorithm>
#include <cstring>

#include "internal.h"

namespace {

const int HEX_EDGES[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int HEX_FACES[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
const int TET_EDGES[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
const int TET_FACES[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};

// Signed permutation of the 2p(p-1) dofs of a quadrilateral face for the 8 ways an element can see it:
// local face dof m == sign[o][m] * base face dof perm[o][m], o = 2*pos0 + flip, pos0 = position of the element's first
// face vertex in the base vertex list, flip = 1 when the element walks the face the other way round.
// Face dof layout: first the dofs tangential to the first face direction (p open x (p-1) interior closed nodes, open
// index fastest), then those tangential to the second.
void quad_orientation_tables(int p, std::vector<int> &perm, std::vector<int> &sign) {
  const int nfd = 2 * p * (p - 1);
  perm.assign((size_t)8 * std::max(nfd, 1), 0);
  sign.assign((size_t)8 * std::max(nfd, 1), 1);
  if (nfd == 0) return;
  struct Dof { int dir, io, ic; };
  std::vector<Dof> layout;
  for (int j = 1; j < p; j++) for (int i = 0; i < p; i++) layout.push_back({0, i, j});
  for (int j = 0; j < p; j++) for (int i = 1; i < p; i++) layout.push_back({1, j, i});
  auto index_of = [&](int dir, int io, int ic) {
    for (int m = 0; m < nfd; m++) if (layout[m].dir == dir && layout[m].io == io && layout[m].ic == ic) return m;
    return -1;
  };
  const int corners[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
  for (int pos0 = 0; pos0 < 4; pos0++)
    for (int flip = 0; flip < 2; flip++) {
      const int step = flip ? -1 : 1;
      int pos[4];
      for (int k = 0; k < 4; k++) pos[k] = ((pos0 + step * k) % 4 + 4) % 4;
      const int *P0 = corners[pos[0]];
      const int es[2] = {corners[pos[1]][0] - P0[0], corners[pos[1]][1] - P0[1]};
      const int et[2] = {corners[pos[3]][0] - P0[0], corners[pos[3]][1] - P0[1]};
      for (int m = 0; m < nfd; m++) {
        const int *tan = layout[m].dir == 0 ? es : et, *oth = layout[m].dir == 0 ? et : es;
        const int bdir = tan[0] != 0 ? 0 : 1;
        const int sg = tan[bdir], so = oth[1 - bdir];
        const int io = sg > 0 ? layout[m].io : p - 1 - layout[m].io, ic = so > 0 ? layout[m].ic : p - layout[m].ic;
        perm[(size_t)(2 * pos0 + flip) * nfd + m] = index_of(bdir, io, ic);
        sign[(size_t)(2 * pos0 + flip) * nfd + m] = sg;
      }
    }
}

}  // namespace

struct hcurl_host_ndspace {
  bool hex = true;
  int nv = 8, nle = 12, nlf = 6, nfv = 4;  // vertices, edges, faces per element; vertices per face
  int p = 1, nd = 0, ned = 0, nfd = 0, nint = 0;
  int64_t n_elem = 0, n_edges = 0, n_faces = 0, ndof = 0, face_off = 0, int_off = 0;
  std::vector<int32_t> el_edges;    // [n_elem][nle]
  std::vector<int32_t> el_faces;    // [n_elem][nlf]
  std::vector<uint8_t> edge_fwd;    // [n_elem][nle] element walks the edge from its lower to its higher vertex
  std::vector<uint8_t> face_code;   // hexes: [n_elem][6] orientation code 2*pos0 + flip relative to the face's base order
  std::vector<int32_t> face_base;   // [n_faces][nfv] vertex order of the first occurrence of every face
  std::vector<int32_t> face_count;  // elements per face (1 = boundary)
  std::vector<int> perm, sign;      // quad orientation tables [8][nfd]
  // lookup chains (hanging off the smallest vertex) kept for queries by vertex list
  std::vector<int32_t> e_head, e_next, e_hi, f_head, f_next, f_s1, f_s2;
};

extern "C" {

static hcurl_host_ndspace *ndspace_create(bool hex, int order, int64_t n_elem, const int32_t *elem_verts) {
  const int max_order = hex ? 4 : 3;
  if (order < 1 || order > max_order || n_elem <= 0 || !elem_verts) { hcurl::set_error("hcurl_host_ndspace_create: bad arguments"); return nullptr; }
  hcurl_host_ndspace *S = new hcurl_host_ndspace;
  const int p = order;
  S->hex = hex; S->p = p; S->n_elem = n_elem;
  S->nv = hex ? 8 : 4; S->nle = hex ? 12 : 6; S->nlf = hex ? 6 : 4; S->nfv = hex ? 4 : 3;
  const int nv_e = S->nv, nle = S->nle, nlf = S->nlf, nfv = S->nfv;
  for (int64_t i = 0; i < n_elem * nv_e; i++)
    if (elem_verts[i] < 0) { hcurl::set_error("hcurl_host_ndspace_create: negative vertex id"); delete S; return nullptr; }
  S->ned = p;
  S->nfd = hex ? 2 * p * (p - 1) : p * (p - 1);
  S->nint = hex ? 3 * p * (p - 1) * (p - 1) : p * (p - 1) * (p - 2) / 2;
  S->nd = hex ? 3 * p * (p + 1) * (p + 1) : p * (p + 2) * (p + 3) / 2;
  S->el_edges.resize((size_t)n_elem * nle); S->edge_fwd.resize((size_t)n_elem * nle);
  S->el_faces.resize((size_t)n_elem * nlf);
  if (hex) S->face_code.resize((size_t)n_elem * nlf);
  auto ledge = [&](int le, int k) { return hex ? HEX_EDGES[le][k] : TET_EDGES[le][k]; };
  auto lface = [&](int lf, int k) { return hex ? HEX_FACES[lf][k] : TET_FACES[lf][k]; };
  // Lookup tables are chains hanging off the smallest vertex of the entity (a vertex has a handful of edges / faces):
  // no hashing, no per-node allocation.
  int64_t nv = 0;
  for (int64_t i = 0; i < n_elem * nv_e; i++) nv = std::max<int64_t>(nv, elem_verts[i] + 1);
  // ---- edges, numbered by first appearance
  {
    std::vector<int32_t> head((size_t)nv, -1), next, hi;
    next.reserve((size_t)n_elem * 4); hi.reserve((size_t)n_elem * 4);
    for (int64_t e = 0; e < n_elem; e++)
      for (int le = 0; le < nle; le++) {
        const int32_t a = elem_verts[e * nv_e + ledge(le, 0)], b = elem_verts[e * nv_e + ledge(le, 1)];
        const int32_t lo = std::min(a, b), h = std::max(a, b);
        int32_t id = head[(size_t)lo];
        while (id >= 0 && hi[(size_t)id] != h) id = next[(size_t)id];
        if (id < 0) {
          id = (int32_t)hi.size();
          hi.push_back(h);
          next.push_back(head[(size_t)lo]);
          head[(size_t)lo] = id;
        }
        S->el_edges[(size_t)e * nle + le] = id;
        S->edge_fwd[(size_t)e * nle + le] = a < b;
      }
    S->n_edges = (int64_t)hi.size();
    S->e_head.swap(head); S->e_next.swap(next); S->e_hi.swap(hi);
  }
  // ---- faces: identified by their three smallest vertices; the first occurrence fixes the base vertex order
  {
    std::vector<int32_t> head((size_t)nv, -1), next, s1, s2;
    std::vector<int32_t> &base = S->face_base;
    next.reserve((size_t)n_elem * 4); s1.reserve((size_t)n_elem * 4); s2.reserve((size_t)n_elem * 4); base.reserve((size_t)n_elem * 4 * nfv);
    for (int64_t e = 0; e < n_elem; e++)
      for (int lf = 0; lf < nlf; lf++) {
        int32_t v[4], s[4];
        for (int k = 0; k < nfv; k++) s[k] = v[k] = elem_verts[e * nv_e + lface(lf, k)];
        std::sort(s, s + nfv);
        int32_t f = head[(size_t)s[0]];
        while (f >= 0 && !(s1[(size_t)f] == s[1] && s2[(size_t)f] == s[2])) f = next[(size_t)f];
        if (f < 0) {
          f = (int32_t)s1.size();
          s1.push_back(s[1]); s2.push_back(s[2]);
          next.push_back(head[(size_t)s[0]]);
          head[(size_t)s[0]] = f;
          for (int k = 0; k < nfv; k++) base.push_back(v[k]);
          S->face_count.push_back(0);
        }
        S->face_count[(size_t)f]++;
        S->el_faces[(size_t)e * nlf + lf] = f;
        if (hex) {
          const int32_t *bv = base.data() + (size_t)f * 4;
          int pos0 = 0;
          while (pos0 < 4 && bv[pos0] != v[0]) pos0++;
          if (pos0 == 4) { hcurl::set_error("hcurl_host_ndspace_create: inconsistent face %d of element %lld", lf, (long long)e); delete S; return nullptr; }
          const int flip = bv[(pos0 + 1) % 4] != v[1];
          S->face_code[(size_t)e * nlf + lf] = (uint8_t)(2 * pos0 + flip);
        }
      }
    S->n_faces = (int64_t)s1.size();
    S->f_head.swap(head); S->f_next.swap(next); S->f_s1.swap(s1); S->f_s2.swap(s2);
  }
  S->face_off = S->n_edges * S->ned;
  S->int_off = S->face_off + S->n_faces * S->nfd;
  S->ndof = S->int_off + n_elem * S->nint;
  if (S->ndof > (int64_t)0x3fffffff) { hcurl::set_error("hcurl_host_ndspace_create: too many dofs for the 30-bit map encoding"); delete S; return nullptr; }
  if (hex) quad_orientation_tables(p, S->perm, S->sign);
  return S;
}

hcurl_host_ndspace *hcurl_host_ndspace_create_hex(int order, int64_t n_elem, const int32_t *elem_verts) {
  return ndspace_create(true, order, n_elem, elem_verts);
}
/* tetrahedra (orders 1..3): face dofs are listed in the frame of the face's base vertex order; an element that sees
 * the face differently applies the triangle-face DofTransformation, for which hcurl_space_create() takes the base
 * vertices of every element face (hcurl_host_ndspace_elem_face_base) */
hcurl_host_ndspace *hcurl_host_ndspace_create_tet(int order, int64_t n_elem, const int32_t *elem_verts) {
  return ndspace_create(false, order, n_elem, elem_verts);
}
/* [n_elem][faces per element][vertices per face]: base vertex order of every element face */
int hcurl_host_ndspace_elem_face_base(const hcurl_host_ndspace *S, int32_t *out) {
  for (int64_t e = 0; e < S->n_elem; e++)
    for (int lf = 0; lf < S->nlf; lf++)
      for (int k = 0; k < S->nfv; k++)
        *out++ = S->face_base[(size_t)S->el_faces[(size_t)e * S->nlf + lf] * S->nfv + k];
  return 0;
}

void hcurl_host_ndspace_destroy(hcurl_host_ndspace *S) { delete S; }
int hcurl_host_ndspace_nd(const hcurl_host_ndspace *S) { return S->nd; }
int64_t hcurl_host_ndspace_ndof(const hcurl_host_ndspace *S) { return S->ndof; }
int64_t hcurl_host_ndspace_n_edges(const hcurl_host_ndspace *S) { return S->n_edges; }
int64_t hcurl_host_ndspace_n_faces(const hcurl_host_ndspace *S) { return S->n_faces; }

/* signed element dofs [n_elem][nd] in MFEM's local order (edges, faces, interior): k, or -1-k when flipped */
int hcurl_host_ndspace_elem_dofs(const hcurl_host_ndspace *S, int32_t *out) {
  const int ned = S->ned, nfd = S->nfd, nint = S->nint, nd = S->nd, nle = S->nle, nlf = S->nlf;
  for (int64_t e = 0; e < S->n_elem; e++) {
    int32_t *o = out + e * nd;
    for (int le = 0; le < nle; le++) {
      const int64_t g = (int64_t)S->el_edges[(size_t)e * nle + le] * ned;
      const bool fwd = S->edge_fwd[(size_t)e * nle + le];
      for (int m = 0; m < ned; m++) {
        const int32_t k = (int32_t)(g + (fwd ? m : ned - 1 - m));
        *o++ = fwd ? k : -1 - k;
      }
    }
    for (int lf = 0; lf < nlf; lf++) {
      const int64_t g = S->face_off + (int64_t)S->el_faces[(size_t)e * nlf + lf] * nfd;
      if (S->hex) {
        const int code = S->face_code[(size_t)e * nlf + lf];
        for (int m = 0; m < nfd; m++) {
          const int32_t k = (int32_t)(g + S->perm[(size_t)code * nfd + m]);
          *o++ = S->sign[(size_t)code * nfd + m] > 0 ? k : -1 - k;
        }
      } else {
        for (int m = 0; m < nfd; m++) *o++ = (int32_t)(g + m);  // orientation handled by the face DofTransformation
      }
    }
    for (int m = 0; m < nint; m++) *o++ = (int32_t)(S->int_off + e * nint + m);
  }
  return 0;
}

/* dofs on boundary faces (faces seen by one element) and on their edges, ascending; returns their number
 * (out may be NULL to query it; capacity needed: ndof entries at most) */
int64_t hcurl_host_ndspace_boundary_dofs(const hcurl_host_ndspace *S, int32_t *out) {
  std::vector<uint8_t> mark((size_t)S->ndof, 0);
  const int nle = S->nle, nlf = S->nlf, nfv = S->nfv;
  int f2e[6][4], nf2e = 0;
  for (int lf = 0; lf < nlf; lf++) {
    int n = 0;
    for (int le = 0; le < nle; le++) {
      bool a = false, b = false;
      for (int k = 0; k < nfv; k++) {
        const int fv = S->hex ? HEX_FACES[lf][k] : TET_FACES[lf][k];
        a |= fv == (S->hex ? HEX_EDGES[le][0] : TET_EDGES[le][0]);
        b |= fv == (S->hex ? HEX_EDGES[le][1] : TET_EDGES[le][1]);
      }
      if (a && b) f2e[lf][n++] = le;
    }
    nf2e = n;
  }
  for (int64_t e = 0; e < S->n_elem; e++)
    for (int lf = 0; lf < nlf; lf++) {
      const int32_t f = S->el_faces[(size_t)e * nlf + lf];
      if (S->face_count[(size_t)f] != 1) continue;
      for (int m = 0; m < S->nfd; m++) mark[(size_t)(S->face_off + (int64_t)f * S->nfd + m)] = 1;
      for (int k = 0; k < nf2e; k++) {
        const int64_t g = (int64_t)S->el_edges[(size_t)e * nle + f2e[lf][k]] * S->ned;
        for (int m = 0; m < S->ned; m++) mark[(size_t)(g + m)] = 1;
      }
    }
  int64_t n = 0;
  for (int64_t d = 0; d < S->ndof; d++)
    if (mark[(size_t)d]) { if (out) out[n] = (int32_t)d; n++; }
  return n;
}

/* Dofs on the given faces (n_faces vertex lists of 4 / 3 corner ids each, any order -- e.g. the side-set faces of
 * hcurl_host_exodus_read) and on their edges, ascending and unique: what BCMap::GetEssentialBdrMarkers ->
 * GetEssentialTrueDofs selects for a set of boundary attributes
 * (/root/reference/src/boundary_conditions/boundary_conditions.cpp:32-51).  Returns the count (out may be NULL),
 * -1 if a face or edge is not part of the mesh. */
int64_t hcurl_host_ndspace_face_dofs(const hcurl_host_ndspace *S, int64_t n_faces, const int32_t *face_verts, int32_t *out) {
  std::vector<uint8_t> mark((size_t)S->ndof, 0);
  const int nfv = S->nfv;
  const int64_t nv = (int64_t)S->e_head.size();
  for (int64_t q = 0; q < n_faces; q++) {
    const int32_t *v = face_verts + q * nfv;
    for (int k = 0; k < nfv; k++)
      if (v[k] < 0 || v[k] >= nv) { hcurl::set_error("hcurl_host_ndspace_face_dofs: vertex id out of range"); return -1; }
    for (int k = 0; k < nfv; k++) {  // edges of the polygon
      const int32_t a = v[k], b = v[(k + 1) % nfv], lo = std::min(a, b), hi = std::max(a, b);
      int32_t id = S->e_head[(size_t)lo];
      while (id >= 0 && S->e_hi[(size_t)id] != hi) id = S->e_next[(size_t)id];
      if (id < 0) { hcurl::set_error("hcurl_host_ndspace_face_dofs: face %lld has an edge that is not in the mesh", (long long)q); return -1; }
      for (int m = 0; m < S->ned; m++) mark[(size_t)((int64_t)id * S->ned + m)] = 1;
    }
    if (S->nfd) {
      int32_t s[4];
      for (int k = 0; k < nfv; k++) s[k] = v[k];
      std::sort(s, s + nfv);
      int32_t f = S->f_head[(size_t)s[0]];
      while (f >= 0 && !(S->f_s1[(size_t)f] == s[1] && S->f_s2[(size_t)f] == s[2])) f = S->f_next[(size_t)f];
      if (f < 0) { hcurl::set_error("hcurl_host_ndspace_face_dofs: face %lld is not in the mesh", (long long)q); return -1; }
      for (int m = 0; m < S->nfd; m++) mark[(size_t)(S->face_off + (int64_t)f * S->nfd + m)] = 1;
    }
  }
  int64_t n = 0;
  for (int64_t d = 0; d < S->ndof; d++)
    if (mark[(size_t)d]) { if (out) out[n] = (int32_t)d; n++; }
  return n;
}

}  // extern "C"
