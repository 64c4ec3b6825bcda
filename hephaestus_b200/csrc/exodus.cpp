// Host-side ExodusII mesh reader (NetCDF classic, CDF-1 / CDF-2 "64-bit offset"), setup time only, no GPU and no
// netCDF library needed.
//
// The reference reads its meshes with mfem::Mesh(mesh_file) (NetCDF-C underneath; e.g.
// /root/reference/examples/team7/Main.cpp:152-155, the data/*.g / *.gen files of SURVEY.md App. C) before
// ProblemBuilder::AddFESpace builds the ND space on them.  This file covers that first step for stand-alone callers
// (SURVEY.md 8 f4): node coordinates, element blocks (TETRA4 / HEX8 / HEX27; block id = element attribute) and side
// sets (-> boundary faces with the side-set id as attribute), with the same conventions as hephaestus_b200/fespace.py's
// read_exodus(), against which tests/test_exodus.py checks it on the reference's own files.
//
// File format (published: "The NetCDF Classic Format Specification"): big-endian; header = magic "CDF" version,
// numrecs, dim list, global attribute list, variable list; every variable entry carries name, dimension ids, attributes,
// type, size and the byte offset of its data (32 bits in version 1, 64 in version 2).  Only fixed-size (non-record)
// variables are needed here.
#include <cctype>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>

#include "internal.h"

namespace {

enum { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6 };
enum { NC_DIMENSION = 10, NC_VARIABLE = 11, NC_ATTRIBUTE = 12 };

struct Var {
  std::vector<int> dims;
  int type = 0;
  uint64_t begin = 0;
  std::map<std::string, std::string> text_attr;
};

struct Reader {
  std::vector<unsigned char> buf;
  size_t pos = 0;
  bool ok = true;
  uint32_t u32() {
    if (pos + 4 > buf.size()) { ok = false; return 0; }
    const unsigned char *p = buf.data() + pos;
    pos += 4;
    return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
  }
  uint64_t u64() { const uint64_t h = u32(); return h << 32 | u32(); }
  std::string name() {
    const uint32_t n = u32();
    if (!ok || pos + n > buf.size()) { ok = false; return ""; }
    std::string s((const char *)buf.data() + pos, n);
    pos += (n + 3u) & ~3u;
    return s;
  }
  static size_t type_size(int t) { return t == NC_BYTE || t == NC_CHAR ? 1 : t == NC_SHORT ? 2 : t == NC_DOUBLE ? 8 : 4; }
  // attribute list; text attributes are returned
  void attrs(std::map<std::string, std::string> *out) {
    const uint32_t tag = u32(), n = u32();
    if (tag == 0 && n == 0) return;
    if (tag != NC_ATTRIBUTE) { ok = false; return; }
    for (uint32_t i = 0; i < n && ok; i++) {
      const std::string nm = name();
      const uint32_t type = u32(), cnt = u32();
      const size_t bytes = (size_t)cnt * type_size((int)type);
      if (pos + bytes > buf.size()) { ok = false; return; }
      if (out && type == NC_CHAR) (*out)[nm] = std::string((const char *)buf.data() + pos, cnt);
      pos += (bytes + 3u) & ~(size_t)3u;
    }
  }
};

const int HEX_SIDES[6][4] = {{0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {0, 4, 7, 3}, {0, 3, 2, 1}, {4, 5, 6, 7}};
const int TET_SIDES[4][3] = {{0, 1, 3}, {1, 2, 3}, {0, 3, 2}, {0, 2, 1}};

}  // namespace

struct hcurl_host_exodus {
  Reader r;
  std::vector<uint64_t> dim_len;
  std::map<std::string, int> dim_id;
  std::map<std::string, Var> vars;
  int geom = HCURL_GEOM_HEX, npe = 0;
  int64_t n_nodes = 0, n_elem = 0, n_side = 0;
  std::vector<int32_t> block_id, ss_id;
  uint64_t count(const Var &v) const {
    uint64_t n = 1;
    for (int d : v.dims) n *= dim_len[(size_t)d];
    return n;
  }
  bool read_i32(const std::string &nm, std::vector<int32_t> &out) const {
    auto it = vars.find(nm);
    if (it == vars.end() || it->second.type != NC_INT) return false;
    const uint64_t n = count(it->second);
    if (it->second.begin + 4 * n > r.buf.size()) return false;
    out.resize((size_t)n);
    const unsigned char *p = r.buf.data() + it->second.begin;
    for (uint64_t i = 0; i < n; i++, p += 4) out[(size_t)i] = (int32_t)((uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]);
    return true;
  }
  bool read_f64(const std::string &nm, std::vector<double> &out) const {
    auto it = vars.find(nm);
    if (it == vars.end() || (it->second.type != NC_DOUBLE && it->second.type != NC_FLOAT)) return false;
    const uint64_t n = count(it->second);
    const size_t ts = it->second.type == NC_DOUBLE ? 8 : 4;
    if (it->second.begin + ts * n > r.buf.size()) return false;
    out.resize((size_t)n);
    const unsigned char *p = r.buf.data() + it->second.begin;
    for (uint64_t i = 0; i < n; i++, p += ts) {
      if (ts == 8) {
        uint64_t v = 0;
        for (int k = 0; k < 8; k++) v = v << 8 | p[k];
        memcpy(&out[(size_t)i], &v, 8);
      } else {
        uint32_t v = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
        float f;
        memcpy(&f, &v, 4);
        out[(size_t)i] = f;
      }
    }
    return true;
  }
};

extern "C" {

hcurl_host_exodus *hcurl_host_exodus_open(const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) { hcurl::set_error("hcurl_host_exodus_open: cannot open %s", path); return nullptr; }
  hcurl_host_exodus *X = new hcurl_host_exodus;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  X->r.buf.resize((size_t)(sz > 0 ? sz : 0));
  const size_t got = fread(X->r.buf.data(), 1, X->r.buf.size(), f);
  fclose(f);
  Reader &r = X->r;
  if (got != r.buf.size() || r.buf.size() < 8 || memcmp(r.buf.data(), "CDF", 3) != 0 || (r.buf[3] != 1 && r.buf[3] != 2)) {
    hcurl::set_error("hcurl_host_exodus_open: %s is not a NetCDF classic (CDF-1/2) file", path);
    delete X;
    return nullptr;
  }
  const bool v2 = r.buf[3] == 2;
  r.pos = 4;
  r.u32();  // numrecs
  {  // dimensions
    const uint32_t tag = r.u32(), n = r.u32();
    if (!(tag == 0 && n == 0)) {
      if (tag != NC_DIMENSION) r.ok = false;
      for (uint32_t i = 0; i < n && r.ok; i++) {
        const std::string nm = r.name();
        X->dim_id[nm] = (int)X->dim_len.size();
        X->dim_len.push_back(r.u32());
      }
    }
  }
  r.attrs(nullptr);  // global attributes
  {  // variables
    const uint32_t tag = r.u32(), n = r.u32();
    if (!(tag == 0 && n == 0)) {
      if (tag != NC_VARIABLE) r.ok = false;
      for (uint32_t i = 0; i < n && r.ok; i++) {
        const std::string nm = r.name();
        Var v;
        const uint32_t nd = r.u32();
        for (uint32_t d = 0; d < nd && r.ok; d++) {
          const uint32_t id = r.u32();
          if (id >= X->dim_len.size()) { r.ok = false; break; }
          v.dims.push_back((int)id);
        }
        r.attrs(&v.text_attr);
        v.type = (int)r.u32();
        r.u32();  // vsize
        v.begin = v2 ? r.u64() : r.u32();
        X->vars[nm] = v;
      }
    }
  }
  if (!r.ok) { hcurl::set_error("hcurl_host_exodus_open: corrupt NetCDF header in %s", path); delete X; return nullptr; }
  // ---- Exodus content: blocks, their type, side sets
  auto dim = [&](const char *nm) -> int64_t { auto it = X->dim_id.find(nm); return it == X->dim_id.end() ? -1 : (int64_t)X->dim_len[(size_t)it->second]; };
  X->n_nodes = dim("num_nodes");
  const int64_t n_blk = dim("num_el_blk");
  if (X->n_nodes <= 0 || n_blk <= 0 || !X->read_i32("eb_prop1", X->block_id) || (int64_t)X->block_id.size() != n_blk) {
    hcurl::set_error("hcurl_host_exodus_open: %s has no nodes / element blocks", path);
    delete X;
    return nullptr;
  }
  for (int64_t b = 0; b < n_blk; b++) {
    const std::string cn = "connect" + std::to_string(b + 1);
    auto it = X->vars.find(cn);
    if (it == X->vars.end() || it->second.dims.size() != 2) { hcurl::set_error("hcurl_host_exodus_open: missing %s", cn.c_str()); delete X; return nullptr; }
    const int64_t ne = (int64_t)X->dim_len[(size_t)it->second.dims[0]];
    const int npe = (int)X->dim_len[(size_t)it->second.dims[1]];
    std::string et = it->second.text_attr.count("elem_type") ? it->second.text_attr["elem_type"] : "";
    for (auto &ch : et) ch = (char)toupper(ch);
    int geom;
    if (et.compare(0, 3, "TET") == 0 && npe == 4) geom = HCURL_GEOM_TET;
    else if (et.compare(0, 3, "HEX") == 0 && (npe == 8 || npe == 27)) geom = HCURL_GEOM_HEX;
    else { hcurl::set_error("hcurl_host_exodus_open: unsupported block %s with %d nodes", et.c_str(), npe); delete X; return nullptr; }
    if (b > 0 && (geom != X->geom || npe != X->npe)) { hcurl::set_error("hcurl_host_exodus_open: mixed element types"); delete X; return nullptr; }
    X->geom = geom; X->npe = npe;
    X->n_elem += ne;
  }
  const int64_t n_ss = dim("num_side_sets");
  if (n_ss > 0 && X->read_i32("ss_prop1", X->ss_id))
    for (int64_t s = 0; s < n_ss; s++) {
      auto it = X->vars.find("elem_ss" + std::to_string(s + 1));
      if (it != X->vars.end()) X->n_side += (int64_t)X->count(it->second);
    }
  return X;
}

void hcurl_host_exodus_close(hcurl_host_exodus *X) { delete X; }
/* geom: HCURL_GEOM_HEX | HCURL_GEOM_TET; nodes_per_elem 4, 8 or 27; n_side boundary faces from all side sets */
int hcurl_host_exodus_counts(const hcurl_host_exodus *X, int64_t *n_nodes, int64_t *n_elem, int *geom, int *nodes_per_elem, int64_t *n_side) {
  *n_nodes = X->n_nodes; *n_elem = X->n_elem; *geom = X->geom; *nodes_per_elem = X->npe; *n_side = X->n_side;
  return 0;
}
/* xyz [n_nodes][3]; conn [n_elem][nodes_per_elem] (0-based, blocks concatenated in file order); attr [n_elem] = block
 * id; bdr [n_side][4|3] corner nodes of every side-set face (side sets concatenated); bdr_attr [n_side] = side-set id.
 * Any output may be NULL. */
int hcurl_host_exodus_read(const hcurl_host_exodus *X, double *xyz, int32_t *conn, int32_t *attr, int32_t *bdr, int32_t *bdr_attr) {
  if (xyz) {
    std::vector<double> c;
    if (X->read_f64("coord", c) && (int64_t)c.size() == 3 * X->n_nodes) {
      for (int64_t i = 0; i < X->n_nodes; i++)
        for (int d = 0; d < 3; d++) xyz[i * 3 + d] = c[(size_t)(d * X->n_nodes + i)];
    } else {
      const char *nm[3] = {"coordx", "coordy", "coordz"};
      for (int d = 0; d < 3; d++) {
        if (!X->read_f64(nm[d], c) || (int64_t)c.size() != X->n_nodes) { hcurl::set_error("hcurl_host_exodus_read: no node coordinates"); return 1; }
        for (int64_t i = 0; i < X->n_nodes; i++) xyz[i * 3 + d] = c[(size_t)i];
      }
    }
  }
  std::vector<int32_t> all_conn;
  if (conn || bdr) {
    all_conn.reserve((size_t)X->n_elem * X->npe);
    int64_t e0 = 0;
    for (size_t b = 0; b < X->block_id.size(); b++) {
      std::vector<int32_t> c;
      if (!X->read_i32("connect" + std::to_string(b + 1), c)) { hcurl::set_error("hcurl_host_exodus_read: cannot read connectivity"); return 1; }
      for (int32_t v : c) all_conn.push_back(v - 1);
      const int64_t ne = (int64_t)c.size() / X->npe;
      if (attr) for (int64_t e = 0; e < ne; e++) attr[e0 + e] = X->block_id[b];
      e0 += ne;
    }
    if (conn) memcpy(conn, all_conn.data(), sizeof(int32_t) * all_conn.size());
  } else if (attr) {
    int64_t e0 = 0;
    for (size_t b = 0; b < X->block_id.size(); b++) {
      const Var &v = X->vars.at("connect" + std::to_string(b + 1));
      const int64_t ne = (int64_t)X->dim_len[(size_t)v.dims[0]];
      for (int64_t e = 0; e < ne; e++) attr[e0 + e] = X->block_id[b];
      e0 += ne;
    }
  }
  if (bdr || bdr_attr) {
    const int nfv = X->geom == HCURL_GEOM_HEX ? 4 : 3;
    int64_t q = 0;
    for (size_t s = 0; s < X->ss_id.size(); s++) {
      std::vector<int32_t> el, sd;
      if (!X->read_i32("elem_ss" + std::to_string(s + 1), el) || !X->read_i32("side_ss" + std::to_string(s + 1), sd) || el.size() != sd.size()) continue;
      for (size_t i = 0; i < el.size(); i++, q++) {
        const int64_t e = el[i] - 1;
        const int side = sd[i] - 1;
        if (e < 0 || e >= X->n_elem || side < 0 || side >= (X->geom == HCURL_GEOM_HEX ? 6 : 4)) { hcurl::set_error("hcurl_host_exodus_read: side set %zu out of range", s + 1); return 1; }
        if (bdr)
          for (int k = 0; k < nfv; k++)
            bdr[q * nfv + k] = all_conn[(size_t)(e * X->npe + (X->geom == HCURL_GEOM_HEX ? HEX_SIDES[side][k] : TET_SIDES[side][k]))];
        if (bdr_attr) bdr_attr[q] = X->ss_id[s];
      }
    }
  }
  return 0;
}

}  // extern "C"
