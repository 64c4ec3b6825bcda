// Host-side reference-element tables for the Nedelec spaces the hot path uses.
//
// Product code (independent of oracle/): 1D Gauss / Gauss-Lobatto nodes, Lagrange
// tables, the ND hexahedron local ordering, and the ND tetrahedron reference
// matrices.  Conventions follow MFEM's ND_HexahedronElement /
// ND_TetrahedronElement as selected by "ND_3D_P<p>"
// (/root/reference/src/problem_builders/problem_builder_base.cpp:110-111).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

#include "internal.h"

namespace hcurl {

static void legendre(int n, double x, double &P, double &dP) {
  // P_n(x) and derivative on [-1,1]
  double p0 = 1.0, p1 = x;
  if (n == 0) { P = 1.0; dP = 0.0; return; }
  for (int k = 2; k <= n; k++) {
    double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1; p1 = pk;
  }
  P = p1;
  dP = n * (x * p1 - p0) / (x * x - 1.0);
}

void gauss_legendre01(int n, double *x, double *w) {
  for (int i = 0; i < (n + 1) / 2; i++) {
    double z = cos(M_PI * (i + 0.75) / (n + 0.5));
    double P, dP;
    for (int it = 0; it < 100; it++) {
      legendre(n, z, P, dP);
      double dz = P / dP;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    legendre(n, z, P, dP);
    double wi = 2.0 / ((1.0 - z * z) * dP * dP);
    x[i] = 0.5 * (1.0 - z); x[n - 1 - i] = 0.5 * (1.0 + z);
    w[i] = w[n - 1 - i] = 0.5 * wi;
  }
  if (n % 2) x[n / 2] = 0.5;
}

void gauss_lobatto01(int n, double *x) {
  // endpoints + roots of P'_{n-1}
  x[0] = 0.0; x[n - 1] = 1.0;
  int m = n - 1;
  for (int i = 1; i < (n + 1) / 2; i++) {
    double z = cos(M_PI * i / m);  // initial guess (Chebyshev-Lobatto)
    for (int it = 0; it < 100; it++) {
      double P, dP;
      legendre(m, z, P, dP);
      // d2P from Legendre ODE: (1-z^2) P'' = 2 z P' - m(m+1) P
      double d2P = (2.0 * z * dP - m * (m + 1) * P) / (1.0 - z * z);
      double dz = dP / d2P;
      z -= dz;
      if (fabs(dz) < 1e-16) break;
    }
    x[i] = 0.5 * (1.0 - z); x[n - 1 - i] = 0.5 * (1.0 + z);
  }
  if (n % 2) x[n / 2] = 0.5;
}

void lagrange_tables(int n, const double *nodes, int nq, const double *xq, double *B, double *G) {
  for (int q = 0; q < nq; q++) {
    for (int i = 0; i < n; i++) {
      double denom = 1.0, val = 1.0, der = 0.0;
      for (int m = 0; m < n; m++) {
        if (m == i) continue;
        denom *= nodes[i] - nodes[m];
        val *= xq[q] - nodes[m];
        double term = 1.0;
        for (int l = 0; l < n; l++)
          if (l != i && l != m) term *= xq[q] - nodes[l];
        der += term;
      }
      B[q * n + i] = val / denom;
      if (G) G[q * n + i] = der / denom;
    }
  }
}

int nd_hex_ndof(int p) { return 3 * p * (p + 1) * (p + 1); }

static const int HV[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int HE[12][2] = {{0, 1}, {1, 2}, {3, 2}, {0, 3}, {4, 5}, {5, 6}, {7, 6}, {4, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int HF[6][4] = {{3, 2, 1, 0}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};

static int tensor_index(int p, int c, int i, int j, int k) {
  int n0 = (c == 0) ? p : p + 1, n1 = (c == 1) ? p : p + 1;
  return c * p * (p + 1) * (p + 1) + i + n0 * (j + n1 * k);
}

static void axis_dir(int va, int vb, int &ax, int &d) {
  for (int a = 0; a < 3; a++)
    if (HV[vb][a] != HV[va][a]) { ax = a; d = HV[vb][a] - HV[va][a]; }
}

void nd_hex_dof_map(int p, std::vector<int> &loc, std::vector<int> &sgn) {
  int nd = nd_hex_ndof(p);
  loc.assign(nd, -1);
  sgn.assign(nd, 1);
  int o = 0;
  // edges: open index runs along the edge va -> vb
  for (int e = 0; e < 12; e++) {
    int ax = 0, d = 1;
    axis_dir(HE[e][0], HE[e][1], ax, d);
    for (int m = 0; m < p; m++) {
      int idx[3] = {HV[HE[e][0]][0] * p, HV[HE[e][0]][1] * p, HV[HE[e][0]][2] * p};
      idx[ax] = d > 0 ? m : p - 1 - m;
      int t = tensor_index(p, ax, idx[0], idx[1], idx[2]);
      loc[t] = o++; sgn[t] = d;
    }
  }
  // faces: frame s = v0->v1, t = v0->v3; s-directed dofs (j=1..p-1 closed in t,
  // i open in s) then t-directed (j open in t, i=1..p-1 closed in s)
  for (int f = 0; f < 6; f++) {
    int as = 0, ds = 1, at = 0, dt = 1;
    axis_dir(HF[f][0], HF[f][1], as, ds);
    axis_dir(HF[f][0], HF[f][3], at, dt);
    int an = 3 - as - at;
    int fixed = HV[HF[f][0]][an] * p;
    for (int j = 1; j < p; j++)
      for (int i = 0; i < p; i++) {
        int idx[3]; idx[an] = fixed;
        idx[as] = ds > 0 ? i : p - 1 - i;
        idx[at] = dt > 0 ? j : p - j;
        int t = tensor_index(p, as, idx[0], idx[1], idx[2]);
        loc[t] = o++; sgn[t] = ds;
      }
    for (int j = 0; j < p; j++)
      for (int i = 1; i < p; i++) {
        int idx[3]; idx[an] = fixed;
        idx[at] = dt > 0 ? j : p - 1 - j;
        idx[as] = ds > 0 ? i : p - i;
        int t = tensor_index(p, at, idx[0], idx[1], idx[2]);
        loc[t] = o++; sgn[t] = dt;
      }
  }
  // interior
  for (int c = 0; c < 3; c++) {
    int lo[3] = {1, 1, 1}, hi[3] = {p, p, p};
    lo[c] = 0;
    for (int k = lo[2]; k < hi[2]; k++)
      for (int j = lo[1]; j < hi[1]; j++)
        for (int i = lo[0]; i < hi[0]; i++) loc[tensor_index(p, c, i, j, k)] = o++;
  }
}

void build_hex_tables(int p, int q, int g, HexTables &t) {
  memset(&t, 0, sizeof(t));
  double xq[MAXQ], wq[MAXQ], cp[MAXP1 + 1], op[MAXP1], ow[MAXP1];
  gauss_legendre01(q, xq, wq);
  gauss_lobatto01(p + 1, cp);
  gauss_legendre01(p, op, ow);
  double B[MAXQ * MAXQ], G[MAXQ * MAXQ];
  lagrange_tables(p, op, q, xq, B, nullptr);
  for (int a = 0; a < q; a++)
    for (int i = 0; i < p; i++) t.Bo[a * MAXP1 + i] = B[a * p + i];
  lagrange_tables(p + 1, cp, q, xq, B, G);
  for (int a = 0; a < q; a++)
    for (int i = 0; i <= p; i++) { t.Bc[a * MAXP1 + i] = B[a * (p + 1) + i]; t.Gc[a * MAXP1 + i] = G[a * (p + 1) + i]; }
  lagrange_tables(q, xq, q, xq, B, G);
  for (int a = 0; a < q; a++)
    for (int m = 0; m < q; m++) t.Dq[a * MAXQ + m] = G[a * q + m];
  double gn[3] = {0.0, 1.0, 0.0};
  if (g == 2) { gn[1] = 0.5; gn[2] = 1.0; }
  lagrange_tables(g + 1, gn, q, xq, B, G);
  for (int a = 0; a < q; a++)
    for (int m = 0; m <= g; m++) { t.Ng[a * 3 + m] = B[a * (g + 1) + m]; t.dNg[a * 3 + m] = G[a * (g + 1) + m]; }
  for (int a = 0; a < q; a++) t.wq[a] = wq[a];
}

// ------------------------------------------------------------------ tets ----
int nd_tet_ndof(int p) { return p * (p + 2) * (p + 3) / 2; }

namespace {
constexpr int MD = 5;  // max degree + 1 of scalar polys handled (p <= 3 -> products up to 6 via integral only)
struct Poly {          // dense coefficients over x^a y^b z^c, a,b,c < MD
  double c[MD][MD][MD];
  Poly() { memset(c, 0, sizeof(c)); }
  double eval(const double *x) const {
    double s = 0;
    for (int a = 0; a < MD; a++)
      for (int b = 0; b < MD; b++)
        for (int d = 0; d < MD; d++)
          if (c[a][b][d] != 0.0) s += c[a][b][d] * pow(x[0], a) * pow(x[1], b) * pow(x[2], d);
    return s;
  }
  Poly deriv(int axis) const {
    Poly r;
    for (int a = 0; a < MD; a++)
      for (int b = 0; b < MD; b++)
        for (int d = 0; d < MD; d++) {
          if (c[a][b][d] == 0.0) continue;
          int e[3] = {a, b, d};
          if (e[axis] == 0) continue;
          double k = e[axis];
          e[axis]--;
          r.c[e[0]][e[1]][e[2]] += k * c[a][b][d];
        }
    return r;
  }
  void axpy(double s, const Poly &o) {
    for (int a = 0; a < MD; a++)
      for (int b = 0; b < MD; b++)
        for (int d = 0; d < MD; d++) c[a][b][d] += s * o.c[a][b][d];
  }
};
double fact(int n) { double f = 1; for (int i = 2; i <= n; i++) f *= i; return f; }
double tet_monomial_integral(int a, int b, int c) { return fact(a) * fact(b) * fact(c) / fact(a + b + c + 3); }
double integrate_product(const Poly &f, const Poly &g) {
  double s = 0;
  for (int a = 0; a < MD; a++)
    for (int b = 0; b < MD; b++)
      for (int d = 0; d < MD; d++) {
        if (f.c[a][b][d] == 0.0) continue;
        for (int a2 = 0; a2 < MD; a2++)
          for (int b2 = 0; b2 < MD; b2++)
            for (int d2 = 0; d2 < MD; d2++)
              if (g.c[a2][b2][d2] != 0.0)
                s += f.c[a][b][d] * g.c[a2][b2][d2] * tet_monomial_integral(a + a2, b + b2, d + d2);
      }
  return s;
}
struct VecPoly { Poly v[3]; };
const double TV[4][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
const int TE[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
const int TF[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
}  // namespace

// shape functions and curls of ND_p on the reference tet as polynomials (MFEM's nodal basis: see below)
static void tet_shape_polys(int p, std::vector<VecPoly> &shp, std::vector<VecPoly> &crl) {
  const int nd = nd_tet_ndof(p);
  // raw basis of ND_p: P_{p-1}^3 (+) m*(y,-x,0), m*(z,0,-x), [a==0] m*(0,z,-y), m hom. deg p-1
  std::vector<VecPoly> raw;
  for (int deg = 0; deg < p; deg++)
    for (int a = 0; a <= deg; a++)
      for (int b = 0; b <= deg - a; b++) {
        int c = deg - a - b;
        for (int d = 0; d < 3; d++) { VecPoly f; f.v[d].c[a][b][c] = 1.0; raw.push_back(f); }
      }
  for (int a = 0; a < p; a++)
    for (int b = 0; b < p - a; b++) {
      int c = p - 1 - a - b;
      { VecPoly f; f.v[0].c[a][b + 1][c] = 1.0; f.v[1].c[a + 1][b][c] = -1.0; raw.push_back(f); }
      { VecPoly f; f.v[0].c[a][b][c + 1] = 1.0; f.v[2].c[a + 1][b][c] = -1.0; raw.push_back(f); }
      if (a == 0) { VecPoly f; f.v[1].c[a][b][c + 1] = 1.0; f.v[2].c[a][b + 1][c] = -1.0; raw.push_back(f); }
    }
  // dof functionals
  std::vector<double> nodes, tang;
  double eop[8], fop[8], iop[8], w8[8];
  gauss_legendre01(p, eop, w8);
  if (p > 1) gauss_legendre01(p - 1, fop, w8);
  if (p > 2) gauss_legendre01(p - 2, iop, w8);
  auto push = [&](const double *x, const double *t) { for (int d = 0; d < 3; d++) { nodes.push_back(x[d]); tang.push_back(t[d]); } };
  for (int e = 0; e < 6; e++)
    for (int i = 0; i < p; i++) {
      double x[3], t[3];
      for (int d = 0; d < 3; d++) { t[d] = TV[TE[e][1]][d] - TV[TE[e][0]][d]; x[d] = TV[TE[e][0]][d] + eop[i] * t[d]; }
      push(x, t);
    }
  for (int f = 0; f < 4; f++)
    for (int j = 0; j <= p - 2; j++)
      for (int i = 0; i + j <= p - 2; i++) {
        double w[3] = {fop[p - 2 - i - j], fop[i], fop[j]};
        double ws = w[0] + w[1] + w[2];
        double x[3], t1[3], t2[3];
        for (int d = 0; d < 3; d++) {
          x[d] = (w[0] * TV[TF[f][0]][d] + w[1] * TV[TF[f][1]][d] + w[2] * TV[TF[f][2]][d]) / ws;
          t1[d] = TV[TF[f][1]][d] - TV[TF[f][0]][d];
          t2[d] = TV[TF[f][2]][d] - TV[TF[f][0]][d];
        }
        push(x, t1); push(x, t2);
      }
  for (int k = 0; k <= p - 3; k++)
    for (int j = 0; j + k <= p - 3; j++)
      for (int i = 0; i + j + k <= p - 3; i++) {
        double w = iop[i] + iop[j] + iop[k] + iop[p - 3 - i - j - k];
        double x[3] = {iop[i] / w, iop[j] / w, iop[k] / w};
        for (int d = 0; d < 3; d++) { double t[3] = {0, 0, 0}; t[d] = 1.0; push(x, t); }
      }
  // Vandermonde V[m][n] = dof_m(raw_n); invert by Gauss-Jordan with partial pivoting
  std::vector<double> V(nd * nd), Vi(nd * nd, 0.0);
  for (int m = 0; m < nd; m++)
    for (int n = 0; n < nd; n++) {
      double s = 0;
      for (int d = 0; d < 3; d++) s += tang[3 * m + d] * raw[n].v[d].eval(&nodes[3 * m]);
      V[m * nd + n] = s;
    }
  for (int i = 0; i < nd; i++) Vi[i * nd + i] = 1.0;
  for (int col = 0; col < nd; col++) {
    int piv = col;
    for (int rr = col + 1; rr < nd; rr++) if (fabs(V[rr * nd + col]) > fabs(V[piv * nd + col])) piv = rr;
    if (piv != col)
      for (int k = 0; k < nd; k++) { std::swap(V[piv * nd + k], V[col * nd + k]); std::swap(Vi[piv * nd + k], Vi[col * nd + k]); }
    double d = 1.0 / V[col * nd + col];
    for (int k = 0; k < nd; k++) { V[col * nd + k] *= d; Vi[col * nd + k] *= d; }
    for (int rr = 0; rr < nd; rr++) {
      if (rr == col) continue;
      double f = V[rr * nd + col];
      if (f == 0.0) continue;
      for (int k = 0; k < nd; k++) { V[rr * nd + k] -= f * V[col * nd + k]; Vi[rr * nd + k] -= f * Vi[col * nd + k]; }
    }
  }
  // shape_m = sum_n Vi[n][m] raw_n
  shp.assign(nd, VecPoly());
  crl.assign(nd, VecPoly());
  for (int m = 0; m < nd; m++) {
    for (int n = 0; n < nd; n++)
      for (int d = 0; d < 3; d++) shp[m].v[d].axpy(Vi[n * nd + m], raw[n].v[d]);
    for (int d = 0; d < 3; d++) {
      int d1 = (d + 1) % 3, d2 = (d + 2) % 3;
      // curl_d = d(v_{d2})/d x_{d1} - d(v_{d1})/d x_{d2}
      crl[m].v[d] = shp[m].v[d2].deriv(d1);
      crl[m].v[d].axpy(-1.0, shp[m].v[d1].deriv(d2));
    }
  }
}

void build_tet_ref(int p, TetRef &r) {
  const int nd = nd_tet_ndof(p);
  r.p = p; r.nd = nd; r.npk = nd * (nd + 1) / 2;
  std::vector<VecPoly> shp, crl;
  tet_shape_polys(p, shp, crl);
  r.Mref.assign(6 * r.npk, 0.0);
  r.Kref.assign(6 * r.npk, 0.0);
  int pr = 0;
  for (int a = 0; a < 3; a++)
    for (int b = a; b < 3; b++, pr++) {
      int k = 0;
      for (int i = 0; i < nd; i++)
        for (int j = i; j < nd; j++, k++) {
          double m = integrate_product(shp[i].v[a], shp[j].v[b]);
          double c = integrate_product(crl[i].v[a], crl[j].v[b]);
          if (a != b) { m += integrate_product(shp[i].v[b], shp[j].v[a]); c += integrate_product(crl[i].v[b], crl[j].v[a]); }
          r.Mref[pr * r.npk + k] = m;
          r.Kref[pr * r.npk + k] = c;
        }
    }
}

// Quadrature-point reference matrices for coefficients that vary inside an element: for every point q of the caller's
// rule (pts [nq][3], w [nq]) and every pair (a,b) as in TetRef, w_q * (f_i^a f_j^b [+ f_i^b f_j^a]) packed upper; f =
// shape functions (curl = false) or their curls (curl = true).  out [nq][6][npk].
void build_tet_ref_q(int p, bool curl, int nq, const double *pts, const double *w, std::vector<double> &out) {
  const int nd = nd_tet_ndof(p), npk = nd * (nd + 1) / 2;
  std::vector<VecPoly> shp, crl;
  tet_shape_polys(p, shp, crl);
  const std::vector<VecPoly> &F = curl ? crl : shp;
  out.assign((size_t)nq * 6 * npk, 0.0);
  std::vector<double> val((size_t)nd * 3);
  for (int q = 0; q < nq; q++) {
    for (int i = 0; i < nd; i++)
      for (int d = 0; d < 3; d++) val[(size_t)i * 3 + d] = F[i].v[d].eval(pts + 3 * q);
    int pr = 0;
    for (int a = 0; a < 3; a++)
      for (int b = a; b < 3; b++, pr++) {
        int k = 0;
        for (int i = 0; i < nd; i++)
          for (int j = i; j < nd; j++, k++) {
            double m = val[(size_t)i * 3 + a] * val[(size_t)j * 3 + b];
            if (a != b) m += val[(size_t)i * 3 + b] * val[(size_t)j * 3 + a];
            out[((size_t)q * 6 + pr) * npk + k] = w[q] * m;
          }
      }
  }
}

void tri_face_transform(int p, const int32_t *lv, const int32_t *bv, double *R) {
  // face nodes (i,j), j outer; weights on (f0,f1,f2) ~ (fop[p-2-i-j], fop[i], fop[j])
  std::vector<std::pair<int, int>> nodes;
  for (int j = 0; j <= p - 2; j++)
    for (int i = 0; i + j <= p - 2; i++) nodes.push_back({i, j});
  int nn = (int)nodes.size(), nf = 2 * nn;
  for (int i = 0; i < nf * nf; i++) R[i] = 0.0;
  int pos[3];
  for (int k = 0; k < 3; k++) {
    pos[k] = -1;
    for (int m = 0; m < 3; m++) if (bv[m] == lv[k]) pos[k] = m;
  }
  const int bt[3][2] = {{0, 0}, {1, 0}, {0, 1}};
  int t1[2] = {bt[pos[1]][0] - bt[pos[0]][0], bt[pos[1]][1] - bt[pos[0]][1]};
  int t2[2] = {bt[pos[2]][0] - bt[pos[0]][0], bt[pos[2]][1] - bt[pos[0]][1]};
  for (int n = 0; n < nn; n++) {
    int wl[3] = {p - 2 - nodes[n].first - nodes[n].second, nodes[n].first, nodes[n].second};
    int wb[3];
    for (int k = 0; k < 3; k++) wb[pos[k]] = wl[k];
    int nb = -1;
    for (int m = 0; m < nn; m++) if (nodes[m].first == wb[1] && nodes[m].second == wb[2]) nb = m;
    R[(2 * n) * nf + 2 * nb] = t1[0];     R[(2 * n) * nf + 2 * nb + 1] = t1[1];
    R[(2 * n + 1) * nf + 2 * nb] = t2[0]; R[(2 * n + 1) * nf + 2 * nb + 1] = t2[1];
  }
}

}  // namespace hcurl

// Nodes and tangents of the Nedelec hex dof functionals in MFEM's local order: dof_l(v) = t_l . v(node_l) on the reference
// element (open points: Gauss-Legendre, closed points: Gauss-Lobatto -- the nodal basis the kernels use).  Host only:
// interpolation of given fields for parity / property checks (hephaestus_b200/fespace.py: interpolate_affine_hex).
extern "C" int hcurl_host_hex_dof_nodes(int p, double *nodes, double *tang) {
  if (p < 1 || p > 4 || !nodes || !tang) { hcurl::set_error("hcurl_host_hex_dof_nodes: order 1..4 and two [nd][3] arrays needed"); return 1; }
  using namespace hcurl;
  std::vector<int> loc, sgn;
  nd_hex_dof_map(p, loc, sgn);
  double op[8], cp[8], w[8];
  gauss_legendre01(p, op, w);
  gauss_lobatto01(p + 1, cp);
  const int nd = nd_hex_ndof(p);
  for (int i = 0; i < nd * 3; i++) { nodes[i] = 0.0; tang[i] = 0.0; }
  for (int c = 0; c < 3; c++) {
    const int n0 = c == 0 ? p : p + 1, n1 = c == 1 ? p : p + 1, n2 = c == 2 ? p : p + 1;
    for (int k = 0; k < n2; k++)
      for (int j = 0; j < n1; j++)
        for (int i = 0; i < n0; i++) {
          const int idx[3] = {i, j, k};
          const int t = tensor_index(p, c, i, j, k), l = loc[t];
          for (int d = 0; d < 3; d++) nodes[l * 3 + d] = d == c ? op[idx[d]] : cp[idx[d]];
          tang[l * 3 + c] = (double)sgn[t];
        }
  }
  return 0;
}
