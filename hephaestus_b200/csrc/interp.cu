// Discrete gradient / curl interpolators on the device: C entry points and the kernel around
// interp_body.h.  HBM-bound by construction: reads the two maps (4 B per entry) and x through L1/L2, writes every output
// dof once; D (at most 4 x 5 doubles) travels in the kernel arguments.
//
// STATUS: the per-thread program is verified on the host against analytic gradients / curls of polynomials and the
// curl(grad) = 0 invariant (tests/test_interp.py); the GPU parity tests are opt-in (HCURL_TEST_INTERP=1) until the kernel
// has been run -- it was written after the round's GPU budget was spent.
#include <cstring>
#include <vector>

#include "internal.h"
#include "interp_body.h"

struct hcurl_interp {
  hcurl_ctx *ctx = nullptr;
  hcurl::InterpArgs a;
  int32_t n_in = 0, n_out = 0;
  uint32_t *d_in = nullptr, *d_out = nullptr;
  double *d_x = nullptr, *d_y = nullptr;  // staging for host vectors
};

namespace hcurl {

__global__ void __launch_bounds__(256) interp_kernel(const InterpArgs a) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid < a.n_elem * a.nd_out) interp_thread(a, gid);
}

// D[i][j] = derivative of the closed (Gauss-Lobatto, p+1 points) Lagrange basis function j at the open
// (Gauss-Legendre, p points) node i
void interp_matrix(int p, double *D) {
  double xo[MAXP1], wo[MAXP1], xc[MAXP1 + 1];
  gauss_legendre01(p, xo, wo);
  gauss_lobatto01(p + 1, xc);
  std::vector<double> B((size_t)p * (p + 1)), G((size_t)p * (p + 1));
  lagrange_tables(p + 1, xc, p, xo, B.data(), G.data());
  for (int i = 0; i < p; i++)
    for (int j = 0; j <= p; j++) D[i * (p + 1) + j] = G[(size_t)i * (p + 1) + j];
}

// tets, order 1: local matrices in MFEM's edge order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3) and face order (1,2,3),(0,3,2),(0,1,3),(0,2,1)
void interp_tet_matrix(int kind, double *M) {
  static const int E[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};
  static const double GL[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};      // gradients of the barycentric coordinates
  static const double NK[4][3] = {{1, 1, 1}, {-1, 0, 0}, {0, -1, 0}, {0, 0, -1}};       // RT_TetrahedronElement::nk
  for (int i = 0; i < 24; i++) M[i] = 0.0;
  if (kind == HCURL_INTERP_GRAD) {
    for (int e = 0; e < 6; e++) { M[e * 4 + E[e][0]] = -1.0; M[e * 4 + E[e][1]] = 1.0; }  // phi(b) - phi(a)
  } else {
    for (int f = 0; f < 4; f++)
      for (int e = 0; e < 6; e++) {
        const double *a = GL[E[e][0]], *b = GL[E[e][1]];  // curl of the Whitney function: 2 grad(l_a) x grad(l_b)
        const double c[3] = {2 * (a[1] * b[2] - a[2] * b[1]), 2 * (a[2] * b[0] - a[0] * b[2]), 2 * (a[0] * b[1] - a[1] * b[0])};
        M[f * 6 + e] = NK[f][0] * c[0] + NK[f][1] * c[1] + NK[f][2] * c[2];
      }
  }
}

int interp_sizes(int kind, int p, int *nd_in, int *nd_out, int geom = HCURL_GEOM_HEX) {
  if (geom == HCURL_GEOM_TET) {
    if (p != 1) { set_error("tet interpolators: order 1 only (got %d)", p); return 1; }
    if (kind == HCURL_INTERP_GRAD) { *nd_in = 4; *nd_out = 6; }
    else if (kind == HCURL_INTERP_CURL) { *nd_in = 6; *nd_out = 4; }
    else { set_error("unknown interpolator kind %d", kind); return 1; }
    return 0;
  }
  if (geom != HCURL_GEOM_HEX) { set_error("unknown geometry %d", geom); return 1; }
  if (p < 1 || p > 4) { set_error("interpolator order must be 1..4 (got %d)", p); return 1; }
  const int nd_h1 = (p + 1) * (p + 1) * (p + 1), nd_nd = 3 * p * (p + 1) * (p + 1), nd_rt = 3 * p * p * (p + 1);
  if (kind == HCURL_INTERP_GRAD) { *nd_in = nd_h1; *nd_out = nd_nd; }
  else if (kind == HCURL_INTERP_CURL) { *nd_in = nd_nd; *nd_out = nd_rt; }
  else { set_error("unknown interpolator kind %d", kind); return 1; }
  return 0;
}

// signed (-1-k) tensor-ordered element dofs -> encoded maps; an output dof is written by the first element that holds it
int interp_encode(int64_t n_elem, int nd_in, int nd_out, const int32_t *in_dofs, const int32_t *out_dofs, int32_t n_in, int32_t n_out,
                  uint32_t *in_map, uint32_t *out_map) {
  for (int64_t i = 0; i < n_elem * nd_in; i++) {
    const int32_t v = in_dofs[i], k = v >= 0 ? v : -1 - v;
    if (k >= n_in) { set_error("interpolator: input dof %d out of range", k); return 1; }
    in_map[i] = (uint32_t)k | (v < 0 ? INTERP_NEG : 0u);
  }
  std::vector<uint8_t> seen((size_t)n_out, 0);
  for (int64_t i = 0; i < n_elem * nd_out; i++) {
    const int32_t v = out_dofs[i], k = v >= 0 ? v : -1 - v;
    if (k >= n_out) { set_error("interpolator: output dof %d out of range", k); return 1; }
    out_map[i] = (uint32_t)k | (v < 0 ? INTERP_NEG : 0u) | (seen[(size_t)k] ? INTERP_SKIP : 0u);
    seen[(size_t)k] = 1;
  }
  return 0;
}

}  // namespace hcurl

using namespace hcurl;

extern "C" {

hcurl_interp *hcurl_interp_create(hcurl_ctx *ctx, int kind, int order, int64_t n_elem, const int32_t *in_dofs, const int32_t *out_dofs,
                                  int32_t n_in, int32_t n_out) {
  return hcurl_interp_create_geom(ctx, HCURL_GEOM_HEX, kind, order, n_elem, in_dofs, out_dofs, n_in, n_out);
}

hcurl_interp *hcurl_interp_create_geom(hcurl_ctx *ctx, int geom, int kind, int order, int64_t n_elem, const int32_t *in_dofs,
                                       const int32_t *out_dofs, int32_t n_in, int32_t n_out) {
  if (!ctx || n_elem <= 0 || !in_dofs || !out_dofs || n_in <= 0 || n_out <= 0) { set_error("hcurl_interp_create: bad arguments"); return nullptr; }
  if ((int64_t)n_in > (int64_t)INTERP_IDX || (int64_t)n_out > (int64_t)INTERP_IDX) { set_error("too many dofs for the 30-bit map encoding"); return nullptr; }
  int nd_in = 0, nd_out = 0;
  if (interp_sizes(kind, order, &nd_in, &nd_out, geom)) return nullptr;
  std::vector<uint32_t> im((size_t)n_elem * nd_in), omap((size_t)n_elem * nd_out);
  if (interp_encode(n_elem, nd_in, nd_out, in_dofs, out_dofs, n_in, n_out, im.data(), omap.data())) return nullptr;
  HC_CUDA_NULL(cudaSetDevice(ctx->device));
  hcurl_interp *it = new hcurl_interp;
  it->ctx = ctx; it->n_in = n_in; it->n_out = n_out;
  memset(&it->a, 0, sizeof(it->a));
  it->a.kind = kind; it->a.geom = geom == HCURL_GEOM_TET ? 1 : 0; it->a.p = order; it->a.n_elem = n_elem; it->a.nd_in = nd_in; it->a.nd_out = nd_out;
  if (geom == HCURL_GEOM_TET) interp_tet_matrix(kind, it->a.Mloc);
  else interp_matrix(order, it->a.D);
  if (!cuda_ok(cudaMalloc(&it->d_in, sizeof(uint32_t) * im.size()), "cudaMalloc") ||
      !cuda_ok(cudaMalloc(&it->d_out, sizeof(uint32_t) * omap.size()), "cudaMalloc") ||
      !cuda_ok(h2d_sync(ctx, it->d_in, im.data(), sizeof(uint32_t) * im.size()), "cudaMemcpy") ||
      !cuda_ok(h2d_sync(ctx, it->d_out, omap.data(), sizeof(uint32_t) * omap.size()), "cudaMemcpy")) {
    cudaFree(it->d_in); cudaFree(it->d_out);
    delete it;
    return nullptr;
  }
  it->a.in_map = it->d_in; it->a.out_map = it->d_out;
  return it;
}

void hcurl_interp_destroy(hcurl_interp *it) {
  if (!it) return;
  cudaSetDevice(it->ctx->device);
  cudaStreamSynchronize(it->ctx->stream);
  cudaFree(it->d_in); cudaFree(it->d_out); cudaFree(it->d_x); cudaFree(it->d_y);
  delete it;
}

int hcurl_interp_apply(hcurl_interp *it, const double *x, double *y, int where) {
  if (!it || !x || !y) { set_error("hcurl_interp_apply: null argument"); return 1; }
  hcurl_ctx *ctx = it->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  InterpArgs a = it->a;
  if (where == HCURL_HOST) {
    if (!it->d_x) HC_CUDA(cudaMalloc(&it->d_x, sizeof(double) * it->n_in));
    if (!it->d_y) HC_CUDA(cudaMalloc(&it->d_y, sizeof(double) * it->n_out));
    HC_CUDA(cudaMemcpyAsync(it->d_x, x, sizeof(double) * it->n_in, cudaMemcpyHostToDevice, ctx->stream));
    a.x = it->d_x; a.y = it->d_y;
  } else { a.x = x; a.y = y; }
  // every output dof that belongs to an element is written; the others (none in a conforming space) are cleared
  HC_CUDA(cudaMemsetAsync(a.y, 0, sizeof(double) * it->n_out, ctx->stream));
  const int64_t n = a.n_elem * a.nd_out;
  interp_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(a);
  HC_CUDA(cudaGetLastError());
  ctx->launches++;
  if (where == HCURL_HOST) {
    HC_CUDA(cudaMemcpyAsync(y, it->d_y, sizeof(double) * it->n_out, cudaMemcpyDeviceToHost, ctx->stream));
    HC_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

/* host-only helpers (no GPU): the 1-D matrix and the encoded maps, for tests */
int hcurl_host_interp_matrix(int order, double *D) {
  if (order < 1 || order > 4 || !D) { set_error("hcurl_host_interp_matrix: bad arguments"); return 1; }
  interp_matrix(order, D);
  return 0;
}
int hcurl_host_interp_tet_matrix(int kind, double *M24) {
  if ((kind != HCURL_INTERP_GRAD && kind != HCURL_INTERP_CURL) || !M24) { set_error("hcurl_host_interp_tet_matrix: bad arguments"); return 1; }
  interp_tet_matrix(kind, M24);
  return 0;
}
int hcurl_host_interp_encode(int kind, int order, int64_t n_elem, const int32_t *in_dofs, const int32_t *out_dofs, int32_t n_in, int32_t n_out,
                             uint32_t *in_map, uint32_t *out_map) {
  int nd_in = 0, nd_out = 0;
  if (interp_sizes(kind, order, &nd_in, &nd_out)) return 1;
  if (n_elem <= 0 || !in_dofs || !out_dofs || !in_map || !out_map) { set_error("hcurl_host_interp_encode: bad arguments"); return 1; }
  return interp_encode(n_elem, nd_in, nd_out, in_dofs, out_dofs, n_in, n_out, in_map, out_map);
}
}
