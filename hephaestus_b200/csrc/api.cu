// C ABI of libhcurl_b200.so: contexts, spaces, operators, PCG (see include/hcurl_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "internal.h"

namespace hcurl {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool cuda_ok(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return false;
}

// ------------------------------------------------------------- small kernels ----
__global__ void encode_map_kernel(const int64_t n_elem, const int nd, const int64_t n_pad, const int transpose,
                                  const int32_t *__restrict__ ldof, const int *__restrict__ loc,
                                  const int *__restrict__ sgn, const int32_t *__restrict__ l2e,
                                  uint32_t *__restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * nd) return;
  const int64_t e = gid / nd;
  const int t = (int)(gid % nd);
  const int l = loc ? loc[t] : t;
  const int32_t v = ldof[e * nd + l];
  const int32_t k = v >= 0 ? v : -1 - v;
  bool neg = v < 0;
  if (sgn && sgn[t] < 0) neg = !neg;
  const uint32_t ext = l2e ? (uint32_t)l2e[k] : (uint32_t)k;
  const uint32_t code = ext | (neg ? MAP_NEG : 0u);
  if (transpose) out[(int64_t)t * n_pad + e] = code;
  else out[gid] = code;
}

__global__ void permute_xyz_kernel(const int64_t n_elem, const double *__restrict__ in, double *__restrict__ out) {
  // MFEM hex corner order -> tensor order (x fastest)
  const int perm[8] = {0, 1, 3, 2, 4, 5, 7, 6};  // tensor node t holds corner perm[t]
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_elem * 24) return;
  const int64_t e = gid / 24;
  const int r = (int)(gid % 24), t = r / 3, d = r % 3;
  out[gid] = in[e * 24 + perm[t] * 3 + d];
}

__global__ void flag_map_kernel(const int64_t n, const uint32_t *__restrict__ in, const double *__restrict__ flag,
                                uint32_t *__restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const uint32_t m = in[gid];
  out[gid] = (flag[m & MAP_IDX] != 0.0) ? (m | MAP_ESS) : m;
}

__global__ void set_flags_kernel(const int32_t n, const int32_t *__restrict__ idx, double *__restrict__ v, double val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[idx[i]] = val;
}
__global__ void copy_idx_kernel(const int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ src,
                                double *__restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[idx[i]] = src[idx[i]];
}
__global__ void axpby_kernel(const int64_t n, const double a, const double *__restrict__ x, const double b,
                             double *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a * x[i] + b * y[i];
}

// ---- deterministic fused reductions ------------------------------------------------
// Every reducing kernel ends with block_finish(): per-block partials are written,
// the last block to arrive sums them in index order (fixed order => run-to-run
// bit-reproducible for a fixed grid) and stores the result in out[slot].
static constexpr int RED_THREADS = 256;

// add_to != nullptr: the result is added to *add_to instead of stored in *out
__device__ __forceinline__ void block_finish(double v, double *partial, unsigned int *counter, double *out, double *add_to = nullptr) {
  __shared__ double sh[RED_THREADS / 32];
  __shared__ bool last;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < RED_THREADS / 32; i++) s += sh[i];
    partial[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicInc(counter, gridDim.x - 1);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence();
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += ((volatile double *)partial)[i];
    // fixed-order tree over lanes
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      if (add_to) *add_to += s;
      else *out = s;
    }
  }
}

__global__ void __launch_bounds__(RED_THREADS)
dot_kernel(const int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *partial,
           unsigned int *counter, double *out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += a[i] * b[i];
  block_finish(s, partial, counter, out);
}
// dst[idx] = src[idx];  *out += sum src[idx]^2
__global__ void __launch_bounds__(RED_THREADS)
copy_idx_dot_kernel(const int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst,
                    double *partial, unsigned int *counter, double *out) {
  double s = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = src[idx[i]];
    dst[idx[i]] = v;
    s += v * v;
  }
  block_finish(s, partial, counter, nullptr, out);
}
// out = sum dinv * b * b   (<C b, b>)
__global__ void __launch_bounds__(RED_THREADS)
cdot_kernel(const int64_t n, const double *__restrict__ dinv, const double *__restrict__ b, double *partial,
            unsigned int *counter, double *out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    s += (dinv ? dinv[i] : 1.0) * bi * bi;
  }
  block_finish(s, partial, counter, out);
}
// r = b - r (optional), p = C r, out = <r, p>
__global__ void __launch_bounds__(RED_THREADS)
pcg_init_kernel(const int64_t n, const int sub, const double *__restrict__ b, const double *__restrict__ dinv,
                double *__restrict__ r, double *__restrict__ p, double *partial, unsigned int *counter, double *out) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = sub ? b[i] - r[i] : b[i];
    const double pi = (dinv ? dinv[i] : 1.0) * ri;
    r[i] = ri;
    p[i] = pi;
    s += ri * pi;
  }
  block_finish(s, partial, counter, out);
}
// alpha = gamma/<s,p>; x += alpha p; r -= alpha s; s = C r; out = <r, s>
// Five read and three write streams: 16-byte accesses, all loads of an iteration issued before its stores (left to
// itself the compiler interleaves them load by load, which cost 30 % of the bandwidth in one build and not in the next).
__global__ void __launch_bounds__(RED_THREADS)
pcg_update_kernel(const int64_t n, const double *__restrict__ gamma, const double *__restrict__ sdotp, const double *__restrict__ dinv,
                  const double *__restrict__ p, double *__restrict__ x, double *__restrict__ r, double *__restrict__ s,
                  double *partial, unsigned int *counter, double *out) {
  // <s, p> = 0 (exact initial guess: r = p = 0): hypre breaks before the update and returns x unchanged
  const double alpha = *sdotp != 0.0 ? *gamma / *sdotp : 0.0;
  double acc0 = 0.0, acc1 = 0.0;
  const int64_t n2 = n >> 1;
  const double2 *p2 = reinterpret_cast<const double2 *>(p), *d2 = reinterpret_cast<const double2 *>(dinv);
  double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r), *s2 = reinterpret_cast<double2 *>(s);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const double2 pv = __ldg(p2 + i), xv = x2[i], rv = r2[i], sv = s2[i];
    const double2 dv = dinv ? __ldg(d2 + i) : make_double2(1.0, 1.0);
    const double2 rn = make_double2(rv.x - alpha * sv.x, rv.y - alpha * sv.y);
    const double2 zn = make_double2(dv.x * rn.x, dv.y * rn.y);
    x2[i] = make_double2(xv.x + alpha * pv.x, xv.y + alpha * pv.y);
    r2[i] = rn;
    s2[i] = zn;
    acc0 = fma(rn.x, zn.x, acc0);
    acc1 = fma(rn.y, zn.y, acc1);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t i = n - 1;
    x[i] += alpha * p[i];
    const double ri = r[i] - alpha * s[i], zi = (dinv ? dinv[i] : 1.0) * ri;
    r[i] = ri;
    s[i] = zi;
    acc0 = fma(ri, zi, acc0);
  }
  block_finish(acc0 + acc1, partial, counter, out);
}
// p = s + beta p, beta = gamma_new / gamma taken from the device (the host has not seen gamma_new yet)
__global__ void pcg_dir_kernel(const int64_t n, const double *__restrict__ gamma_new, const double *__restrict__ gamma,
                               const double *__restrict__ s, double *__restrict__ p) {
  const double beta = *gamma_new / *gamma;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = s[i] + beta * p[i];
}
// End of one iteration of the graph-captured loop (pcg_graph_build): hypre's stopping tests on the device, in the order
// of the host loop below -- <s, p> = 0: stop, x unchanged; gamma_new / bi_prod < eps: converged; |gamma_new| < 1e-292:
// stop; iteration limit -- and the value of the WHILE node's condition.  S[0] gamma, S[1] <s,p>, S[2] gamma_new,
// S[3] bi_prod, S[4] eps; state[0] iterations done, [1] converged, [2] limit.
__global__ void pcg_check_kernel(double *S, int *state, cudaGraphConditionalHandle handle) {
  const double sdotp = S[1], gamma_new = S[2];
  const int it = state[0] + 1;
  state[0] = it;
  bool stop = false;
  if (sdotp == 0.0) stop = true;
  else {
    S[0] = gamma_new;
    if (gamma_new / S[3] < S[4]) { state[1] = 1; stop = true; }
    else if (fabs(gamma_new) < 1e-292) stop = true;
  }
  if (it >= state[2]) stop = true;
  cudaGraphSetConditional(handle, stop ? 0u : 1u);
}
// ---- Chebyshev preconditioner (HCURL_PC_CHEBYSHEV) ----------------------------------------------------------
// deterministic pseudo-random start vector of the power method, values in [0.5, 1.5) (oracle/pcg.py: start_vector)
__global__ void cheb_seed_kernel(const int64_t n, const uint32_t offset, double *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t h = ((uint32_t)i + offset) * 2654435761u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    v[i] = 0.5 + (double)h / 4294967296.0;
  }
}
// out[0] = <v, w>, out[1] = <v, D v>  (D = 1 / dinv)
__global__ void __launch_bounds__(RED_THREADS)
cheb_rayleigh_kernel(const int64_t n, const double *__restrict__ v, const double *__restrict__ w, const double *__restrict__ dinv,
                     double *partial, unsigned int *counter, double *out, const int which) {
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s += which == 0 ? v[i] * w[i] : v[i] * v[i] / dinv[i];
  block_finish(s, partial, counter, out);
}
// v = scale * dinv * w
__global__ void cheb_scale_kernel(const int64_t n, const double scale, const double *__restrict__ dinv, const double *__restrict__ w,
                                  double *__restrict__ v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = scale * dinv[i] * w[i];
}
// res = r; d = inv_theta * dinv * r; z = d
__global__ void cheb_start_kernel(const int64_t n, const double inv_theta, const double *__restrict__ dinv, const double *__restrict__ r,
                                  double *__restrict__ res, double *__restrict__ d, double *__restrict__ z) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i], di = inv_theta * dinv[i] * ri;
    res[i] = ri;
    d[i] = di;
    z[i] = di;
  }
}
// res -= t; d = c1 d + c2 dinv res; z += d
__global__ void cheb_step_kernel(const int64_t n, const double c1, const double c2, const double *__restrict__ dinv,
                                 const double *__restrict__ t, double *__restrict__ res, double *__restrict__ d, double *__restrict__ z) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = res[i] - t[i], di = c1 * d[i] + c2 * dinv[i] * ri;
    res[i] = ri;
    d[i] = di;
    z[i] += di;
  }
}
// alpha = gamma / <s,p>; x += alpha p; r -= alpha s
__global__ void pcg_xr_kernel(const int64_t n, const double *__restrict__ gamma, const double *__restrict__ sdotp, const double *__restrict__ p,
                              const double *__restrict__ s, double *__restrict__ x, double *__restrict__ r) {
  const double alpha = *sdotp != 0.0 ? *gamma / *sdotp : 0.0;  // breakdown: leave x, r unchanged (as hypre does)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    r[i] -= alpha * s[i];
  }
}
__global__ void invert_kernel(const int64_t n, double *__restrict__ d) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = 1.0 / d[i];
}

// FP64 FMA peak micro-benchmark
__global__ void dfma_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
    a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

static int red_grid(hcurl_ctx *ctx, int64_t n) {
  int64_t g = (n + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
  if (g < 1) g = 1;
  const int64_t cap = (int64_t)ctx->sm_count * 8;
  if (g > cap) g = cap;
  if (g > ctx->n_partial) g = ctx->n_partial;
  return (int)g;
}

// ---- apply on extended device vectors -------------------------------------------------
int op_apply_ext(hcurl_op *op, double *d_x, double *d_y, bool constrained, double scale, bool accumulate, double *d_dot,
                 bool *dot_done) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  const int64_t next = sp->n_ext();
  // the patch kernel writes every dof the elements touch (owner stores, the others add after it) and the identity on
  // the essential dofs: nothing to clear, nothing to fix up; the other kernels RED-add everything
  const bool patch_path = hex_patch_eligible(op) && (((uintptr_t)d_x | (uintptr_t)d_y) & 15) == 0;
  const bool fuse_dot = patch_path && d_dot && !accumulate;
  if (dot_done) *dot_done = fuse_dot;
  if (exchange_broadcast(sp, d_x)) return 1;
  if (!patch_path) {
    if (!accumulate) HC_CUDA(cudaMemsetAsync(d_y, 0, sizeof(double) * next, ctx->stream));
    else if (sp->n_ghost) HC_CUDA(cudaMemsetAsync(d_y + sp->n_tdof, 0, sizeof(double) * sp->n_ghost, ctx->stream));
  }
  const uint32_t *map = (constrained && op->d_map_ess) ? op->d_map_ess : sp->d_map;
  cudaEvent_t e0 = op->ev0, e1 = op->ev1;
  if (op->ring_ok) { e0 = op->ring0[op->ring_n % hcurl_op::RING]; e1 = op->ring1[op->ring_n % hcurl_op::RING]; op->ring_n++; op->ev0 = e0; op->ev1 = e1; }
  if (op->timed) HC_CUDA(cudaEventRecord(e0, ctx->stream));
  int rc;
  if (patch_path) { op->last_kernel = 3; rc = hex_patch_launch(op, constrained, d_x, d_y, scale, accumulate, fuse_dot ? d_dot : nullptr); }
  else if (hex_fast_eligible(op)) { op->last_kernel = 2; rc = hex_fast_launch(op, constrained, d_x, d_y, scale); }
  else { op->last_kernel = 1; rc = (sp->geom == HCURL_GEOM_HEX) ? hex_apply_launch(op, map, d_x, d_y, scale) : tet_apply_launch(op, map, d_x, d_y, scale); }
  if (rc) return rc;
  if (op->timed) HC_CUDA(cudaEventRecord(e1, ctx->stream));
  if (exchange_reduce(sp, d_y)) return 1;
  if (patch_path && sp->n_nbr == 0) return 0;  // essential rows were written by the kernel
  if (constrained && op->n_ess > 0) {
    if (fuse_dot) {  // essential rows are identity: y = x there, and <x, y> gains x^2
      int g = (op->n_ess + RED_THREADS * 4 - 1) / (RED_THREADS * 4);
      if (g > ctx->n_partial) g = ctx->n_partial;
      copy_idx_dot_kernel<<<g, RED_THREADS, 0, ctx->stream>>>(op->n_ess, op->d_ess, d_x, d_y, ctx->d_partial, ctx->d_counter, d_dot);
    } else copy_idx_kernel<<<(op->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op->n_ess, op->d_ess, d_x, d_y);
    ctx->launches++;
  }
  HC_CUDA(cudaGetLastError());
  return 0;
}

int op_diag_ext(hcurl_op *op, double *d_diag) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  HC_CUDA(cudaMemsetAsync(d_diag, 0, sizeof(double) * sp->n_ext(), ctx->stream));
  int rc = (sp->geom == HCURL_GEOM_HEX) ? hex_diag_launch(op, d_diag) : tet_diag_launch(op, d_diag);
  if (rc) return rc;
  if (exchange_reduce(sp, d_diag, false)) return 1;  // diagonal entries are orientation-independent
  if (op->n_ess > 0) set_flags_kernel<<<(op->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op->n_ess, op->d_ess, d_diag, 1.0);
  HC_CUDA(cudaGetLastError());
  return 0;
}

static int ensure_work(hcurl_op *op) {
  const size_t bytes = sizeof(double) * (size_t)op->sp->n_ext();
  if (!op->d_xin) HC_CUDA(cudaMalloc(&op->d_xin, bytes));
  if (!op->d_yout) HC_CUDA(cudaMalloc(&op->d_yout, bytes));
  return 0;
}

static cudaMemcpyKind kind_in(int where) { return where != HCURL_HOST ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice; }
static cudaMemcpyKind kind_out(int where) { return where != HCURL_HOST ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

}  // namespace hcurl

using namespace hcurl;

extern "C" {

const char *hcurl_version(void) { return "hcurl_b200 0.1 (sm_100a, FP64)"; }
const char *hcurl_last_error(void) { return g_err; }

int hcurl_comm_unique_id(void *id128) { return comm_unique_id(id128); }

hcurl_ctx *hcurl_ctx_create(int device, const void *uid, int rank, int nranks) {
  int ndev = 0;
  if (!cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev <= 0) {
    set_error("no CUDA device available (this library has no CPU fallback)");
    return nullptr;
  }
  if (device < 0 || device >= ndev) { set_error("invalid device %d (have %d)", device, ndev); return nullptr; }
  // one process per GPU: constant-memory tables and cached device buffers of the kernels are per process
  static int first_device = -1;
  if (first_device >= 0 && device != first_device) {
    set_error("this process already uses CUDA device %d: one process per GPU (requested %d)", first_device, device);
    return nullptr;
  }
  first_device = device;
  HC_CUDA_NULL(cudaSetDevice(device));
  hcurl_ctx *ctx = new hcurl_ctx;
  ctx->device = device; ctx->rank = rank; ctx->nranks = nranks < 1 ? 1 : nranks;
  cudaDeviceProp prop;
  HC_CUDA_NULL(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  HC_CUDA_NULL(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  HC_CUDA_NULL(cudaMalloc(&ctx->d_red, sizeof(double) * 16));
  HC_CUDA_NULL(cudaMemset(ctx->d_red, 0, sizeof(double) * 16));
  HC_CUDA_NULL(cudaMallocHost(&ctx->h_red, sizeof(double) * 16));
  HC_CUDA_NULL(cudaMalloc(&ctx->d_counter, sizeof(unsigned int)));
  HC_CUDA_NULL(cudaMemset(ctx->d_counter, 0, sizeof(unsigned int)));
  ctx->n_partial = ctx->sm_count * 8;
  HC_CUDA_NULL(cudaMalloc(&ctx->d_partial, sizeof(double) * ctx->n_partial));
  HC_CUDA_NULL(cudaEventCreate(&ctx->mark[0]));
  HC_CUDA_NULL(cudaEventCreate(&ctx->mark[1]));
  int coop = 0;
  HC_CUDA_NULL(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  ctx->coop_ok = coop != 0;
  if (comm_init(ctx, uid)) { delete ctx; return nullptr; }
  if (ctx->nranks > 1) {  // second stream + hand-off events: the shared-dof exchange overlaps with interior patches
    HC_CUDA_NULL(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) HC_CUDA_NULL(cudaEventCreateWithFlags(&ctx->ev_x[i], cudaEventDisableTiming));
  }
  return ctx;
}

void hcurl_ctx_destroy(hcurl_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->comm_stream) { cudaStreamSynchronize(ctx->comm_stream); cudaStreamDestroy(ctx->comm_stream); }
  if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
  if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
  for (int i = 0; i < 4; i++) if (ctx->ev_x[i]) cudaEventDestroy(ctx->ev_x[i]);
  comm_destroy(ctx);
  cudaFree(ctx->d_red); cudaFreeHost(ctx->h_red); cudaFree(ctx->d_counter); cudaFree(ctx->d_partial);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}
int hcurl_ctx_device(const hcurl_ctx *ctx) { return ctx ? ctx->device : -1; }
int hcurl_ctx_synchronize(hcurl_ctx *ctx) { HC_CUDA(cudaStreamSynchronize(ctx->stream)); return 0; }
int hcurl_ctx_mark(hcurl_ctx *ctx, int which) {
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaEventRecord(ctx->mark[which ? 1 : 0], ctx->stream));
  return 0;
}
double hcurl_ctx_elapsed_ms(hcurl_ctx *ctx) {
  float ms = 0.f;
  cudaSetDevice(ctx->device);
  if (cudaEventSynchronize(ctx->mark[1]) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, ctx->mark[0], ctx->mark[1]) != cudaSuccess) return -1.0;
  return ms;
}
int64_t hcurl_ctx_launch_count(const hcurl_ctx *ctx) { return ctx->launches; }

hcurl_space *hcurl_space_create(hcurl_ctx *ctx, int geom, int geom_order, int order, int64_t n_elem, const double *elem_xyz,
                                const int32_t *elem_ldof, const int32_t *elem_verts, const int32_t *elem_face_base,
                                int32_t n_ldof, const int32_t *ldof_ltdof, int32_t n_tdof, int n_nbr, const int32_t *nbr_rank,
                                const int32_t *own_off, const int32_t *own_ldof, const int32_t *ghost_off,
                                const int32_t *ghost_ldof, const int8_t *ghost_sign) {
  if (!ctx) { set_error("null context"); return nullptr; }
  if (geom == HCURL_GEOM_HEX) {
    if (order < 1 || order > 4) { set_error("hex order must be 1..4 (got %d)", order); return nullptr; }
    if (geom_order != 1 && geom_order != 2) { set_error("hex geometry order must be 1 or 2"); return nullptr; }
  } else if (geom == HCURL_GEOM_TET) {
    if (order < 1 || order > 3) { set_error("tet order must be 1..3 (got %d)", order); return nullptr; }
    if (geom_order != 1) { set_error("tets must be straight (geometry order 1)"); return nullptr; }
    if (order >= 2 && (!elem_verts || !elem_face_base)) {
      set_error("tets of order >= 2 need elem_verts and elem_face_base (triangle-face DofTransformation)");
      return nullptr;
    }
  } else { set_error("unknown geometry %d", geom); return nullptr; }
  if (n_elem <= 0 || !elem_xyz || !elem_ldof) { set_error("empty mesh or null arrays"); return nullptr; }
  if ((int64_t)n_ldof >= (int64_t)MAP_IDX) { set_error("too many local dofs for the 30-bit map encoding"); return nullptr; }
  HC_CUDA_NULL(cudaSetDevice(ctx->device));
  hcurl_space *sp = new hcurl_space;
  sp->ctx = ctx; sp->geom = geom; sp->geom_order = geom_order; sp->order = order; sp->n_elem = n_elem;
  sp->n_ldof = n_ldof;
  const bool hex = geom == HCURL_GEOM_HEX;
  sp->nd = hex ? nd_hex_ndof(order) : nd_tet_ndof(order);
  sp->ngn = hex ? (geom_order == 1 ? 8 : 27) : 4;
  sp->n_elem_pad = (n_elem + 31) / 32 * 32;
  // ---- parallel layout
  std::vector<int32_t> l2e;
  if (n_nbr > 0 && (ctx->nranks <= 1 || !ldof_ltdof || !nbr_rank || !own_off || !ghost_off)) {
    set_error("neighbour lists given (n_nbr = %d) but the context has one rank / the parallel layout is incomplete", n_nbr);
    delete sp;
    return nullptr;
  }
  if (ldof_ltdof) {
    for (int q = 0; q < n_nbr; q++)
      if (nbr_rank[q] < 0 || nbr_rank[q] >= ctx->nranks || nbr_rank[q] == ctx->rank || own_off[q + 1] < own_off[q] || ghost_off[q + 1] < ghost_off[q]) {
        set_error("bad neighbour entry %d (rank %d of %d)", q, nbr_rank[q], ctx->nranks);
        delete sp;
        return nullptr;
      }
    for (int32_t i = 0; n_nbr > 0 && i < own_off[n_nbr]; i++)
      if (own_ldof[i] < 0 || own_ldof[i] >= n_ldof) { set_error("own_ldof[%d] out of range", i); delete sp; return nullptr; }
    for (int32_t i = 0; n_nbr > 0 && i < ghost_off[n_nbr]; i++)
      if (ghost_ldof[i] < 0 || ghost_ldof[i] >= n_ldof) { set_error("ghost_ldof[%d] out of range", i); delete sp; return nullptr; }
    sp->n_tdof = n_tdof;
    sp->n_nbr = n_nbr;
    sp->nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
    sp->own_off.assign(own_off, own_off + n_nbr + 1);
    sp->ghost_off.assign(ghost_off, ghost_off + n_nbr + 1);
    sp->n_ghost = n_nbr ? ghost_off[n_nbr] : 0;
    l2e.assign(n_ldof, -1);
    for (int32_t i = 0; i < n_ldof; i++) if (ldof_ltdof[i] >= 0) l2e[i] = ldof_ltdof[i];
    for (int32_t g = 0; g < sp->n_ghost; g++) l2e[ghost_ldof[g]] = n_tdof + g;
    for (int32_t i = 0; i < n_ldof; i++)
      if (l2e[i] < 0) { set_error("ldof %d is neither owned nor listed as a ghost", i); delete sp; return nullptr; }
    const int32_t n_own = n_nbr ? own_off[n_nbr] : 0;
    if (n_own > 0) {
      std::vector<int32_t> own_ext(n_own);
      for (int32_t i = 0; i < n_own; i++) own_ext[i] = l2e[own_ldof[i]];
      HC_CUDA_NULL(cudaMalloc(&sp->d_own_idx, sizeof(int32_t) * n_own));
      HC_CUDA_NULL(h2d_sync(ctx, sp->d_own_idx, own_ext.data(), sizeof(int32_t) * n_own));
      sp->own_ext = own_ext;
      HC_CUDA_NULL(cudaMalloc(&sp->d_sendbuf, sizeof(double) * n_own));
      HC_CUDA_NULL(cudaMalloc(&sp->d_recvbuf, sizeof(double) * n_own));
    }
    if (ghost_sign && sp->n_ghost > 0) {
      HC_CUDA_NULL(cudaMalloc(&sp->d_ghost_sign, sp->n_ghost));
      HC_CUDA_NULL(h2d_sync(ctx, sp->d_ghost_sign, ghost_sign, sp->n_ghost));
    }
    sp->ldof_to_ext = l2e;
  } else {
    sp->n_tdof = n_ldof;
  }
  // ---- geometry
  const size_t nxyz = (size_t)n_elem * sp->ngn * 3;
  HC_CUDA_NULL(cudaMalloc(&sp->d_xyz, sizeof(double) * nxyz));
  if (hex && geom_order == 1) {
    double *tmp = nullptr;
    HC_CUDA_NULL(cudaMalloc(&tmp, sizeof(double) * nxyz));
    HC_CUDA_NULL(h2d_sync(ctx, tmp, elem_xyz, sizeof(double) * nxyz));
    permute_xyz_kernel<<<(unsigned)((nxyz + 255) / 256), 256, 0, ctx->stream>>>(n_elem, tmp, sp->d_xyz);
    HC_CUDA_NULL(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp);
  } else {
    HC_CUDA_NULL(h2d_sync(ctx, sp->d_xyz, elem_xyz, sizeof(double) * nxyz));
  }
  // ---- element -> dof map
  {
    const size_t nmap = (size_t)n_elem * sp->nd;
    int32_t *d_ldof = nullptr;
    HC_CUDA_NULL(cudaMalloc(&d_ldof, sizeof(int32_t) * nmap));
    HC_CUDA_NULL(h2d_sync(ctx, d_ldof, elem_ldof, sizeof(int32_t) * nmap));
    int *d_loc = nullptr, *d_sgn = nullptr;
    int32_t *d_l2e = nullptr;
    if (hex) {
      std::vector<int> loc, sgn;
      nd_hex_dof_map(order, loc, sgn);
      HC_CUDA_NULL(cudaMalloc(&d_loc, sizeof(int) * sp->nd));
      HC_CUDA_NULL(cudaMalloc(&d_sgn, sizeof(int) * sp->nd));
      HC_CUDA_NULL(h2d_sync(ctx, d_loc, loc.data(), sizeof(int) * sp->nd));
      HC_CUDA_NULL(h2d_sync(ctx, d_sgn, sgn.data(), sizeof(int) * sp->nd));
    }
    if (!l2e.empty()) {
      HC_CUDA_NULL(cudaMalloc(&d_l2e, sizeof(int32_t) * n_ldof));
      HC_CUDA_NULL(h2d_sync(ctx, d_l2e, l2e.data(), sizeof(int32_t) * n_ldof));
    }
    const size_t nout = hex ? nmap : (size_t)sp->nd * sp->n_elem_pad;
    HC_CUDA_NULL(cudaMalloc(&sp->d_map, sizeof(uint32_t) * nout));
    HC_CUDA_NULL(cudaMemsetAsync(sp->d_map, 0, sizeof(uint32_t) * nout, ctx->stream));
    encode_map_kernel<<<(unsigned)((nmap + 255) / 256), 256, 0, ctx->stream>>>(n_elem, sp->nd, sp->n_elem_pad, hex ? 0 : 1, d_ldof,
                                                                               d_loc, d_sgn, d_l2e, sp->d_map);
    HC_CUDA_NULL(cudaGetLastError());
    HC_CUDA_NULL(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_ldof); cudaFree(d_loc); cudaFree(d_sgn); cudaFree(d_l2e);
  }
  // ---- patch decomposition for the patch-based hex kernels (hex_patch.cu)
  if (hex_patch_build(sp, elem_xyz, elem_ldof, l2e.empty() ? nullptr : l2e.data())) { hcurl_space_destroy(sp); return nullptr; }
  // ---- triangle-face dof transformations (tets, p >= 2)
  if (!hex && order >= 2) {
    const int nn = order * (order - 1) / 2;  // face nodes
    sp->nf = 2 * nn;
    std::vector<double> R((size_t)n_elem * 4 * sp->nf * sp->nf);
    static const int TF[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    for (int64_t e = 0; e < n_elem; e++)
      for (int f = 0; f < 4; f++) {
        int32_t lv[3] = {elem_verts[e * 4 + TF[f][0]], elem_verts[e * 4 + TF[f][1]], elem_verts[e * 4 + TF[f][2]]};
        tri_face_transform(order, lv, elem_face_base + (e * 4 + f) * 3, &R[((size_t)e * 4 + f) * sp->nf * sp->nf]);
      }
    HC_CUDA_NULL(cudaMalloc(&sp->d_faceR, sizeof(double) * R.size()));
    HC_CUDA_NULL(h2d_sync(ctx, sp->d_faceR, R.data(), sizeof(double) * R.size()));
  }
  // ---- several ranks: map the neighbours' landing buffers for the peer-memory exchange (collective; comm.cu)
  if (ctx->nranks > 1 && p2p_setup(sp)) { hcurl_space_destroy(sp); return nullptr; }
  return sp;
}

int64_t hcurl_space_ext_size(const hcurl_space *sp) { return sp->n_ext(); }
void hcurl_space_destroy(hcurl_space *sp) {
  if (!sp) return;
  cudaSetDevice(sp->ctx->device);
  cudaStreamSynchronize(sp->ctx->stream);
  hex_patch_free(sp);
  cudaFree(sp->d_mapT); cudaFree(sp->d_xyzT);
  cudaFree(sp->d_xyz); cudaFree(sp->d_map); cudaFree(sp->d_faceR); cudaFree(sp->d_own_idx);
  p2p_free(sp);
  cudaFree(sp->d_ghost_sign); cudaFree(sp->d_sendbuf); cudaFree(sp->d_recvbuf);
  delete sp;
}
int32_t hcurl_space_ndof_local(const hcurl_space *sp) { return sp->n_ldof; }
int32_t hcurl_space_ndof_true(const hcurl_space *sp) { return sp->n_tdof; }
int hcurl_space_elem_ndof(const hcurl_space *sp) { return sp->nd; }

hcurl_op *hcurl_op_create(hcurl_ctx *ctx, hcurl_space *sp, int q) {
  if (!ctx || !sp) { set_error("null context or space"); return nullptr; }
  hcurl_op *op = new hcurl_op;
  op->ctx = ctx; op->sp = sp;
  const int p = sp->order;
  if (q == HCURL_QUAD_AUTO) {  // one p+1 rule where it is exact (all elements parallelepipeds), MFEM's rules otherwise
    q = 0;
    if (sp->geom == HCURL_GEOM_HEX && sp->geom_order == 1) {
      bool affine = false;
      if (cudaSetDevice(ctx->device) != cudaSuccess || hex_all_affine(sp, &affine)) { delete op; return nullptr; }
      if (affine) q = p + 1;
    }
  }
  if (sp->geom == HCURL_GEOM_HEX) {
    if (q == 0) {
      op->qc = p + 1;
      op->qm = (sp->geom_order == 1) ? p + 2 : p + 3;  // order 3g-1+2p -> n = order/2+1 (odd-rounded)
      op->single_rule = false;
    } else if (q == p + 1 || q == p + 2) {
      op->qc = op->qm = q;
      op->single_rule = true;
    } else {
      set_error("quad_pts_per_dim must be 0, p+1, p+2 or HCURL_QUAD_AUTO (got %d for p=%d)", q, p);
      delete op;
      return nullptr;
    }
  }
  HC_CUDA_NULL(cudaSetDevice(ctx->device));
  for (int i = 0; i < hcurl_op::RING; i++) {
    HC_CUDA_NULL(cudaEventCreate(&op->ring0[i]));
    HC_CUDA_NULL(cudaEventCreate(&op->ring1[i]));
  }
  op->ring_ok = true;
  op->ev0 = op->ring0[0]; op->ev1 = op->ring1[0];
  op->timed = true;
  return op;
}

void hcurl_op_destroy(hcurl_op *op) {
  if (!op) return;
  cudaSetDevice(op->ctx->device);
  cudaStreamSynchronize(op->ctx->stream);
  cudaFree(op->d_alpha); cudaFree(op->d_beta); cudaFree(op->d_map_ess); cudaFree(op->d_ess);
  cudaFree(op->d_edata); cudaFree(op->d_mapT_ess); cudaFree(op->d_edataP); cudaFree(op->d_blob_ess);
  cudaFree(op->d_emat); cudaFree(op->d_xin); cudaFree(op->d_yout); cudaFree(op->d_xin2); cudaFree(op->d_yout2); cudaFree(op->d_tmp);
  for (int i = 0; i < 2; i++) { if (op->ev_h2d[i]) cudaEventDestroy(op->ev_h2d[i]); if (op->ev_run[i]) cudaEventDestroy(op->ev_run[i]); if (op->ev_d2h[i]) cudaEventDestroy(op->ev_d2h[i]); }
  cudaFree(op->d_refq_k); cudaFree(op->d_refq_m);
  for (int i = 0; i < hcurl_op::RING; i++) { cudaEventDestroy(op->ring0[i]); cudaEventDestroy(op->ring1[i]); }
  delete op;
}
int hcurl_op_set_kernel(hcurl_op *op, int which) {
  if (which < 0 || which > 3) { set_error("kernel selector must be 0 (auto), 1 (generic), 2 (element-wise fast) or 3 (patch)"); return 1; }
  op->force_generic = which == 1;
  op->kernel_sel = which;
  op->generation++;
  return 0;
}
int hcurl_op_last_kernel(const hcurl_op *op) { return op->last_kernel; }
int hcurl_op_qpts_curl(const hcurl_op *op) { return op->sp->geom == HCURL_GEOM_HEX ? op->qc * op->qc * op->qc : op->tet_nqc; }
int hcurl_op_qpts_mass(const hcurl_op *op) { return op->sp->geom == HCURL_GEOM_HEX ? op->qm * op->qm * op->qm : op->tet_nqm; }

int hcurl_op_set_tet_rules(hcurl_op *op, int nq_curl, const double *pts_curl, const double *w_curl, int nq_mass,
                           const double *pts_mass, const double *w_mass) {
  if (!op || op->sp->geom != HCURL_GEOM_TET) { set_error("hcurl_op_set_tet_rules: not a tet operator"); return 1; }
  if (nq_curl < 1 || nq_mass < 1 || nq_curl > 64 || nq_mass > 64 || !pts_curl || !w_curl || !pts_mass || !w_mass) {
    set_error("hcurl_op_set_tet_rules: rules of 1..64 points needed");
    return 1;
  }
  HC_CUDA(cudaSetDevice(op->ctx->device));
  return tet_set_rules(op, nq_curl, pts_curl, w_curl, nq_mass, pts_mass, w_mass);
}

int hcurl_op_set_coefficients(hcurl_op *op, int layout, const double *alpha, const double *beta, int where) {
  hcurl_space *sp = op->sp;
  op->generation++;
  HC_CUDA(cudaSetDevice(op->ctx->device));
  if (layout == HCURL_COEF_PER_QPT && sp->geom != HCURL_GEOM_HEX && (op->tet_nqc < 1 || op->tet_nqm < 1)) {
    set_error("PER_QPT coefficients on tets need the quadrature rules first (hcurl_op_set_tet_rules)");
    return 1;
  }
  if (layout != HCURL_COEF_PER_ELEM && layout != HCURL_COEF_PER_QPT) { set_error("bad coefficient layout %d", layout); return 1; }
  op->coef_layout = layout;
  const int64_t na = sp->n_elem * (layout == HCURL_COEF_PER_QPT ? hcurl_op_qpts_curl(op) : 1);
  const int64_t nb = sp->n_elem * (layout == HCURL_COEF_PER_QPT ? hcurl_op_qpts_mass(op) : 1);
  auto put = [&](const double *src, double *&dst, int64_t &cur, int64_t n) -> int {
    if (!src) return 0;
    if (cur != n) { cudaFree(dst); dst = nullptr; HC_CUDA(cudaMalloc(&dst, sizeof(double) * n)); cur = n; }
    HC_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, kind_in(where), op->ctx->stream));
    return 0;
  };
  if (put(alpha, op->d_alpha, op->n_alpha, na)) return 1;
  if (put(beta, op->d_beta, op->n_beta, nb)) return 1;
  op->has_alpha = alpha != nullptr;
  op->has_beta = beta != nullptr;
  HC_CUDA(cudaStreamSynchronize(op->ctx->stream));  // host buffers are borrowed for the call only
  op->emat_valid = false;
  op->edata_valid = false;
  op->edataP_valid = false;
  return 0;
}

int hcurl_op_set_essential_tdofs(hcurl_op *op, const int32_t *ess, int32_t n_ess) {
  hcurl_space *sp = op->sp;
  op->generation++;
  hcurl_ctx *ctx = op->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  for (int32_t i = 0; i < n_ess; i++)
    if (ess[i] < 0 || ess[i] >= sp->n_tdof) { set_error("essential tdof %d out of range", ess[i]); return 1; }
  cudaFree(op->d_ess); op->d_ess = nullptr;
  cudaFree(op->d_map_ess); op->d_map_ess = nullptr;
  cudaFree(op->d_mapT_ess); op->d_mapT_ess = nullptr;
  op->n_ess = n_ess;
  // flags live on the extended vector so that ghost copies of essential dofs are masked too
  double *flag = nullptr;
  HC_CUDA(cudaMalloc(&flag, sizeof(double) * sp->n_ext()));
  HC_CUDA(cudaMemsetAsync(flag, 0, sizeof(double) * sp->n_ext(), ctx->stream));
  if (n_ess > 0) {
    HC_CUDA(cudaMalloc(&op->d_ess, sizeof(int32_t) * n_ess));
    HC_CUDA(cudaMemcpyAsync(op->d_ess, ess, sizeof(int32_t) * n_ess, cudaMemcpyHostToDevice, ctx->stream));
    set_flags_kernel<<<(n_ess + 255) / 256, 256, 0, ctx->stream>>>(n_ess, op->d_ess, flag, 1.0);
  }
  if (exchange_broadcast(sp, flag)) return 1;
  // any rank may hold flagged ghosts even when it owns no essential dof
  const size_t nmap = sp->geom == HCURL_GEOM_HEX ? (size_t)sp->n_elem * sp->nd : (size_t)sp->nd * sp->n_elem_pad;
  HC_CUDA(cudaMalloc(&op->d_map_ess, sizeof(uint32_t) * nmap));
  flag_map_kernel<<<(unsigned)((nmap + 255) / 256), 256, 0, ctx->stream>>>((int64_t)nmap, sp->d_map, flag, op->d_map_ess);
  HC_CUDA(cudaGetLastError());
  if (hex_patch_flag_essential(op, flag)) return 1;
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(flag);
  return 0;
}

static int apply_common(hcurl_op *op, const double *x, double *y, int where, bool constrained, double scale, int accumulate) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  if (!op->has_alpha && !op->has_beta) { set_error("operator has no coefficients set"); return 1; }
  const int64_t nt = sp->n_tdof;
  if ((where == HCURL_DEVICE && sp->n_ghost == 0) || where == HCURL_DEVICE_EXT) {
    // zero-copy: the caller's device vectors are (or have room for) the extended vectors
    return op_apply_ext(op, const_cast<double *>(x), y, constrained, scale, accumulate != 0);
  }
  if (ensure_work(op)) return 1;
  HC_CUDA(cudaMemcpyAsync(op->d_xin, x, sizeof(double) * nt, kind_in(where), ctx->stream));
  if (accumulate) HC_CUDA(cudaMemcpyAsync(op->d_yout, y, sizeof(double) * nt, kind_in(where), ctx->stream));
  if (op_apply_ext(op, op->d_xin, op->d_yout, constrained, scale, accumulate != 0)) return 1;
  HC_CUDA(cudaMemcpyAsync(y, op->d_yout, sizeof(double) * nt, kind_out(where), ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hcurl_op_apply(hcurl_op *op, const double *x, double *y, int where) {
  return apply_common(op, x, y, where, true, 1.0, 0);
}
int hcurl_op_apply_unconstrained(hcurl_op *op, const double *x, double *y, double scale, int accumulate, int where) {
  return apply_common(op, x, y, where, false, scale, accumulate);
}

// y_j = A x_j for a batch of vectors (mfem::Operator::ArrayMult, which EquationSystem inherits and MFEM's block / multi-
// vector callers use).  With HOST buffers the three stages of consecutive vectors run concurrently on three streams --
// x_{j+1} host -> device  |  apply of x_j  |  y_{j-1} device -> host -- through two staging pairs, so that both directions
// of the PCIe link are busy at once (a single host-buffer apply is one direction at a time: 2 x 406 MB at ~54 GB/s for the
// bench vector, against 0.44 ms of compute).  Pinned host buffers are needed for the copies to be asynchronous.
int hcurl_op_apply_batch(hcurl_op *op, int n, const double *const *x, double *const *y, int where) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  if (n < 0 || (n > 0 && (!x || !y))) { set_error("hcurl_op_apply_batch: bad arguments"); return 1; }
  if (where != HCURL_HOST) {
    for (int j = 0; j < n; j++) if (apply_common(op, x[j], y[j], where, true, 1.0, 0)) return 1;
    return 0;
  }
  HC_CUDA(cudaSetDevice(ctx->device));
  if (!op->has_alpha && !op->has_beta) { set_error("operator has no coefficients set"); return 1; }
  if (ensure_work(op)) return 1;
  const size_t bytes_ext = sizeof(double) * (size_t)sp->n_ext(), bytes = sizeof(double) * (size_t)sp->n_tdof;
  if (!op->d_xin2) HC_CUDA(cudaMalloc(&op->d_xin2, bytes_ext));
  if (!op->d_yout2) HC_CUDA(cudaMalloc(&op->d_yout2, bytes_ext));
  if (!ctx->h2d_stream) HC_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) HC_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    if (!op->ev_h2d[i]) HC_CUDA(cudaEventCreateWithFlags(&op->ev_h2d[i], cudaEventDisableTiming));
    if (!op->ev_run[i]) HC_CUDA(cudaEventCreateWithFlags(&op->ev_run[i], cudaEventDisableTiming));
    if (!op->ev_d2h[i]) HC_CUDA(cudaEventCreateWithFlags(&op->ev_d2h[i], cudaEventDisableTiming));
  }
  double *xin[2] = {op->d_xin, op->d_xin2}, *yout[2] = {op->d_yout, op->d_yout2};
  HC_CUDA(cudaStreamSynchronize(ctx->stream));  // earlier work on the staging buffers
  for (int j = 0; j < n; j++) {
    const int b = j & 1;
    if (j >= 2) HC_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, op->ev_run[b], 0));   // the apply of x_{j-2} has read its staging copy
    HC_CUDA(cudaMemcpyAsync(xin[b], x[j], bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
    HC_CUDA(cudaEventRecord(op->ev_h2d[b], ctx->h2d_stream));
    HC_CUDA(cudaStreamWaitEvent(ctx->stream, op->ev_h2d[b], 0));
    if (j >= 2) HC_CUDA(cudaStreamWaitEvent(ctx->stream, op->ev_d2h[b], 0));       // y_{j-2} has left its staging copy
    if (op_apply_ext(op, xin[b], yout[b], true, 1.0, false)) return 1;
    HC_CUDA(cudaEventRecord(op->ev_run[b], ctx->stream));
    HC_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, op->ev_run[b], 0));
    HC_CUDA(cudaMemcpyAsync(y[j], yout[b], bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
    HC_CUDA(cudaEventRecord(op->ev_d2h[b], ctx->d2h_stream));
  }
  HC_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hcurl_op_apply_complex(hcurl_op *op_re, hcurl_op *op_im, const double *xr, const double *xi, double *yr, double *yi, int where) {
  if (op_re->sp != op_im->sp) { set_error("complex apply: operators must share a space"); return 1; }
  hcurl_space *sp = op_re->sp;
  hcurl_ctx *ctx = op_re->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  const int64_t nt = sp->n_tdof, next = sp->n_ext();
  // work vectors: xr, xi, yr, yi on the extended layout
  if (!op_re->d_tmp) HC_CUDA(cudaMalloc(&op_re->d_tmp, sizeof(double) * 4 * next));
  double *dxr = op_re->d_tmp, *dxi = dxr + next, *dyr = dxi + next, *dyi = dyr + next;
  HC_CUDA(cudaMemcpyAsync(dxr, xr, sizeof(double) * nt, kind_in(where), ctx->stream));
  HC_CUDA(cudaMemcpyAsync(dxi, xi, sizeof(double) * nt, kind_in(where), ctx->stream));
  if (sp->geom == HCURL_GEOM_TET) {
    // tets: ONE fused pass over the elements (tet_apply_complex_kernel): xr, xi gathered once, both packed element
    // matrices streamed once, yr, yi scattered once; essential rows of the block system are the identity on both parts
    if (op_re->n_ess != op_im->n_ess) { set_error("complex apply: both operators need the same essential dofs"); return 1; }
    const uint32_t *map = op_re->d_map_ess ? op_re->d_map_ess : sp->d_map;
    if (exchange_broadcast(sp, dxr) || exchange_broadcast(sp, dxi)) return 1;
    HC_CUDA(cudaMemsetAsync(dyr, 0, sizeof(double) * 2 * next, ctx->stream));  // (dyr, dyi are adjacent)
    if (tet_apply_complex_launch(op_re, op_im, map, dxr, dxi, dyr, dyi)) return 1;
    if (exchange_reduce(sp, dyr) || exchange_reduce(sp, dyi)) return 1;
    if (op_re->n_ess > 0) {
      copy_idx_kernel<<<(op_re->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op_re->n_ess, op_re->d_ess, dxr, dyr);
      copy_idx_kernel<<<(op_re->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op_re->n_ess, op_re->d_ess, dxi, dyi);
      ctx->launches += 2;
    }
    HC_CUDA(cudaGetLastError());
  } else {
  // hexes: essential rows of the block system are identity on both parts: constrain Ar, and
  // apply Ai with the same essential masking but no identity contribution
  if (op_apply_ext(op_re, dxr, dyr, true, 1.0, false)) return 1;   // yr = Ar xr (ess: yr = xr)
  if (op_apply_ext(op_re, dxi, dyi, true, 1.0, false)) return 1;   // yi = Ar xi (ess: yi = xi)
  // masked Ai: use op_im's essential map if set, without the identity fix-up
  {
    const uint32_t *map = op_im->d_map_ess ? op_im->d_map_ess : sp->d_map;
    if (exchange_broadcast(sp, dxr)) return 1;  // ghosts already valid, cheap for n_nbr = 0
    auto run = [&](const double *xin, double *yacc, double scale) -> int {
      if (sp->n_ghost) HC_CUDA(cudaMemsetAsync(yacc + sp->n_tdof, 0, sizeof(double) * sp->n_ghost, ctx->stream));
      int rc = (sp->geom == HCURL_GEOM_HEX) ? hex_apply_launch(op_im, map, xin, yacc, scale) : tet_apply_launch(op_im, map, xin, yacc, scale);
      if (rc) return rc;
      return exchange_reduce(sp, yacc);
    };
    if (run(dxi, dyr, -1.0)) return 1;  // yr -= Ai xi
    if (run(dxr, dyi, +1.0)) return 1;  // yi += Ai xr
  }
  }
  HC_CUDA(cudaMemcpyAsync(yr, dyr, sizeof(double) * nt, kind_out(where), ctx->stream));
  HC_CUDA(cudaMemcpyAsync(yi, dyi, sizeof(double) * nt, kind_out(where), ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hcurl_op_diagonal(hcurl_op *op, double *diag, int where) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  if (ensure_work(op)) return 1;
  if (op_diag_ext(op, op->d_yout)) return 1;
  HC_CUDA(cudaMemcpyAsync(diag, op->d_yout, sizeof(double) * sp->n_tdof, kind_out(where), ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hcurl_op_form_rhs(hcurl_op *op, const double *x_bc, const double *b, double *B, int where) {
  hcurl_space *sp = op->sp;
  hcurl_ctx *ctx = op->ctx;
  HC_CUDA(cudaSetDevice(ctx->device));
  if (ensure_work(op)) return 1;
  const int64_t nt = sp->n_tdof, next = sp->n_ext();
  // d_xin = x_bc restricted to essential dofs, d_yout = b - A_unconstrained d_xin, then B[ess] = x_bc[ess]
  // (x_bc is staged in the operator's second work vector, allocated once: this runs every time step)
  if (!op->d_xin2) HC_CUDA(cudaMalloc(&op->d_xin2, sizeof(double) * next));
  double *xfull = op->d_xin2;
  HC_CUDA(cudaMemcpyAsync(xfull, x_bc, sizeof(double) * nt, kind_in(where), ctx->stream));
  HC_CUDA(cudaMemsetAsync(op->d_xin, 0, sizeof(double) * next, ctx->stream));
  if (op->n_ess > 0) copy_idx_kernel<<<(op->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op->n_ess, op->d_ess, xfull, op->d_xin);
  HC_CUDA(cudaMemcpyAsync(op->d_yout, b, sizeof(double) * nt, kind_in(where), ctx->stream));
  if (op_apply_ext(op, op->d_xin, op->d_yout, false, -1.0, true)) return 1;
  if (op->n_ess > 0) copy_idx_kernel<<<(op->n_ess + 255) / 256, 256, 0, ctx->stream>>>(op->n_ess, op->d_ess, xfull, op->d_yout);
  HC_CUDA(cudaMemcpyAsync(B, op->d_yout, sizeof(double) * nt, kind_out(where), ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int hcurl_op_kernel_stats_reset(hcurl_op *op) { op->ring_n = 0; return 0; }
double hcurl_op_kernel_stats_ms(hcurl_op *op, int *count) {
  cudaSetDevice(op->ctx->device);
  const int n = op->ring_n < hcurl_op::RING ? op->ring_n : hcurl_op::RING;
  if (count) *count = n;
  if (n == 0) return 0.0;
  cudaStreamSynchronize(op->ctx->stream);
  double tot = 0.0;
  for (int i = 0; i < n; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, op->ring0[i], op->ring1[i]) != cudaSuccess) return -1.0;
    tot += ms;
  }
  return tot / n;
}
double hcurl_op_last_kernel_ms(hcurl_op *op) {
  float ms = 0.f;
  cudaSetDevice(op->ctx->device);
  if (cudaEventSynchronize(op->ev1) != cudaSuccess) return -1.0;
  if (cudaEventElapsedTime(&ms, op->ev0, op->ev1) != cudaSuccess) return -1.0;
  return ms;
}

// ------------------------------------------------------------------------ PCG ----
hcurl_pcg *hcurl_pcg_create(hcurl_ctx *ctx, hcurl_op *op, int precond) {
  if (!ctx || !op) { set_error("null context or operator"); return nullptr; }
  if (precond != HCURL_PC_NONE && precond != HCURL_PC_JACOBI && precond != HCURL_PC_CHEBYSHEV) { set_error("unknown preconditioner %d", precond); return nullptr; }
  HC_CUDA_NULL(cudaSetDevice(ctx->device));
  hcurl_pcg *s = new hcurl_pcg;
  s->ctx = ctx; s->op = op; s->precond = precond;
  const size_t bytes = sizeof(double) * (size_t)op->sp->n_ext();
  HC_CUDA_NULL(cudaMalloc(&s->d_r, bytes));
  HC_CUDA_NULL(cudaMalloc(&s->d_p, bytes));
  HC_CUDA_NULL(cudaMalloc(&s->d_s, bytes));
  HC_CUDA_NULL(cudaMalloc(&s->d_x, bytes));
  HC_CUDA_NULL(cudaMalloc(&s->d_b, bytes));
  if (precond == HCURL_PC_JACOBI || precond == HCURL_PC_CHEBYSHEV) HC_CUDA_NULL(cudaMalloc(&s->d_dinv, bytes));
  if (precond == HCURL_PC_CHEBYSHEV) {
    HC_CUDA_NULL(cudaMalloc(&s->d_cres, bytes));
    HC_CUDA_NULL(cudaMalloc(&s->d_cd, bytes));
    HC_CUDA_NULL(cudaMalloc(&s->d_ct, bytes));
    HC_CUDA_NULL(cudaMemset(s->d_cd, 0, bytes));
  }
  HC_CUDA_NULL(cudaEventCreate(&s->ev0));
  HC_CUDA_NULL(cudaEventCreate(&s->ev1));
  return s;
}
void hcurl_pcg_destroy(hcurl_pcg *s) {
  if (!s) return;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaFree(s->d_r); cudaFree(s->d_p); cudaFree(s->d_s); cudaFree(s->d_x); cudaFree(s->d_b); cudaFree(s->d_dinv);
  cudaFree(s->d_cres); cudaFree(s->d_cd); cudaFree(s->d_ct);
  cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1);
  if (s->ev_red) cudaEventDestroy(s->ev_red);
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  cudaFree(s->d_state);
  delete s;
}
int hcurl_pcg_set(hcurl_pcg *s, double rtol, double atol, int max_iter, int print_level) {
  s->rtol = rtol; s->atol = atol; s->max_iter = max_iter; s->print_level = print_level;
  return 0;
}
int hcurl_pcg_update(hcurl_pcg *s) { s->diag_valid = false; return 0; }

int hcurl_pcg_set_chebyshev(hcurl_pcg *s, int order, int power_iters) {
  if (!s || order < 1 || order > 16 || power_iters < 1) { set_error("hcurl_pcg_set_chebyshev: order must be 1..16, power_iters >= 1"); return 1; }
  s->cheb_order = order; s->cheb_power_iters = power_iters; s->diag_valid = false;
  return 0;
}
double hcurl_pcg_lambda_max(const hcurl_pcg *s) { return s ? s->lambda_max : 0.0; }

// read n doubles of the reduction scratch on the host (stream-ordered)
static int read_red(hcurl_ctx *ctx, int n) {
  HC_CUDA(cudaMemcpyAsync(ctx->h_red, ctx->d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// lambda_max(D^-1 A) by the power method: Rayleigh quotient <v, A v> / <v, D v> (oracle/pcg.py: power_lambda_max)
static int cheb_estimate(hcurl_pcg *s) {
  hcurl_ctx *ctx = s->ctx;
  hcurl_op *op = s->op;
  const int64_t nt = op->sp->n_tdof;
  cudaStream_t st = ctx->stream;
  const int g = red_grid(ctx, nt);
  double *S = ctx->d_red, *v = s->d_cd, *w = s->d_ct;
  // distinct streams of the hash on different ranks (the serial case matches the oracle's start vector)
  cheb_seed_kernel<<<g, RED_THREADS, 0, st>>>(nt, (uint32_t)ctx->rank * 0x9e3779b9u, v);
  double lam = 0.0;
  for (int it = 0; it < s->cheb_power_iters; it++) {
    if (op_apply_ext(op, v, w, true, 1.0, false)) return 1;
    cheb_rayleigh_kernel<<<g, RED_THREADS, 0, st>>>(nt, v, w, s->d_dinv, ctx->d_partial, ctx->d_counter, S + 0, 0);
    cheb_rayleigh_kernel<<<g, RED_THREADS, 0, st>>>(nt, v, w, s->d_dinv, ctx->d_partial, ctx->d_counter, S + 1, 1);
    if (allreduce_sum(ctx, S, 2)) return 1;
    if (read_red(ctx, 2)) return 1;
    lam = ctx->h_red[0] / ctx->h_red[1];
    // v = D^-1 w, normalised in the D inner product: <D^-1 w, w>
    cheb_scale_kernel<<<g, RED_THREADS, 0, st>>>(nt, 1.0, s->d_dinv, w, v);
    cheb_rayleigh_kernel<<<g, RED_THREADS, 0, st>>>(nt, v, w, s->d_dinv, ctx->d_partial, ctx->d_counter, S + 0, 0);
    if (allreduce_sum(ctx, S, 1)) return 1;
    if (read_red(ctx, 1)) return 1;
    if (!(ctx->h_red[0] > 0.0)) { set_error("Chebyshev: power method broke down (operator not positive definite?)"); return 1; }
    cheb_scale_kernel<<<g, RED_THREADS, 0, st>>>(nt, 1.0 / sqrt(ctx->h_red[0]), s->d_dinv, w, v);
    ctx->launches += 5;
  }
  HC_CUDA(cudaGetLastError());
  if (!(lam > 0.0)) { set_error("Chebyshev: eigenvalue estimate %g is not positive", lam); return 1; }
  s->lambda_max = lam;
  return 0;
}

// z = q(D^-1 A) D^-1 r: three-term Chebyshev iteration from z = 0 on [0.3, 1.2] * lambda_max (oracle/pcg.py: chebyshev)
static int cheb_apply(hcurl_pcg *s, const double *r, double *z) {
  hcurl_ctx *ctx = s->ctx;
  const int64_t nt = s->op->sp->n_tdof;
  cudaStream_t st = ctx->stream;
  const int g = red_grid(ctx, nt);
  const double lo = 0.3 * s->lambda_max, hi = 1.2 * s->lambda_max;
  const double theta = 0.5 * (hi + lo), delta = 0.5 * (hi - lo), sigma = theta / delta;
  cheb_start_kernel<<<g, RED_THREADS, 0, st>>>(nt, 1.0 / theta, s->d_dinv, r, s->d_cres, s->d_cd, z);
  ctx->launches++;
  double rho = 1.0 / sigma;
  for (int j = 1; j < s->cheb_order; j++) {
    if (op_apply_ext(s->op, s->d_cd, s->d_ct, true, 1.0, false)) return 1;
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    cheb_step_kernel<<<g, RED_THREADS, 0, st>>>(nt, rho_new * rho, 2.0 * rho_new / delta, s->d_dinv, s->d_ct, s->d_cres, s->d_cd, z);
    ctx->launches++;
    rho = rho_new;
  }
  HC_CUDA(cudaGetLastError());
  return 0;
}

// hypre's PCG loop with the Chebyshev polynomial as C; plain sequence, one host read per iteration
static int pcg_solve_chebyshev(hcurl_pcg *s, const double *b, double *x, int zero_guess, int where, int *iters, double *final_rel,
                               int *converged) {
  hcurl_ctx *ctx = s->ctx;
  hcurl_op *op = s->op;
  hcurl_space *sp = op->sp;
  const int64_t nt = sp->n_tdof;
  cudaStream_t st = ctx->stream;
  double *S = ctx->d_red;  // S[0]=gamma S[1]=<s,p> S[2]=gamma_new S[3]=bi_prod
  const int g = red_grid(ctx, nt);
  if (!s->diag_valid) {
    if (op_diag_ext(op, s->d_dinv)) return 1;
    invert_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_dinv);
    if (cheb_estimate(s)) return 1;
    s->diag_valid = true;
  }
  HC_CUDA(cudaMemcpyAsync(s->d_b, b, sizeof(double) * nt, kind_in(where), st));
  if (zero_guess) HC_CUDA(cudaMemsetAsync(s->d_x, 0, sizeof(double) * sp->n_ext(), st));
  else HC_CUDA(cudaMemcpyAsync(s->d_x, x, sizeof(double) * nt, kind_in(where), st));
  // bi_prod = <C b, b>
  if (cheb_apply(s, s->d_b, s->d_s)) return 1;
  dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_s, s->d_b, ctx->d_partial, ctx->d_counter, S + 3);
  if (allreduce_sum(ctx, S + 3, 1)) return 1;
  if (read_red(ctx, 4)) return 1;
  const double bi_prod = ctx->h_red[3];
  if (!(bi_prod > 0.0)) {
    HC_CUDA(cudaMemcpyAsync(x, s->d_b, sizeof(double) * nt, kind_out(where), st));
    HC_CUDA(cudaStreamSynchronize(st));
    if (converged) *converged = 1;
    return 0;
  }
  const double eps = fmax(s->rtol * s->rtol, s->atol * s->atol / bi_prod);
  // r = b - A x ; p = C r ; gamma = <r,p>
  if (!zero_guess) {
    if (op_apply_ext(op, s->d_x, s->d_r, true, 1.0, false)) return 1;
    pcg_init_kernel<<<g, RED_THREADS, 0, st>>>(nt, 1, s->d_b, nullptr, s->d_r, s->d_p, ctx->d_partial, ctx->d_counter, S + 0);  // r = b - r
  } else HC_CUDA(cudaMemcpyAsync(s->d_r, s->d_b, sizeof(double) * nt, cudaMemcpyDeviceToDevice, st));
  if (cheb_apply(s, s->d_r, s->d_p)) return 1;
  dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_r, s->d_p, ctx->d_partial, ctx->d_counter, S + 0);
  if (allreduce_sum(ctx, S + 0, 1)) return 1;
  if (read_red(ctx, 4)) return 1;
  double gamma = ctx->h_red[0];
  int i = 0, conv = 0;
  HC_CUDA(cudaEventRecord(s->ev0, st));
  while (i + 1 <= s->max_iter) {
    i++;
    if (op_apply_ext(op, s->d_p, s->d_s, true, 1.0, false)) return 1;
    dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_s, s->d_p, ctx->d_partial, ctx->d_counter, S + 1);
    if (allreduce_sum(ctx, S + 1, 1)) return 1;
    pcg_xr_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + 0, S + 1, s->d_p, s->d_s, s->d_x, s->d_r);
    if (cheb_apply(s, s->d_r, s->d_s)) return 1;
    dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_r, s->d_s, ctx->d_partial, ctx->d_counter, S + 2);
    ctx->launches += 3;
    if (allreduce_sum(ctx, S + 2, 1)) return 1;
    if (read_red(ctx, 4)) return 1;
    const double sdotp = ctx->h_red[1], gamma_new = ctx->h_red[2];
    if (sdotp == 0.0) break;
    if (s->print_level > 1 && ctx->rank == 0) printf("  hcurl pcg (chebyshev) it %d  rel res %.6e\n", i, sqrt(fabs(gamma_new) / bi_prod));
    if (gamma_new / bi_prod < eps) { gamma = gamma_new; conv = 1; break; }
    if (fabs(gamma_new) < 1e-292) { gamma = gamma_new; break; }
    pcg_dir_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + 2, S + 0, s->d_s, s->d_p);  // p = s + (gamma_new / gamma) p
    HC_CUDA(cudaMemcpyAsync(S + 0, S + 2, sizeof(double), cudaMemcpyDeviceToDevice, st));
    ctx->launches++;
    gamma = gamma_new;
  }
  HC_CUDA(cudaEventRecord(s->ev1, st));
  HC_CUDA(cudaGetLastError());
  HC_CUDA(cudaMemcpyAsync(x, s->d_x, sizeof(double) * nt, kind_out(where), st));
  HC_CUDA(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->last_loop_ms, s->ev0, s->ev1);
  if (iters) *iters = i;
  if (final_rel) *final_rel = sqrt(fabs(gamma) / bi_prod);
  if (converged) *converged = conv;
  if (s->print_level > 0 && ctx->rank == 0)
    printf("hcurl pcg (chebyshev order %d, lambda_max %.4f): %d iterations, final relative residual %.6e%s\n", s->cheb_order, s->lambda_max, i,
           sqrt(fabs(gamma) / bi_prod), conv ? "" : " (not converged)");
  return 0;
}

// The Krylov loop as ONE graph launch: a WHILE conditional node (CUDA 12.4+) whose body is one iteration -- apply,
// <p, A p>, fused update, direction update, pcg_check_kernel -- captured from the same enqueue code the host loop runs.
// No host round trip per iteration: at the reference's problem sizes (TEAM7: 68 547 dofs) an iteration is a handful of
// kernels of a few microseconds each, and the host loop's per-iteration event wait and launch latencies dominate it.
// Single rank, Jacobi / no preconditioner, operators whose apply does not go through the patch kernels (their zeroing
// watermark takes a per-launch host counter).  HCURL_PCG_GRAPH=0 keeps the host loop.
static bool pcg_graph_wanted(const hcurl_pcg *s) {
  static const bool on = [] { const char *e = getenv("HCURL_PCG_GRAPH"); return !(e && e[0] == '0'); }();
  return on && s->ctx->nranks == 1 && s->op->sp->n_nbr == 0 && s->precond != HCURL_PC_CHEBYSHEV && s->print_level <= 1 && !hex_patch_eligible(s->op) &&
         !hex_fast_eligible(s->op);
}
static int pcg_graph_build(hcurl_pcg *s) {
  hcurl_ctx *ctx = s->ctx;
  hcurl_op *op = s->op;
  const int64_t nt = op->sp->n_tdof;
  cudaStream_t st = ctx->stream;
  double *S = ctx->d_red;
  const int g = red_grid(ctx, nt);
  const double *dinv = s->precond == HCURL_PC_JACOBI ? s->d_dinv : nullptr;
  if (s->graph_exec) { cudaGraphExecDestroy(s->graph_exec); s->graph_exec = nullptr; }
  if (s->graph) { cudaGraphDestroy(s->graph); s->graph = nullptr; }
  if (!s->d_state) HC_CUDA(cudaMalloc(&s->d_state, sizeof(int) * 4));
  // one ordinary apply first: lazily built device data (element matrices, constant tables, work vectors) must exist
  // before the capture, which cannot allocate or synchronise
  if (op_apply_ext(op, s->d_p, s->d_s, true, 1.0, false)) return 1;
  HC_CUDA(cudaStreamSynchronize(st));
  HC_CUDA(cudaGraphCreate(&s->graph, 0));
  cudaGraphConditionalHandle handle;
  HC_CUDA(cudaGraphConditionalHandleCreate(&handle, s->graph, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  HC_CUDA(cudaGraphAddNode(&node, s->graph, nullptr, 0, &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  const bool timed = op->timed;
  op->timed = false;  // (timing events cannot be read back from a captured launch)
  HC_CUDA(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  int rc = op_apply_ext(op, s->d_p, s->d_s, true, 1.0, false);
  if (!rc) {
    dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_s, s->d_p, ctx->d_partial, ctx->d_counter, S + 1);
    pcg_update_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + 0, S + 1, dinv, s->d_p, s->d_x, s->d_r, s->d_s, ctx->d_partial, ctx->d_counter, S + 2);
    pcg_dir_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + 2, S + 0, s->d_s, s->d_p);
    pcg_check_kernel<<<1, 1, 0, st>>>(S, s->d_state, handle);
  }
  cudaGraph_t captured = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(st, &captured);
  op->timed = timed;
  if (rc) return 1;
  if (ce != cudaSuccess) { set_error("PCG graph capture failed: %s", cudaGetErrorString(ce)); return 1; }
  HC_CUDA(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
  s->graph_generation = op->generation;
  return 0;
}

int hcurl_pcg_solve(hcurl_pcg *s, const double *b, double *x, int zero_guess, int where, int *iters, double *final_rel,
                    int *converged) {
  if (s->precond == HCURL_PC_CHEBYSHEV) {
    HC_CUDA(cudaSetDevice(s->ctx->device));
    if (iters) *iters = 0;
    if (final_rel) *final_rel = 0.0;
    if (converged) *converged = 0;
    return pcg_solve_chebyshev(s, b, x, zero_guess, where, iters, final_rel, converged);
  }
  hcurl_ctx *ctx = s->ctx;
  hcurl_op *op = s->op;
  hcurl_space *sp = op->sp;
  HC_CUDA(cudaSetDevice(ctx->device));
  const int64_t nt = sp->n_tdof;
  cudaStream_t st = ctx->stream;
  double *S = ctx->d_red;  // S[0]=gamma S[1]=<s,p> S[2]=gamma_new S[3]=bi_prod
  const int g = red_grid(ctx, nt);
  if (s->precond == HCURL_PC_JACOBI && !s->diag_valid) {
    if (op_diag_ext(op, s->d_dinv)) return 1;
    invert_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_dinv);
    s->diag_valid = true;
  }
  const double *dinv = s->precond == HCURL_PC_JACOBI ? s->d_dinv : nullptr;
  HC_CUDA(cudaMemcpyAsync(s->d_b, b, sizeof(double) * nt, kind_in(where), st));
  if (zero_guess) HC_CUDA(cudaMemsetAsync(s->d_x, 0, sizeof(double) * sp->n_ext(), st));
  else HC_CUDA(cudaMemcpyAsync(s->d_x, x, sizeof(double) * nt, kind_in(where), st));
  if (iters) *iters = 0;
  if (final_rel) *final_rel = 0.0;
  if (converged) *converged = 0;
  // bi_prod = <C b, b>
  cdot_kernel<<<g, RED_THREADS, 0, st>>>(nt, dinv, s->d_b, ctx->d_partial, ctx->d_counter, S + 3);
  if (allreduce_sum(ctx, S + 3, 1)) return 1;
  HC_CUDA(cudaMemcpyAsync(ctx->h_red, S, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  const double bi_prod = ctx->h_red[3];
  if (!(bi_prod > 0.0)) {
    // hypre: zero right-hand side -> x = b
    HC_CUDA(cudaMemcpyAsync(x, s->d_b, sizeof(double) * nt, kind_out(where), st));
    HC_CUDA(cudaStreamSynchronize(st));
    if (converged) *converged = 1;
    return 0;
  }
  const double eps = fmax(s->rtol * s->rtol, s->atol * s->atol / bi_prod);
  // r = b - A x ; p = C r ; gamma = <r,p>
  if (!zero_guess) { if (op_apply_ext(op, s->d_x, s->d_r, true, 1.0, false)) return 1; }
  pcg_init_kernel<<<g, RED_THREADS, 0, st>>>(nt, zero_guess ? 0 : 1, s->d_b, dinv, s->d_r, s->d_p, ctx->d_partial, ctx->d_counter, S + 0);
  if (allreduce_sum(ctx, S + 0, 1)) return 1;
  HC_CUDA(cudaMemcpyAsync(ctx->h_red, S, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
  HC_CUDA(cudaStreamSynchronize(st));
  double gamma = ctx->h_red[0];
  int i = 0, conv = 0;
  // The loop follows hypre's PCG step by step; what differs is where the numbers live:
  //  * <p, A p> comes out of the patch kernel itself when it runs (summed per element in modal space);
  //  * gamma alternates between S[0] and S[2], so that beta = gamma_new / gamma is formed on the device and the
  //    direction update is enqueued BEFORE the host has read gamma_new: the host's convergence test (same test,
  //    same iteration count) overlaps with it instead of draining the GPU every iteration.  When the test says
  //    stop, the direction update already in flight only touched p, which is dead by then.
  if (!s->ev_red) HC_CUDA(cudaEventCreateWithFlags(&s->ev_red, cudaEventDisableTiming));
  s->last_used_graph = false;
  if (s->max_iter >= 1 && pcg_graph_wanted(s)) {
    if (!s->graph_exec || s->graph_generation != op->generation) {
      // (the warm-up apply of the build overwrites d_s only; p, r, x and S are as the prelude left them)
      if (pcg_graph_build(s)) return 1;
    }
    const int state[4] = {0, 0, s->max_iter, 0};
    HC_CUDA(cudaMemcpyAsync(s->d_state, state, sizeof(state), cudaMemcpyHostToDevice, st));
    HC_CUDA(cudaMemcpyAsync(S + 4, &eps, sizeof(double), cudaMemcpyHostToDevice, st));
    HC_CUDA(cudaEventRecord(s->ev0, st));
    HC_CUDA(cudaGraphLaunch(s->graph_exec, st));
    HC_CUDA(cudaEventRecord(s->ev1, st));
    int hstate[4];
    HC_CUDA(cudaMemcpyAsync(hstate, s->d_state, sizeof(hstate), cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaMemcpyAsync(ctx->h_red, S, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaMemcpyAsync(x, s->d_x, sizeof(double) * nt, kind_out(where), st));
    HC_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&s->last_loop_ms, s->ev0, s->ev1);
    i = hstate[0]; conv = hstate[1]; gamma = ctx->h_red[0];
    ctx->launches += (int64_t)i * 5;
    s->last_used_graph = true;
    if (iters) *iters = i;
    if (final_rel) *final_rel = sqrt(fabs(gamma) / bi_prod);
    if (converged) *converged = conv;
    if (s->print_level > 0 && ctx->rank == 0)
      printf("hcurl pcg: %d iterations, final relative residual %.6e%s\n", i, sqrt(fabs(gamma) / bi_prod), conv ? "" : " (not converged)");
    return 0;
  }
  HC_CUDA(cudaEventRecord(s->ev0, st));
  int go = 0;  // S[go] holds gamma, S[2 - go] receives gamma_new
  while (i + 1 <= s->max_iter) {
    i++;
    bool fused = false;
    static const bool want_fused = [] { const char *e = getenv("HCURL_PCG_FUSED_DOT"); return !(e && e[0] == '0'); }();
    if (op_apply_ext(op, s->d_p, s->d_s, true, 1.0, false, want_fused ? S + 1 : nullptr, &fused)) return 1;
    if (!fused) { dot_kernel<<<g, RED_THREADS, 0, st>>>(nt, s->d_s, s->d_p, ctx->d_partial, ctx->d_counter, S + 1); ctx->launches++; }
    if (allreduce_sum(ctx, S + 1, 1)) return 1;
    pcg_update_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + go, S + 1, dinv, s->d_p, s->d_x, s->d_r, s->d_s, ctx->d_partial, ctx->d_counter,
                                                 S + 2 - go);
    ctx->launches++;
    if (allreduce_sum(ctx, S + 2 - go, 1)) return 1;
    HC_CUDA(cudaMemcpyAsync(ctx->h_red, S, sizeof(double) * 4, cudaMemcpyDeviceToHost, st));
    HC_CUDA(cudaEventRecord(s->ev_red, st));
    static const bool nospec = [] { const char *e = getenv("HCURL_PCG_NOSPEC"); return e && e[0] == '1'; }();
    if (nospec) HC_CUDA(cudaStreamSynchronize(st));
    pcg_dir_kernel<<<g, RED_THREADS, 0, st>>>(nt, S + 2 - go, S + go, s->d_s, s->d_p);
    ctx->launches++;
    HC_CUDA(cudaEventSynchronize(s->ev_red));
    const double sdotp = ctx->h_red[1], gamma_new = ctx->h_red[2 - go];
    if (s->print_level > 2) printf("  [rank %d] it %d  <p,Ap> %.17g  gamma_new %.17g  gamma %.17g  bi_prod %.17g  eps %.3g\n", ctx->rank, i, sdotp, gamma_new, gamma, bi_prod, eps);
    if (sdotp == 0.0) { break; }
    if (s->print_level > 1 && ctx->rank == 0) printf("  hcurl pcg it %d  rel res %.6e\n", i, sqrt(fabs(gamma_new) / bi_prod));
    if (gamma_new / bi_prod < eps) { gamma = gamma_new; conv = 1; break; }
    if (fabs(gamma_new) < 1e-292) { gamma = gamma_new; break; }
    gamma = gamma_new;
    go = 2 - go;
  }
  HC_CUDA(cudaEventRecord(s->ev1, st));
  HC_CUDA(cudaGetLastError());
  HC_CUDA(cudaMemcpyAsync(x, s->d_x, sizeof(double) * nt, kind_out(where), st));
  HC_CUDA(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&s->last_loop_ms, s->ev0, s->ev1);
  if (iters) *iters = i;
  if (final_rel) *final_rel = sqrt(fabs(gamma) / bi_prod);
  if (converged) *converged = conv;
  if (s->print_level > 0 && ctx->rank == 0)
    printf("hcurl pcg: %d iterations, final relative residual %.6e%s\n", i, sqrt(fabs(gamma) / bi_prod), conv ? "" : " (not converged)");
  return 0;
}
double hcurl_pcg_last_loop_ms(hcurl_pcg *s) { return s->last_loop_ms; }
int hcurl_pcg_last_used_graph(const hcurl_pcg *s) { return s && s->last_used_graph ? 1 : 0; }

// --------------------------------------------------------------- vector helpers ----
double *hcurl_device_alloc(hcurl_ctx *ctx, int64_t n) {
  double *p = nullptr;
  HC_CUDA_NULL(cudaSetDevice(ctx->device));
  HC_CUDA_NULL(cudaMalloc(&p, sizeof(double) * (size_t)(n > 0 ? n : 1)));
  return p;
}
void hcurl_device_free(hcurl_ctx *ctx, double *p) { cudaSetDevice(ctx->device); cudaFree(p); }
int hcurl_copy_to_device(hcurl_ctx *ctx, double *dst, const double *src, int64_t n) {
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int hcurl_copy_to_host(hcurl_ctx *ctx, double *dst, const double *src, int64_t n) {
  HC_CUDA(cudaSetDevice(ctx->device));
  HC_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}
double *hcurl_host_alloc_pinned(int64_t n) {
  double *p = nullptr;
  HC_CUDA_NULL(cudaMallocHost(&p, sizeof(double) * (size_t)(n > 0 ? n : 1)));
  return p;
}
void hcurl_host_free_pinned(double *p) { cudaFreeHost(p); }
int hcurl_dot(hcurl_ctx *ctx, const double *a, const double *b, int64_t n, double *result) {
  HC_CUDA(cudaSetDevice(ctx->device));
  dot_kernel<<<red_grid(ctx, n), RED_THREADS, 0, ctx->stream>>>(n, a, b, ctx->d_partial, ctx->d_counter, ctx->d_red + 8);
  if (allreduce_sum(ctx, ctx->d_red + 8, 1)) return 1;
  HC_CUDA(cudaMemcpyAsync(ctx->h_red + 8, ctx->d_red + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  *result = ctx->h_red[8];
  return 0;
}

double hcurl_fp64_peak_tflops(hcurl_ctx *ctx) {
  if (cudaSetDevice(ctx->device) != cudaSuccess) return -1.0;
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
  double *out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0, ctx->stream);
    dfma_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
  return flops / (best * 1e-3) / 1e12;
}

}  // extern "C"
