// Shared-DOF exchange and scalar all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// Replaces the MPI traffic hidden inside hypre/MFEM on the reference's hot
// path: the ParCSR halo exchange of
// hypre_ParCSRMatrixMatvec, the prolongation P / P^T of
// ParFiniteElementSpace (EquationSystem::FormLinearSystem / RecoverFEMSolution,
// /root/reference/src/equation_systems/equation_system.cpp:132-222) and the
// MPI_Allreduce behind hypre_ParVectorInnerProd in the PCG loop.
//
// NCCL is dlopen'ed so that the library loads and runs on one GPU without it.
// Extended vectors are laid out [owned true dofs | ghosts grouped by owner
// rank], so broadcast receives land in place (zero-copy) and reduce sends
// leave from the ghost block directly; only the owner side packs / unpacks.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "internal.h"

namespace hcurl {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t;
enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

struct Nccl {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  ncclResult_t (*CommInitRank)(void **, int, ncclUniqueId_t, int) = nullptr;
  ncclResult_t (*CommDestroy)(void *) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static Nccl *nccl() {
  static Nccl n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (n.handle) {
#define LD(sym) *(void **)(&n.sym) = dlsym(n.handle, "nccl" #sym)
      LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(AllReduce); LD(AllGather); LD(Send); LD(Recv);
      LD(GroupStart); LD(GroupEnd); LD(GetErrorString);
#undef LD
      if (!n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.Send || !n.Recv) n.handle = nullptr;
    }
  }
  return n.handle ? &n : nullptr;
}

#define HC_NCCL(call)                                                               \
  do {                                                                              \
    ncclResult_t r_ = (call);                                                       \
    if (r_ != 0) {                                                                  \
      set_error("NCCL error %d (%s) in %s", r_, N->GetErrorString ? N->GetErrorString(r_) : "?", #call); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

int comm_unique_id(void *out128) {
  Nccl *N = nccl();
  if (!N) { set_error("NCCL library (libnccl.so.2) not found: multi-GPU unavailable"); return 1; }
  ncclUniqueId_t id;
  HC_NCCL(N->GetUniqueId(&id));
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init(hcurl_ctx *ctx, const void *uid) {
  if (ctx->nranks <= 1) return 0;
  Nccl *N = nccl();
  if (!N) { set_error("NCCL library (libnccl.so.2) not found: multi-GPU unavailable"); return 1; }
  if (!uid) { set_error("nranks > 1 requires a NCCL unique id"); return 1; }
  ncclUniqueId_t id;
  memcpy(&id, uid, 128);
  HC_NCCL(N->CommInitRank(&ctx->comm, ctx->nranks, id, ctx->rank));
  return 0;
}

void comm_destroy(hcurl_ctx *ctx) {
  Nccl *N = nccl();
  if (N && ctx->comm) N->CommDestroy(ctx->comm);
  ctx->comm = nullptr;
}

int allreduce_sum(hcurl_ctx *ctx, double *d_vals, int n) {
  if (ctx->nranks <= 1) return 0;
  Nccl *N = nccl();
  HC_NCCL(N->AllReduce(d_vals, d_vals, (size_t)n, NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream));
  return 0;
}

__global__ void pack_kernel(const int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ v,
                            double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = v[idx[i]];
}
__global__ void unpack_add_kernel(const int32_t n, const int32_t *__restrict__ idx, const double *__restrict__ buf,
                                  double *__restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(v + idx[i], buf[i]);  // an owned dof may be shared with several neighbours
}
__global__ void sign_kernel(const int32_t n, const int8_t *__restrict__ sgn, double *__restrict__ v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && sgn[i] < 0) v[i] = -v[i];
}

// ---- shared-dof exchange through NVLink peer memory ----------------------------------------------------------------------
// The two NCCL send / recv groups of an apply cost ~45 us each on NVSwitch for 1-2 MB of payload: pure latency (rendezvous,
// proxy, several kernels).  Here every rank maps its neighbours' landing buffers and flag words once (CUDA IPC, handles
// all-gathered over the communicator at space creation) and an exchange is two kernels per rank:
//   push:  wait until the neighbour has acknowledged the previous exchange of this kind (its landing area is free), write
//          my values straight into ITS landing buffer over NVLink, system fence, the last block raises the neighbour's
//          `ready` word to the exchange's sequence number;
//   pull:  wait for `ready` from every neighbour that sends to me, move my landing area into place (ghost block of x, or
//          signed add into the owned dofs of y), the last block raises the neighbours' `ack` words.
// No host involvement, no rendezvous; sequence numbers instead of reset flags, so nothing is ever cleared.  Ranks issue
// their exchanges in the same order (SPMD), every wait is for a kernel enqueued earlier on the peer's stream: no cycle.
// A wait that lasts longer than ~4 s traps instead of hanging the device.
struct P2PNbr {
  double *peer_bcast;             // neighbour's landing area for MY owned values (its ghost block order)
  double *peer_reduce;            // neighbour's landing area for MY ghost contributions (its own-list order)
  unsigned long long *peer_flags; // neighbour's flag words for me: [0] bcast ready, [1] bcast ack, [2] reduce ready, [3] reduce ack
  int32_t own0, own1;             // my own-list range for this neighbour
  int32_t gh0, gh1;               // my ghost range owned by this neighbour
};
struct P2PArgs {
  int n_nbr;
  P2PNbr nb[8];
  unsigned long long *flags;      // my flag words: [r * 4 + k] written by neighbour r
  double *land_bcast, *land_reduce;  // my landing areas: [n_ghost], [n_own]
  unsigned int *done;             // last-block counters [4]
};
struct P2PExchange {
  P2PArgs a;
  unsigned long long seq_bcast = 0, seq_reduce = 0;
  double *d_land = nullptr;
  unsigned long long *d_flags = nullptr;
  unsigned int *d_done = nullptr;
  void *opened[16] = {nullptr};
  int n_opened = 0;
};

__device__ __forceinline__ void p2p_wait(const unsigned long long *flag, unsigned long long seq) {
  unsigned long long v;
  for (long long tries = 0;; tries++) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= seq) return;
    __nanosleep(200);
    if (tries > 20000000LL) __trap();  // a peer that never arrives: fail instead of hanging the device
  }
}
__device__ __forceinline__ void p2p_raise(unsigned long long *flag, unsigned long long seq) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
}
__device__ __forceinline__ bool p2p_last_block(unsigned int *counter) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicInc(counter, gridDim.x - 1) == gridDim.x - 1;
  __syncthreads();
  return last;
}

// owner -> ghosts, my side as OWNER: x[own_idx[i]] into the neighbours' ghost landing areas
__global__ void p2p_push_bcast_kernel(const P2PArgs a, const int32_t *__restrict__ own_idx, const double *__restrict__ x,
                                      const unsigned long long seq) {
  if (threadIdx.x < a.n_nbr && a.nb[threadIdx.x].own1 > a.nb[threadIdx.x].own0) p2p_wait(a.flags + threadIdx.x * 4 + 1, seq - 1);
  __syncthreads();
  const int32_t n = a.nb[a.n_nbr - 1].own1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = 0;
    while (i >= a.nb[r].own1) r++;
    a.nb[r].peer_bcast[i - a.nb[r].own0] = x[own_idx[i]];
  }
  if (p2p_last_block(a.done + 0) && threadIdx.x < a.n_nbr && a.nb[threadIdx.x].own1 > a.nb[threadIdx.x].own0)
    p2p_raise(a.nb[threadIdx.x].peer_flags + 0, seq);
}
// ... my side as GHOST HOLDER: landing area -> ghost block of x (with the orientation signs), acknowledge
__global__ void p2p_pull_bcast_kernel(const P2PArgs a, const int8_t *__restrict__ sgn, double *__restrict__ x_ghost, const unsigned long long seq) {
  if (threadIdx.x < a.n_nbr && a.nb[threadIdx.x].gh1 > a.nb[threadIdx.x].gh0) p2p_wait(a.flags + threadIdx.x * 4 + 0, seq);
  __syncthreads();
  const int32_t n = a.nb[a.n_nbr - 1].gh1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double v = __ldcv(a.land_bcast + i);
    x_ghost[i] = (sgn && sgn[i] < 0) ? -v : v;
  }
  if (p2p_last_block(a.done + 1) && threadIdx.x < a.n_nbr && a.nb[threadIdx.x].gh1 > a.nb[threadIdx.x].gh0)
    p2p_raise(a.nb[threadIdx.x].peer_flags + 1, seq);
}
// ghosts -> owner (sum), my side as GHOST HOLDER: signed ghost values into the owners' landing areas
__global__ void p2p_push_reduce_kernel(const P2PArgs a, const int8_t *__restrict__ sgn, const double *__restrict__ y_ghost, const unsigned long long seq) {
  if (threadIdx.x < a.n_nbr && a.nb[threadIdx.x].gh1 > a.nb[threadIdx.x].gh0) p2p_wait(a.flags + threadIdx.x * 4 + 3, seq - 1);
  __syncthreads();
  const int32_t n = a.nb[a.n_nbr - 1].gh1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int r = 0;
    while (i >= a.nb[r].gh1) r++;
    const double v = y_ghost[i];
    a.nb[r].peer_reduce[i - a.nb[r].gh0] = (sgn && sgn[i] < 0) ? -v : v;
  }
  if (p2p_last_block(a.done + 2) && threadIdx.x < a.n_nbr && a.nb[threadIdx.x].gh1 > a.nb[threadIdx.x].gh0)
    p2p_raise(a.nb[threadIdx.x].peer_flags + 2, seq);
}
// ... my side as OWNER: landing area added into the owned dofs (an owned dof may be shared with several neighbours)
__global__ void p2p_pull_reduce_kernel(const P2PArgs a, const int32_t *__restrict__ own_idx, double *__restrict__ y, const unsigned long long seq) {
  if (threadIdx.x < a.n_nbr && a.nb[threadIdx.x].own1 > a.nb[threadIdx.x].own0) p2p_wait(a.flags + threadIdx.x * 4 + 2, seq);
  __syncthreads();
  const int32_t n = a.nb[a.n_nbr - 1].own1;
  for (int32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) atomicAdd(y + own_idx[i], __ldcv(a.land_reduce + i));
  if (p2p_last_block(a.done + 3) && threadIdx.x < a.n_nbr && a.nb[threadIdx.x].own1 > a.nb[threadIdx.x].own0)
    p2p_raise(a.nb[threadIdx.x].peer_flags + 3, seq);
}

// Collective: every rank of the communicator calls it once per space (ranks without neighbours take part in the all-gather).
// On any failure the space keeps the NCCL exchange (returns 0 with sp->p2p == nullptr) unless the collective itself fails.
int p2p_setup(hcurl_space *sp) {
  hcurl_ctx *ctx = sp->ctx;
  if (ctx->nranks <= 1) return 0;
  static const bool off = [] { const char *e = getenv("HCURL_P2P_EXCHANGE"); return e && e[0] == '0'; }();
  Nccl *N = nccl();
  if (off || !N || !N->AllGather || ctx->nranks > 64) return 0;  // (the same on every rank)
  const int32_t n_own = sp->n_nbr ? sp->own_off[sp->n_nbr] : 0, n_ghost = sp->n_ghost;
  P2PExchange *X = new P2PExchange;
  memset(&X->a, 0, sizeof(X->a));
  static const bool dbg = [] { const char *e = getenv("HCURL_P2P_DEBUG"); return e && e[0] == '1'; }();
  auto why = [&](const char *what, cudaError_t e = cudaSuccess) { if (dbg) fprintf(stderr, "[hcurl p2p rank %d] %s%s%s\n", ctx->rank, what, e != cudaSuccess ? ": " : "", e != cudaSuccess ? cudaGetErrorString(e) : ""); };
  bool ok = sp->n_nbr <= 8;  // (a per-rank condition: decided together below)
  ok = ok && cudaMalloc(&X->d_land, sizeof(double) * (size_t)(n_ghost + n_own + 1)) == cudaSuccess;
  ok = ok && cudaMalloc(&X->d_flags, sizeof(unsigned long long) * 32) == cudaSuccess;
  ok = ok && cudaMalloc(&X->d_done, sizeof(unsigned int) * 4) == cudaSuccess;
  if (ok) { cudaMemset(X->d_flags, 0, sizeof(unsigned long long) * 32); cudaMemset(X->d_done, 0, sizeof(unsigned int) * 4); }
  // what every rank publishes: 64 doubles = 512 bytes: [0] ok, [1] n_nbr, [2..9] nbr_rank, [10..18] ghost_off, [19..27] own_off,
  // bytes 256.. : the two IPC handles
  double rec[64];
  memset(rec, 0, sizeof(rec));
  cudaIpcMemHandle_t h_land, h_flags;
  ok = ok && cudaIpcGetMemHandle(&h_land, X->d_land) == cudaSuccess && cudaIpcGetMemHandle(&h_flags, X->d_flags) == cudaSuccess;
  rec[0] = ok ? 1.0 : 0.0;
  rec[1] = ok ? sp->n_nbr : 0;
  for (int r = 0; ok && r < sp->n_nbr; r++) rec[2 + r] = sp->nbr_rank[r];
  for (int r = 0; ok && sp->n_nbr > 0 && r <= sp->n_nbr; r++) { rec[10 + r] = sp->ghost_off[r]; rec[19 + r] = sp->own_off[r]; }
  if (ok) { memcpy((char *)rec + 256, &h_land, sizeof(h_land)); memcpy((char *)rec + 256 + 64, &h_flags, sizeof(h_flags)); }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  double *d_all = nullptr;
  HC_CUDA(cudaMalloc(&d_all, sizeof(double) * 64 * (size_t)(ctx->nranks + 1)));
  HC_CUDA(cudaMemcpyAsync(d_all + 64 * (size_t)ctx->nranks, rec, sizeof(rec), cudaMemcpyHostToDevice, ctx->stream));
  HC_NCCL(N->AllGather(d_all + 64 * (size_t)ctx->nranks, d_all, 64, NCCL_FLOAT64, ctx->comm, ctx->stream));
  std::vector<double> all((size_t)64 * ctx->nranks);
  HC_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(double) * all.size(), cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_all);
  if (!ok) why("local set-up failed (allocation / cudaIpcGetMemHandle / more than 8 neighbours)", cudaGetLastError());
  for (int r = 0; r < ctx->nranks; r++) ok = ok && all[(size_t)64 * r] == 1.0;  // everybody or nobody
  P2PArgs &a = X->a;
  a.n_nbr = sp->n_nbr;
  a.flags = X->d_flags; a.land_bcast = X->d_land; a.land_reduce = X->d_land + n_ghost; a.done = X->d_done;
  for (int r = 0; ok && r < sp->n_nbr; r++) {
    const double *pr = all.data() + (size_t)64 * sp->nbr_rank[r];
    const int pn = (int)pr[1];
    int slot = -1;  // my slot in the neighbour's lists
    for (int q = 0; q < pn; q++) if ((int)pr[2 + q] == ctx->rank) slot = q;
    if (slot < 0) { why("neighbour does not list me"); ok = false; break; }
    const int32_t p_gh0 = (int32_t)pr[10 + slot], p_gh1 = (int32_t)pr[10 + slot + 1], p_ow0 = (int32_t)pr[19 + slot], p_ow1 = (int32_t)pr[19 + slot + 1];
    P2PNbr &nb = a.nb[r];
    nb.own0 = sp->own_off[r]; nb.own1 = sp->own_off[r + 1]; nb.gh0 = sp->ghost_off[r]; nb.gh1 = sp->ghost_off[r + 1];
    if (p_gh1 - p_gh0 != nb.own1 - nb.own0 || p_ow1 - p_ow0 != nb.gh1 - nb.gh0) { why("list lengths of a pair differ"); ok = false; break; }  // lists of a pair must match
    cudaIpcMemHandle_t ph_land, ph_flags;
    memcpy(&ph_land, (const char *)pr + 256, 64);
    memcpy(&ph_flags, (const char *)pr + 256 + 64, 64);
    void *pl = nullptr, *pf = nullptr;
    { const cudaError_t e = cudaIpcOpenMemHandle(&pl, ph_land, cudaIpcMemLazyEnablePeerAccess); if (e != cudaSuccess) { why("cudaIpcOpenMemHandle", e); cudaGetLastError(); ok = false; break; } }
    X->opened[X->n_opened++] = pl;
    { const cudaError_t e = cudaIpcOpenMemHandle(&pf, ph_flags, cudaIpcMemLazyEnablePeerAccess); if (e != cudaSuccess) { why("cudaIpcOpenMemHandle (flags)", e); cudaGetLastError(); ok = false; break; } }
    X->opened[X->n_opened++] = pf;
    int p_nghost = (int)pr[10 + pn];
    nb.peer_bcast = (double *)pl + p_gh0;
    nb.peer_reduce = (double *)pl + p_nghost + p_ow0;
    nb.peer_flags = (unsigned long long *)pf + slot * 4;
  }
  // agree on the outcome (a rank that could not map a peer must not leave the others waiting for it)
  double *d_ok = nullptr;
  HC_CUDA(cudaMalloc(&d_ok, sizeof(double)));
  const double mine = ok ? 0.0 : 1.0;
  HC_CUDA(cudaMemcpyAsync(d_ok, &mine, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  HC_NCCL(N->AllReduce(d_ok, d_ok, 1, NCCL_FLOAT64, NCCL_SUM, ctx->comm, ctx->stream));
  double bad = 0.0;
  HC_CUDA(cudaMemcpyAsync(&bad, d_ok, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  HC_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_ok);
  sp->p2p = X;
  if (bad != 0.0) { why("falling back to the NCCL exchange"); p2p_free(sp); }
  else why("peer-memory exchange active");
  return 0;
}

void p2p_free(hcurl_space *sp) {
  P2PExchange *X = sp->p2p;
  if (!X) return;
  for (int i = 0; i < X->n_opened; i++) cudaIpcCloseMemHandle(X->opened[i]);
  cudaFree(X->d_land); cudaFree(X->d_flags); cudaFree(X->d_done);
  delete X;
  sp->p2p = nullptr;
}

// x_ext: owner -> ghosts
int exchange_broadcast(hcurl_space *sp, double *d_ext, cudaStream_t stream) {
  hcurl_ctx *ctx = sp->ctx;
  if (sp->n_nbr == 0) return 0;
  const cudaStream_t st = stream ? stream : ctx->stream;
  if (sp->p2p) {  // peer-memory path: two kernels, no NCCL
    P2PExchange *X = sp->p2p;
    const unsigned long long seq = ++X->seq_bcast;
    const int32_t no = sp->own_off[sp->n_nbr], ng = sp->n_ghost;
    if (no > 0) p2p_push_bcast_kernel<<<min(4 * ctx->sm_count, (no + 255) / 256), 256, 0, st>>>(X->a, sp->d_own_idx, d_ext, seq);
    if (ng > 0) p2p_pull_bcast_kernel<<<min(4 * ctx->sm_count, (ng + 255) / 256), 256, 0, st>>>(X->a, sp->d_ghost_sign, d_ext + sp->n_tdof, seq);
    ctx->launches += (no > 0) + (ng > 0);
    HC_CUDA(cudaGetLastError());
    return 0;
  }
  Nccl *N = nccl();
  const int32_t n_own = sp->own_off[sp->n_nbr];
  if (n_own > 0) ctx->launches++;
  if (n_own > 0) pack_kernel<<<(n_own + 255) / 256, 256, 0, st>>>(n_own, sp->d_own_idx, d_ext, sp->d_sendbuf);
  HC_NCCL(N->GroupStart());
  for (int r = 0; r < sp->n_nbr; r++) {
    const int32_t ns = sp->own_off[r + 1] - sp->own_off[r], ng = sp->ghost_off[r + 1] - sp->ghost_off[r];
    if (ns > 0) HC_NCCL(N->Send(sp->d_sendbuf + sp->own_off[r], (size_t)ns, NCCL_FLOAT64, sp->nbr_rank[r], ctx->comm, st));
    if (ng > 0) HC_NCCL(N->Recv(d_ext + sp->n_tdof + sp->ghost_off[r], (size_t)ng, NCCL_FLOAT64, sp->nbr_rank[r], ctx->comm, st));
  }
  HC_NCCL(N->GroupEnd());
  if (sp->d_ghost_sign && sp->n_ghost > 0)
    sign_kernel<<<(sp->n_ghost + 255) / 256, 256, 0, st>>>(sp->n_ghost, sp->d_ghost_sign, d_ext + sp->n_tdof);
  HC_CUDA(cudaGetLastError());
  return 0;
}

// y_ext: ghosts -> owner (sum).  signed_values: the entries are components of a vector in the local orientation of
// each dof (flip the ghosts that are oriented against their owner); false for quantities that do not change sign with
// the orientation (the operator diagonal).
int exchange_reduce(hcurl_space *sp, double *d_ext, bool signed_values, cudaStream_t stream) {
  hcurl_ctx *ctx = sp->ctx;
  if (sp->n_nbr == 0) return 0;
  const cudaStream_t st = stream ? stream : ctx->stream;
  if (sp->p2p) {
    P2PExchange *X = sp->p2p;
    const unsigned long long seq = ++X->seq_reduce;
    const int32_t no = sp->own_off[sp->n_nbr], ng = sp->n_ghost;
    if (ng > 0) p2p_push_reduce_kernel<<<min(4 * ctx->sm_count, (ng + 255) / 256), 256, 0, st>>>(X->a, signed_values ? sp->d_ghost_sign : nullptr, d_ext + sp->n_tdof, seq);
    if (no > 0) p2p_pull_reduce_kernel<<<min(4 * ctx->sm_count, (no + 255) / 256), 256, 0, st>>>(X->a, sp->d_own_idx, d_ext, seq);
    ctx->launches += (no > 0) + (ng > 0);
    HC_CUDA(cudaGetLastError());
    return 0;
  }
  Nccl *N = nccl();
  if (signed_values && sp->d_ghost_sign && sp->n_ghost > 0)
    sign_kernel<<<(sp->n_ghost + 255) / 256, 256, 0, st>>>(sp->n_ghost, sp->d_ghost_sign, d_ext + sp->n_tdof);
  HC_NCCL(N->GroupStart());
  for (int r = 0; r < sp->n_nbr; r++) {
    const int32_t ns = sp->own_off[r + 1] - sp->own_off[r], ng = sp->ghost_off[r + 1] - sp->ghost_off[r];
    if (ng > 0) HC_NCCL(N->Send(d_ext + sp->n_tdof + sp->ghost_off[r], (size_t)ng, NCCL_FLOAT64, sp->nbr_rank[r], ctx->comm, st));
    if (ns > 0) HC_NCCL(N->Recv(sp->d_recvbuf + sp->own_off[r], (size_t)ns, NCCL_FLOAT64, sp->nbr_rank[r], ctx->comm, st));
  }
  HC_NCCL(N->GroupEnd());
  const int32_t n_own = sp->own_off[sp->n_nbr];
  if (n_own > 0) ctx->launches++;
  if (n_own > 0) unpack_add_kernel<<<(n_own + 255) / 256, 256, 0, st>>>(n_own, sp->d_own_idx, sp->d_recvbuf, d_ext);
  HC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hcurl
