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

#include <cstring>

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
      LD(GetUniqueId); LD(CommInitRank); LD(CommDestroy); LD(AllReduce); LD(Send); LD(Recv);
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

// x_ext: owner -> ghosts
int exchange_broadcast(hcurl_space *sp, double *d_ext, cudaStream_t stream) {
  hcurl_ctx *ctx = sp->ctx;
  if (sp->n_nbr == 0) return 0;
  const cudaStream_t st = stream ? stream : ctx->stream;
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
