// Internal declarations of libhcurl_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/hcurl_b200.h"

namespace hcurl {

void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what);
#define HC_CUDA(call)                                   \
  do {                                                  \
    if (!::hcurl::cuda_ok((call), #call)) return 1;     \
  } while (0)
#define HC_CUDA_NULL(call)                              \
  do {                                                  \
    if (!::hcurl::cuda_ok((call), #call)) return nullptr; \
  } while (0)

// encoded element->dof map entry: bits 0..29 index into the extended vector
// [owned true dofs | ghosts], bit 30 essential, bit 31 sign flip
static constexpr uint32_t MAP_IDX = 0x3fffffffu;
static constexpr uint32_t MAP_ESS = 0x40000000u;
static constexpr uint32_t MAP_NEG = 0x80000000u;
// patch blobs (patch.cpp): unit index in bits 0..27, bit 28 marks the shared units this patch OWNS (it is the
// lowest-numbered patch touching them: it pre-fills them), bits 29/30 essential flags of the unit's first / second
// double, bit 31 marks units shared with other patches.  Slot codes handed to build_patches carry a 29-bit unit index
// and orientation flags in bits 29..31 instead.
static constexpr uint32_t PATCH_IDX = 0x0fffffffu;
static constexpr uint32_t PATCH_OWNED = 0x10000000u;
static constexpr uint32_t PATCH_ESS0 = 0x20000000u, PATCH_ESS1 = 0x40000000u;
static constexpr uint32_t PATCH_SHARED = 0x80000000u;
static constexpr uint32_t SLOT_IDX = 0x1fffffffu;  // unit index of a slot code
static constexpr uint32_t PAIR_SWAP = 0x20000000u, PAIR_NEG1 = 0x40000000u, PAIR_NEG0 = 0x80000000u;

// ---------------------------------------------------------------- basis ----
void gauss_legendre01(int n, double *x, double *w);
void gauss_lobatto01(int n, double *x);
// B[q*n+i] = l_i(xq[q]), G = derivative
void lagrange_tables(int n, const double *nodes, int nq, const double *xq, double *B, double *G);
int nd_hex_ndof(int p);
// loc[t] = MFEM local index of tensor-ordered basis function t, sgn[t] = +-1
void nd_hex_dof_map(int p, std::vector<int> &loc, std::vector<int> &sgn);

static constexpr int MAXQ = 7;
static constexpr int MAXP1 = 5;
struct HexTables {  // passed by value to kernels (copied to shared memory)
  double Bo[MAXQ * MAXP1];   // [q][i], i < P      open basis at quad points
  double Bc[MAXQ * MAXP1];   // [q][i], i < P+1    closed basis
  double Gc[MAXQ * MAXP1];   // [q][i]             closed basis derivative
  double Dq[MAXQ * MAXQ];    // [q][m] collocation derivative on the quad points
  double Ng[MAXQ * 3];       // [q][a], a < G+1    geometry basis
  double dNg[MAXQ * 3];
  double wq[MAXQ];
};
void build_hex_tables(int p, int q, int g, HexTables &t);

int nd_tet_ndof(int p);
struct TetRef {
  int p, nd, npk;  // npk = nd(nd+1)/2
  // reference matrices, packed upper (i<=j) row-major: [6][npk]
  // pair order (a,b): (0,0),(0,1),(0,2),(1,1),(1,2),(2,2); off-diagonal pairs
  // hold the symmetrised sum  M^{ab}+M^{ba}
  std::vector<double> Mref, Kref;
};
void build_tet_ref(int p, TetRef &r);
void build_tet_ref_q(int p, bool curl, int nq, const double *pts, const double *w, std::vector<double> &out);
// triangle-face dof transformation: R (nf x nf, row-major), x_local = R x_base
void tri_face_transform(int p, const int32_t *local_verts, const int32_t *base_verts, double *R);

// ---------------------------------------------------------------- patches ----
// Host result of build_patches() (patch.cpp); uploaded into a PatchSet by hex_patch.cu.
struct PatchHost {
  int FE = 0, stride = 0, nd = 0;
  int64_t n_patch = 0;
  int64_t n_boundary = 0;         // patches [0, n_boundary) touch units of the inter-rank exchange
  static constexpr int HDR = 32;  // header words per patch blob
  std::vector<int32_t> perm;      // [n_patch*FE] element at each patch position, -1 = padding
  std::vector<uint32_t> blob_off; // [n_patch+1] start of each patch blob in 16-byte units
  std::vector<uint32_t> blob;     // see patch.cpp for the layout
  uint32_t max_blob_words = 0;
  uint32_t max_dofs = 0;          // largest number of unique units of one patch
  // the blob of patch k by lifetime: [blob_off[k], part_off[2k]) header + unit list, [part_off[2k], part_off[2k+1])
  // slot-offset / overflow / dependency lists, [part_off[2k+1], blob_off[k+1]) per-slot table (16-byte units); largest
  // parts in words
  std::vector<uint32_t> part_off;
  uint32_t max_d_words = 0, max_l_words = 0, max_t_words = 0;
  // E-vector swizzle: slot s of patch position j sits at unit s*stride + (j ^ swz_row[s] ^ swz_grp[j / NB]), NB = 128 /
  // unit_bytes (values < NB, so positions stay inside their aligned run of NB: the element side remains conflict-free)
  std::vector<uint8_t> swz_row, swz_grp;
  std::vector<uint32_t> own_off, own_idx;  // the shared units each patch owns: [n_patch+1] offsets into own_idx
};
int build_patches(int nd, int64_t n_elem, int FE, int stride, int unit_bytes, int64_t n_ext, const double *centroid,
                  const uint32_t *slot_code, const uint8_t *exchange_units, PatchHost &out);
// slot s of the rotated per-component layout (hex_patch.cu) -> standard tensor index
int hex_rot_slot_to_tensor(int p, int slot);
// dof pairs of an order-2 hex space (patch.cpp); returns 1 when the map is not made of aligned pairs
int hex_p2_pair_codes(int64_t n_elem, const int32_t *elem_ldof, const int32_t *l2e, uint32_t *code);
// single-dof slot codes of a hex space of order p in tensor order (component-major, x fastest): extended index | bit 31 = sign flip
void hex_slot_codes(int p, int64_t n_elem, const int32_t *elem_ldof, const int32_t *l2e, uint32_t *code);

struct PatchSet {  // device copy
  int FE = 0;
  int unit = 16;                 // bytes per unit of the dof side: 16 = aligned dof pair (order 2), 8 = single dof (order 1)
  int64_t n_patch = 0, n_boundary = 0;
  int32_t *d_perm = nullptr;
  int stride = 0;
  uint32_t *d_blob_off = nullptr, *d_part_off = nullptr, *d_blob = nullptr;
  int64_t blob_words = 0;
  uint32_t max_d_words = 0, max_l_words = 0, max_t_words = 0, max_dofs = 0;
  bool general_orient = false;   // order 2: some element slot pair is flagged other than "reversed line" (all three bits)
  bool stream = false;           // order 2, 64-element patches: the warp-specialised kernel (patch_stream.h)
  uint8_t swz_row[27] = {0}, swz_grp[16] = {0};
  uint32_t *d_own_off = nullptr, *d_own_idx = nullptr;  // owned shared units per patch (zeroed ahead of the sums)
  // split-phase watermark of the zeroing (patch_pipeline.h): one arrival counter per batch of pipeline iterations,
  // never reset: every launch adds its grid size to each, `arrivals` is the running total
  uint32_t *d_count = nullptr;
  int64_t n_count = 0;
  uint32_t arrivals = 0;
};

// 1-D transform tables of the order-2 modal element operator (hex_patch.cu, hex_patch_p2_body.h)
struct ModalTables {
  // closed basis (Gauss-Lobatto nodes 0, 1/2, 1) -> orthonormal Legendre L0, L1, L2:
  //   T = [[a0, a1, a0], [-b, 0, b], [c0, c1, c0]]   (symmetry of the nodes)
  // open basis (2 Gauss-Legendre nodes) -> L0, L1 is [[1/2, 1/2], [-1/2, 1/2]]; the kernel applies
  // [[1, 1], [-1, 1]] (adds only) and folds the two factors 1/2 (trial and test side) into the
  // geometry factors.
  double a0, a1, b, c0, c1;
  double d1, d2;  // L1' = d1 L0, L2' = d2 L1
};
int compute_modal_tables(ModalTables &T);

// Kernel arguments of the patch kernels (patch_pipeline.h; order 2: 16-byte units, order 1: 8-byte units)
struct PatchArgs {
  int64_t n_patch;           // patches [k_begin, n_patch) are processed by this launch
  int64_t k_begin;           // (0 unless accumulating: the zeroing watermark counts whole launches)
  int dot_add;               // fused dot product: add to *dot_out instead of storing
  int ess_identity;          // 1: essential dofs get y = x (and x^2 enters the fused dot); 0: they get 0 (several ranks: the caller writes the rows after the exchange)
  const uint32_t *blob_off;  // [n_patch+1] in 16-byte units
  const uint32_t *part_off;  // [2 n_patch] begin of the lists and of the per-slot table of every patch (16-byte units, absolute)
  const uint32_t *blob;      // patch blobs (patch.cpp)
  const double *edata;       // [n_patch][12][FE] affine geometry factors in patch order
  const double *x;
  double *y;
  double scale;
  uint32_t d_cap_bytes, l_cap_bytes, t_cap_bytes;  // shared-memory room for one header + unit list / set of lists / table (multiples of 16)
  uint32_t xs_cap_bytes;     // room for the x staging vector
  const uint32_t *own_off, *own_idx;  // the shared units each patch owns
  uint32_t *count;           // arrival counters of the zeroing batches
  uint32_t count_target;     // value every counter reaches when all CTAs of this launch have arrived
  // fused <x, y> (PCG's <p, A p>): per-CTA partial sums, arrival counter, result (null = not wanted)
  double *dot_partial;
  unsigned int *dot_counter;
  double *dot_out;
  // EY position of (slot row s, element e) is s*(FE+1) + (e ^ swz_row[s] ^ swz_grp[e / NB]): chosen by patch.cpp so
  // that the dof side's reads of EY spread over the banks
  uint8_t swz_row[27], swz_grp[16];
  ModalTables modal;         // order 2: 1-D transform tables (kernel parameters live in the constant bank)
};

// ----------------------------------------------------------------- NCCL ----
struct Nccl;  // dlopen'ed function table (comm.cpp)

}  // namespace hcurl

namespace hcurl { struct P2PExchange; }

struct hcurl_ctx {
  int device = 0;
  int rank = 0, nranks = 1;
  cudaStream_t stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // copy streams of the pipelined batch apply, created on first use
  void *comm = nullptr;  // ncclComm_t
  double *d_red = nullptr;   // device reduction scratch (few doubles)
  double *h_red = nullptr;   // pinned host mirror
  unsigned int *d_counter = nullptr;
  double *d_partial = nullptr;  // per-block partial sums
  int n_partial = 0;
  int sm_count = 148;
  cudaEvent_t mark[2] = {nullptr, nullptr};
  cudaEvent_t ev_x[4] = {nullptr, nullptr, nullptr, nullptr};  // main <-> comm stream hand-offs of an overlapped apply
  int64_t launches = 0;
  // per-context (= per-device) launch state of the patch kernels: shared memory opted into / CTAs per SM, by variant
  static constexpr int N_KVAR = 16;
  size_t kattr_smem[N_KVAR] = {0};
  int kattr_per_sm[N_KVAR] = {0};
  bool coop_ok = false;
  int max_smem_optin = -1;  // cudaDevAttrMaxSharedMemoryPerBlockOptin, queried on first use
};

struct hcurl_space {
  hcurl_ctx *ctx = nullptr;
  int geom = 0, geom_order = 1, order = 1;
  int64_t n_elem = 0;
  int nd = 0;
  int ngn = 0;  // geometry nodes per element
  int32_t n_ldof = 0, n_tdof = 0, n_ghost = 0;
  int64_t n_ext() const { return (int64_t)n_tdof + n_ghost; }
  double *d_xyz = nullptr;     // [n_elem][ngn][3]
  uint32_t *d_map = nullptr;   // hex: [n_elem][nd] tensor order; tet: [nd][n_elem_pad]
  int64_t n_elem_pad = 0;
  // tets, order>=2: per element face transformation blocks, [n_elem][4][nf*nf]
  double *d_faceR = nullptr;
  int nf = 0;
  // exchange
  int n_nbr = 0;
  std::vector<int> nbr_rank;
  std::vector<int32_t> own_off, ghost_off;  // [n_nbr+1]
  int32_t *d_own_idx = nullptr;             // ext indices of own-shared dofs (concatenated)
  int8_t *d_ghost_sign = nullptr;           // [n_ghost] or null
  double *d_sendbuf = nullptr, *d_recvbuf = nullptr;  // size own_off.back()
  hcurl::P2PExchange *p2p = nullptr;  // shared-dof exchange through NVLink peer memory (comm.cu), null = NCCL send / recv
  std::vector<int32_t> ldof_to_ext;  // host copy (for essential dofs etc.)
  std::vector<int32_t> own_ext;      // host copy of d_own_idx
  // fast order-2 path (hex_fast.cu): dof-major map / vertices, affine classification
  bool fast_ready = false, all_affine = false;
  uint32_t *d_mapT = nullptr;
  double *d_xyzT = nullptr;
  // patch path (hex_patch.cu): built at creation for order-2 trilinear hexes
  bool has_patches = false;
  hcurl::PatchSet patch;
  double *d_xyzP = nullptr;  // [24][n_patch*FE] vertices in patch order (non-affine meshes)
};

struct hcurl_op {
  hcurl_ctx *ctx = nullptr;
  hcurl_space *sp = nullptr;
  int qc = 0, qm = 0;      // points per dim for curl / mass (hex)
  bool single_rule = false;
  int coef_layout = HCURL_COEF_PER_ELEM;
  bool has_alpha = false, has_beta = false;
  double *d_alpha = nullptr, *d_beta = nullptr;
  int64_t n_alpha = 0, n_beta = 0;
  uint32_t *d_map_ess = nullptr;  // copy of the map with essential flags (or null = none set)
  int32_t *d_ess = nullptr;       // essential ext indices
  int32_t n_ess = 0;
  // tets: packed element matrices [npk][n_elem_pad]
  double *d_emat = nullptr;
  bool emat_valid = false;
  // tets, PER_QPT coefficients: the caller's quadrature rules (hcurl_op_set_tet_rules) as per-point reference matrices
  int tet_nqc = 0, tet_nqm = 0;
  double *d_refq_k = nullptr, *d_refq_m = nullptr;  // [nq][6][npk]
  // fast order-2 hex path: per-element affine data [12][n_pad], dof-major essential map
  double *d_edata = nullptr;
  bool edata_valid = false;
  uint32_t *d_mapT_ess = nullptr;
  uint32_t *d_blob_ess = nullptr;  // copy of the patch blobs with essential flags on the dof words
  double *d_edataP = nullptr;      // [n_patch][12][FE] affine geometry factors in patch order
  bool edataP_valid = false;
  int kernel_sel = 0;              // 0 auto, 1 generic, 2 element-wise fast path, 3 patch path
  int last_kernel = 0;             // which of them the most recent apply launched
  bool force_generic = false;
  // work vectors (extended length)
  int generation = 0;  // bumped whenever device buffers the launches refer to may have been reallocated (coefficients, essential dofs, kernel choice)
  double *d_xin = nullptr, *d_yout = nullptr, *d_tmp = nullptr;
  double *d_xin2 = nullptr, *d_yout2 = nullptr;  // second staging pair of the pipelined batch apply (hcurl_op_apply_batch)
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_run[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  static constexpr int RING = 256;
  cudaEvent_t ring0[RING], ring1[RING];
  int ring_n = 0;      // samples recorded since reset (may exceed RING)
  bool ring_ok = false;
};

struct hcurl_pcg {
  hcurl_ctx *ctx = nullptr;
  hcurl_op *op = nullptr;
  int precond = HCURL_PC_JACOBI;
  double rtol = 1e-9, atol = 1e-16;
  int max_iter = 1000, print_level = 0;
  double *d_dinv = nullptr, *d_r = nullptr, *d_p = nullptr, *d_s = nullptr, *d_x = nullptr, *d_b = nullptr;
  bool diag_valid = false;
  // Chebyshev preconditioner: polynomial order, power iterations, eigenvalue estimate, work vectors (extended length)
  int cheb_order = 3, cheb_power_iters = 10;
  double lambda_max = 0.0;
  double *d_cres = nullptr, *d_cd = nullptr, *d_ct = nullptr;
  float last_loop_ms = 0.f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_red = nullptr;
  // whole Krylov loop as one CUDA graph with a device-side WHILE node (api.cu: pcg_graph_*): built on first use for the
  // operator's current generation, replayed by later solves
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  int graph_generation = -1;
  int *d_state = nullptr;  // [0] iterations done, [1] converged, [2] iteration limit
  bool last_used_graph = false;
};

namespace hcurl {
// Set-up upload from pageable host memory, ordered with the context's (non-blocking) stream: the kernels that consume
// the buffer are launched on that stream, which has no implicit ordering with the legacy default stream.
inline cudaError_t h2d_sync(hcurl_ctx *ctx, void *dst, const void *src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
  return e != cudaSuccess ? e : cudaStreamSynchronize(ctx->stream);
}
// kernels / drivers implemented in the .cu files (all enqueue on ctx->stream)
int hex_apply_launch(hcurl_op *op, const uint32_t *d_map, const double *d_x, double *d_y, double scale);
int hex_diag_launch(hcurl_op *op, double *d_diag);
bool hex_fast_eligible(const hcurl_op *op);
int hex_fast_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale);
int hex_patch_build(hcurl_space *sp, const double *elem_xyz, const int32_t *elem_ldof, const int32_t *l2e);
void hex_patch_free(hcurl_space *sp);
bool hex_patch_eligible(const hcurl_op *op);
// y (+)= scale * A x through the patch path: every dof of y the space's elements touch is written (nothing has to be
// cleared before), essential dofs get y = x when `constrained`
// d_dot (optional, device): receives <x, y> over this rank's elements and owned essential dofs, computed inside the kernel
int hex_patch_launch(hcurl_op *op, bool constrained, const double *d_x, double *d_y, double scale, bool accumulate,
                     double *d_dot = nullptr);
int hex_patch_flag_essential(hcurl_op *op, const double *d_flag);
// the patch kernels (hex_patch_kernels.cu): device environment of patch_pipeline.h, cooperative launch
int hex_patch_kernel_launch(hcurl_op *op, const hcurl::PatchArgs &a, bool accumulate);
size_t patch_smem_bytes(int unit, int FE, uint32_t d_cap, uint32_t l_cap, uint32_t t_cap, uint32_t xs_cap);
size_t patch_stream_smem_bytes(int FE, uint32_t d_cap, uint32_t l_cap, uint32_t t_cap, uint32_t xs_cap);
int hex_fast_prepare_space(hcurl_space *sp);
int hex_all_affine(hcurl_space *sp, bool *all_affine);
int tet_build_emat(hcurl_op *op);
int tet_set_rules(hcurl_op *op, int nqc, const double *pc, const double *wc, int nqm, const double *pm, const double *wm);
int tet_apply_launch(hcurl_op *op, const uint32_t *d_map, const double *d_x, double *d_y, double scale);
int tet_diag_launch(hcurl_op *op, double *d_diag);
int tet_apply_complex_launch(hcurl_op *op_re, hcurl_op *op_im, const uint32_t *d_map, const double *d_xr, const double *d_xi, double *d_yr, double *d_yi);
// exchange (comm.cu): broadcast owner->ghost on x_ext, reduce ghost->owner on y_ext
// stream: where the exchange is enqueued (null = the context's main stream)
int exchange_broadcast(hcurl_space *sp, double *d_ext, cudaStream_t stream = nullptr);
int exchange_reduce(hcurl_space *sp, double *d_ext, bool signed_values = true, cudaStream_t stream = nullptr);
int allreduce_sum(hcurl_ctx *ctx, double *d_vals, int n);
int p2p_setup(hcurl_space *sp);   // collective over the communicator: every rank calls it at space creation
void p2p_free(hcurl_space *sp);
int p2p_setup(hcurl_space *sp);   // collective over the communicator: every rank calls it at space creation
void p2p_free(hcurl_space *sp);
int comm_init(hcurl_ctx *ctx, const void *uid);
void comm_destroy(hcurl_ctx *ctx);
int comm_unique_id(void *out128);
// the apply on extended device vectors, including exchange and essential handling
// d_dot (optional, device scalar): on return of the stream work holds <x, y> over this rank's share (elements + owned
// essential dofs) when the patch kernel ran (*dot_done = true); otherwise untouched
int op_apply_ext(hcurl_op *op, double *d_x_ext, double *d_y_ext, bool constrained, double scale, bool accumulate,
                 double *d_dot = nullptr, bool *dot_done = nullptr);
int op_diag_ext(hcurl_op *op, double *d_diag_ext);
}  // namespace hcurl
