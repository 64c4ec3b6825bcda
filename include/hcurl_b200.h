/*
 * hcurl_b200.h -- C ABI of libhcurl_b200.so
 *
 * B200-native (sm_100a) matrix-free apply of the Nedelec H(curl) operator
 *
 *      A = (alpha curl u, curl v) + (beta u, v)
 *
 * and a device-resident Jacobi-preconditioned CG around it.  This is the
 * drop-in boundary for ONE hot path of aurora-multiphysics/hephaestus
 * (SURVEY.md section 8).  Every entry point names the reference interface it
 * replaces (file:line under /root/reference).  Plain pointers and sizes only;
 * no C++ or torch types cross this boundary.
 *
 * Conventions
 *  - All functions returning int return 0 on success, non-zero on error;
 *    hcurl_last_error() then describes the failure (thread-local string).
 *    The reference's error behaviour is MFEM_ABORT; the C++ adaptor
 *    (adaptor/heph_gpu_operator.hpp) turns non-zero into MFEM_ABORT.
 *  - Handles are opaque, one set per rank/GPU, not thread-safe (the reference
 *    is single-threaded per MPI rank), freed by the matching *_destroy.
 *  - Host pointers are borrowed for the duration of the call only.
 *  - Vectors are FP64.  Local indices are int32, global sizes int64.
 *  - There is NO CPU fallback: every compute entry point fails if no CUDA
 *    device is usable.
 */
#ifndef HCURL_B200_H
#define HCURL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcurl_ctx hcurl_ctx;
typedef struct hcurl_space hcurl_space;
typedef struct hcurl_op hcurl_op;
typedef struct hcurl_pcg hcurl_pcg;
typedef struct hcurl_interp hcurl_interp;

/* element geometry kinds */
enum { HCURL_GEOM_HEX = 0, HCURL_GEOM_TET = 1 };
/* coefficient layouts: one value per element, or one per quadrature point (curl rule for alpha, mass rule for beta;
 * hexes: x-fastest tensor order of the operator's rule; tets: the caller's rules, hcurl_op_set_tet_rules). */
enum { HCURL_COEF_PER_ELEM = 0, HCURL_COEF_PER_QPT = 1 };
/* preconditioners: HCURL_PC_JACOBI reproduces mfem::HypreDiagScale; HCURL_PC_CHEBYSHEV is a Jacobi-scaled Chebyshev
 * polynomial (no reference counterpart: the device-side stand-in for the AMS preconditioner, which is out of scope) */
enum { HCURL_PC_NONE = 0, HCURL_PC_JACOBI = 1, HCURL_PC_CHEBYSHEV = 2 };
/* where a vector argument lives.  HCURL_DEVICE_EXT (apply entry points only): device vectors with room for
 * hcurl_space_ext_size() entries -- the owned true dofs followed by the rank's ghost dofs -- which the
 * library uses in place (the ghost tail of x is overwritten by the shared-dof exchange, that of y is
 * scratch); on one GPU it equals HCURL_DEVICE.  It saves the two device copies per apply that
 * HCURL_DEVICE needs on ranks that hold ghosts. */
enum { HCURL_HOST = 0, HCURL_DEVICE = 1, HCURL_DEVICE_EXT = 2 };

/* library / build info; never touches the GPU */
const char *hcurl_version(void);
const char *hcurl_last_error(void);

/* ---- context -------------------------------------------------------------
 * One per rank.  `device` is the CUDA ordinal.  For multi-GPU runs pass the
 * 128-byte NCCL unique id obtained from hcurl_comm_unique_id() on rank 0 and
 * broadcast by the host (MPI_Bcast in the reference's world,
 * /root/reference/examples/team7/Main.cpp:155 creates the ParMesh on
 * MPI_COMM_WORLD); pass NULL / nranks = 1 for a single GPU.
 * Replaces: the implicit MPI communicator inside mfem::ParMesh /
 * HypreParMatrix on the hot path. */
int hcurl_comm_unique_id(void *id128);
hcurl_ctx *hcurl_ctx_create(int device, const void *nccl_unique_id128, int rank, int nranks);
void hcurl_ctx_destroy(hcurl_ctx *ctx);
int hcurl_ctx_device(const hcurl_ctx *ctx);
int hcurl_ctx_synchronize(hcurl_ctx *ctx);
/* Measurement helpers (bench.py): CUDA events recorded on the stream every kernel of
 * this context is launched on; elapsed time between mark 0 and mark 1 in ms (syncs). */
int hcurl_ctx_mark(hcurl_ctx *ctx, int which);
double hcurl_ctx_elapsed_ms(hcurl_ctx *ctx);
/* number of library kernels launched on this context since creation */
int64_t hcurl_ctx_launch_count(const hcurl_ctx *ctx);

/* ---- finite element space --------------------------------------------------
 * Replaces: the data `ParFiniteElementSpace` feeds to the assembly in
 * `EquationSystem::BuildBilinearForms`
 * (/root/reference/src/equation_systems/equation_system.cpp:324-347) for an
 * "ND_3D_P<order>" collection
 * (/root/reference/src/problem_builders/problem_builder_base.cpp:98-127).
 *
 *  geom        HCURL_GEOM_HEX | HCURL_GEOM_TET
 *  geom_order  1 (8 / 4 corner nodes per element) or 2 (hex only: 27 nodes,
 *              tensor order x-fastest on {0,1/2,1}^3)
 *  order       Nedelec order p (hex 1..4, tet 1..3)
 *  elem_xyz    [n_elem][n_geom_nodes][3] doubles: element node coordinates
 *              (corner order = MFEM / Exodus HEX8 / TETRA4)
 *  elem_ldof   [n_elem][nd] signed local dof ids in MFEM's element-local order
 *              (edges | faces | interior); k >= 0 plain, -1-k sign flipped --
 *              exactly FiniteElementSpace::GetElementDofs.
 *  elem_face_base  tets with order >= 2 only, else NULL: [n_elem][4][3] global
 *              vertex ids of each local face in the mesh's own (base) face
 *              vertex order (Mesh::GetFaceVertices), and
 *  elem_verts  [n_elem][4] global vertex ids of the element; the library
 *              derives MFEM's ND triangle-face DofTransformation from them.
 *  n_ldof      number of local (L-vector) dofs on this rank
 *
 * Parallel layout (ParFiniteElementSpace::GetLocalTDofNumber /
 * GroupCommunicator semantics).  Pass n_nbr = 0 and
 * ldof_ltdof = NULL on a single rank (T-vector == L-vector).
 *  ldof_ltdof  [n_ldof] local true-dof index of each ldof this rank OWNS,
 *              -1 for ldofs owned by another rank
 *  n_tdof      number of true dofs owned by this rank
 *  nbr_rank    [n_nbr] neighbour ranks
 *  own_off     [n_nbr+1] offsets into own_ldof: ldofs this rank owns that
 *              neighbour r also holds (this rank SENDS their values on
 *              broadcast and RECEIVES partial sums on reduce)
 *  ghost_off   [n_nbr+1] offsets into ghost_ldof/ghost_sign: ldofs owned by
 *              neighbour r (received on broadcast, sent on reduce).  Both
 *              sides must list a shared pair of ranks' dofs in the same order.
 *  ghost_sign  [..] +1/-1: orientation of the local dof relative to the
 *              owner's (ND sign flips across ranks), or NULL for all +1
 */
hcurl_space *hcurl_space_create(hcurl_ctx *ctx, int geom, int geom_order, int order, int64_t n_elem,
                                const double *elem_xyz, const int32_t *elem_ldof,
                                const int32_t *elem_verts, const int32_t *elem_face_base,
                                int32_t n_ldof, const int32_t *ldof_ltdof, int32_t n_tdof, int n_nbr,
                                const int32_t *nbr_rank, const int32_t *own_off, const int32_t *own_ldof,
                                const int32_t *ghost_off, const int32_t *ghost_ldof,
                                const int8_t *ghost_sign);
/* length of the extended vector [owned true dofs | ghosts] of this rank */
int64_t hcurl_space_ext_size(const hcurl_space *sp);
void hcurl_space_destroy(hcurl_space *sp);
int32_t hcurl_space_ndof_local(const hcurl_space *sp);
int32_t hcurl_space_ndof_true(const hcurl_space *sp);
int hcurl_space_elem_ndof(const hcurl_space *sp);

/* ---- operator ----------------------------------------------------------------
 * A = K(alpha) + M(beta) with essential rows/cols replaced by identity.
 * Replaces: CurlCurlKernel::Apply (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24),
 * VectorFEMassKernel::Apply (/root/reference/src/kernels/vector_fe_mass_kernel.cpp:21-24),
 * blf->Assemble() + FormLinearSystem + HypreParMatrixFromBlocks
 * (/root/reference/src/equation_systems/equation_system.cpp:132-190, :345) and the
 * operator the result is used as: EquationSystem::Mult / GetGradient (ibid. :200-210).
 *
 *  quad_pts_per_dim  0 = MFEM's full-assembly rules (curl-curl: p+1 points per
 *                    dimension; mass: p+2 on trilinear hexes, p+3 on HEX27;
 *                    exact rules on straight tets), or p+1 / p+2 to force one
 *                    Gauss rule for both terms on hexes (p+1 is exact on
 *                    affine hexes with per-element-constant coefficients);
 *                    HCURL_QUAD_AUTO = the p+1 rule when every element of a
 *                    trilinear hex mesh is a parallelepiped (there it equals
 *                    MFEM's rules to round-off and unlocks the patch kernels),
 *                    MFEM's rules otherwise.
 */
enum { HCURL_QUAD_MFEM = 0, HCURL_QUAD_AUTO = -1 };
hcurl_op *hcurl_op_create(hcurl_ctx *ctx, hcurl_space *sp, int quad_pts_per_dim);
void hcurl_op_destroy(hcurl_op *op);
/* Kernel selection: 0 = automatic (fastest eligible: patch kernels for order-2 trilinear hexes
 * with one p+1 rule and per-element coefficients), 1 = always the generic sum-factorised
 * kernel, 2 = element-wise register-blocked kernel, 3 = patch kernel.  All compute the same
 * operator; tests compare them. */
int hcurl_op_set_kernel(hcurl_op *op, int which);
/* which kernel the most recent apply launched (1, 2, 3 as above; 0 = none yet) */
int hcurl_op_last_kernel(const hcurl_op *op);
/* Tets with coefficients that vary inside an element (mfem::FunctionCoefficient, as in the reference's MMS tests,
 * /root/reference/test/integration/test_aform_source.cpp:32-37): the caller hands over the quadrature rules MFEM's
 * integrators use on the element -- IntRules.Get(Geometry::TETRAHEDRON, 2p-2) for CurlCurlIntegrator, order 2p for
 * VectorFEMassIntegrator -- so that MFEM's simplex tables are not restated in the library; then PER_QPT
 * values alpha [n_elem][nq_curl], beta [n_elem][nq_mass] through hcurl_op_set_coefficients.  pts [nq][3] reference
 * coordinates, w [nq] weights (sum 1/6).  The element matrices are rebuilt on the device from per-point reference
 * matrices whenever the coefficients change. */
int hcurl_op_set_tet_rules(hcurl_op *op, int nq_curl, const double *pts_curl, const double *w_curl, int nq_mass,
                           const double *pts_mass, const double *w_mass);
/* Number of coefficient values expected per element for PER_QPT layout. */
int hcurl_op_qpts_curl(const hcurl_op *op);
int hcurl_op_qpts_mass(const hcurl_op *op);
/* alpha or beta may be NULL = term absent (magnetostatic: beta NULL,
 * /root/reference/src/formulations/Magnetostatic/statics_formulation.cpp:145-148;
 * mass solves: alpha NULL, /root/reference/src/auxsolvers/scaled_vector_gridfunction_aux.cpp:51-60).
 * Cheap to call every time step (Coefficients::SetTime,
 * /root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:43;
 * dt change, equation_system.cpp:413-426): no rebuild, values only. */
int hcurl_op_set_coefficients(hcurl_op *op, int layout, const double *alpha, const double *beta,
                              int where);
/* Essential true dofs (EquationSystem::_ess_tdof_lists,
 * /root/reference/src/equation_systems/equation_system.hpp:94): rows/cols -> identity. */
int hcurl_op_set_essential_tdofs(hcurl_op *op, const int32_t *ess_tdof, int32_t n_ess);
/* y = A x on T-vectors (EquationSystem::Mult, equation_system.cpp:200-204). */
int hcurl_op_apply(hcurl_op *op, const double *x, double *y, int where);
/* y = K(alpha)+M(beta) applied WITHOUT essential elimination, scaled and
 * accumulated: y = scale * A x + (accumulate ? y : 0).  This is the
 * WeakCurlCurlKernel RHS term  lf -= K(alpha) u_n
 * (/root/reference/src/kernels/weak_curl_curl_kernel.cpp:14-31) and the
 * boundary lift  B -= A_e x_bc  of FormLinearSystem. */
int hcurl_op_apply_unconstrained(hcurl_op *op, const double *x, double *y, double scale,
                                 int accumulate, int where);
/* y_j = A x_j, j = 0 .. n-1 (mfem::Operator::ArrayMult -- EquationSystem is an mfem::Operator,
 * /root/reference/src/equation_systems/equation_system.hpp:15-24; its Mult is equation_system.cpp:200-204).
 * With HCURL_HOST buffers (pinned for the copies to be asynchronous) the upload of x_{j+1}, the apply of
 * x_j and the download of y_{j-1} overlap, so both directions of the host link are busy at once. */
int hcurl_op_apply_batch(hcurl_op *op, int n, const double *const *x, double *const *y, int where);
/* Complex (HERMITIAN convention) block apply [[Ar, -Ai], [Ai, Ar]]
 * (/root/reference/src/formulations/ComplexMaxwell/complex_maxwell_formulation.cpp:181-217):
 * yr = Ar xr - Ai xi,  yi = Ai xr + Ar xi.  op_re and op_im share a space. */
int hcurl_op_apply_complex(hcurl_op *op_re, hcurl_op *op_im, const double *xr, const double *xi,
                           double *yr, double *yi, int where);
/* diag(A) with essential entries = 1 (what mfem::HypreDiagScale reads). */
int hcurl_op_diagonal(hcurl_op *op, double *diag, int where);
/* FormLinearSystem RHS (equation_system.cpp:149-152):
 * B = b - A_unconstrained * x_bc(ess only), B[ess] = x[ess]. */
int hcurl_op_form_rhs(hcurl_op *op, const double *x_bc, const double *b, double *B, int where);
/* average device time (ms) of the element kernel(s) of the most recent apply */
double hcurl_op_last_kernel_ms(hcurl_op *op);
/* element-kernel time statistics since the last reset: CUDA events bracket the element
 * kernel(s) of every apply (ring of 256 pairs).  Returns the mean ms, *count = samples. */
int hcurl_op_kernel_stats_reset(hcurl_op *op);
double hcurl_op_kernel_stats_ms(hcurl_op *op, int *count);

/* ---- Krylov solver ---------------------------------------------------------------
 * hypre PCG semantics (two_norm = 0) with optional Jacobi preconditioner.
 * Replaces: mfem::HyprePCG + mfem::HypreDiagScale as configured by
 * DefaultJacobiPCGSolver (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57) and
 * ProblemBuilder::ConstructJacobianSolverWithOptions
 * (/root/reference/src/problem_builders/problem_builder_base.cpp:220-254), called from
 * TimeDomainEquationSystemProblemOperator::ImplicitSolve
 * (/root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:46-48). */
hcurl_pcg *hcurl_pcg_create(hcurl_ctx *ctx, hcurl_op *op, int precond);
void hcurl_pcg_destroy(hcurl_pcg *pcg);
int hcurl_pcg_set(hcurl_pcg *pcg, double rtol, double atol, int max_iter, int print_level);
/* Re-read the operator (new coefficients / essential dofs): recomputes the
 * Jacobi diagonal.  mfem::Solver::SetOperator. */
int hcurl_pcg_update(hcurl_pcg *pcg);
/* Solve A x = b.  zero_initial_guess = !iterative_mode.  Outputs iteration
 * count, final relative residual sqrt(<Cr,r>/<Cb,b>), converged flag. */
int hcurl_pcg_solve(hcurl_pcg *pcg, const double *b, double *x, int zero_initial_guess, int where,
                    int *iters, double *final_rel_res, int *converged);
/* HCURL_PC_CHEBYSHEV: z = q(D^-1 A) D^-1 r, q the Chebyshev polynomial of degree order-1 for the interval
 * [0.3, 1.2] * lambda_max(D^-1 A) (the interval of mfem::OperatorChebyshevSmoother / hypre's Chebyshev smoother),
 * order-1 operator applies per preconditioner application; lambda_max by power_iters steps of the power method whenever
 * the operator changes (hcurl_pcg_update).  Defaults: order 3, 10 power iterations. */
int hcurl_pcg_set_chebyshev(hcurl_pcg *pcg, int order, int power_iters);
/* the estimate of lambda_max(D^-1 A) used by the most recent solve (0 before the first) */
double hcurl_pcg_lambda_max(const hcurl_pcg *pcg);
/* device time (ms) spent inside the iteration loop of the most recent solve */
double hcurl_pcg_last_loop_ms(hcurl_pcg *pcg);
/* 1 when the last solve ran its Krylov loop as one CUDA graph (device-side WHILE node: no host round trip per
 * iteration; single rank, Jacobi / no preconditioner, element-wise apply kernels), 0 for the host-driven loop. */
int hcurl_pcg_last_used_graph(const hcurl_pcg *pcg);

/* ---- discrete gradient / curl (hexes, orders 1..4) -------------------------------------------
 * Replaces: mfem::ParDiscreteLinearOperator + CurlInterpolator of CurlAuxSolver (B = curl A every step,
 * /root/reference/src/auxsolvers/curl_aux.cpp:13-30) and + GradientInterpolator of ScalarPotentialSource::BuildGrad
 * (/root/reference/src/sources/scalar_potential_source.cpp:89-97, applied at :130).  Like the reference's, the operator
 * acts on L-vectors (grid functions) of one rank: y = G x (H1 -> ND) or y = C x (ND -> RT), matrix-free and
 * geometry-free (the bases are nodal for the pulled-back fields).
 *   in_dofs  [n_elem][nd_in], out_dofs [n_elem][nd_out]: element dofs in TENSOR order (x fastest, component-major;
 *            H1 (p+1)^3 | ND component c: p points along c, p+1 along the others | RT component c: p+1 along c, p
 *            along the others), encoded like FiniteElementSpace::GetElementDofs (k, or -1-k for a sign flip).  From
 *            MFEM: GetElementDofs composed with TensorBasisElement::GetDofMap() of the element.
 *   order p: H1_FECollection(p) -> ND_FECollection(p) -> RT_FECollection(p-1). */
enum { HCURL_INTERP_GRAD = 0, HCURL_INTERP_CURL = 1 };
hcurl_interp *hcurl_interp_create(hcurl_ctx *ctx, int kind, int order, int64_t n_elem, const int32_t *in_dofs,
                                  const int32_t *out_dofs, int32_t n_in, int32_t n_out);
/* the same for a given geometry: HCURL_GEOM_HEX as above; HCURL_GEOM_TET: order 1 only (ND1 -> RT0, H1 P1 -> ND1, the
 * TEAM7 configuration), element dofs in MFEM's own local order (GetElementDofs as is): 4 vertices / 6 edges / 4 faces */
hcurl_interp *hcurl_interp_create_geom(hcurl_ctx *ctx, int geom, int kind, int order, int64_t n_elem, const int32_t *in_dofs,
                                       const int32_t *out_dofs, int32_t n_in, int32_t n_out);
void hcurl_interp_destroy(hcurl_interp *it);
/* y [n_out] = Interp x [n_in]; where = HCURL_HOST or HCURL_DEVICE */
int hcurl_interp_apply(hcurl_interp *it, const double *x, double *y, int where);
/* host-only (no GPU): the 1-D matrix D[i*(p+1)+j] = derivative of the closed basis function j at the open node i, and
 * the encoded maps (index | bit 31 sign | bit 30 on outputs: another element writes this dof) */
int hcurl_host_interp_matrix(int order, double *D);
/* tets, order 1: the local matrix [n_out_local * n_in_local] (gradient 6 x 4, curl 4 x 6), 24 doubles */
int hcurl_host_interp_tet_matrix(int kind, double *M24);
int hcurl_host_interp_encode(int kind, int order, int64_t n_elem, const int32_t *in_dofs, const int32_t *out_dofs,
                             int32_t n_in, int32_t n_out, uint32_t *in_map, uint32_t *out_map);

/* ---- vector helpers (device memory owned by the library) --------------------- */
double *hcurl_device_alloc(hcurl_ctx *ctx, int64_t n_doubles);
void hcurl_device_free(hcurl_ctx *ctx, double *p);
int hcurl_copy_to_device(hcurl_ctx *ctx, double *dst_dev, const double *src_host, int64_t n);
int hcurl_copy_to_host(hcurl_ctx *ctx, double *dst_host, const double *src_dev, int64_t n);
/* page-locked host buffers for the host<->device legs of the HOST entry points */
double *hcurl_host_alloc_pinned(int64_t n_doubles);
void hcurl_host_free_pinned(double *p);
/* global dot product over true dofs (all-reduced across ranks) */
int hcurl_dot(hcurl_ctx *ctx, const double *a_dev, const double *b_dev, int64_t n, double *result);

/* ---- host-side setup helper: patch decomposition (no GPU needed) ---------------------
 * Groups elements into patches of `elems_per_patch` spatially adjacent elements (Morton
 * order of the centroids) and lists, per patch, its unique dofs with the element slots
 * each one feeds -- the structure the patch kernels use instead of the per-element
 * gather / atomic scatter through FiniteElementSpace::GetElementDofs
 * (what blf->Assemble() resolves through the CSR sparsity in the reference,
 * /root/reference/src/equation_systems/equation_system.cpp:345).
 *  unit_bytes  8: the gathered / scattered unit is one double; 16: an aligned pair of doubles (the
 *             order-2 hex kernel works on pairs, see hephaestus_b200/csrc/hex_patch.cu)
 *  slot_code  [n_elem][nd]: unit index (bits 0..28) | orientation flags (bits 29..31; for 8-byte units
 *             bit 31 = sign flip) of each element slot
 *  slot_stride  >= elems_per_patch: E-vector index (in units) of (slot s, patch position j) = s*slot_stride + j',
 *             j' = j up to the bank swizzle reported by hcurl_host_patches_swizzle()
 *  exchange_units  NULL, or one flag per unit: patches touching a flagged unit (ghosts, units other ranks hold as
 *             ghosts) are placed first -- hcurl_host_patches_boundary_count() of them
 *  outputs    perm [n_patch*epp] element at each patch position (-1 padding);
 *             blob_off [n_patch+1] start of each patch blob in 16-byte units; blob: per patch a
 *             32-word header + the unit list (units the patch owns and stores first, units it adds to second), the
 *             16/32/64-bit slot-entry lists + overflow list (dof side) and the per-slot local unit
 *             index table (element side); layout in hephaestus_b200/csrc/patch.cpp. */
/* Order-2 hex spaces: the 27 dof pairs of every element (the 16-byte units of the patch kernel): codes[e][r] = pair
 * index q (doubles 2q, 2q+1 of the vector) | bit 29 swap | bit 30 sign flip of the second local dof | bit 31 sign
 * flip of the first.  elem_ldof [n_elem][54] as for hcurl_space_create.  Returns 0, or 1 when the map does not
 * consist of aligned pairs (such spaces use the element-wise kernels). */
int hcurl_host_hex_p2_pair_codes(int64_t n_elem, const int32_t *elem_ldof, uint32_t *codes);
/* Hex spaces of any order 1..4: single-dof slot codes in the kernels' tensor order (component-major, x fastest, the
 * order of the device map): codes[e][t] = dof index | bit 31 sign flip (MFEM's -1-k encoding of
 * FiniteElementSpace::GetElementDofs combined with the fixed sign of the local basis function).  The 8-byte units of
 * the order-1 patch kernel.  elem_ldof [n_elem][nd] as for hcurl_space_create. */
int hcurl_host_hex_slot_codes(int order, int64_t n_elem, const int32_t *elem_ldof, uint32_t *codes);
typedef struct hcurl_host_patches hcurl_host_patches;
hcurl_host_patches *hcurl_host_patches_create(int nd, int64_t n_elem, int elems_per_patch, int slot_stride,
                                              int unit_bytes, int64_t n_dof, const double *centroid,
                                              const uint32_t *slot_code, const uint8_t *exchange_units);
void hcurl_host_patches_destroy(hcurl_host_patches *p);
int64_t hcurl_host_patches_count(const hcurl_host_patches *p);
int64_t hcurl_host_patches_boundary_count(const hcurl_host_patches *p);
int64_t hcurl_host_patches_blob_words(const hcurl_host_patches *p);
int64_t hcurl_host_patches_max_blob_words(const hcurl_host_patches *p);
/* largest number of unique dofs in one patch (sizes the shared-memory x staging vector) */
int64_t hcurl_host_patches_max_dofs(const hcurl_host_patches *p);
/* E-vector swizzle chosen by the builder: (slot s, position j) sits at unit s*slot_stride + (j ^ row[s] ^ grp[j / NB]),
 * NB = 128 / unit_bytes; row has nd entries, grp ceil(elems_per_patch / NB) */
int hcurl_host_patches_swizzle(const hcurl_host_patches *p, uint8_t *row, uint8_t *grp);
/* the three parts of every patch blob by lifetime in the kernel's pipeline: begin of the lists and of the per-slot
 * table (part_off [2*n_patch], 16-byte units) and the largest parts in 32-bit words */
int hcurl_host_patches_parts(const hcurl_host_patches *p, uint32_t *part_off, int64_t *max_d_words, int64_t *max_l_words,
                             int64_t *max_t_words);
/* the 1-D transform tables of the order-2 modal element operator (a0, a1, b, c0, c1, d1, d2), for the host-side check
 * of the kernel program (tests/test_patch_p2_emu.py) */
int hcurl_host_modal_tables_p2(double *out7);
int hcurl_host_patches_copy(const hcurl_host_patches *p, int32_t *perm, uint32_t *blob_off, uint32_t *blob);
/* the shared units every patch owns (the lowest-numbered patch touching a unit owns it; the kernel zeroes them ahead of
 * the sums instead of clearing y): own_off [n_patch+1], own_idx [returned count]; null pointers are skipped */
int64_t hcurl_host_patches_owned(const hcurl_host_patches *p, uint32_t *own_off, uint32_t *own_idx);

/* ---- host-side setup helper: ND space on a hex or tet mesh (no GPU needed) ------------------
 * What mfem::ParFiniteElementSpace hands the assembly once ProblemBuilder::AddFESpace("ND_3D_P<p>") has run
 * (/root/reference/src/problem_builders/problem_builder_base.cpp:98-127): global edge / face enumeration and the
 * signed FiniteElementSpace::GetElementDofs map (k, or -1-k when the element sees the dof flipped), in MFEM's
 * conventions.  For callers that are not handed MFEM's arrays (stand-alone benchmarks, SURVEY.md
 * 8 f4); the result feeds hcurl_space_create().  elem_verts: [n_elem][8] corner ids in MFEM's hexahedron order. */
typedef struct hcurl_host_ndspace hcurl_host_ndspace;
hcurl_host_ndspace *hcurl_host_ndspace_create_hex(int order, int64_t n_elem, const int32_t *elem_verts);
/* tetrahedra, orders 1..3: elem_verts [n_elem][4]; face dofs stay in the frame of the face's first occurrence (its
 * base vertex order) -- hcurl_space_create() applies the triangle-face DofTransformation from
 * hcurl_host_ndspace_elem_face_base() ([n_elem][4][3]; for hexes [n_elem][6][4], informational) */
hcurl_host_ndspace *hcurl_host_ndspace_create_tet(int order, int64_t n_elem, const int32_t *elem_verts);
int hcurl_host_ndspace_elem_face_base(const hcurl_host_ndspace *s, int32_t *out);
void hcurl_host_ndspace_destroy(hcurl_host_ndspace *s);
int hcurl_host_ndspace_nd(const hcurl_host_ndspace *s);          /* dofs per element, 3p(p+1)^2 */
int64_t hcurl_host_ndspace_ndof(const hcurl_host_ndspace *s);
int64_t hcurl_host_ndspace_n_edges(const hcurl_host_ndspace *s);
int64_t hcurl_host_ndspace_n_faces(const hcurl_host_ndspace *s);
int hcurl_host_ndspace_elem_dofs(const hcurl_host_ndspace *s, int32_t *signed_dofs /* [n_elem][nd] */);
/* dofs on boundary faces and their edges, ascending; returns the count (out may be NULL; room for ndof entries) */
int64_t hcurl_host_ndspace_boundary_dofs(const hcurl_host_ndspace *s, int32_t *out);
/* dofs on the given faces ([n_faces][4|3] corner ids, any order: e.g. side-set faces) and their edges -- the essential
 * dofs BCMap::GetEssentialBdrMarkers selects (/root/reference/src/boundary_conditions/boundary_conditions.cpp:32-51);
 * ascending; returns the count (out may be NULL), -1 if a face is not part of the mesh */
int64_t hcurl_host_ndspace_face_dofs(const hcurl_host_ndspace *s, int64_t n_faces, const int32_t *face_verts, int32_t *out);

/* ---- host-side setup helper: rank-local space under a host-supplied element partition (no GPU needed) ---------
 * What mfem::ParMesh(comm, mesh, partitioning) + ParFiniteElementSpace derive from METIS output
 * (/root/reference/examples/team7/Main.cpp:155), for callers without MFEM: from the SERIAL space's signed element -> dof
 * map (elem_dofs [n_elem][nd], e.g. hcurl_host_ndspace_elem_dofs) and an element -> rank array, the arrays of rank
 * `rank` that hcurl_space_create() takes.  Every rank keeps the serial orientation of a shared dof (ghost signs +1), the
 * lowest rank holding a dof owns it, and the two lists of a neighbour pair are in ascending serial dof number.
 *   copy: elems [n_elem_local] serial element ids; l2g [n_ldof] serial dof of every local dof; elem_ldof
 *         [n_elem_local][nd] signed local dofs; ldof_ltdof [n_ldof] (-1 = not owned); nbr_rank [n_nbr]; own_off /
 *         ghost_off [n_nbr+1]; own_ldof [n_own]; ghost_ldof [n_ghost]  (any pointer may be NULL) */
typedef struct hcurl_host_partition hcurl_host_partition;
hcurl_host_partition *hcurl_host_partition_create(int64_t n_elem, int nd, int64_t n_dof, const int32_t *elem_dofs,
                                                  const int32_t *elem_rank, int rank);
void hcurl_host_partition_destroy(hcurl_host_partition *p);
int hcurl_host_partition_sizes(const hcurl_host_partition *p, int64_t *n_elem_local, int32_t *n_ldof, int32_t *n_tdof,
                               int *n_nbr, int32_t *n_own, int32_t *n_ghost);
int hcurl_host_partition_copy(const hcurl_host_partition *p, int64_t *elems, int32_t *l2g, int32_t *elem_ldof,
                              int32_t *ldof_ltdof, int32_t *nbr_rank, int32_t *own_off, int32_t *own_ldof,
                              int32_t *ghost_off, int32_t *ghost_ldof);

/* ---- host-side setup helper: ExodusII mesh reader (no GPU, no netCDF library needed) ----
 * What mfem::Mesh(mesh_file) does for the .g and .gen files in the reference's data directory (NetCDF classic,
 * CDF-1/2) before the ND space is built (/root/reference/examples/team7/Main.cpp:152-155): node coordinates, element blocks (TETRA4 / HEX8 /
 * HEX27; block id = element attribute), side sets (-> boundary faces, side-set id = attribute). */
typedef struct hcurl_host_exodus hcurl_host_exodus;
hcurl_host_exodus *hcurl_host_exodus_open(const char *path);
void hcurl_host_exodus_close(hcurl_host_exodus *x);
int hcurl_host_exodus_counts(const hcurl_host_exodus *x, int64_t *n_nodes, int64_t *n_elem, int *geom, int *nodes_per_elem,
                             int64_t *n_side);
/* xyz [n_nodes][3]; conn [n_elem][nodes_per_elem] 0-based, blocks in file order; attr [n_elem]; bdr [n_side][4|3]
 * corner nodes of the side-set faces; bdr_attr [n_side].  Any output may be NULL. */
int hcurl_host_exodus_read(const hcurl_host_exodus *x, double *xyz, int32_t *conn, int32_t *attr, int32_t *bdr, int32_t *bdr_attr);

/* FP64 FMA micro-benchmark used for the roofline co-bound:
 * returns achieved TFLOP/s of dependent-chain-free DFMA on the device. */
double hcurl_fp64_peak_tflops(hcurl_ctx *ctx);

/* Host-side helper (no GPU): nodes[nd][3] and tangents[nd][3] of the Nedelec hex dof functionals in MFEM's local order,
 * dof_l(v) = t_l . v(node_l) on the reference element (ND_HexahedronElement's Gauss-Legendre / Gauss-Lobatto nodal basis;
 * what GridFunction::ProjectCoefficient evaluates).  Used to interpolate given fields for parity and property checks. */
int hcurl_host_hex_dof_nodes(int order, double *nodes, double *tang);

#ifdef __cplusplus
}
#endif
#endif /* HCURL_B200_H */
