// heph_gpu_operator.hpp -- header-only C++ adaptor that plugs libhcurl_b200.so into Hephaestus
// behind its existing plug points.  Host code stays C++ / MFEM; everything
// numerical happens behind the C ABI in include/hcurl_b200.h.
//
//   HephGpuOperator : mfem::Operator   -- what EquationSystem::FormLinearSystem hands out in
//       place of HypreParMatrixFromBlocks (/root/reference/src/equation_systems/equation_system.cpp:132-190);
//       EquationSystem::Mult / GetGradient (ibid. :200-210) then run on the GPU unchanged.
//   HephGpuPCG : mfem::Solver          -- replaces mfem::HyprePCG + HypreDiagScale as configured by
//       DefaultJacobiPCGSolver (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57) and
//       ProblemBuilder::ConstructJacobianSolverWithOptions
//       (/root/reference/src/problem_builders/problem_builder_base.cpp:220-254).
//   MakeGpuSpace(...)                  -- extracts from a ParFiniteElementSpace exactly the arrays
//       the reference's assembly consumes (element vertices, signed GetElementDofs, true-dof
//       numbering, shared-dof groups) and hands them to hcurl_space_create.
//
// Include AFTER mfem.hpp (or adaptor/mock_mfem.hpp, a minimal stand-in used by the compile test:
// MFEM itself is not installable in the build container, so this header has been compiled
// against the mock only -- see DESIGN.md "unverified against real MFEM").
#pragma once
#include <cstdint>
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../include/hcurl_b200.h"

#ifndef MFEM_ABORT
#error "include mfem.hpp (or adaptor/mock_mfem.hpp) before heph_gpu_operator.hpp"
#endif

namespace hephaestus
{

#define HEPH_GPU_CHECK(call)                                                                  \
  do                                                                                          \
  {                                                                                           \
    if ((call) != 0)                                                                          \
      MFEM_ABORT(std::string(#call) + " failed: " + hcurl_last_error()); /* reference: abort */ \
  } while (0)

/// One per MPI rank.  `nccl_unique_id` comes from hcurl_comm_unique_id on rank 0 + MPI_Bcast.
class HephGpuContext
{
public:
  HephGpuContext(int device, const void * nccl_unique_id, int rank, int nranks)
    : _ctx(hcurl_ctx_create(device, nccl_unique_id, rank, nranks))
  {
    if (!_ctx)
      MFEM_ABORT(std::string("hcurl_ctx_create failed: ") + hcurl_last_error());
  }
  ~HephGpuContext() { hcurl_ctx_destroy(_ctx); }
  HephGpuContext(const HephGpuContext &) = delete;
  hcurl_ctx * Get() const { return _ctx; }

private:
  hcurl_ctx * _ctx;
};

/// Device copy of an ND_3D_P<p> ParFiniteElementSpace.
class HephGpuSpace
{
public:
  HephGpuSpace(HephGpuContext & ctx, mfem::ParFiniteElementSpace & fes) : _fes(fes)
  {
    mfem::ParMesh * mesh = fes.GetParMesh();
    const int ne = mesh->GetNE();
    const int order = fes.FEColl()->GetOrder(); // "ND_3D_P<p>" -> p
    const bool tet = mesh->GetElementGeometry(0) == mfem::Geometry::TETRAHEDRON;
    const mfem::GridFunction * nodes = mesh->GetNodes();
    const int geom_order = nodes ? nodes->FESpace()->GetMaxElementOrder() : 1;
    const int ngn = tet ? 4 : (geom_order == 1 ? 8 : 27);

    // ---- element node coordinates (corner order = MFEM's; HEX27: x-fastest tensor order)
    std::vector<double> xyz(size_t(ne) * ngn * 3);
    mfem::Array<int> v;
    for (int e = 0; e < ne; e++)
    {
      if (geom_order == 1)
      {
        mesh->GetElementVertices(e, v);
        for (int k = 0; k < ngn; k++)
          for (int d = 0; d < 3; d++)
            xyz[(size_t(e) * ngn + k) * 3 + d] = mesh->GetVertex(v[k])[d];
      }
      else
      {
        // sample the (curved) element map at the 27 tensor points {0,1/2,1}^3
        mfem::ElementTransformation * T = mesh->GetElementTransformation(e);
        mfem::Vector x(3);
        for (int k = 0; k < 27; k++)
        {
          mfem::IntegrationPoint ip;
          ip.Set3(0.5 * (k % 3), 0.5 * ((k / 3) % 3), 0.5 * (k / 9));
          T->Transform(ip, x);
          for (int d = 0; d < 3; d++)
            xyz[(size_t(e) * 27 + k) * 3 + d] = x(d);
        }
      }
    }
    // ---- signed element dofs: FiniteElementSpace::GetElementDofs (k, or -1-k when flipped)
    mfem::Array<int> dofs;
    fes.GetElementDofs(0, dofs);
    const int nd = dofs.Size();
    std::vector<int32_t> ldof(size_t(ne) * nd);
    for (int e = 0; e < ne; e++)
    {
      fes.GetElementDofs(e, dofs);
      for (int i = 0; i < nd; i++)
        ldof[size_t(e) * nd + i] = dofs[i];
    }
    // ---- tets of order >= 2: vertex ids of the element and of its faces in the mesh's own
    // face-vertex order; the library derives ND_DofTransformation from them
    std::vector<int32_t> ev, fb;
    if (tet && order >= 2)
    {
      ev.resize(size_t(ne) * 4);
      fb.resize(size_t(ne) * 12);
      mfem::Array<int> faces, ori, fv;
      for (int e = 0; e < ne; e++)
      {
        mesh->GetElementVertices(e, v);
        for (int k = 0; k < 4; k++)
          ev[size_t(e) * 4 + k] = v[k];
        mesh->GetElementFaces(e, faces, ori);
        for (int f = 0; f < 4; f++)
        {
          mesh->GetFaceVertices(faces[f], fv);
          for (int k = 0; k < 3; k++)
            fb[(size_t(e) * 4 + f) * 3 + k] = fv[k];
        }
      }
    }
    // ---- parallel layout: owned true dofs + per-neighbour shared-dof lists (GroupTopology)
    const int n_ldof = fes.GetVSize(), n_tdof = fes.GetTrueVSize();
    std::vector<int32_t> ltdof(n_ldof);
    for (int i = 0; i < n_ldof; i++)
      ltdof[i] = fes.GetLocalTDofNumber(i); // -1 when another rank owns it
    std::vector<int32_t> nbr, own_off(1, 0), own_ldof, ghost_off(1, 0), ghost_ldof;
    std::vector<int8_t> ghost_sign;
    if (fes.GetNRanks() > 1)
    {
#ifndef HEPH_GPU_ALLOW_UNVERIFIED_MULTIRANK
      // This branch has only ever been compiled against the mock (MFEM is not installable in the build container).  The
      // lists below follow MFEM's GroupCommunicator conventions as documented, but until they have been checked against
      // a real ParFiniteElementSpace a silent mis-pairing of shared dofs is worse than an abort.  The library's own
      // multi-GPU path (hcurl_host_partition_*, tests/test_gpu_multi.py) does not go through here.
      MFEM_ABORT("HephGpuSpace: several MPI ranks need -DHEPH_GPU_ALLOW_UNVERIFIED_MULTIRANK (unverified against MFEM)");
#endif
      // A group g is a set of ranks sharing some ldofs; its master owns them.  GroupCommunicator::GroupLDofTable() lists
      // the ldofs of every group in the order BOTH sides agree on (shared edges, then triangles, then quadrilaterals of
      // the group; an entry -1-k marks a dof this rank sees with the opposite orientation).  Local group ids differ
      // between ranks, so the groups of a neighbour are ordered by a rank-independent key: (master rank, the master's
      // own id of the group, GroupTopology::GetGroupMasterGroup).
      mfem::GroupTopology & gt = fes.GetGroupTopo();
      const mfem::Table & gl = fes.GroupComm().GroupLDofTable();
      const int me = fes.GetMyRank();
      struct Seg { int master, master_group, g; };
      std::map<int, std::vector<Seg>> segs; // neighbour rank -> groups shared with it
      for (int g = 1; g < gt.NGroups(); g++)
      {
        const int master = gt.GetNeighborRank(gt.GetGroupMaster(g));
        const Seg sg{master, gt.GetGroupMasterGroup(g), g};
        if (master == me)
        {
          const int * gr = gt.GetGroup(g);
          for (int m = 0; m < gt.GetGroupSize(g); m++)
            if (gt.GetNeighborRank(gr[m]) != me)
              segs[gt.GetNeighborRank(gr[m])].push_back(sg);
        }
        else
          segs[master].push_back(sg);
      }
      for (auto & kv : segs)
      {
        std::sort(kv.second.begin(), kv.second.end(), [](const Seg & a, const Seg & b)
                  { return a.master != b.master ? a.master < b.master : a.master_group < b.master_group; });
        nbr.push_back(kv.first);
        for (const Seg & sg : kv.second)
        {
          const int * row = gl.GetRow(sg.g);
          for (int k = 0; k < gl.RowSize(sg.g); k++)
          {
            const int ld = row[k] >= 0 ? row[k] : -1 - row[k];
            if (sg.master == me)
              own_ldof.push_back(ld);
            else
            {
              ghost_ldof.push_back(ld);
              // the master's value enters with the relative orientation of the two ranks' copies; the master's own
              // entry is non-negative by MFEM's convention, so the sign of this rank's entry is that orientation
              ghost_sign.push_back(row[k] >= 0 ? 1 : -1);
            }
          }
        }
        own_off.push_back((int32_t)own_ldof.size());
        ghost_off.push_back((int32_t)ghost_ldof.size());
      }
    }
    _space = hcurl_space_create(ctx.Get(), tet ? HCURL_GEOM_TET : HCURL_GEOM_HEX, geom_order, order, ne,
                                xyz.data(), ldof.data(), ev.empty() ? nullptr : ev.data(),
                                fb.empty() ? nullptr : fb.data(), n_ldof,
                                fes.GetNRanks() > 1 ? ltdof.data() : nullptr, n_tdof, (int)nbr.size(),
                                nbr.data(), own_off.data(), own_ldof.data(), ghost_off.data(),
                                ghost_ldof.data(), ghost_sign.empty() ? nullptr : ghost_sign.data());
    if (!_space)
      MFEM_ABORT(std::string("hcurl_space_create failed: ") + hcurl_last_error());
  }
  ~HephGpuSpace() { hcurl_space_destroy(_space); }
  HephGpuSpace(const HephGpuSpace &) = delete;
  hcurl_space * Get() const { return _space; }
  mfem::ParFiniteElementSpace & FESpace() const { return _fes; }

private:
  mfem::ParFiniteElementSpace & _fes;
  hcurl_space * _space{nullptr};
};

/// A = K(alpha) + M(beta) on true dofs, essential rows/cols -> identity.
/// alpha / beta are the coefficient objects the reference's kernels hold:
/// CurlCurlKernel::_coef (/root/reference/src/kernels/curl_curl_kernel.hpp:22-23) and
/// VectorFEMassKernel::_coef (/root/reference/src/kernels/vector_fe_mass_kernel.hpp).
/// quad_pts_per_dim: HCURL_QUAD_AUTO (default) = the p+1 rule wherever it is exact (all hexes parallelepipeds with
/// per-element coefficients: equal to MFEM's full-assembly rules to round-off -- tests/test_gpu_apply.py::
/// test_quad_auto_picks_the_exact_rule_on_affine_meshes -- and the patch kernels apply), MFEM's rules otherwise;
/// 0 = MFEM's full-assembly rules everywhere.
class HephGpuOperator : public mfem::Operator
{
public:
  /// How alpha / beta are sampled: once per element at its centre (PWCoefficient / ConstantCoefficient and
  /// TransformedCoefficients of those -- TEAM7, TEAM4, the rod and coil examples), or at every quadrature point of the
  /// integrators' rules (FunctionCoefficients, as in the reference's MMS tests,
  /// /root/reference/test/integration/test_aform_source.cpp:32-37).  PER_QUAD_POINT keeps MFEM's rules (the patch
  /// kernels need per-element coefficients).
  enum Sampling { PER_ELEMENT = 0, PER_QUAD_POINT = 1 };

  HephGpuOperator(HephGpuContext & ctx, HephGpuSpace & space, mfem::Coefficient * alpha,
                  mfem::Coefficient * beta, const mfem::Array<int> & ess_tdof_list,
                  int quad_pts_per_dim = HCURL_QUAD_AUTO, Sampling sampling = PER_ELEMENT)
    : mfem::Operator(space.FESpace().GetTrueVSize()), _space(space), _alpha(alpha), _beta(beta), _sampling(sampling)
  {
    _op = hcurl_op_create(ctx.Get(), space.Get(), sampling == PER_QUAD_POINT ? 0 : quad_pts_per_dim);
    if (!_op)
      MFEM_ABORT(std::string("hcurl_op_create failed: ") + hcurl_last_error());
    UpdateCoefficients();
    SetEssentialTrueDofs(ess_tdof_list);
  }
  ~HephGpuOperator() override { hcurl_op_destroy(_op); }

  /// EquationSystem::_ess_tdof_lists (equation_system.hpp:94): rows / columns -> identity
  void SetEssentialTrueDofs(const mfem::Array<int> & ess_tdof_list)
  {
    std::vector<int32_t> ess(ess_tdof_list.Size());
    for (int i = 0; i < ess_tdof_list.Size(); i++)
      ess[i] = ess_tdof_list[i];
    HEPH_GPU_CHECK(hcurl_op_set_essential_tdofs(_op, ess.data(), (int32_t)ess.size()));
  }

  /// Re-sample alpha / beta (Coefficients::SetTime, dt change:
  /// /root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:43,
  /// /root/reference/src/equation_systems/equation_system.cpp:413-426).  Values only, no rebuild.
  void UpdateCoefficients()
  {
    mfem::ParMesh * mesh = _space.FESpace().GetParMesh();
    const int ne = mesh->GetNE();
    if (_sampling == PER_ELEMENT)
    {
      std::vector<double> a(_alpha ? ne : 0), b(_beta ? ne : 0);
      for (int e = 0; e < ne; e++)
      {
        mfem::ElementTransformation * T = mesh->GetElementTransformation(e);
        const mfem::IntegrationPoint & ip = mfem::Geometries.GetCenter(mesh->GetElementGeometry(e));
        T->SetIntPoint(&ip);
        if (_alpha) a[e] = _alpha->Eval(*T, ip);
        if (_beta) b[e] = _beta->Eval(*T, ip);
      }
      HEPH_GPU_CHECK(hcurl_op_set_coefficients(_op, HCURL_COEF_PER_ELEM, _alpha ? a.data() : nullptr,
                                               _beta ? b.data() : nullptr, HCURL_HOST));
      return;
    }
    // the rules MFEM's integrators pick on these elements: CurlCurlIntegrator order 2p-2 (simplices) /
    // 2p (tensor elements), VectorFEMassIntegrator order OrderW + 2p; hex rules are tensor rules, x fastest, which is
    // the library's order; tet rules are handed to the library point by point
    const int geom = mesh->GetElementGeometry(0);
    const int p = _space.FESpace().FEColl()->GetOrder();
    mfem::ElementTransformation * T0 = mesh->GetElementTransformation(0);
    const bool tet = geom == mfem::Geometry::TETRAHEDRON;
    const mfem::IntegrationRule & irc = mfem::IntRules.Get(geom, tet ? 2 * p - 2 : 2 * p);
    const mfem::IntegrationRule & irm = mfem::IntRules.Get(geom, T0->OrderW() + 2 * p);
    if (tet)
    {
      auto flat = [](const mfem::IntegrationRule & ir, std::vector<double> & pts, std::vector<double> & w)
      {
        for (int q = 0; q < ir.GetNPoints(); q++)
        {
          const mfem::IntegrationPoint & ip = ir.IntPoint(q);
          pts.push_back(ip.x); pts.push_back(ip.y); pts.push_back(ip.z);
          w.push_back(ip.weight);
        }
      };
      std::vector<double> pc, wc, pm, wm;
      flat(irc, pc, wc);
      flat(irm, pm, wm);
      HEPH_GPU_CHECK(hcurl_op_set_tet_rules(_op, (int)wc.size(), pc.data(), wc.data(), (int)wm.size(), pm.data(), wm.data()));
    }
    if (irc.GetNPoints() != hcurl_op_qpts_curl(_op) || irm.GetNPoints() != hcurl_op_qpts_mass(_op))
      MFEM_ABORT("HephGpuOperator: MFEM's integration rules differ from the library's (curl / mass points per element)");
    std::vector<double> a(_alpha ? (size_t)ne * irc.GetNPoints() : 0), b(_beta ? (size_t)ne * irm.GetNPoints() : 0);
    for (int e = 0; e < ne; e++)
    {
      mfem::ElementTransformation * T = mesh->GetElementTransformation(e);
      if (_alpha)
        for (int q = 0; q < irc.GetNPoints(); q++)
        {
          T->SetIntPoint(&irc.IntPoint(q));
          a[(size_t)e * irc.GetNPoints() + q] = _alpha->Eval(*T, irc.IntPoint(q));
        }
      if (_beta)
        for (int q = 0; q < irm.GetNPoints(); q++)
        {
          T->SetIntPoint(&irm.IntPoint(q));
          b[(size_t)e * irm.GetNPoints() + q] = _beta->Eval(*T, irm.IntPoint(q));
        }
    }
    HEPH_GPU_CHECK(hcurl_op_set_coefficients(_op, HCURL_COEF_PER_QPT, _alpha ? a.data() : nullptr,
                                             _beta ? b.data() : nullptr, HCURL_HOST));
  }

  /// y = A x on T-vectors: EquationSystem::Mult (equation_system.cpp:200-204)
  void Mult(const mfem::Vector & x, mfem::Vector & y) const override
  {
    HEPH_GPU_CHECK(hcurl_op_apply(_op, x.HostRead(), y.HostWrite(), HCURL_HOST));
  }
  /// Y_j = A X_j (mfem::Operator::ArrayMult; with MFEM proper the arguments are mfem::Array<const Vector *> /
  /// mfem::Array<Vector *> -- same body): uploads, applies and downloads of consecutive vectors overlap
  template <class ConstVecPtrArray, class VecPtrArray>
  void ArrayMultT(const ConstVecPtrArray & X, VecPtrArray & Y, int n) const
  {
    std::vector<const double *> px((size_t)n);
    std::vector<double *> py((size_t)n);
    for (int j = 0; j < n; j++)
    {
      px[(size_t)j] = X[j]->HostRead();
      py[(size_t)j] = Y[j]->HostWrite();
    }
    HEPH_GPU_CHECK(hcurl_op_apply_batch(_op, n, px.data(), py.data(), HCURL_HOST));
  }
  void ArrayMult(const std::vector<const mfem::Vector *> & X, std::vector<mfem::Vector *> & Y) const override
  {
    ArrayMultT(X, Y, (int)X.size());
  }
  /// B = b - A x_bc|ess, B[ess] = x_bc[ess]: the right-hand side FormLinearSystem produces
  void FormRHS(const mfem::Vector & x_bc, const mfem::Vector & b, mfem::Vector & B) const
  {
    HEPH_GPU_CHECK(hcurl_op_form_rhs(_op, x_bc.HostRead(), b.HostRead(), B.HostWrite(), HCURL_HOST));
  }
  /// y += scale * (K + M) x without elimination: WeakCurlCurlKernel's  lf -= K(alpha) u_n
  /// (/root/reference/src/kernels/weak_curl_curl_kernel.cpp:14-31)
  void AddMultUnconstrained(const mfem::Vector & x, mfem::Vector & y, double scale) const
  {
    HEPH_GPU_CHECK(hcurl_op_apply_unconstrained(_op, x.HostRead(), y.HostReadWrite(), scale, 1, HCURL_HOST));
  }
  hcurl_op * Get() const { return _op; }

private:
  HephGpuSpace & _space;
  mfem::Coefficient * _alpha;
  mfem::Coefficient * _beta;
  Sampling _sampling;
  hcurl_op * _op{nullptr};
};

/// Discrete gradient (H1 -> ND) or discrete curl (ND -> RT) on grid functions (L-vectors), hex meshes.
/// Drop-in for the mfem::ParDiscreteLinearOperator the reference assembles with a GradientInterpolator
/// (ScalarPotentialSource::BuildGrad, /root/reference/src/sources/scalar_potential_source.cpp:89-97) or a
/// CurlInterpolator (CurlAuxSolver::Init, /root/reference/src/auxsolvers/curl_aux.cpp:19-23); Mult replaces
/// _grad->Mult(*_phi, *_grad_phi) (:130) and _curl->Mult(*_u, *_curl_u) (curl_aux.cpp:29).
/// The library wants element dofs in tensor order: GetElementDofs composed with the element's own
/// TensorBasisElement::GetDofMap(), so no local ordering convention is restated on this side.
class HephGpuDiscreteInterpolator : public mfem::Operator
{
public:
  enum Kind { GRADIENT = HCURL_INTERP_GRAD, CURL = HCURL_INTERP_CURL };
  HephGpuDiscreteInterpolator(HephGpuContext & ctx, Kind kind, mfem::ParFiniteElementSpace & domain,
                              mfem::ParFiniteElementSpace & range)
    : mfem::Operator(range.GetVSize(), domain.GetVSize())
  {
    // ND_FECollection(p) and H1_FECollection(p) report p; RT_FECollection(p-1) pairs with them
    const int p = domain.FEColl()->GetOrder();
    mfem::ParMesh * mesh = domain.GetParMesh();
    const bool tets = mesh->GetNE() > 0 && mesh->GetElementGeometry(0) == mfem::Geometry::TETRAHEDRON;
    std::vector<int32_t> in, out;
    if (tets)
    { // lowest order only (the TEAM7 configuration): MFEM's own local order, no dof map involved
      NativeDofs(domain, in);
      NativeDofs(range, out);
    }
    else
    {
      TensorDofs(domain, in);
      TensorDofs(range, out);
    }
    _it = hcurl_interp_create_geom(ctx.Get(), tets ? HCURL_GEOM_TET : HCURL_GEOM_HEX, (int)kind, p, mesh->GetNE(), in.data(),
                                   out.data(), domain.GetVSize(), range.GetVSize());
    if (!_it)
      MFEM_ABORT(std::string("hcurl_interp_create_geom failed: ") + hcurl_last_error());
  }
  ~HephGpuDiscreteInterpolator() override { hcurl_interp_destroy(_it); }
  void Mult(const mfem::Vector & x, mfem::Vector & y) const override
  {
    HEPH_GPU_CHECK(hcurl_interp_apply(_it, x.HostRead(), y.HostWrite(), HCURL_HOST));
  }

  /// [ne][nd] element dofs exactly as FiniteElementSpace::GetElementDofs returns them
  static void NativeDofs(const mfem::ParFiniteElementSpace & fes, std::vector<int32_t> & out)
  {
    const int ne = fes.GetParMesh()->GetNE();
    mfem::Array<int> dofs;
    out.clear();
    for (int e = 0; e < ne; e++)
    {
      fes.GetElementDofs(e, dofs);
      for (int t = 0; t < dofs.Size(); t++)
        out.push_back(dofs[t]);
    }
  }

  /// [ne][nd] element dofs in tensor order, FiniteElementSpace::GetElementDofs encoding (k or -1-k)
  static void TensorDofs(const mfem::ParFiniteElementSpace & fes, std::vector<int32_t> & out)
  {
    const int ne = fes.GetParMesh()->GetNE();
    mfem::Array<int> dofs;
    out.clear();
    for (int e = 0; e < ne; e++)
    {
      const auto * tb = dynamic_cast<const mfem::TensorBasisElement *>(fes.GetFE(e));
      if (!tb)
        MFEM_ABORT("HephGpuDiscreteInterpolator: not a tensor-product element (hex meshes only)");
      const mfem::Array<int> & dm = tb->GetDofMap();
      fes.GetElementDofs(e, dofs);
      for (int t = 0; t < dofs.Size(); t++)
      {
        const int m = dm.Size() ? dm[t] : t;
        const int d = dofs[m >= 0 ? m : -1 - m];
        const int k = d >= 0 ? d : -1 - d;
        const bool flip = (m < 0) != (d < 0);
        out.push_back(flip ? -1 - k : k);
      }
    }
  }

private:
  hcurl_interp * _it{nullptr};
};

/// Jacobi-preconditioned CG with hypre's PCG semantics, on the device.
/// Options mirror DefaultJacobiPCGSolver (hephaestus_solvers.hpp:39-42): Tolerance / AbsTolerance
/// arrive as float, MaxIter as unsigned int, PrintLevel as int.
class HephGpuPCG : public mfem::Solver
{
public:
  /// chebyshev_order > 0 replaces the Jacobi preconditioner by a Jacobi-scaled Chebyshev polynomial of that order
  /// (no reference counterpart; the reference's AMS preconditioner is out of scope)
  HephGpuPCG(HephGpuContext & ctx, float tol = 1.0e-9f, float abstol = 1e-16f,
             unsigned int max_iter = 1000, int print_level = 0, int chebyshev_order = 0)
    : _ctx(ctx), _tol(tol), _abstol(abstol), _max_iter((int)max_iter), _print_level(print_level),
      _cheb_order(chebyshev_order)
  {
    iterative_mode = true; // like mfem::HyprePCG; NewtonSolver::SetSolver sets it to false
  }
  ~HephGpuPCG() override { hcurl_pcg_destroy(_pcg); }

  /// NewtonSolver hands over EquationSystem::GetGradient() (equation_system.cpp:206-210).
  /// Like HyprePCG::SetOperator insisting on a HypreParMatrix, this insists on a HephGpuOperator.
  void SetOperator(const mfem::Operator & op) override
  {
    const auto * gop = dynamic_cast<const HephGpuOperator *>(&op);
    if (!gop)
      MFEM_ABORT("HephGpuPCG::SetOperator: operator is not a HephGpuOperator");
    height = op.Height();
    width = op.Width();
    if (_pcg && gop == _gop)
    {
      HEPH_GPU_CHECK(hcurl_pcg_update(_pcg)); // new coefficients: rebuild the Jacobi diagonal
      return;
    }
    hcurl_pcg_destroy(_pcg);
    _gop = gop;
    _pcg = hcurl_pcg_create(_ctx.Get(), gop->Get(), _cheb_order > 0 ? HCURL_PC_CHEBYSHEV : HCURL_PC_JACOBI);
    if (!_pcg)
      MFEM_ABORT(std::string("hcurl_pcg_create failed: ") + hcurl_last_error());
    HEPH_GPU_CHECK(hcurl_pcg_set(_pcg, _tol, _abstol, _max_iter, _print_level));
    if (_cheb_order > 0)
      HEPH_GPU_CHECK(hcurl_pcg_set_chebyshev(_pcg, _cheb_order, 10));
  }
  void Mult(const mfem::Vector & b, mfem::Vector & x) const override
  {
    if (!_pcg)
      MFEM_ABORT("HephGpuPCG::Mult called before SetOperator");
    int iters = 0, converged = 0;
    double rel = 0.0;
    HEPH_GPU_CHECK(hcurl_pcg_solve(_pcg, b.HostRead(), x.HostReadWrite(), iterative_mode ? 0 : 1, HCURL_HOST,
                                   &iters, &rel, &converged));
    _num_iterations = iters;
    _final_rel_res = rel;
    _converged = converged != 0;
  }
  int GetNumIterations() const { return _num_iterations; }
  double GetFinalRelResidual() const { return _final_rel_res; }
  bool GetConverged() const { return _converged; }

private:
  HephGpuContext & _ctx;
  double _tol, _abstol;
  int _max_iter, _print_level, _cheb_order;
  const HephGpuOperator * _gop{nullptr};
  hcurl_pcg * _pcg{nullptr};
  mutable int _num_iterations{0};
  mutable double _final_rel_res{0.0};
  mutable bool _converged{false};
};

} // namespace hephaestus
