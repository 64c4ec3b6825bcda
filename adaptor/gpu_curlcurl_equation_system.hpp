// gpu_curlcurl_equation_system.hpp -- the equation system a Hephaestus maintainer drops next to
// src/formulations/HCurl/hcurl_formulation.hpp: CurlCurlEquationSystem with its operator, right-hand-side kernel and
// linear-system formation on the GPU.  Everything else of the time step
// (/root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:30-59) runs unchanged:
//   SetTimeStep -> UpdateEquationSystem (BuildLinearForms / BuildBilinearForms) -> BuildJacobian (FormLinearSystem)
//   -> NewtonSolver(jacobian solver = HephGpuPCG).Mult -> RecoverFEMSolution.
// What it replaces, per step:
//   * CurlCurlKernel + VectorFEMassKernel + blf->Assemble()  (equation_system.cpp:324-347, :413-426): coefficient VALUES are
//     re-sampled into the device operator (hcurl_op_set_coefficients); the operator and the space live across steps;
//   * WeakCurlCurlKernel  lf -= K(alpha) u_n  (src/kernels/weak_curl_curl_kernel.cpp:14-31): one unconstrained apply of
//     K(alpha) on the device;
//   * blf->FormLinearSystem + HypreParMatrixFromBlocks (equation_system.cpp:132-190): essential elimination of the
//     right-hand side on the device (hcurl_op_form_rhs); `op` becomes the HephGpuOperator.
// Compiled and run here against adaptor/mock_hephaestus.hpp (-DHEPH_GPU_MOCK: MFEM cannot be built in this container,
// tests/test_adaptor.py::test_time_steps_through_the_gpu_equation_system); in the reference tree it includes the real
// headers.
#pragma once
#ifdef HEPH_GPU_MOCK
#include "mock_hephaestus.hpp"
#else
#include "hcurl_formulation.hpp"
#endif
#include "heph_gpu_operator.hpp"

namespace hephaestus
{

class GpuCurlCurlEquationSystem : public CurlCurlEquationSystem
{
public:
  GpuCurlCurlEquationSystem(const hephaestus::InputParameters & params, std::shared_ptr<HephGpuContext> gpu_ctx,
                            HephGpuOperator::Sampling sampling = HephGpuOperator::PER_ELEMENT)
    : CurlCurlEquationSystem(params), _gpu_ctx(std::move(gpu_ctx)), _sampling(sampling)
  {
  }

  /// the coefficient objects arrive here (the base class registers "dt_<alpha>"); they are looked up again when the
  /// operators are built, so that the kernels' InputParameters are not needed
  void Init(hephaestus::GridFunctions & gridfunctions, const hephaestus::FESpaces & fespaces, hephaestus::BCMap & bc_map,
            hephaestus::Coefficients & coefficients) override
  {
    CurlCurlEquationSystem::Init(gridfunctions, fespaces, bc_map, coefficients);
    _coefs = &coefficients;
    _u = gridfunctions.Get(_h_curl_var_name);
  }

  /// register the variables only: the three kernels of hcurl_formulation.cpp:130-157 are what the device operators replace
  void AddKernels() override
  {
    AddTrialVariableNameIfMissing(_h_curl_var_name);
    AddTestVariableNameIfMissing(GetTimeDerivativeName(_h_curl_var_name));
  }

  /// equation_system.cpp:413-426 without the re-assembly: the new dt reaches the operator through "dt_<alpha>"
  void SetTimeStep(double dt) override
  {
    if (fabs(dt - _dt_coef.constant) > 1.0e-12 * dt)
    {
      _dt_coef.constant = dt;
      _coefs_dirty = true;
    }
  }

  /// equation_system.cpp:324-347 (new ParBilinearForm + kernels + Assemble every step) -> coefficient values only.
  /// time_dependent_coefficients: re-sample every step (Coefficients::SetTime); otherwise only after a change of dt.
  void BuildBilinearForms() override
  {
    EnsureOperators();
    if (_coefs_dirty || time_dependent_coefficients)
    {
      _op->UpdateCoefficients();
      _kop->UpdateCoefficients();
      _coefs_dirty = false;
    }
  }

  /// replaces equation_system.cpp:132-190 for the single-variable H(curl) system
  void FormLinearSystem(mfem::OperatorHandle & op, mfem::BlockVector & trueX, mfem::BlockVector & trueRHS) override
  {
    EnsureOperators();
    const auto & var = _test_var_names.at(0);
    mfem::ParFiniteElementSpace * fes = _test_pfespaces.at(0);
    const int n = fes->GetTrueVSize();
    mfem::Vector b(n), xbc(n), un(n);
    _lfs.Get(var)->ParallelAssemble(b);        // sources and integrated BCs: assembled on the host by the reference
    _u->GetTrueDofs(un);
    _kop->AddMultUnconstrained(un, b, -1.0);   // b -= K(alpha) u_n  (WeakCurlCurlKernel)
    _xs.at(0)->GetTrueDofs(xbc);               // carries the essential values
    _op->SetEssentialTrueDofs(_ess_tdof_lists.at(0));
    _op->FormRHS(xbc, b, trueRHS.GetBlock(0)); // B = b - A x_bc|ess, B[ess] = x_bc[ess]
    trueX.GetBlock(0) = xbc;
    op.Reset(_op.get(), false);                // EquationSystem::Mult / GetGradient now run on the GPU
    height = width = n;
  }

  HephGpuOperator * GpuOperator() const { return _op.get(); }
  bool time_dependent_coefficients{false};

private:
  void EnsureOperators()
  {
    if (_op)
      return;
    mfem::ParFiniteElementSpace * fes = _test_pfespaces.at(0);
    _gpu_space = std::make_unique<HephGpuSpace>(*_gpu_ctx, *fes);
    mfem::Array<int> none;
    // A = K(dt alpha) + M(beta), and K(alpha) for the right-hand side
    _op = std::make_unique<HephGpuOperator>(*_gpu_ctx, *_gpu_space, _coefs->_scalars.Get(_dtalpha_coef_name),
                                            _coefs->_scalars.Get(_beta_coef_name), none, HCURL_QUAD_AUTO, _sampling);
    _kop = std::make_unique<HephGpuOperator>(*_gpu_ctx, *_gpu_space, _coefs->_scalars.Get(_alpha_coef_name), nullptr, none,
                                             HCURL_QUAD_AUTO, _sampling);
    _coefs_dirty = false;
  }

  std::shared_ptr<HephGpuContext> _gpu_ctx;
  HephGpuOperator::Sampling _sampling;
  hephaestus::Coefficients * _coefs{nullptr};
  mfem::ParGridFunction * _u{nullptr};
  std::unique_ptr<HephGpuSpace> _gpu_space;
  std::unique_ptr<HephGpuOperator> _op, _kop;
  bool _coefs_dirty{true};
};

} // namespace hephaestus
