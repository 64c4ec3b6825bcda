// mock_hephaestus.hpp -- the part of the reference's own class hierarchy that adaptor/gpu_curlcurl_equation_system.hpp
// derives from and talks to, restated on top of mock_mfem.hpp so that the glue can be COMPILED AND RUN without MFEM
// (which cannot be built in this container).  TEST infrastructure.  Names, members and call order follow
//   /root/reference/src/equation_systems/equation_system.hpp:15-142, equation_system.cpp:116-222, 324-347, 413-435
//   /root/reference/src/formulations/HCurl/hcurl_formulation.hpp:36-47, hcurl_formulation.cpp:108-157
//   /root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:30-59
// Only what the single-variable H(curl) system uses is present (no mixed / nonlinear forms); the linear form is a
// plain vector the test fills (in the reference it is assembled by MFEM's linear-form kernels on the host, outside the
// GPU path).
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "mock_mfem.hpp"

namespace mfem
{
class ConstantCoefficient : public Coefficient
{
public:
  double constant;
  explicit ConstantCoefficient(double c = 1.0) : constant(c) {}
  double Eval(ElementTransformation &, const IntegrationPoint &) override { return constant; }
};
/// f(a, b) evaluated point-wise (mfem::TransformedCoefficient)
class TransformedCoefficient : public Coefficient
{
public:
  TransformedCoefficient(Coefficient * a, Coefficient * b, std::function<double(double, double)> f)
    : _a(a), _b(b), _f(std::move(f)) {}
  double Eval(ElementTransformation & T, const IntegrationPoint & ip) override { return _f(_a->Eval(T, ip), _b->Eval(T, ip)); }

private:
  Coefficient *_a, *_b;
  std::function<double(double, double)> _f;
};
/// value = f(x(ip)) (mfem::FunctionCoefficient)
class FunctionCoefficient : public Coefficient
{
public:
  explicit FunctionCoefficient(std::function<double(const Vector &)> f) : _f(std::move(f)) {}
  double Eval(ElementTransformation & T, const IntegrationPoint & ip) override
  {
    Vector x;
    T.Transform(ip, x);
    return _f(x);
  }

private:
  std::function<double(const Vector &)> _f;
};

class ParGridFunction : public Vector
{
public:
  explicit ParGridFunction(ParFiniteElementSpace * f) : Vector(f->GetVSize()), _fes(f) {}
  ParFiniteElementSpace * ParFESpace() const { return _fes; }
  ParGridFunction & operator=(double v) { for (auto & e : _d) e = v; return *this; }
  void GetTrueDofs(Vector & tv) const { tv._d = _d; }        // serial mock: T-vector == L-vector
  void Distribute(const Vector * tv) { _d = tv->_d; }

private:
  ParFiniteElementSpace * _fes;
};
class ParLinearForm : public Vector
{
public:
  explicit ParLinearForm(ParFiniteElementSpace * f) : Vector(f->GetVSize()) {}
  ParLinearForm & operator=(double v) { for (auto & e : _d) e = v; return *this; }
  void ParallelAssemble(Vector & tv) const { tv._d = _d; }
};
/// one block is all the H(curl) systems have
class BlockVector : public Vector
{
public:
  explicit BlockVector(int n) : Vector(n) {}
  Vector & GetBlock(int) { return *this; }
  BlockVector & operator=(double v) { for (auto & e : _d) e = v; return *this; }
};
class OperatorHandle
{
public:
  ~OperatorHandle() { if (_own) delete _op; }
  void Reset(Operator * op, bool own = true) { if (_own) delete _op; _op = op; _own = own; }
  Operator * operator->() const { return _op; }
  Operator & operator*() const { return *_op; }
  Operator * Ptr() const { return _op; }

private:
  Operator * _op{nullptr};
  bool _own{false};
};
/// the one Newton iteration the reference runs per step (ProblemBuilder::ConstructNonlinearSolver: MaxIter 1, zero start)
class NewtonSolver
{
public:
  void SetSolver(Solver & s) { _prec = &s; _prec->iterative_mode = false; }
  void SetOperator(const Operator & op) { _oper = &op; }
  void Mult(const Vector & b, Vector & x) const;
  Solver * _prec{nullptr};
  const Operator * _oper{nullptr};
};
} // namespace mfem

namespace hephaestus
{
inline double prodFunc(double a, double b) { return a * b; }   // src/coefficients/coefficients.cpp:8-11

template <class T>
class NamedFieldsMap
{
public:
  void Register(const std::string & name, std::shared_ptr<T> p) { _m[name] = std::move(p); }
  bool Has(const std::string & name) const { return _m.count(name) != 0; }
  T * Get(const std::string & name) const
  {
    auto it = _m.find(name);
    if (it == _m.end()) MFEM_ABORT("NamedFieldsMap: no field named " + name);
    return it->second.get();
  }
  T & GetRef(const std::string & name) const { return *Get(name); }

private:
  std::map<std::string, std::shared_ptr<T>> _m;
};
using GridFunctions = NamedFieldsMap<mfem::ParGridFunction>;
using FESpaces = NamedFieldsMap<mfem::ParFiniteElementSpace>;

class Coefficients
{
public:
  NamedFieldsMap<mfem::Coefficient> _scalars;
  double _t{0.0};
  void SetTime(double t) { _t = t; }
};
class InputParameters
{
public:
  void SetParam(const std::string & k, const std::string & v) { _s[k] = v; }
  template <class T> T GetParam(const std::string & k) const { return _s.at(k); }

private:
  std::map<std::string, std::string> _s;
};
/// essential data of one variable: the true-dof list and the boundary values (BCMap::ApplyEssentialBCs projects the
/// boundary coefficient's tangential trace into the grid function; here the test supplies the projected values)
class BCMap
{
public:
  std::vector<int> ess_tdofs;
  std::function<void(double, mfem::Vector &)> boundary_values;   // (time, grid function) -> sets the essential entries
  double time{0.0};
  void ApplyEssentialBCs(const std::string &, mfem::Array<int> & ess_tdof_list, mfem::ParGridFunction & gf, mfem::ParMesh *)
  {
    ess_tdof_list.SetSize((int)ess_tdofs.size());
    for (size_t i = 0; i < ess_tdofs.size(); i++) ess_tdof_list[(int)i] = ess_tdofs[i];
    if (boundary_values) boundary_values(time, gf);
  }
  void ApplyIntegratedBCs(const std::string &, mfem::ParLinearForm &, mfem::ParMesh *) {}
};
/// the domain load (s_{n+1}, v): assembled on the host by MFEM's linear-form integrators in the reference
class Sources
{
public:
  std::function<void(double, mfem::Vector &)> load;   // (time, lf) -> lf += (s(t), v)
  double time{0.0};
  void Apply(mfem::ParLinearForm * lf) { if (load) load(time, *lf); }
};

class EquationSystem : public mfem::Operator
{
public:
  std::vector<std::string> _trial_var_names, _test_var_names;
  std::vector<mfem::ParFiniteElementSpace *> _test_pfespaces;
  NamedFieldsMap<mfem::ParLinearForm> _lfs;
  std::vector<mfem::Array<int>> _ess_tdof_lists;

  virtual void AddTestVariableNameIfMissing(const std::string & n) { if (!Has(_test_var_names, n)) _test_var_names.push_back(n); }
  virtual void AddTrialVariableNameIfMissing(const std::string & n) { if (!Has(_trial_var_names, n)) _trial_var_names.push_back(n); }
  virtual void AddKernels() {}
  virtual void ApplyBoundaryConditions(BCMap & bc_map)   // equation_system.cpp:116-131
  {
    _ess_tdof_lists.resize(_test_var_names.size());
    for (size_t i = 0; i < _test_var_names.size(); i++)
    {
      *(_xs.at(i)) = 0.0;
      bc_map.ApplyEssentialBCs(_test_var_names[i], _ess_tdof_lists.at(i), *(_xs.at(i)), _test_pfespaces.at(i)->GetParMesh());
      bc_map.ApplyIntegratedBCs(_test_var_names[i], _lfs.GetRef(_test_var_names[i]), _test_pfespaces.at(i)->GetParMesh());
    }
  }
  virtual void Init(GridFunctions & gridfunctions, const FESpaces &, BCMap &, Coefficients &)   // equation_system.cpp:224-285
  {
    for (auto & n : _test_var_names)
    {
      _test_pfespaces.push_back(gridfunctions.Get(n)->ParFESpace());
      _xs.emplace_back(std::make_unique<mfem::ParGridFunction>(gridfunctions.Get(n)->ParFESpace()));
    }
  }
  virtual void BuildLinearForms(BCMap & bc_map, Sources & sources)   // equation_system.cpp:287-322
  {
    for (size_t i = 0; i < _test_var_names.size(); i++)
    {
      _lfs.Register(_test_var_names[i], std::make_shared<mfem::ParLinearForm>(_test_pfespaces.at(i)));
      _lfs.GetRef(_test_var_names[i]) = 0.0;
    }
    ApplyBoundaryConditions(bc_map);
    for (auto & n : _test_var_names) sources.Apply(_lfs.Get(n));
    ApplyLinearFormKernels();
  }
  virtual void ApplyLinearFormKernels() {}   // (the lf kernels: WeakCurlCurlKernel in the H(curl) system)
  virtual void BuildBilinearForms() {}
  virtual void BuildMixedBilinearForms() {}
  virtual void BuildEquationSystem(BCMap & bc_map, Sources & sources)
  {
    BuildLinearForms(bc_map, sources);
    BuildBilinearForms();
    BuildMixedBilinearForms();
  }
  virtual void FormLinearSystem(mfem::OperatorHandle & op, mfem::BlockVector & trueX, mfem::BlockVector & trueRHS) = 0;
  virtual void BuildJacobian(mfem::BlockVector & trueX, mfem::BlockVector & trueRHS)   // equation_system.cpp:192-198
  {
    height = trueX.Size();
    width = trueRHS.Size();
    FormLinearSystem(_jacobian, trueX, trueRHS);
  }
  void Mult(const mfem::Vector & x, mfem::Vector & residual) const override { _jacobian->Mult(x, residual); }
  virtual mfem::Operator & GetGradient(const mfem::Vector &) const { return *_jacobian; }
  virtual void RecoverFEMSolution(mfem::BlockVector & trueX, GridFunctions & gridfunctions)
  {
    for (auto & n : _test_var_names) gridfunctions.Get(n)->Distribute(&trueX.GetBlock(0));
  }

protected:
  static bool Has(const std::vector<std::string> & v, const std::string & n)
  {
    for (auto & s : v) if (s == n) return true;
    return false;
  }
  std::vector<std::unique_ptr<mfem::ParGridFunction>> _xs;
  mutable mfem::OperatorHandle _jacobian;
};

class TimeDependentEquationSystem : public EquationSystem
{
public:
  static std::string GetTimeDerivativeName(std::string name) { return std::string("d") + name + std::string("_dt"); }
  void AddTrialVariableNameIfMissing(const std::string & n) override
  {
    EquationSystem::AddTrialVariableNameIfMissing(n);
    if (!Has(_trial_var_time_derivative_names, GetTimeDerivativeName(n))) _trial_var_time_derivative_names.push_back(GetTimeDerivativeName(n));
  }
  virtual void SetTimeStep(double dt) { if (std::abs(dt - _dt_coef.constant) > 1.0e-12 * dt) _dt_coef.constant = dt; }
  virtual void UpdateEquationSystem(BCMap & bc_map, Sources & sources)   // equation_system.cpp:428-435
  {
    BuildLinearForms(bc_map, sources);
    BuildBilinearForms();
    BuildMixedBilinearForms();
  }
  mfem::ConstantCoefficient _dt_coef{1.0};
  std::vector<std::string> _trial_var_time_derivative_names;
};

class CurlCurlEquationSystem : public TimeDependentEquationSystem   // hcurl_formulation.hpp:36-47
{
public:
  explicit CurlCurlEquationSystem(const InputParameters & params)
    : _h_curl_var_name(params.GetParam<std::string>("HCurlVarName")), _alpha_coef_name(params.GetParam<std::string>("AlphaCoefName")),
      _beta_coef_name(params.GetParam<std::string>("BetaCoefName")), _dtalpha_coef_name(std::string("dt_") + _alpha_coef_name)
  {
  }
  void Init(GridFunctions & gridfunctions, const FESpaces & fespaces, BCMap & bc_map, Coefficients & coefficients) override
  {
    coefficients._scalars.Register(_dtalpha_coef_name, std::make_shared<mfem::TransformedCoefficient>(
                                                           &_dt_coef, coefficients._scalars.Get(_alpha_coef_name), prodFunc));
    TimeDependentEquationSystem::Init(gridfunctions, fespaces, bc_map, coefficients);
  }
  void AddKernels() override   // hcurl_formulation.cpp:130-157 (the kernels themselves are what the GPU system replaces)
  {
    AddTrialVariableNameIfMissing(_h_curl_var_name);
    AddTestVariableNameIfMissing(GetTimeDerivativeName(_h_curl_var_name));
  }
  std::string _h_curl_var_name, _alpha_coef_name, _beta_coef_name, _dtalpha_coef_name;
};
} // namespace hephaestus

inline void
mfem::NewtonSolver::Mult(const Vector & b, Vector & x) const
{
  // mfem::NewtonSolver with iterative_mode = false, max_iter = 1 on a linear operator: x = 0; r = A x - b;
  // c = J^-1 r; x -= c  -- the operator is applied once for the residual, the solver once
  // (time_domain_equation_system_problem_operator.cpp:46-48)
  const int n = b.Size();
  x.SetSize(n);
  Vector r(n), c(n);
  for (int i = 0; i < n; i++) x(i) = 0.0;
  _oper->Mult(x, r);
  for (int i = 0; i < n; i++) r(i) -= b(i);
  const auto * es = dynamic_cast<const hephaestus::EquationSystem *>(_oper);
  _prec->SetOperator(es ? es->GetGradient(x) : *_oper);
  _prec->Mult(r, c);
  for (int i = 0; i < n; i++) x(i) -= c(i);
}
