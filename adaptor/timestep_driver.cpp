// Test driver for adaptor/gpu_curlcurl_equation_system.hpp: the per-step call sequence of
// TimeDomainEquationSystemProblemOperator::ImplicitSolve
// (/root/reference/src/problem_operators/time_domain_equation_system_problem_operator.cpp:30-59) and of
// mfem::BackwardEulerSolver around it, on mock MFEM / Hephaestus objects built from a dump written by
// tests/test_adaptor.py.  Usage: timestep_driver <in.bin> <out.bin>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define HEPH_GPU_MOCK
#include "gpu_curlcurl_equation_system.hpp"

template <class T>
static std::vector<T> rd(FILE * f)
{
  int64_t n = 0;
  if (fread(&n, sizeof(n), 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
  std::vector<T> v((size_t)n);
  if (n && fread(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}
static void wr(FILE * f, const double * p, int64_t n)
{
  fwrite(&n, sizeof(n), 1, f);
  fwrite(p, sizeof(double), (size_t)n, f);
}

int main(int argc, char ** argv)
{
  if (argc < 3) return 2;
  try
  {
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto hdr = rd<int64_t>(f);  // kind (5 hex, 4 tet), order, nd, ndofs, nsteps
    auto dtv = rd<double>(f);   // dt
    mfem::ParMesh mesh;
    mesh.kind = (int)hdr[0];
    mesh.verts = rd<double>(f);
    { auto e = rd<int32_t>(f); mesh.elems.assign(e.begin(), e.end()); }
    { auto a = rd<int32_t>(f); mesh.attr.assign(a.begin(), a.end()); }
    { auto a = rd<int32_t>(f); mesh.elem_faces.assign(a.begin(), a.end()); }
    { auto a = rd<int32_t>(f); mesh.face_verts.assign(a.begin(), a.end()); }
    mfem::FiniteElementCollection fec((int)hdr[1]);
    auto fes = std::make_shared<mfem::ParFiniteElementSpace>();
    fes->mesh = &mesh; fes->fec = &fec; fes->nd = (int)hdr[2]; fes->ndofs = (int)hdr[3];
    { auto a = rd<int32_t>(f); fes->elem_dofs.assign(a.begin(), a.end()); }
    auto nu_v = rd<double>(f), sigma_v = rd<double>(f);  // per attribute
    auto ess = rd<int32_t>(f);
    const int nsteps = (int)hdr[4], n = fes->ndofs;
    auto loads = rd<double>(f), kbc = rd<double>(f);     // [nsteps][n] each
    fclose(f);
    const double dt = dtv[0];

    // what the problem builder registers (A-formulation: alpha = 1/mu = nu, beta = sigma; a_formulation.cpp:1-14)
    hephaestus::Coefficients coefficients;
    coefficients._scalars.Register("magnetic_reluctivity", std::make_shared<mfem::PWConstCoefficient>(nu_v));
    coefficients._scalars.Register("electrical_conductivity", std::make_shared<mfem::PWConstCoefficient>(sigma_v));
    hephaestus::FESpaces fespaces;
    fespaces.Register("_HCurlFESpace", fes);
    hephaestus::GridFunctions gridfunctions;
    gridfunctions.Register("magnetic_vector_potential", std::make_shared<mfem::ParGridFunction>(fes.get()));
    gridfunctions.Register("dmagnetic_vector_potential_dt", std::make_shared<mfem::ParGridFunction>(fes.get()));
    hephaestus::BCMap bc_map;
    bc_map.ess_tdofs.assign(ess.begin(), ess.end());
    int step = 0;
    bc_map.boundary_values = [&](double, mfem::Vector & gf) { for (int32_t d : ess) gf(d) = kbc[(size_t)step * n + d]; };
    hephaestus::Sources sources;
    sources.load = [&](double, mfem::Vector & lf) { for (int i = 0; i < n; i++) lf(i) += loads[(size_t)step * n + i]; };

    auto gpu_ctx = std::make_shared<hephaestus::HephGpuContext>(0, nullptr, 0, 1);
    hephaestus::InputParameters params;
    params.SetParam("HCurlVarName", "magnetic_vector_potential");
    params.SetParam("AlphaCoefName", "magnetic_reluctivity");
    params.SetParam("BetaCoefName", "electrical_conductivity");
    hephaestus::GpuCurlCurlEquationSystem eq(params, gpu_ctx);
    hephaestus::HephGpuPCG pcg(*gpu_ctx, 1e-12f, 0.0f, 5000u, 0);

    // TimeDomainEquationSystemProblemOperator::Init (:18-28)
    eq.AddKernels();
    eq.Init(gridfunctions, fespaces, bc_map, coefficients);
    *gridfunctions.Get("magnetic_vector_potential") = 0.0;
    *gridfunctions.Get("dmagnetic_vector_potential_dt") = 0.0;
    eq.BuildEquationSystem(bc_map, sources);

    mfem::BlockVector true_x(n), true_rhs(n);
    mfem::NewtonSolver newton;
    std::vector<double> iters;
    const hephaestus::HephGpuOperator * first_op = nullptr;
    for (step = 0; step < nsteps; step++)
    {
      const double t = (step + 1) * dt;
      // ImplicitSolve (:30-51)
      coefficients.SetTime(t);
      bc_map.time = sources.time = t;
      eq.SetTimeStep(dt);                                  // BuildEquationSystemOperator (:53-59)
      eq.UpdateEquationSystem(bc_map, sources);
      eq.BuildJacobian(true_x, true_rhs);
      newton.SetSolver(pcg);
      newton.SetOperator(eq);
      newton.Mult(true_rhs, true_x);
      eq.RecoverFEMSolution(true_x, gridfunctions);
      iters.push_back((double)pcg.GetNumIterations());
      if (!pcg.GetConverged()) { fprintf(stderr, "PCG did not converge in step %d\n", step); return 4; }
      // the operator and the device space live across steps
      if (step == 0) first_op = eq.GpuOperator();
      else if (first_op != eq.GpuOperator()) { fprintf(stderr, "operator rebuilt\n"); return 5; }
      // mfem::BackwardEulerSolver::Step: x += dt * k
      mfem::ParGridFunction & u = *gridfunctions.Get("magnetic_vector_potential");
      const mfem::ParGridFunction & k = *gridfunctions.Get("dmagnetic_vector_potential_dt");
      for (int i = 0; i < n; i++) u(i) += dt * k(i);
    }
    FILE * o = fopen(argv[2], "wb");
    wr(o, gridfunctions.Get("magnetic_vector_potential")->_d.data(), n);
    wr(o, gridfunctions.Get("dmagnetic_vector_potential_dt")->_d.data(), n);
    wr(o, iters.data(), (int64_t)iters.size());
    fclose(o);
    return 0;
  }
  catch (const std::exception & e)
  {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
