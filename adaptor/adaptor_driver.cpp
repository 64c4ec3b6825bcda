// Test driver: builds mock MFEM objects from a dump written by tests/test_adaptor.py, then runs
// the adaptor exactly as Hephaestus would (operator apply, FormRHS, NewtonSolver-style solve)
// and writes the results back.  Usage: adaptor_driver <in.bin> <out.bin>
//        adaptor_driver --tensor-dofs <in.bin> <out.bin>   host only: HephGpuDiscreteInterpolator::TensorDofs of a mock
//        space (header: nd, ne; arrays: dof_map [nd], elem_dofs [ne][nd]) -> [ne][nd] as doubles
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "mock_mfem.hpp"
#include "heph_gpu_operator.hpp"

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

static int tensor_dofs_mode(const char * in, const char * out)
{
  FILE * f = fopen(in, "rb");
  if (!f) return 2;
  auto hdr = rd<int64_t>(f);
  auto dm = rd<int32_t>(f), ed = rd<int32_t>(f);
  fclose(f);
  mfem::ParMesh mesh;
  mesh.attr.assign((size_t)hdr[1], 1);
  mfem::ParFiniteElementSpace fes;
  fes.mesh = &mesh; fes.nd = (int)hdr[0];
  fes.elem_dofs.assign(ed.begin(), ed.end());
  for (int32_t v : dm) fes.fe._dof_map.Append(v);
  std::vector<int32_t> t;
  hephaestus::HephGpuDiscreteInterpolator::TensorDofs(fes, t);
  std::vector<double> d(t.begin(), t.end());
  FILE * o = fopen(out, "wb");
  wr(o, d.data(), (int64_t)d.size());
  fclose(o);
  return 0;
}

int main(int argc, char ** argv)
{
  if (argc < 3) return 2;
  try
  {
    if (std::string(argv[1]) == "--tensor-dofs" && argc == 4) return tensor_dofs_mode(argv[2], argv[3]);
    FILE * f = fopen(argv[1], "rb");
    if (!f) return 2;
    auto hdr = rd<int64_t>(f);  // kind (5 hex, 4 tet), order, nd, ndofs, quad
    mfem::ParMesh mesh;
    mesh.kind = (int)hdr[0];
    mesh.verts = rd<double>(f);
    { auto e = rd<int32_t>(f); mesh.elems.assign(e.begin(), e.end()); }
    { auto a = rd<int32_t>(f); mesh.attr.assign(a.begin(), a.end()); }
    { auto a = rd<int32_t>(f); mesh.elem_faces.assign(a.begin(), a.end()); }
    { auto a = rd<int32_t>(f); mesh.face_verts.assign(a.begin(), a.end()); }
    mfem::FiniteElementCollection fec((int)hdr[1]);
    mfem::ParFiniteElementSpace fes;
    fes.mesh = &mesh; fes.fec = &fec; fes.nd = (int)hdr[2]; fes.ndofs = (int)hdr[3];
    { auto a = rd<int32_t>(f); fes.elem_dofs.assign(a.begin(), a.end()); }
    auto alpha_v = rd<double>(f), beta_v = rd<double>(f);  // per attribute
    auto ess = rd<int32_t>(f);
    auto x = rd<double>(f), xbc = rd<double>(f), b = rd<double>(f);
    fclose(f);

    hephaestus::HephGpuContext ctx(0, nullptr, 0, 1);
    hephaestus::HephGpuSpace space(ctx, fes);
    mfem::PWConstCoefficient alpha(alpha_v), beta(beta_v);
    mfem::Array<int> ess_tdof;
    for (int32_t e : ess) ess_tdof.Append(e);
    hephaestus::HephGpuOperator A(ctx, space, &alpha, &beta, ess_tdof, (int)hdr[4]);

    mfem::Vector X(fes.ndofs), Y(fes.ndofs), XBC(fes.ndofs), B(fes.ndofs), RHS(fes.ndofs), SOL(fes.ndofs);
    X._d = x; XBC._d = xbc; B._d = b;
    A.Mult(X, Y);                      // EquationSystem::Mult
    {                                  // mfem::Operator::ArrayMult: pipelined batch, same results
      mfem::Vector Y2(fes.ndofs), Y3(fes.ndofs), Y4(fes.ndofs);
      std::vector<const mfem::Vector *> XS{&X, &XBC, &X};
      std::vector<mfem::Vector *> YS{&Y2, &Y3, &Y4};
      A.ArrayMult(XS, YS);
      double d = 0.0, nrm = 0.0;
      for (int i = 0; i < fes.ndofs; i++) { d = std::max(d, std::fabs(Y2._d[(size_t)i] - Y._d[(size_t)i])); d = std::max(d, std::fabs(Y4._d[(size_t)i] - Y._d[(size_t)i])); nrm = std::max(nrm, std::fabs(Y._d[(size_t)i])); }
      if (!(d <= 1e-12 * nrm)) { fprintf(stderr, "ArrayMult differs from Mult: %g (scale %g)\n", d, nrm); return 4; }
    }
    A.FormRHS(XBC, B, RHS);            // FormLinearSystem right-hand side
    hephaestus::HephGpuPCG pcg(ctx, 1e-9f, 1e-16f, 2000u, 0);
    pcg.iterative_mode = false;        // NewtonSolver::SetSolver
    pcg.SetOperator(A);                // NewtonSolver: prec->SetOperator(GetGradient(x))
    pcg.Mult(RHS, SOL);
    double info[3] = {(double)pcg.GetNumIterations(), pcg.GetFinalRelResidual(), pcg.GetConverged() ? 1.0 : 0.0};

    FILE * o = fopen(argv[2], "wb");
    wr(o, Y._d.data(), fes.ndofs);
    wr(o, RHS._d.data(), fes.ndofs);
    wr(o, SOL._d.data(), fes.ndofs);
    wr(o, info, 3);
    fclose(o);
    // a non-GPU operator must be refused like HyprePCG refuses a non-HypreParMatrix
    struct Dummy : mfem::Operator { Dummy() : mfem::Operator(3) {} void Mult(const mfem::Vector &, mfem::Vector &) const override {} } dummy;
    bool threw = false;
    try { pcg.SetOperator(dummy); } catch (const std::runtime_error &) { threw = true; }
    return threw ? 0 : 3;
  }
  catch (const std::exception & e)
  {
    fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
