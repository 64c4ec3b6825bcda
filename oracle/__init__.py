"""CPU oracle for the H(curl) hot path -- TEST INFRASTRUCTURE ONLY.

This package is a from-scratch CPU restatement (numpy/scipy + a little C) of
the algorithm the reference runs for the path in SURVEY.md section 8:

  * ``mfem::CurlCurlIntegrator`` / ``mfem::VectorFEMassIntegrator`` element
    matrices as selected by ``CurlCurlKernel::Apply``
    (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24) and
    ``VectorFEMassKernel::Apply``
    (/root/reference/src/kernels/vector_fe_mass_kernel.cpp:21-24);
  * full assembly + essential-BC elimination as in
    ``EquationSystem::FormLinearSystem``
    (/root/reference/src/equation_systems/equation_system.cpp:132-190);
  * the assembled operator apply ``EquationSystem::Mult`` (ibid. :200-204);
  * hypre PCG + diagonal scaling as configured by ``DefaultJacobiPCGSolver``
    (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57).

The arithmetic itself lives in third-party dependencies that are NOT vendored
under /root/reference: MFEM (github.com/Heinrich-BR/mfem, branch
SubmeshBoundary, no pinned SHA; a fork of MFEM ~4.5/4.6 --
/root/reference/docker/hephaestus-deps/Dockerfile:116-138) and hypre 2.28.x
(ibid. :77-104).  Their published algorithms are restated here.

PARITY PINNING STATUS:
  * pinned by reference-held tests: the known answer of TestComplexIrisWaveguide
    for ND1 hexes (/root/reference/test/integration/test_complex_iris_wg.cpp:191-195,
    reproduced to 7 digits by iris_kat.py), the curl-curl element-matrix formula
    (/root/reference/test/unit/test_nlintegrators.cpp:47-59) and the ND2-tet
    MMS convergence windows of TestAForm / TestHFormSource
    (/root/reference/test/integration/test_aform_source.cpp:130-208);
  * hex order >= 2 operator values (same order-generic code as the pinned
    order 1, no reference test of their own), MFEM's global DOF numbering and
    the discrete curl / gradient local matrices (interp.py): PARITY UNPINNED
    (no reference test holds values for them; MFEM cannot be built here).
    They are validated by mathematical invariants instead.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product package
(hephaestus_b200/) never does.
"""
