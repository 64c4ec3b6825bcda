"""hypre PCG with diagonal scaling, as driven by MFEM (oracle; test
infrastructure).

Restates the Krylov loop the reference configures in
``DefaultJacobiPCGSolver`` (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57:
``mfem::HyprePCG`` + ``mfem::HypreDiagScale``; SetTol/SetAbsTol/SetMaxIter) and
in ``ProblemBuilder::ConstructJacobianSolverWithOptions``
(/root/reference/src/problem_builders/problem_builder_base.cpp:240-254)
[hypre 2.28 ``hypre_PCGSolve``, SURVEY.md Appendix A.6]:

    two_norm = 0 (MFEM never sets it):
        bi_prod = <C b, b>;  eps = max(rtol^2, atol^2 / bi_prod)
    r = b - A x;  p = C r;  gamma = <r, p>
    loop i = 1..max_iter:
        s = A p;  alpha = gamma / <s, p>;  x += alpha p;  r -= alpha s
        s = C r;  gamma_new = <r, s>
        if gamma_new / bi_prod < eps: break
        beta = gamma_new / gamma;  p = s + beta p

``Tolerance`` / ``AbsTolerance`` arrive as C ``float`` and are widened to
double, exactly as the reference reads them
(problem_builder_base.cpp:225-233).
"""
import numpy as np


def as_float_param(v):
    """The reference stores solver tolerances as `float`."""
    return float(np.float32(v))


def pcg(A, b, x0=None, diag=None, rtol=1e-9, atol=1e-16, max_iter=1000, matvec=None):
    """Returns (x, iters, final_rel_res, converged, history).
    A: anything supporting `A @ v`, or pass matvec. diag: Jacobi diagonal or None."""
    mv = matvec if matvec is not None else (lambda v: A @ v)
    n = len(b)
    x = np.zeros(n) if x0 is None else x0.copy()

    def C(v):
        return v / diag if diag is not None else v.copy()

    bi_prod = float(np.dot(C(b), b))
    if not bi_prod > 0.0:
        # hypre: rhs is zero -> x = b (= 0)
        return b.copy(), 0, 0.0, True, []
    eps = max(rtol * rtol, atol * atol / bi_prod)
    r = b - mv(x)
    p = C(r)
    gamma = float(np.dot(r, p))
    hist = [np.sqrt(abs(gamma) / bi_prod)]
    i = 0
    converged = False
    while i + 1 <= max_iter:
        i += 1
        s = mv(p)
        sdotp = float(np.dot(s, p))
        if sdotp == 0.0:
            break
        alpha = gamma / sdotp
        x += alpha * p
        r -= alpha * s
        s = C(r)
        gamma_new = float(np.dot(r, s))
        hist.append(np.sqrt(abs(gamma_new) / bi_prod))
        if gamma_new / bi_prod < eps:
            gamma = gamma_new
            converged = True
            break
        if abs(gamma_new) < 1e-292:
            gamma = gamma_new
            break
        beta = gamma_new / gamma
        gamma = gamma_new
        p = s + beta * p
    return x, i, float(np.sqrt(abs(gamma) / bi_prod)), converged, hist
