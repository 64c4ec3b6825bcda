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


def start_vector(n, offset=0):
    """deterministic pseudo-random start vector of the power method, values in [0.5, 1.5): the same integer hash as
    cheb_seed_kernel (hephaestus_b200/csrc/api.cu)"""
    i = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) & np.uint64(0xFFFFFFFF)
    h = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return 0.5 + h.astype(np.float64) / 4294967296.0


def power_lambda_max(A, diag, n_iter=10, matvec=None):
    """Largest eigenvalue of D^-1 A by n_iter steps of the power method, Rayleigh quotient in the D inner product
    (a lower bound that the Chebyshev interval then widens by 1.2, as MFEM's OperatorChebyshevSmoother and hypre do)."""
    mv = matvec if matvec is not None else (lambda v: A @ v)
    v = start_vector(len(diag))
    lam = 0.0
    for _ in range(n_iter):
        w = mv(v)
        lam = float(np.dot(v, w) / np.dot(v, diag * v))
        v = w / diag
        v /= np.sqrt(float(np.dot(v, diag * v)))
    return lam


def chebyshev(A, diag, order, lam_max, matvec=None):
    """z = q(D^-1 A) D^-1 r with q the degree-(order-1) Chebyshev polynomial for the interval
    [0.3, 1.2] * lam_max (the interval MFEM's OperatorChebyshevSmoother / hypre's Chebyshev smoother use), written as
    the three-term Chebyshev iteration from z = 0 (Saad, Iterative Methods, Alg. 12.1): order-1 operator applies.
    The B200 build's stand-in for the reference's AMS preconditioner, which is out of scope."""
    mv = matvec if matvec is not None else (lambda v: A @ v)
    lo, hi = 0.3 * lam_max, 1.2 * lam_max
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta

    def C(r):
        res = r.copy()
        d = res / diag / theta
        z = d.copy()
        rho = 1.0 / sigma
        for _ in range(order - 1):
            res -= mv(d)
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * (res / diag)
            z += d
            rho = rho_new
        return z
    return C


def pcg(A, b, x0=None, diag=None, rtol=1e-9, atol=1e-16, max_iter=1000, matvec=None, precond=None):
    """Returns (x, iters, final_rel_res, converged, history).
    A: anything supporting `A @ v`, or pass matvec. diag: Jacobi diagonal or None; precond: callable r -> C r
    (overrides diag)."""
    mv = matvec if matvec is not None else (lambda v: A @ v)
    n = len(b)
    x = np.zeros(n) if x0 is None else x0.copy()

    def C(v):
        if precond is not None:
            return precond(v)
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
