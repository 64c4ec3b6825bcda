"""GPU PCG (hypre semantics + Jacobi) against the oracle's restatement of
DefaultJacobiPCGSolver (/root/reference/src/solvers/hephaestus_solvers.hpp:34-57)."""
import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import operators
from oracle import assemble as oa
from oracle import pcg as opcg

from helpers import coefficients, make_pair, rel_err

pytestmark = pytest.mark.gpu


def _setup(ctx, kind, n, p, alpha_scale=1.0):
    m, sp, mo, so = make_pair(kind, n, p, True)
    a, b = coefficients(m)
    a = a * alpha_scale
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    op.SetCoefficients(a, b)
    ess = sp.boundary_dofs()
    op.SetEssentialTrueDofs(ess)
    A = oa.assemble(so, a, b)
    Ac, _ = oa.eliminate(A, ess)
    return op, Ac, sp


@pytest.mark.parametrize("kind,n,p", [("hex", 4, 1), ("hex", 4, 2), ("hex", 3, 3), ("tet", 3, 1), ("tet", 2, 2)])
def test_pcg_matches_oracle(ctx, kind, n, p):
    op, Ac, sp = _setup(ctx, kind, n, p, alpha_scale=0.05)
    b = Ac @ np.random.default_rng(54321).standard_normal(sp.ndof)
    s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=5000)
    s.SetOperator(op)
    x = s.Mult(b)
    xo, it, res, conv, hist = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=5000)
    assert s.converged and conv
    assert abs(s.num_iterations - it) <= max(2, it // 50), (s.num_iterations, it)
    assert rel_err(x, xo) < 1e-6
    # true residual meets the tolerance in the C-norm
    r = b - Ac @ x
    assert np.sqrt((r * r / Ac.diagonal()).sum() / (b * b / Ac.diagonal()).sum()) < 10 * s.tol


def test_pcg_mass_only_and_unpreconditioned(ctx):
    """The configuration the reference itself runs with Jacobi-PCG: H(curl) mass solves
    (/root/reference/src/auxsolvers/scaled_vector_gridfunction_aux.cpp:51-60)."""
    m, sp, mo, so = make_pair("hex", 4, 2, True)
    _, b_ = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    op.SetCoefficients(None, b_)
    op.SetEssentialTrueDofs(np.zeros(0, np.int32))
    M = oa.assemble(so, None, b_)
    rhs = M @ np.random.default_rng(1).standard_normal(sp.ndof)
    for jac in (True, False):
        s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=1000, jacobi=jac)
        s.SetOperator(op)
        x = s.Mult(rhs)
        xo, it, _, conv, _ = opcg.pcg(M, rhs, diag=M.diagonal() if jac else None, rtol=s.tol, atol=s.abstol)
        assert s.converged and conv and abs(s.num_iterations - it) <= 2
        assert rel_err(x, xo) < 1e-7


def test_pcg_max_iter_and_zero_rhs(ctx):
    op, Ac, sp = _setup(ctx, "hex", 3, 2)
    s = operators.HephGpuPCG(ctx, tol=1e-14, max_iter=3)
    s.SetOperator(op)
    b = Ac @ np.ones(sp.ndof)
    s.Mult(b)
    assert s.num_iterations == 3 and not s.converged
    x = s.Mult(np.zeros(sp.ndof))
    assert s.num_iterations == 0 and np.all(x == 0)


def test_pcg_reuse_after_coefficient_update(ctx):
    """dt / Coefficients::SetTime change: same operator handle, new values
    (equation_system.cpp:413-435)."""
    m, sp, mo, so = make_pair("hex", 3, 2, True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    ess = sp.boundary_dofs()
    op.SetEssentialTrueDofs(ess)
    s = operators.HephGpuPCG(ctx, tol=1e-10, max_iter=5000)
    for dt in (0.05, 0.2):
        op.SetCoefficients(dt * a, b)
        s.SetOperator(op)
        Ac, _ = oa.eliminate(oa.assemble(so, dt * a, b), ess)
        rhs = Ac @ np.random.default_rng(4).standard_normal(sp.ndof)
        x = s.Mult(rhs)
        xo, it, _, _, _ = opcg.pcg(Ac, rhs, diag=Ac.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=5000)
        assert abs(s.num_iterations - it) <= 2 and rel_err(x, xo) < 1e-6


def test_pcg_order1_patch_kernel_fused_dot(ctx, monkeypatch):
    """Jacobi-PCG on affine order-1 hexes through the order-1 patch kernel: <p, A p> comes from the kernel's fused dot
    product; iteration count and solution as the oracle's hypre-semantics PCG."""
    m, sp, mo, so = make_pair("hex", (12, 10, 9), 1, False)
    a, b = coefficients(m)
    a = a * 0.05
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 2)
    op.SetCoefficients(a, b)
    ess = sp.boundary_dofs()
    op.SetEssentialTrueDofs(ess)
    Ac, _ = oa.eliminate(oa.assemble(so, a, b, q_override=2), ess)
    rhs = Ac @ np.random.default_rng(54321).standard_normal(sp.ndof)
    s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=5000)
    s.SetOperator(op)
    x = s.Mult(rhs)
    assert op.LastKernel() == 3
    xo, it, res, conv, hist = opcg.pcg(Ac, rhs, diag=Ac.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=5000)
    assert s.converged and conv
    assert abs(s.num_iterations - it) <= max(2, it // 50), (s.num_iterations, it)
    assert rel_err(x, xo) < 1e-6


@pytest.mark.parametrize("kind,n,p,order", [("hex", 6, 2, 3), ("hex", 8, 1, 2), ("tet", 3, 2, 4), ("hex", 6, 2, 1)])
def test_pcg_chebyshev_matches_oracle(ctx, kind, n, p, order):
    """HCURL_PC_CHEBYSHEV against the oracle's restatement (oracle/pcg.py: power_lambda_max, chebyshev): eigenvalue
    estimate, iteration count, solution; fewer iterations than Jacobi for order >= 2; order 1 with the interval's
    1/theta scaling is Jacobi up to a constant, i.e. the same iteration count."""
    op, Ac, sp = _setup(ctx, kind, n, p, alpha_scale=0.05)
    b = Ac @ np.random.default_rng(54321).standard_normal(sp.ndof)
    d = Ac.diagonal()
    s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=5000, chebyshev_order=order)
    s.SetOperator(op)
    x = s.Mult(b)
    lam = opcg.power_lambda_max(Ac, d)
    assert abs(s.lambda_max() - lam) < 1e-9 * lam
    xo, it, res, conv, hist = opcg.pcg(Ac, b, precond=opcg.chebyshev(Ac, d, order, lam), rtol=s.tol, atol=s.abstol, max_iter=5000)
    assert s.converged and conv
    assert abs(s.num_iterations - it) <= max(2, it // 50), (s.num_iterations, it)
    assert rel_err(x, xo) < 1e-6
    _, it_j, _, _, _ = opcg.pcg(Ac, b, diag=d, rtol=s.tol, atol=s.abstol, max_iter=5000)
    if order >= 2:
        assert s.num_iterations < it_j
    else:
        assert abs(s.num_iterations - it_j) <= 2
    # reuse with new coefficients: the estimate is refreshed
    lam0 = s.lambda_max()
    op.SetCoefficients(*[c * f for c, f in zip(coefficients(sp.mesh), (0.5, 3.0))])
    s.SetOperator(op)
    s.Mult(b)
    assert s.converged and s.lambda_max() != lam0


def test_graph_loop_is_used_and_equals_the_host_loop(ctx):
    """Element-wise apply kernels (tets; perturbed hexes), single rank, Jacobi: the Krylov loop runs as one CUDA graph
    with a device-side WHILE node.  Same iteration count and solution as the oracle's hypre-semantics loop, the iteration
    limit is honoured, a second solve replays the graph, and a change of the essential dofs rebuilds it."""
    op, Ac, sp = _setup(ctx, "tet", 3, 1, alpha_scale=0.05)
    b = Ac @ np.random.default_rng(54321).standard_normal(sp.ndof)
    s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=5000)
    s.SetOperator(op)
    x = s.Mult(b)
    assert s.last_used_graph()
    xo, it, res, conv, hist = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=5000)
    assert s.converged and conv and abs(s.num_iterations - it) <= 2
    assert rel_err(x, xo) < 1e-6
    assert s.final_rel_res < s.tol
    x2 = s.Mult(2.0 * b)                      # replay
    assert s.last_used_graph() and s.converged and rel_err(x2, 2.0 * xo) < 1e-6
    s.max_iter = 7                            # iteration limit
    L.check(L.lib().hcurl_pcg_set(s.h, s.tol, s.abstol, 7, 0), "hcurl_pcg_set")
    s.Mult(b)
    assert s.num_iterations == 7 and not s.converged
    _, it7, res7, _, _ = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=7)
    assert abs(s.final_rel_res - res7) <= 1e-6 * res7
    # other essential dofs: the operator's device maps are reallocated, the graph is rebuilt
    ess2 = sp.boundary_dofs()[::2]
    op.SetEssentialTrueDofs(ess2)
    L.check(L.lib().hcurl_pcg_set(s.h, s.tol, s.abstol, 5000, 0), "hcurl_pcg_set")
    s.SetOperator(op)
    from oracle import assemble as oa2
    m, sp_, mo, so = make_pair("tet", 3, 1, True)
    # (the same mesh and coefficients as _setup: rebuild the eliminated matrix for the new essential set)
    a, bb = coefficients(m)
    A2, _ = oa2.eliminate(oa2.assemble(so, 0.05 * a, bb), ess2)
    b2 = A2 @ np.random.default_rng(5).standard_normal(sp.ndof)
    x3 = s.Mult(b2)
    xo3, it3, _, conv3, _ = opcg.pcg(A2, b2, diag=A2.diagonal(), rtol=s.tol, atol=s.abstol, max_iter=5000)
    assert s.last_used_graph() and s.converged and conv3 and abs(s.num_iterations - it3) <= 2
    assert rel_err(x3, xo3) < 1e-6
