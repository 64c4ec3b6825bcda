"""The reference's own meshes (committed fixtures, scripts/make_golden.py): BASELINE.json configs
"TEAM7 ... order-1 Nedelec, PCG", "cylinder-hex-q2.gen ... orders 2-3" and the beam-tet MMS mesh.
CPU part: the product's DOF-map generator equals the oracle's.  GPU part: apply / diagonal / PCG
through the C ABI against the oracle (1e-12) and against the committed oracle vectors."""
import numpy as np
import pytest

from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import pcg as opcg
from oracle import space as osp

from helpers import coef_by_attr, golden_mesh, golden_vectors, rel_err

TOL = 1e-12


@pytest.mark.parametrize("name,orders", [("beam_tet", (1, 2, 3)), ("cylinder_hex_q2", (1, 2, 3)), ("team7", (1,))])
def test_dof_maps_identical_on_reference_meshes(name, orders):
    m, mo = golden_mesh(name)
    for p in orders:
        if name == "cylinder_hex_q2" and p == 3:
            continue  # 22 275 dofs through the python-loop oracle space: covered on the GPU side
        sp, so = fespace.NDSpace(m, p), osp.NDSpace(mo, p)
        assert sp.ndof == so.ndof
        assert np.array_equal(sp.signed_elem_dofs(), so.signed_elem_dofs())
        assert np.array_equal(np.sort(sp.boundary_dofs()), so.boundary_dofs())


def _gpu_op(ctx, m, p, alpha, beta, q=0):
    from hephaestus_b200 import operators
    sp = fespace.NDSpace(m, p)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, q)
    op.SetCoefficients(alpha, beta)
    return sp, op


@pytest.mark.gpu
@pytest.mark.parametrize("name,p", [("beam_tet", 1), ("beam_tet", 2), ("beam_tet", 3), ("cylinder_hex_q2", 1),
                                    ("cylinder_hex_q2", 2), ("cylinder_hex_q2", 3), ("team7", 1)])
def test_apply_matches_committed_oracle_vectors(ctx, name, p):
    """unconstrained y = (K(alpha) + M(beta)) x, MFEM's full-assembly rules (HEX27: curl p+1, mass p+3 points)"""
    m, _ = golden_mesh(name)
    gv = golden_vectors()
    alpha, beta = coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
    sp, op = _gpu_op(ctx, m, p, alpha, beta)
    assert sp.ndof == int(gv["%s_p%d_ndof" % (name, p)])
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    y = op.MultUnconstrained(x, np.zeros(sp.ndof), scale=1.0, accumulate=False)
    assert rel_err(y, gv["%s_p%d_y" % (name, p)]) < TOL
    op.SetEssentialTrueDofs(np.zeros(0, np.int32))
    assert rel_err(op.Diagonal(), gv["%s_p%d_diag" % (name, p)]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name,p", [("cylinder_hex_q2", 2), ("team7", 1)])
def test_constrained_apply_and_pcg_on_reference_meshes(ctx, name, p):
    """A-formulation step operator  M(sigma) + K(dt nu)  with the essential tangential BC on every
    boundary face; Jacobi-PCG with hypre semantics against the oracle's, same configuration."""
    from hephaestus_b200 import operators
    m, mo = golden_mesh(name)
    dt = float(np.float32(0.001)) if name == "team7" else 0.05   # examples/team7/Main.cpp:226 (float dt)
    alpha, beta = dt * coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
    sp, op = _gpu_op(ctx, m, p, alpha, beta)
    so = osp.NDSpace(mo, p)
    ess = so.boundary_dofs()
    op.SetEssentialTrueDofs(ess)
    A = oa.assemble(so, alpha, beta)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, ess, x)) < TOL
    Ac, _ = oa.eliminate(A, ess)
    b = Ac @ np.random.default_rng(54321).standard_normal(sp.ndof)
    tol = opcg.as_float_param(1e-8)
    solver = operators.HephGpuPCG(ctx, tol=tol, abstol=0.0, max_iter=5000)
    solver.SetOperator(op)
    xs = solver.Mult(b)
    xo, it_o, _, conv_o, _ = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=tol, atol=0.0, max_iter=5000)
    assert solver.converged and conv_o
    assert abs(solver.num_iterations - it_o) <= max(2, it_o // 100), (solver.num_iterations, it_o)
    assert rel_err(xs, xo) < 1e-5


# ---- BASELINE configs with the reference's own coefficients -------------------------------------------------------------
MU0 = np.pi * 4.0e-7
SIGMA_PLATE = 3.526e7     # /root/reference/examples/team7/Main.cpp:68-89: sigma = 3.526e7 in the plate (attribute 2), 1 elsewhere


@pytest.mark.gpu
@pytest.mark.parametrize("name,p", [("team7", 1), ("beam_tet", 2), ("beam_tet", 3)])
def test_complex_block_apply_fused_tet_kernel(ctx, name, p):
    """ComplexMaxwell block operator [[Ar, -Ai], [Ai, Ar]] (complex_maxwell_formulation.cpp:181-217) through the fused
    one-pass tet kernel.  team7: the coefficients of TestComplexTeam7 (test/integration/test_complex_team7.cpp:70-95:
    f = 200 Hz, mu0, eps = 0, sigma in {1, 3.526e7}) -- Ar = K(1/mu0), Ai = omega M(sigma); beam_tet at orders 2, 3: the
    nd = 20 / nd = 45 variants of the kernel with a dielectric term."""
    from hephaestus_b200 import operators
    m, mo = golden_mesh(name)
    so = osp.NDSpace(mo, p)
    sp = fespace.NDSpace(m, p)
    gsp = operators.GpuNDSpace(ctx, sp)
    w = 2.0 * np.pi * 200.0
    if name == "team7":
        nu, eps, sig = np.full(m.n_elem, 1.0 / MU0), None, np.where(m.attr == 2, SIGMA_PLATE, 1.0)
    else:
        nu, eps, sig = coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1e-7, 3e-7), coef_by_attr(m.attr, 1.0, 50.0)
    beta_re = None if eps is None else -w * w * eps
    op_re = operators.HephGpuOperator(ctx, gsp, 0)
    op_re.SetCoefficients(nu, beta_re)
    op_im = operators.HephGpuOperator(ctx, gsp, 0)
    op_im.SetCoefficients(None, w * sig)
    ess = so.boundary_dofs()
    if name == "team7":
        ess = ess[::3]        # an arbitrary essential set: the reference's TEAM7 problems have natural boundaries
    op_re.SetEssentialTrueDofs(ess)
    op_im.SetEssentialTrueDofs(ess)
    rng = np.random.default_rng(11)
    xr, xi = rng.standard_normal(sp.ndof), rng.standard_normal(sp.ndof)
    yr, yi = operators.complex_mult(op_re, op_im, xr, xi)
    Ar, Ai = oa.assemble(so, nu, beta_re), oa.assemble(so, None, w * sig)
    Arc, _ = oa.eliminate(Ar, ess)
    keep = np.ones(sp.ndof)
    keep[ess] = 0
    Aic = Ai.multiply(keep[:, None]).multiply(keep[None, :]).tocsr()
    assert rel_err(yr, Arc @ xr - Aic @ xi) < TOL
    assert rel_err(yi, Aic @ xr + Arc @ xi) < TOL
    assert np.array_equal(yr[ess], xr[ess]) and np.array_equal(yi[ess], xi[ess])


@pytest.mark.gpu
@pytest.mark.parametrize("max_iter,res_tol,sol_tol", [(200, 0.05, 1e-3), (1000, 1.0, 2e-2)])
def test_team7_transient_step_with_the_reference_coefficients(ctx, max_iter, res_tol, sol_tol):
    """The TEAM7 A-formulation step operator  M(sigma) + K(dt / mu0)  with the example's coefficients and solver options
    (/root/reference/examples/team7/Main.cpp:68-89, 217-229: sigma in {1, 3.526e7}, mu0, float dt = 0.001, float tolerance
    1e-16, 1000 iterations; natural boundary) on data/team7.g, order 1: apply to 1e-12, then Jacobi-PCG with hypre
    semantics against the oracle's.  With a conductivity contrast of 3.5e7 Jacobi-PCG does not converge within the
    iteration limit on either side (the reference uses AMS, out of scope), so both run to the limit; CG is sensitive to
    round-off here -- perturbing b by 1e-15 moves the oracle's own final residual by 1 % after 200 iterations and by 30 %
    after 1000, its solution by 1e-5 / 9e-4 -- and the bounds below are those sensitivities with a margin."""
    from hephaestus_b200 import operators
    m, mo = golden_mesh("team7")
    dt = float(np.float32(0.001))
    alpha, beta = np.full(m.n_elem, dt / MU0), np.where(m.attr == 2, SIGMA_PLATE, 1.0)
    sp, op = _gpu_op(ctx, m, 1, alpha, beta)
    so = osp.NDSpace(mo, 1)
    op.SetEssentialTrueDofs(np.zeros(0, np.int32))
    A = oa.assemble(so, alpha, beta).tocsr()
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    assert rel_err(op.Mult(x), A @ x) < TOL
    b = A @ np.random.default_rng(54321).standard_normal(sp.ndof)
    tol = opcg.as_float_param(1e-16)
    solver = operators.HephGpuPCG(ctx, tol=tol, abstol=0.0, max_iter=max_iter)
    solver.SetOperator(op)
    xs = solver.Mult(b)
    xo, it_o, rel_o, conv_o, _ = opcg.pcg(A, b, diag=A.diagonal(), rtol=tol, atol=0.0, max_iter=max_iter)
    assert not conv_o and not solver.converged
    assert solver.num_iterations == it_o == max_iter
    assert abs(solver.final_rel_res - rel_o) <= res_tol * rel_o, (solver.final_rel_res, rel_o)
    assert rel_err(xs, xo) < sol_tol


@pytest.mark.gpu
def test_team4_order2_apply_against_oracle_samples(ctx):
    """BASELINE.json configs[2]: data/team4.g, order-2 Nedelec tets (127 074 elements, 808 658 dofs -- the size the
    blueprint quotes), batched element-matrix kernel.  The oracle takes two minutes for this matrix, so the fixture holds
    4096 sampled entries of its y = (K(alpha) + M(beta)) x and diagonal plus x.y and |y|^2 (scripts/make_golden.py team4)."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "team4_samples.npz"))
    m, _ = golden_mesh("team4")
    alpha, beta = coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
    sp = fespace.NativeNDSpace(m, 2)
    assert sp.ndof == int(g["ndof"]) == 808658
    from hephaestus_b200 import operators
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    op.SetCoefficients(alpha, beta)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    y = op.MultUnconstrained(x, np.zeros(sp.ndof), scale=1.0, accumulate=False)
    idx = g["idx"]
    assert np.abs(y[idx] - g["y"]).max() <= TOL * float(g["ymax"])
    assert abs(x @ y - float(g["xdoty"])) <= 1e-11 * abs(float(g["xdoty"]))
    assert abs(y @ y - float(g["ynorm2"])) <= 1e-11 * float(g["ynorm2"])
    op.SetEssentialTrueDofs(np.zeros(0, np.int32))
    assert np.abs(op.Diagonal()[idx] - g["diag"]).max() <= TOL * np.abs(g["diag"]).max()
