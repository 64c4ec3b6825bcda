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
