"""One tested time step of the hot path through the C ABI, on the reference's own meshes and in
the setting its MMS tests pin (/root/reference/test/integration/test_aform_source.cpp:19-49: sigma = 1, nu = 1/mu(x) a
FunctionCoefficient, exact dA/dt = (sin pi y sin pi z, 0, 0), dt = 0.05, essential tangential trace on the whole boundary):

    set_coefficients (per quadrature point)  ->  b = (s_{n+1}, v) - K(nu) u_n  [apply_unconstrained]  ->  form_rhs
    ->  Jacobi-PCG  ->  u += dt k

for two backward-Euler steps against the oracle's restatement (oracle/mms.py + oracle/assemble.py), iteration counts
included.  What stays on the host in the reference stays on the host here: the domain load vector and the projection of
the boundary data (computed with the oracle's routines -- in plugged-in use MFEM's)."""
import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import mms
from oracle import pcg as opcg
from oracle import space as osp

from helpers import golden_mesh, rel_err

DT = 0.05


def _qpoints(so, rule):
    """physical coordinates of a reference rule's points in every element (the ElementTransformation the adaptor uses)"""
    mesh = so.mesh
    X = mesh.elem_coords()
    pts = rule[0]
    if mesh.kind == "tet":
        return oa.tet_jacobians(X, pts)[1]
    if mesh.geom_nodes is not None:
        return oa.hex27_jacobians(mesh.geom_nodes, oa.hex27_ref_coords(mesh.geom_nodes), pts)[1]
    return oa.hex8_jacobians(X, pts)[1]


@pytest.mark.gpu
@pytest.mark.parametrize("name,p,refine", [("beam_tet", 2, 1), ("beam_tet", 1, 1), ("cylinder_hex_q2", 2, 0)])
def test_two_backward_euler_steps_of_the_mms_problem(ctx, name, p, refine):
    from hephaestus_b200 import operators
    m, mo = golden_mesh(name)
    for _ in range(refine):
        mo = mms.refine_tets(mo)
        m = fespace.Mesh("tet", mo.verts, mo.elems, mo.attr)
    sp, so = fespace.NDSpace(m, p), osp.NDSpace(mo, p)
    assert np.array_equal(sp.signed_elem_dofs(), so.signed_elem_dofs())
    nu = lambda X, attr=None: 1.0 / mms.mu_expr(X)  # noqa: E731
    rc, rm = oa.default_rules(so)                    # MFEM's rules on this element type
    xc, xm = _qpoints(so, rc), _qpoints(so, rm)
    ne = mo.n_elem
    nu_c = nu(xc.reshape(-1, 3)).reshape(ne, -1)
    gsp = operators.GpuNDSpace(ctx, sp)
    op, kop = operators.HephGpuOperator(ctx, gsp, 0), operators.HephGpuOperator(ctx, gsp, 0)
    if m.kind == "tet":
        for o in (op, kop):
            o.SetTetRules(rc[0], rc[1], rm[0], rm[1])
    assert op.qpts_curl == len(rc[1]) and op.qpts_mass == len(rm[1])
    op.SetCoefficients(DT * nu_c, np.ones((ne, len(rm[1]))), layout=L.PER_QPT)      # A = K(dt nu) + M(sigma = 1)
    kop.SetCoefficients(nu_c, None, layout=L.PER_QPT)                               # K(nu): the weak curl-curl term
    ess = so.boundary_dofs()
    op.SetEssentialTrueDofs(ess)
    kop.SetEssentialTrueDofs(np.zeros(0, np.int32))
    solver = operators.HephGpuPCG(ctx, tol=float(np.float32(1e-12)), abstol=0.0, max_iter=5000)
    solver.SetOperator(op)
    # oracle
    A = oa.assemble(so, lambda X, a: DT * nu(X), 1.0)
    K = oa.assemble(so, lambda X, a: nu(X), None)
    k_bc = oa.project_tangential(so, mms.adot_exact, dofs=ess)
    u_gpu, u_ref = np.zeros(sp.ndof), np.zeros(sp.ndof)
    for step in range(2):
        t = (step + 1) * DT
        load = mms.load_vector(so, lambda X: mms.source_field(X, t))
        # ---- device: b = load - K u_n; B = FormLinearSystem rhs; solve; u += dt k
        b = kop.MultUnconstrained(u_gpu, load.copy(), scale=-1.0, accumulate=True)
        B = op.FormRHS(k_bc, b)
        k_gpu = solver.Mult(B)
        assert solver.converged
        # ---- oracle
        Ac, Bo = oa.eliminate(A, ess, k_bc, load - K @ u_ref)
        assert rel_err(B, Bo) < 1e-11
        k_ref, it_o, _, conv, _ = opcg.pcg(Ac, Bo, diag=Ac.diagonal(), rtol=float(np.float32(1e-12)), atol=0.0, max_iter=5000)
        assert conv and abs(solver.num_iterations - it_o) <= 2, (step, solver.num_iterations, it_o)
        assert rel_err(k_gpu, k_ref) < 1e-8
        u_gpu = u_gpu + DT * k_gpu
        u_ref = u_ref + DT * k_ref
    assert rel_err(u_gpu, u_ref) < 1e-8
    # and the discretisation error is what the reference's test measures: small against the exact solution
    err = mms.l2_error(so, u_gpu, lambda X: mms.a_exact(X, 2 * DT))
    ref = mms.l2_error(so, np.zeros(sp.ndof), lambda X: mms.a_exact(X, 2 * DT))
    assert err < (0.35 if p >= 2 else 0.8) * ref, (err, ref)
