"""GPU parity tests: the CUDA operator (through the C ABI) against the CPU oracle.
Bar: FP64 relative error <= 1e-12 (BASELINE.json north_star)."""
import os

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import operators
from oracle import assemble as oa

from helpers import coefficients, make_pair, rel_err, star_mesh

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _check_apply(ctx, kind, n, p, q, perturb, scramble=False, ess=True, alpha_on=True, beta_on=True):
    m, sp, mo, so = make_pair(kind, n, p, perturb, scramble=scramble)
    assert np.array_equal(sp.signed_elem_dofs(), so.signed_elem_dofs())
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, q)
    op.SetCoefficients(a if alpha_on else None, b if beta_on else None)
    essd = sp.boundary_dofs(lambda c: c[:, 2] < 1e-9) if ess else np.zeros(0, np.int32)
    op.SetEssentialTrueDofs(essd)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    y = op.Mult(x)
    A = oa.assemble(so, a if alpha_on else None, b if beta_on else None, q_override=q if q else None)
    y_ref = oa.constrained_apply(A, essd, x)
    err = rel_err(y, y_ref)
    assert err < TOL, (kind, n, p, q, perturb, err)
    # diagonal
    d = op.Diagonal()
    Ac, _ = oa.eliminate(A, essd)
    assert rel_err(d, Ac.diagonal()) < TOL
    # unconstrained scaled accumulate
    y0 = np.random.default_rng(1).standard_normal(sp.ndof)
    y2 = op.MultUnconstrained(x, y0.copy(), scale=-0.5, accumulate=True)
    assert rel_err(y2, y0 - 0.5 * (A @ x)) < TOL
    return op, A, essd, sp


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("perturb", [False, True])
def test_hex_apply_mfem_rule(ctx, p, perturb):
    _check_apply(ctx, "hex", 3 if p < 4 else 2, p, 0, perturb)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_hex_apply_fast_rule(ctx, p):
    _check_apply(ctx, "hex", 3 if p < 4 else 2, p, p + 1, True, scramble=True)


@pytest.mark.parametrize("p", [1, 2])
def test_hex_apply_q_p2(ctx, p):
    _check_apply(ctx, "hex", 3, p, p + 2, True)


def test_hex_affine_rules_agree(ctx):
    """On affine hexes with per-element-constant coefficients p+1 points are exact:
    the fast rule and MFEM's rule must give the same operator."""
    m, sp, mo, so = make_pair("hex", 4, 2, perturb=False)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    x = np.random.default_rng(0).standard_normal(sp.ndof)
    ys = []
    for q in (0, 3):
        op = operators.HephGpuOperator(ctx, gsp, q)
        op.SetCoefficients(a, b)
        ys.append(op.Mult(x))
    assert rel_err(ys[0], ys[1]) < TOL


@pytest.mark.parametrize("terms", ["curl", "mass"])
def test_hex_single_term(ctx, terms):
    _check_apply(ctx, "hex", 3, 2, 0, True, alpha_on=terms == "curl", beta_on=terms == "mass")


def test_hex_ragged_batch(ctx):
    """element count not a multiple of the CTA batch; single element."""
    _check_apply(ctx, "hex", (5, 3, 1), 2, 0, True)
    _check_apply(ctx, "hex", (1, 1, 1), 2, 0, False, ess=False)
    _check_apply(ctx, "hex", (1, 1, 1), 1, 0, False, ess=False)


def test_hex_per_qpt_coefficients(ctx):
    m, sp, mo, so = make_pair("hex", 3, 2, True)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)

    def mu(X, attr=None):
        return 1.0 / (1.0 + 0.5 * np.cos(np.pi * X[:, 0]) * np.cos(np.pi * X[:, 1]))

    def sig(X, attr=None):
        return 1.0 + X[:, 2] ** 2
    # sample at the operator's quadrature points (x-fastest tensor order): the oracle's rule order
    (cp, _), (mp, _) = oa.default_rules(so)
    _, xc = oa.hex8_jacobians(mo.elem_coords(), cp)
    _, xm = oa.hex8_jacobians(mo.elem_coords(), mp)
    a = mu(xc.reshape(-1, 3)).reshape(m.n_elem, -1)
    b = sig(xm.reshape(-1, 3)).reshape(m.n_elem, -1)
    assert a.shape[1] == op.qpts_curl and b.shape[1] == op.qpts_mass
    op.SetCoefficients(a, b, layout=L.PER_QPT)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    y = op.Mult(x)
    A = oa.assemble(so, mu, sig)
    assert rel_err(y, A @ x) < TOL
    assert rel_err(op.Diagonal(), A.diagonal()) < TOL


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("scramble", [False, True])
def test_tet_apply(ctx, p, scramble):
    _check_apply(ctx, "tet", 2, p, 0, True, scramble=scramble)


def test_form_rhs(ctx):
    op, A, essd, sp = _check_apply(ctx, "hex", 3, 2, 0, True)
    rng = np.random.default_rng(7)
    xbc, b = rng.standard_normal(sp.ndof), rng.standard_normal(sp.ndof)
    B = op.FormRHS(xbc, b)
    _, B_ref = oa.eliminate(A, essd, xbc, b)
    assert rel_err(B, B_ref) < TOL


def test_complex_apply(ctx):
    """[[Ar,-Ai],[Ai,Ar]] with Ar = K(1/mu) - w^2 M(eps), Ai = w M(sigma)
    (complex_maxwell_formulation.cpp:181-190)."""
    m, sp, mo, so = make_pair("tet", 2, 1, True)
    gsp = operators.GpuNDSpace(ctx, sp)
    rng = np.random.default_rng(11)
    nu, eps, sig = rng.uniform(.5, 2, m.n_elem), rng.uniform(.5, 2, m.n_elem), rng.uniform(.5, 2, m.n_elem)
    w = 2.5
    op_re = operators.HephGpuOperator(ctx, gsp, 0)
    op_re.SetCoefficients(nu, -w * w * eps)
    op_im = operators.HephGpuOperator(ctx, gsp, 0)
    op_im.SetCoefficients(None, w * sig)
    essd = sp.boundary_dofs(lambda c: c[:, 0] < 1e-9)
    op_re.SetEssentialTrueDofs(essd)
    op_im.SetEssentialTrueDofs(essd)
    xr, xi = rng.standard_normal(sp.ndof), rng.standard_normal(sp.ndof)
    yr, yi = operators.complex_mult(op_re, op_im, xr, xi)
    Ar = oa.assemble(so, nu, -w * w * eps)
    Ai = oa.assemble(so, None, w * sig)
    Arc, _ = oa.eliminate(Ar, essd)
    keep = np.ones(sp.ndof)
    keep[essd] = 0
    Aic = Ai.multiply(keep[:, None]).multiply(keep[None, :]).tocsr()
    assert rel_err(yr, Arc @ xr - Aic @ xi) < TOL
    assert rel_err(yi, Aic @ xr + Arc @ xi) < TOL


def test_device_resident_apply_matches_host_path(ctx):
    m, sp, mo, so = make_pair("hex", 4, 2, True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    op.SetCoefficients(a, b)
    op.SetEssentialTrueDofs(sp.boundary_dofs())
    x = np.random.default_rng(2).standard_normal(sp.ndof)
    y_host = op.Mult(x)
    dx, dy = ctx.alloc(sp.ndof).upload(x), ctx.alloc(sp.ndof)
    op.Mult(dx, dy)
    ctx.synchronize()
    assert rel_err(dy.download(), y_host) < 1e-14
    assert op.last_kernel_ms() > 0


def test_errors_are_loud(ctx):
    m, sp, mo, so = make_pair("hex", 2, 1, False)
    gsp = operators.GpuNDSpace(ctx, sp)
    with pytest.raises(L.HcurlError):
        operators.HephGpuOperator(ctx, gsp, 7)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    with pytest.raises(L.HcurlError):
        op.Mult(np.zeros(sp.ndof))  # no coefficients yet
    with pytest.raises(L.HcurlError):
        op.SetEssentialTrueDofs(np.array([sp.ndof + 5], np.int32))


@pytest.mark.parametrize("perturb", [False, True])
@pytest.mark.parametrize("n", [(3, 3, 3), (5, 4, 3), (7, 5, 1)])
def test_hex_p2_fast_paths_match_generic_and_oracle(ctx, perturb, n):
    """Order-2 fast kernels -- patch kernel (hex_patch.cu: modal affine variant) and the
    element-wise register-blocked kernel (hex_fast.cu: affine and trilinear variants) --
    against the generic kernel and the oracle, incl. ragged CTA batches / partial patches."""
    m, sp, mo, so = make_pair("hex", n, 2, perturb, scramble=True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    essd = sp.boundary_dofs(lambda c: c[:, 0] < 1e-9)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    A = oa.assemble(so, a, b, q_override=3)
    y_ref = oa.constrained_apply(A, essd, x)
    y0 = np.random.default_rng(1).standard_normal(sp.ndof)
    yu_ref = y0 - 0.5 * (A @ x)
    for which in (0, 1, 2, 3):
        op = operators.HephGpuOperator(ctx, gsp, 3)
        op.SetKernel(which)
        op.SetCoefficients(a, b)
        op.SetEssentialTrueDofs(essd)
        assert rel_err(op.Mult(x), y_ref) < TOL, which
        # the kernel asked for is the one that ran (patch kernel: affine meshes only)
        if which in (1, 2) or not perturb:
            assert op.LastKernel() == which or (which == 0 and op.LastKernel() == 3), (which, op.LastKernel())
        yu = op.MultUnconstrained(x, y0.copy(), scale=-0.5, accumulate=True)
        assert rel_err(yu, yu_ref) < TOL, which
    # coefficient update reuses the handle
    op.SetKernel(0)
    op.SetCoefficients(2.0 * a, 0.5 * b)
    A2 = oa.assemble(so, 2.0 * a, 0.5 * b, q_override=3)
    assert rel_err(op.Mult(x), oa.constrained_apply(A2, essd, x)) < TOL


@pytest.mark.parametrize("fe", [32, 64, 128])
def test_hex_p2_patch_kernel_patch_sizes(ctx, fe, monkeypatch):
    """Patch kernel with 32 / 64 / 128 elements per patch (128: one stream of four element warps) on a mesh with several patches,
    a skewed (non axis-aligned) affine geometry and a device-resident repeat (determinism of the
    private-dof stores, RED only on patch-boundary dofs)."""
    monkeypatch.setenv("HCURL_PATCH_FE", str(fe))
    from hephaestus_b200 import fespace
    from oracle import mesh as om, space as osp
    n = (9, 8, 6)
    shear = np.array([[1.0, 0.3, 0.1], [0.0, 0.8, 0.2], [0.1, 0.0, 1.2]])
    m = fespace.cartesian_hex_mesh(*n)
    m = fespace.Mesh("hex", m.verts @ shear.T, m.elems, m.attr)
    mo = om.cartesian_hex(*n)
    mo = om.Mesh("hex", mo.verts @ shear.T, mo.elems, mo.attr)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    op.SetKernel(3)
    op.SetCoefficients(a, b)
    essd = sp.boundary_dofs()
    op.SetEssentialTrueDofs(essd)
    x = np.random.default_rng(7).standard_normal(sp.ndof)
    A = oa.assemble(so, a, b, q_override=3)
    y_ref = oa.constrained_apply(A, essd, x)
    y = op.Mult(x)
    assert rel_err(y, y_ref) < TOL
    assert op.LastKernel() == 3
    dx, dy = ctx.alloc(sp.ndof).upload(x), ctx.alloc(sp.ndof)
    for _ in range(3):
        op.Mult(dx, dy)
    ctx.synchronize()
    assert rel_err(dy.download(), y_ref) < TOL
    # an arbitrary essential set splits dof pairs (one double constrained, its partner free)
    ess2 = np.sort(np.random.default_rng(11).choice(sp.ndof, sp.ndof // 5, replace=False)).astype(np.int32)
    op.SetEssentialTrueDofs(ess2)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, ess2, x)) < TOL
    assert op.LastKernel() == 3
    y0 = np.random.default_rng(2).standard_normal(sp.ndof)
    assert rel_err(op.MultUnconstrained(x, y0.copy(), scale=2.0, accumulate=True), y0 + 2.0 * (A @ x)) < TOL


@pytest.mark.parametrize("which", [1, 2, 3])
def test_hex_p2_unstructured_affine_valence6(ctx, which, monkeypatch):
    """Unstructured affine hex mesh (six rhombic prisms around each central edge): patch dofs with
    more than four element slots take the overflow path of the patch kernel."""
    monkeypatch.setenv("HCURL_PATCH_FE", "32")
    from hephaestus_b200 import fespace
    from oracle import mesh as om, space as osp
    m, mo = star_mesh(6), star_mesh(6, cls=om.Mesh)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    assert np.array_equal(sp.signed_elem_dofs(), so.signed_elem_dofs())
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    op.SetKernel(which)
    op.SetCoefficients(a, b)
    essd = sp.boundary_dofs(lambda c: c[:, 2] < 1e-9)
    op.SetEssentialTrueDofs(essd)
    x = np.random.default_rng(3).standard_normal(sp.ndof)
    A = oa.assemble(so, a, b, q_override=3)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, essd, x)) < TOL
    assert op.LastKernel() == which


@pytest.mark.parametrize("n,fe", [((9, 8, 6), 128), ((5, 4, 3), 32), ((12, 11, 7), 64), ((16, 16, 16), 256)])
def test_hex_p1_patch_kernel(ctx, n, fe, monkeypatch):
    """hex_patch_p1_kernel (patch_pipeline.h + hex_patch_p1_elem.h): constrained apply, accumulate, arbitrary essential set, scrambled
    numbering (sign flips), repeated device-resident applies -- same parity bar as the order-2 patch kernel."""
    monkeypatch.setenv("HCURL_PATCH_FE", str(fe))
    m, sp, mo, so = make_pair("hex", n, 1, False, scramble=True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 2)
    op.SetKernel(3)
    op.SetCoefficients(a, b)
    essd = sp.boundary_dofs(lambda c: c[:, 0] < 1e-9)
    op.SetEssentialTrueDofs(essd)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    A = oa.assemble(so, a, b, q_override=2)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, essd, x)) < TOL
    assert op.LastKernel() == 3
    y0 = np.random.default_rng(1).standard_normal(sp.ndof)
    assert rel_err(op.MultUnconstrained(x, y0.copy(), scale=-0.5, accumulate=True), y0 - 0.5 * (A @ x)) < TOL
    ess2 = np.sort(np.random.default_rng(11).choice(sp.ndof, sp.ndof // 5, replace=False)).astype(np.int32)
    op.SetEssentialTrueDofs(ess2)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, ess2, x)) < TOL
    dx, dy = ctx.alloc(sp.ndof).upload(x), ctx.alloc(sp.ndof)
    for _ in range(3):
        op.Mult(dx, dy)
    ctx.synchronize()
    assert rel_err(dy.download(), oa.constrained_apply(A, ess2, x)) < TOL
    # the generic sum-factorised kernel agrees (the path non-affine order-1 meshes take)
    op.SetKernel(1)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, ess2, x)) < TOL
    assert op.LastKernel() == 1


@pytest.mark.parametrize("general", ["0", "1"])
@pytest.mark.parametrize("fe", ["32", "64", "128"])
@pytest.mark.parametrize("n", [(9, 8, 6), (5, 4, 3)])
def test_hex_p2_patch_kernel_variants(ctx, n, fe, general, monkeypatch):
    """hex_patch_p2_kernel with 32- / 64-element patches, reversed-line and general orientation handling, on a scrambled
    vertex numbering (reversed lines)."""
    monkeypatch.setenv("HCURL_PATCH_GENERAL", general)
    monkeypatch.setenv("HCURL_PATCH_FE", fe)
    _p2_patch_variant_parity(ctx, n)


def _p2_patch_variant_parity(ctx, n):
    m, sp, mo, so = make_pair("hex", n, 2, False, scramble=True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    op.SetKernel(3)
    op.SetCoefficients(a, b)
    essd = sp.boundary_dofs(lambda c: c[:, 0] < 1e-9)
    op.SetEssentialTrueDofs(essd)
    x = np.random.default_rng(12345).standard_normal(sp.ndof)
    A = oa.assemble(so, a, b, q_override=3)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, essd, x)) < TOL
    assert op.LastKernel() == 3
    y0 = np.random.default_rng(1).standard_normal(sp.ndof)
    assert rel_err(op.MultUnconstrained(x, y0.copy(), scale=-0.5, accumulate=True), y0 - 0.5 * (A @ x)) < TOL
    ess2 = np.sort(np.random.default_rng(11).choice(sp.ndof, sp.ndof // 5, replace=False)).astype(np.int32)
    op.SetEssentialTrueDofs(ess2)
    assert rel_err(op.Mult(x), oa.constrained_apply(A, ess2, x)) < TOL
    dx, dy = ctx.alloc(sp.ndof).upload(x), ctx.alloc(sp.ndof)
    for _ in range(3):
        op.Mult(dx, dy)
    ctx.synchronize()
    assert rel_err(dy.download(), oa.constrained_apply(A, ess2, x)) < TOL
    # Jacobi-PCG through it (fused <p, A p>)
    s = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=3000)
    op.SetEssentialTrueDofs(sp.boundary_dofs())
    s.SetOperator(op)
    rhs = np.random.default_rng(3).standard_normal(sp.ndof)
    rhs[sp.boundary_dofs()] = 0.0
    xs = s.Mult(rhs)
    assert s.converged
    Ac, _ = oa.eliminate(A, sp.boundary_dofs())
    r = rhs - Ac @ xs
    assert np.sqrt((r * r / Ac.diagonal()).sum() / (rhs * rhs / Ac.diagonal()).sum()) < 10 * s.tol


def test_quad_auto_picks_the_exact_rule_on_affine_meshes(ctx):
    """HCURL_QUAD_AUTO (-1): p+1 rule (patch kernel) on parallelepiped meshes -- equal to MFEM's rules to round-off --
    and MFEM's rules (generic kernel) on a perturbed mesh."""
    for perturb, kernel in ((False, 3), (True, 1)):
        m, sp, mo, so = make_pair("hex", (5, 4, 3), 2, perturb)
        a, b = coefficients(m)
        gsp = operators.GpuNDSpace(ctx, sp)
        op = operators.HephGpuOperator(ctx, gsp, -1)
        op.SetCoefficients(a, b)
        ess = sp.boundary_dofs()
        op.SetEssentialTrueDofs(ess)
        x = np.random.default_rng(4).standard_normal(sp.ndof)
        A = oa.assemble(so, a, b)                       # MFEM's default rules
        assert rel_err(op.Mult(x), oa.constrained_apply(A, ess, x)) < TOL
        assert op.LastKernel() == kernel


@pytest.mark.parametrize("which", ["curl", "mass"])
def test_patch_kernel_single_term_operators(ctx, which):
    """statics (curl-curl only, statics_formulation.cpp:131-161) and mass solves (mass only,
    scaled_vector_gridfunction_aux.cpp:51-60) on affine order-2 hexes take the patch kernel: the absent term's geometry
    factors are zero; a coefficient set earlier and then dropped must not leak in"""
    m, sp, mo, so = make_pair("hex", (6, 5, 4), 2, False, scramble=True)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    op.SetCoefficients(a, b)                                   # both first ...
    x = np.random.default_rng(9).standard_normal(sp.ndof)
    op.SetEssentialTrueDofs(np.zeros(0, np.int32))
    op.Mult(x)
    aa, bb = (a, None) if which == "curl" else (None, b)
    op.SetCoefficients(aa, bb)                                 # ... then only one
    A = oa.assemble(so, aa, bb, q_override=3)
    assert rel_err(op.Mult(x), A @ x) < TOL
    assert op.LastKernel() == 3
    if which == "mass":
        s = operators.HephGpuPCG(ctx, tol=1e-10, max_iter=500)
        s.SetOperator(op)
        rhs = A @ x
        assert rel_err(s.Mult(rhs), x) < 1e-7 and s.converged


def test_batch_apply_matches_single_applies(ctx):
    """hcurl_op_apply_batch (mfem::Operator::ArrayMult): pipelined uploads / applies / downloads give exactly the results
    of one hcurl_op_apply per vector, for more vectors than staging buffers, pinned and pageable host memory."""
    m, sp, mo, so = make_pair("hex", 5, 2, False)
    a, b = coefficients(m)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    op.SetCoefficients(a, b)
    op.SetEssentialTrueDofs(sp.boundary_dofs())
    rng = np.random.default_rng(21)
    xs = [rng.standard_normal(sp.ndof) for _ in range(5)]
    ref = [op.Mult(x) for x in xs]
    ys = op.ArrayMult(xs)
    for y, r in zip(ys, ref):
        assert rel_err(y, r) < 1e-13   # (the patch kernel's three-patch sums are not order-reproducible)
    px = [operators.pinned_array(sp.ndof) for _ in range(2)]
    py = [operators.pinned_array(sp.ndof) for _ in range(2)]
    px[0][:], px[1][:] = xs[0], xs[1]
    op.ArrayMult([px[0], px[1], px[0]], [py[0], py[1], py[0]])
    assert rel_err(py[0], ref[0]) < 1e-13 and rel_err(py[1], ref[1]) < 1e-13
    assert op.ArrayMult([]) == []


def test_patch_apply_run_to_run(ctx):
    """What repeats bit for bit and what does not: a unit (dof pair) written by one patch is stored, one
    shared by two patches receives 0 + a + b -- both independent of the order in which the persistent grid gets to the
    patches; only the units shared by three or more patches (the dof lines along patch edges, ~2 % of the units) receive
    their RED-adds in arrival order and may differ in the last bits between two applies."""
    from hephaestus_b200 import fespace
    from patch_emu_util import build_patches
    n = 24
    m = fespace.cartesian_hex_mesh(n, n, n)
    sp = fespace.NativeNDSpace(m, 2)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, 3)
    a, b = coefficients(m)
    op.SetCoefficients(a, b)
    op.SetEssentialTrueDofs(sp.boundary_dofs())
    x = np.random.default_rng(99).standard_normal(sp.ndof)
    y0 = op.Mult(x)
    assert op.LastKernel() == 3
    P = build_patches(sp, m, 2, 64)
    touched = np.zeros(sp.ndof // 2, np.int32)
    for k in range(len(P["off"]) - 1):
        B = P["blob"][int(P["off"][k]) * 4:]
        np.add.at(touched, (B[32:32 + int(B[0])] & 0x0FFFFFFF).astype(np.int64), 1)
    assert touched.min() >= 1
    multi = np.repeat(touched >= 3, 2)
    assert 0 < multi.mean() < 0.1
    worst = 0.0
    for _ in range(6):
        y = op.Mult(x)
        assert np.array_equal(y[~multi], y0[~multi])
        worst = max(worst, float(np.abs(y[multi] - y0[multi]).max()))
    assert worst <= 64 * np.finfo(float).eps * np.abs(y0).max()
