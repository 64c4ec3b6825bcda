"""Pinning the CPU oracle (no GPU).  The reference holds no operator-level golden vectors for this
path; what its own tests do hold, and what is checked here:

  * the curl-curl element-matrix formula spelled out in
    /root/reference/test/unit/test_nlintegrators.cpp:36-59  (w = weight/detJ * q;
    curlshape * J^T; elmat += w * A A^T), restated point by point;
  * the MMS convergence window of TestAForm (ND2 tets on beam-tet.mesh x {0,1,2} refinements,
    dt = 0.05, one backward-Euler step, rate in (1.85, 3.0)):
    /root/reference/test/integration/test_aform_source.cpp:130-208;
  * mathematical invariants of the discretisation -- these are what
    validate hex orders >= 2, whose VALUES no reference test pins ("parity unpinned");
  * the committed regression vectors tests/golden/oracle_vectors.npz (oracle vs its own earlier
    output on the reference's meshes: guards against drift, not an MFEM pin).
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import assemble as oa
from oracle import mesh as om
from oracle import mms
from oracle import pcg as opcg
from oracle import space as osp

from helpers import coef_by_attr, golden_mesh, golden_vectors


def _small(kind, perturb, n=2):
    m = om.cartesian_hex(n, n, n, 1.0, 0.7, 1.3)
    if perturb:
        m = om.perturb_interior(m)
    return om.hex_to_tets(m) if kind == "tet" else m


@pytest.mark.parametrize("kind,p", [("hex", 1), ("hex", 2), ("hex", 3), ("tet", 1), ("tet", 2), ("tet", 3)])
def test_curlcurl_element_formula_of_reference_unit_test(kind, p):
    """elmat = sum_q (w_q / detJ_q) q (C J^T)(C J^T)^T, looping points as test_nlintegrators.cpp:47-59 does."""
    mesh = _small(kind, True)
    space = osp.NDSpace(mesh, p)
    q = np.random.default_rng(0).uniform(0.5, 2.0, mesh.n_elem)
    (cp, cw), _ = oa.default_rules(space)
    _, C = space.fe.shape_and_curl(cp)
    X = mesh.elem_coords()
    got = {e0: A for e0, _, A in oa.element_matrices(space, q, None, chunk=1)}
    for e in range(0, mesh.n_elem, max(1, mesh.n_elem // 7)):
        elmat = np.zeros((space.fe.nd, space.fe.nd))
        for i in range(len(cw)):
            J = (oa.tet_jacobians(X[e:e + 1], cp[i:i + 1]) if kind == "tet" else oa.hex8_jacobians(X[e:e + 1], cp[i:i + 1]))[0][0, 0]
            w = cw[i] / np.linalg.det(J) * q[e]
            dFt = C[i] @ J.T                      # MultABt(curlshape, J, curlshape_dFt)
            elmat += w * dFt @ dFt.T              # AddMult_a_AAt
        assert np.allclose(got[e][0], elmat, rtol=1e-13, atol=1e-13 * np.abs(elmat).max())


@pytest.mark.parametrize("kind,p,perturb", [("hex", 1, True), ("hex", 2, True), ("hex", 3, False), ("hex", 4, False),
                                            ("tet", 1, True), ("tet", 2, True), ("tet", 3, True)])
def test_constants_reproduced_and_energy(kind, p, perturb):
    """A.7 (iii): the interpolant of a constant field c has curl 0 and mass energy beta |c|^2 |Omega|."""
    mesh = _small(kind, perturb)
    space = osp.NDSpace(mesh, p)
    c = np.array([0.3, -1.1, 0.7])
    x = oa.project_tangential(space, lambda X: np.broadcast_to(c, X.shape))
    K = oa.assemble(space, 1.0, None)
    M = oa.assemble(space, None, 2.5)
    assert np.abs(K @ x).max() < 1e-11 * np.abs(K).max()
    vol = 1.0 * 0.7 * 1.3
    assert abs(x @ (M @ x) - 2.5 * (c @ c) * vol) < 1e-11
    # symmetry / definiteness (A.7 ii)
    assert abs(K - K.T).max() < 1e-12 * abs(K).max() and abs(M - M.T).max() < 1e-12 * abs(M).max()
    if space.ndof < 700:
        assert np.linalg.eigvalsh(M.toarray()).min() > 0
        assert np.linalg.eigvalsh(K.toarray()).min() > -1e-10 * abs(K).max()


@pytest.mark.parametrize("kind,p", [("hex", 1), ("hex", 2), ("hex", 3), ("tet", 1), ("tet", 2), ("tet", 3)])
def test_curl_of_gradient_vanishes(kind, p):
    """A.7 (i): gradients of degree-p polynomials lie in ND_p on affine elements; K grad(phi) = 0."""
    mesh = _small(kind, False)
    space = osp.NDSpace(mesh, p)

    def grad(X):  # phi = x^p y + z^p - x y z (only terms of per-variable degree <= p for p >= 1)
        x, y, z = X[:, 0], X[:, 1], X[:, 2]
        if p == 1:
            if kind == "tet":  # P1: phi linear
                return np.broadcast_to(np.array([0.2, -0.4, 1.0]), X.shape)
            return np.stack([y * z, x * z, x * y], 1)  # Q1: phi = x y z
        if kind == "tet":      # P_p: total degree p
            return np.stack([p * x ** (p - 1) + y, x + 0 * y, p * z ** (p - 1)], 1)
        return np.stack([p * x ** (p - 1) * y - y * z, x ** p - x * z, p * z ** (p - 1) - x * y], 1)

    x = oa.project_tangential(space, grad)
    K = oa.assemble(space, 1.0, None)
    assert np.abs(K @ x).max() < 1e-10 * np.abs(K).max() * max(1.0, np.abs(x).max())
    assert np.abs(x).max() > 0.1


def test_tangential_continuity_under_vertex_renumbering():
    """A.7 (v): x^T A x of an interpolant does not depend on vertex numbering (orientation handling)."""
    for kind, p in (("hex", 2), ("tet", 2), ("tet", 3)):
        mesh = _small(kind, True)
        perm = np.random.default_rng(5).permutation(len(mesh.verts))
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        mesh2 = om.Mesh(kind, mesh.verts[perm], inv[mesh.elems], mesh.attr)
        f = lambda X: np.stack([np.sin(X[:, 1]), X[:, 0] * X[:, 2], np.cos(X[:, 0])], 1)  # noqa: E731
        e = []
        for m in (mesh, mesh2):
            s = osp.NDSpace(m, p)
            x = oa.project_tangential(s, f)
            A = oa.assemble(s, 0.7, 1.3)
            e.append(x @ (A @ x))
        assert abs(e[0] - e[1]) < 1e-11 * abs(e[0])


def _mms_step(mesh, p, dt=0.05):
    """One backward-Euler step of the A-formulation from A_0 = 0 (test_aform_source.cpp: sigma = 1,
    nu = 1/mu(x), all boundary attributes essential with the tangential trace of dA/dt)."""
    space = osp.NDSpace(mesh, p)
    nu = lambda X, attr: 1.0 / mms.mu_expr(X)  # noqa: E731
    A = oa.assemble(space, lambda X, a: dt * nu(X, a), 1.0)
    b = mms.load_vector(space, lambda X: mms.source_field(X, dt))
    ess = space.boundary_dofs()
    k_bc = oa.project_tangential(space, mms.adot_exact, dofs=ess)
    Ac, B = oa.eliminate(A, ess, k_bc, b)
    k = spla.spsolve(Ac.tocsc(), B)
    # the reference's solver configuration for this path, on the same system
    kp, it, _, conv, _ = opcg.pcg(Ac, B, diag=Ac.diagonal(), rtol=1e-12, atol=0.0, max_iter=5000)
    assert conv and np.linalg.norm(kp - k) < 1e-8 * np.linalg.norm(k)
    return space.ndof, mms.l2_error(space, dt * k, lambda X: mms.a_exact(X, dt))


def test_mms_rate_window_of_TestAForm():
    """ND2 tets, beam-tet.mesh x {0, 1, 2} uniform refinements; the reference accepts 2 - 0.15 < r < 2 + 1.0 for BOTH
    level pairs (test/integration/test_aform_source.cpp:198-207), and so does this test: 1.921 and 1.977.
    The coarse pair (h = 1 -> 1/2, half a sine period per element: pre-asymptotic) depends on HOW the inner octahedron of
    every refined tet is split: with MFEM's rule -- best child aspect ratio measured against the equilateral tetrahedron
    (Mesh::UniformRefinement3D_base, "best aspect ratio" via JacToPerfJac [MFEM-upstream]) -- it is 1.92; measuring the
    aspect ratio against the right-angled reference tetrahedron instead (this oracle's first restatement) or taking the
    shortest diagonal gives 1.80, outside the window, and a fixed second or third diagonal 1.37."""
    _, mesh0 = golden_mesh("beam_tet")
    mesh, nd, err = mesh0, [], []
    for lvl in range(3):
        n, e = _mms_step(mesh, 2)
        nd.append(n)
        err.append(e)
        if lvl < 2:
            mesh = mms.refine_tets(mesh)
    assert nd == [494, 3152, 22256]
    rate = [np.log(err[i] / err[i - 1]) / np.log((nd[i - 1] / nd[i]) ** (1.0 / 3)) for i in (1, 2)]  # EstimateConvergenceRate
    for r in rate:
        assert 2 - 0.15 < r < 2 + 1.0, (nd, err, rate)
    # the sensitivity that makes the coarse pair a pin of the refinement rule as well
    for rule, lo, hi in (("best", 1.75, 1.85), (1, 1.30, 1.45)):
        n1, e1 = _mms_step(mms.refine_tets(mesh0, rule), 2)
        r01 = np.log(e1 / err[0]) / np.log((nd[0] / n1) ** (1.0 / 3))
        assert n1 == nd[1] and lo < r01 < hi, (rule, r01)


def test_mms_TestHFormSource_is_the_same_discrete_problem():
    """Second MMS pin of the reference (test_hform_source.cpp:143-220, same mesh / order / dt / window): the
    H-formulation's alpha = rho = 1/sigma(x) and beta = mu = 1 are the A-form test's alpha = nu = 1/mu(x) and
    beta = sigma = 1, with the same source, exact solution and essential data -- so the window asserted in
    test_mms_rate_window_of_TestAForm pins this test too."""
    X = np.random.default_rng(0).uniform(0, 1, (1000, 3)) * [8.0, 1.0, 1.0]
    assert np.array_equal(1.0 / mms.sigma_expr_hform(X), 1.0 / mms.mu_expr(X))
    assert mms.hdot_exact is mms.adot_exact and mms.source_field_hform is mms.source_field


def test_mms_rate_hex_orders():
    """Same manufactured problem on hex cubes: rate ~ p (hex orders >= 2 are not pinned by any
    reference test; this is the strongest available check of their discretisation)."""
    for p, ns in ((1, (4, 8)), (2, (2, 4)), (3, (2, 4))):
        out = [_mms_step(om.perturb_interior(om.cartesian_hex(n, n, n), frac=0.1), p) for n in ns]
        r = np.log(out[1][1] / out[0][1]) / np.log((out[0][0] / out[1][0]) ** (1.0 / 3))
        assert p - 0.3 < r < p + 1.2, (p, out, r)


@pytest.mark.parametrize("name,orders", [("beam_tet", (1, 2, 3)), ("cylinder_hex_q2", (1, 2)), ("team7", (1,))])
def test_oracle_regression_vectors_on_reference_meshes(name, orders):
    _, mesh = golden_mesh(name)
    gv = golden_vectors()
    alpha, beta = coef_by_attr(mesh.attr, 0.5, 2.0), coef_by_attr(mesh.attr, 1.0, 1.0e3)
    for p in orders:
        s = osp.NDSpace(mesh, p)
        assert s.ndof == int(gv["%s_p%d_ndof" % (name, p)])
        A = oa.assemble(s, alpha, beta)
        x = np.random.default_rng(12345).standard_normal(s.ndof)
        y = gv["%s_p%d_y" % (name, p)]
        assert np.linalg.norm(A @ x - y) <= 1e-13 * np.linalg.norm(y)


def test_reference_mesh_dof_counts():
    """ND dof counts of the reference's meshes."""
    _, t7 = golden_mesh("team7")
    topo = om.Topology(t7)
    assert (t7.n_elem, topo.n_edges, topo.n_faces) == (57905, 68547, 116434)
    _, cyl = golden_mesh("cylinder_hex_q2")
    topo = om.Topology(cyl)
    assert (cyl.n_elem, topo.n_edges, topo.n_faces) == (252, 969, 858)
    assert osp.NDSpace(cyl, 2, topo).ndof == 6882


def test_hypre_pcg_semantics():
    """Stopping rule on <C r, r> / <C b, b> < tol^2 with tol arriving as a float (Appendix A.6)."""
    mesh = _small("hex", True, 3)
    s = osp.NDSpace(mesh, 1)
    A = oa.assemble(s, 0.05, 1.0)
    ess = s.boundary_dofs()
    Ac, _ = oa.eliminate(A, ess)
    b = Ac @ np.random.default_rng(1).standard_normal(s.ndof)
    tol = opcg.as_float_param(1e-6)
    x, it, rel, conv, hist = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=tol, atol=0.0, max_iter=500)
    assert conv and rel < tol
    r = b - Ac @ x
    d = Ac.diagonal()
    assert np.sqrt((r @ (r / d)) / (b @ (b / d))) < tol * 1.0001
    x1, it1, _, conv1, _ = opcg.pcg(Ac, b, diag=Ac.diagonal(), rtol=tol, atol=0.0, max_iter=it - 1)
    assert not conv1 and it1 == it - 1


def test_chebyshev_preconditioner_is_the_chebyshev_residual_polynomial():
    """oracle/pcg.py: chebyshev() -- the three-term recurrence equals z = q(D^-1 A) D^-1 r with
    1 - t q(t) = T_k((theta - t) / delta) / T_k(theta / delta) on [0.3, 1.2] * lambda_max, evaluated here through the
    eigen-decomposition of D^-1/2 A D^-1/2; the power-method estimate is a lower bound within 15 % of lambda_max; and
    Chebyshev-PCG needs fewer iterations than Jacobi-PCG for the same solution."""
    from oracle import assemble as oa, pcg as opcg
    from helpers import make_pair, coefficients
    m, sp, mo, so = make_pair("hex", 3, 2, True)
    a, b = coefficients(m)
    ess = sp.boundary_dofs()
    Ac, _ = oa.eliminate(oa.assemble(so, a * 0.05, b), ess)
    d = Ac.diagonal()
    S = Ac.toarray() / np.sqrt(d)[:, None] / np.sqrt(d)[None, :]
    w, V = np.linalg.eigh(S)
    lam = opcg.power_lambda_max(Ac, d)
    assert 0.85 * w.max() < lam <= w.max() * (1 + 1e-12)
    r = np.random.default_rng(3).standard_normal(len(d))
    lo, hi = 0.3 * lam, 1.2 * lam
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    T = np.polynomial.chebyshev.Chebyshev
    for k in (1, 2, 3, 5):
        z = opcg.chebyshev(Ac, d, k, lam)(r)
        Tk = T.basis(k)
        q = (1.0 - Tk((theta - w) / delta) / Tk(theta / delta)) / w          # q(t) on the spectrum
        z_ref = (V @ (q * (V.T @ (r / np.sqrt(d))))) / np.sqrt(d)
        assert np.abs(z - z_ref).max() < 1e-11 * np.abs(z_ref).max()
        assert np.all(q > 0)                                                  # a positive definite preconditioner
    rhs = Ac @ np.random.default_rng(1).standard_normal(len(d))
    xj, itj, _, cj, _ = opcg.pcg(Ac, rhs, diag=d, rtol=1e-10, max_iter=5000)
    xc, itc, _, cc, _ = opcg.pcg(Ac, rhs, precond=opcg.chebyshev(Ac, d, 3, lam), rtol=1e-10, max_iter=5000)
    assert cj and cc and itc < itj
    assert np.linalg.norm(xc - xj) < 1e-7 * np.linalg.norm(xj)
