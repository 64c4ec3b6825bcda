"""Element matrices, full assembly, essential-BC elimination (oracle; test
infrastructure).

Restates, without sum factorisation or any GPU-side trick, what the reference
does on this path:

  * ``mfem::CurlCurlIntegrator(coef)`` added by ``CurlCurlKernel::Apply``
    (/root/reference/src/kernels/curl_curl_kernel.cpp:21-24):
        K_e[i,j] = sum_q w_q a(x_q)/detJ_q (J_q c_i).(J_q c_j)
    -- the same formula the reference spells out in
    /root/reference/test/unit/test_nlintegrators.cpp:47-59
    (w = ip.weight/Ttr.Weight(); MultABt(curlshape, J, curlshape_dFt);
    AddMult_a_AAt(w, curlshape_dFt, elmat));
  * ``mfem::VectorFEMassIntegrator(coef)`` added by
    ``VectorFEMassKernel::Apply``
    (/root/reference/src/kernels/vector_fe_mass_kernel.cpp:21-24):
        M_e[i,j] = sum_q w_q b(x_q) detJ_q (J_q^-T v_i).(J_q^-T v_j);
  * ``blf->Assemble()`` (/root/reference/src/equation_systems/equation_system.cpp:345);
  * ``FormLinearSystem`` elimination (ibid. :132-190): essential rows/cols
    zeroed, diagonal one, B -= A_e x_bc, B[ess] = x_bc.

Quadrature follows MFEM's full-assembly defaults:
curl-curl order 2p (hex) / 2p-2 (tet); mass order OrderW + 2p with
OrderW = 3g-1 (hex, geometry order g) / 3(g-1) (tet).
"""
import numpy as np
import scipy.sparse as sp

from .nd_hex import HEX_VERTS
from .quadrature import gauss_legendre, hex_rule_points, lagrange_eval, tet_rule


# ----------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------
def hex8_jacobians(X, pts):
    """X: (ne,8,3) corner coords; pts (nq,3).  Returns J (ne,nq,3,3) with
    J[i,j] = d x_i / d xi_j, and mapped points (ne,nq,3)."""
    pts = np.atleast_2d(pts)
    N = np.ones((len(pts), 8))
    dN = np.ones((len(pts), 8, 3))
    for v in range(8):
        f = [pts[:, d] if HEX_VERTS[v, d] else 1.0 - pts[:, d] for d in range(3)]
        df = [np.ones(len(pts)) * (1.0 if HEX_VERTS[v, d] else -1.0) for d in range(3)]
        N[:, v] = f[0] * f[1] * f[2]
        dN[:, v, 0] = df[0] * f[1] * f[2]
        dN[:, v, 1] = f[0] * df[1] * f[2]
        dN[:, v, 2] = f[0] * f[1] * df[2]
    J = np.einsum("evi,qvj->eqij", X, dN)
    xq = np.einsum("evi,qv->eqi", X, N)
    return J, xq


def hex27_ref_coords(X27):
    """Infer the reference position (in {0,1/2,1}^3) of each of the 27 geometry
    nodes from the first element by matching against the trilinear image of
    the candidates (robust as long as curvature is small vs element size)."""
    cand = np.array([[i, j, k] for k in (0, .5, 1) for j in (0, .5, 1) for i in (0, .5, 1)])
    _, xc = hex8_jacobians(X27[:1, :8], cand)
    ref = np.zeros((27, 3))
    for n in range(27):
        d = np.linalg.norm(xc[0] - X27[0, n], axis=1)
        ref[n] = cand[np.argmin(d)]
    assert len({tuple(r) for r in ref}) == 27
    return ref


def hex27_jacobians(X27, ref, pts):
    pts = np.atleast_2d(pts)
    nodes1d = np.array([0.0, 0.5, 1.0])
    tabs = [lagrange_eval(nodes1d, pts[:, d]) for d in range(3)]
    idx = np.rint(ref * 2).astype(int)
    N = np.zeros((len(pts), 27))
    dN = np.zeros((len(pts), 27, 3))
    for n in range(27):
        b = [tabs[d][0][:, idx[n, d]] for d in range(3)]
        g = [tabs[d][1][:, idx[n, d]] for d in range(3)]
        N[:, n] = b[0] * b[1] * b[2]
        dN[:, n, 0] = g[0] * b[1] * b[2]
        dN[:, n, 1] = b[0] * g[1] * b[2]
        dN[:, n, 2] = b[0] * b[1] * g[2]
    J = np.einsum("evi,qvj->eqij", X27, dN)
    xq = np.einsum("evi,qv->eqi", X27, N)
    return J, xq


def tet_jacobians(X, pts):
    pts = np.atleast_2d(pts)
    J1 = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 3] - X[:, 0]], axis=-1)  # (ne,3,3)
    J = np.repeat(J1[:, None], len(pts), axis=1)
    xq = X[:, None, 0, :] + np.einsum("eij,qj->eqi", J1, pts)
    return J, xq


def hex_tensor_rule(n):
    x, w = gauss_legendre(n)
    pts = np.array([[x[i], x[j], x[k]] for k in range(n) for j in range(n) for i in range(n)])
    wts = np.array([w[i] * w[j] * w[k] for k in range(n) for j in range(n) for i in range(n)])
    return pts, wts


def default_rules(space, q_override=None):
    """(curl rule, mass rule) as (pts, wts) following MFEM's defaults, or a
    single n-points-per-dim Gauss rule for both if q_override is given (hex)."""
    mesh, p = space.mesh, space.p
    if mesh.kind == "hex":
        g = 2 if mesh.geom_nodes is not None else 1
        if q_override:
            return hex_tensor_rule(q_override), hex_tensor_rule(q_override)
        nc = hex_rule_points(2 * p)
        nm = hex_rule_points(3 * g - 1 + 2 * p)
        return hex_tensor_rule(nc), hex_tensor_rule(nm)
    return tet_rule(max(2 * p - 2, 0)), tet_rule(2 * p)


def _coef_at(coef, e_slice, xq, attr):
    """coef: scalar | (ne,) per element | callable(x (n,3), attr (n,)) -> (n,)."""
    ne, nq = xq.shape[:2]
    if callable(coef):
        a = np.repeat(attr[e_slice, None], nq, 1).reshape(-1)
        return np.asarray(coef(xq.reshape(-1, 3), a)).reshape(ne, nq)
    c = np.asarray(coef, dtype=np.float64)
    if c.ndim == 0:
        return np.full((ne, nq), float(c))
    return np.repeat(c[e_slice, None], nq, 1)


def element_matrices(space, alpha, beta, q_override=None, chunk=4096):
    """Yield (e0, e1, A_e) with A_e = K_e(alpha) + M_e(beta) in *local* dofs,
    shape (e1-e0, nd, nd).  alpha/beta None skips the term."""
    mesh, fe = space.mesh, space.fe
    (cp, cw), (mp, mw) = default_rules(space, q_override)
    Vc, Cc = fe.shape_and_curl(cp)
    Vm, _ = fe.shape_and_curl(mp)
    X = mesh.elem_coords()
    if mesh.geom_nodes is not None:
        ref27 = hex27_ref_coords(mesh.geom_nodes)
    for e0 in range(0, mesh.n_elem, chunk):
        e1 = min(mesh.n_elem, e0 + chunk)
        sl = slice(e0, e1)

        def jac(pts):
            if mesh.kind == "tet":
                return tet_jacobians(X[sl], pts)
            if mesh.geom_nodes is not None:
                return hex27_jacobians(mesh.geom_nodes[sl], ref27, pts)
            return hex8_jacobians(X[sl], pts)

        A = np.zeros((e1 - e0, fe.nd, fe.nd))
        if alpha is not None:
            J, xq = jac(cp)
            det = np.linalg.det(J)
            a = _coef_at(alpha, sl, xq, mesh.attr)
            JC = np.einsum("eqij,qnj->eqni", J, Cc)
            A += np.einsum("eq,eqni,eqmi->enm", cw[None] * a / det, JC, JC)
        if beta is not None:
            J, xq = jac(mp)
            det = np.linalg.det(J)
            b = _coef_at(beta, sl, xq, mesh.attr)
            Jinv = np.linalg.inv(J)
            JV = np.einsum("eqji,qnj->eqni", Jinv, Vm)  # J^-T v
            A += np.einsum("eq,eqni,eqmi->enm", mw[None] * b * det, JV, JV)
        yield e0, e1, A


def to_global(space, e0, e1, A):
    """Local element matrices -> matrices in global-dof frame (R^T A R)."""
    if space.elem_R is not None:
        R = space.elem_R[e0:e1]
        return np.einsum("eki,ekl,elj->eij", R, A, R)
    s = space.elem_sign[e0:e1].astype(np.float64)
    return A * s[:, :, None] * s[:, None, :]


def assemble(space, alpha, beta, q_override=None):
    """Assembled CSR of K(alpha) + M(beta) on the global dofs (no BCs)."""
    n = space.ndof
    acc = sp.csr_matrix((n, n))
    for e0, e1, A in element_matrices(space, alpha, beta, q_override):
        G = to_global(space, e0, e1, A)
        d = space.elem_dofs[e0:e1]
        nd = d.shape[1]
        rows = np.repeat(d[:, :, None], nd, 2).reshape(-1)
        cols = np.repeat(d[:, None, :], nd, 1).reshape(-1)
        acc = acc + sp.coo_matrix((G.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    acc.sum_duplicates()
    acc.sort_indices()
    return acc


def eliminate(A, ess, x_bc=None, b=None):
    """FormLinearSystem semantics (diag policy ONE).  Returns (A_c, B)."""
    n = A.shape[0]
    keep = np.ones(n)
    keep[ess] = 0.0
    Dk = sp.diags(keep)
    Ac = (Dk @ A @ Dk + sp.diags(1.0 - keep)).tocsr()
    Ac.sort_indices()
    B = None
    if b is not None:
        xb = np.zeros(n)
        if x_bc is not None:
            xb[ess] = x_bc[ess]
        B = b - A @ xb
        B[ess] = xb[ess]
    return Ac, B


def constrained_apply(A, ess, x):
    """y = A_c x computed the matrix-free way:
    zero x[ess], apply, y[ess] = x[ess]."""
    z = x.copy()
    z[ess] = 0.0
    y = A @ z
    y[ess] = x[ess]
    return y


def project_tangential(space, func, dofs=None):
    """Nodal interpolation of a vector field into the ND space: dof value =
    (J t) . F(x(node)) (what ProjectBdrCoefficientTangent / ProjectCoefficient
    compute for these nodal ND bases).  Returns a global vector (entries not
    touched stay 0); if `dofs` is given only those global dofs are written."""
    mesh, fe = space.mesh, space.fe
    nodes, tang = fe.dof_nodes()
    X = mesh.elem_coords()
    if mesh.kind == "tet":
        J, xq = tet_jacobians(X, nodes)
    elif mesh.geom_nodes is not None:
        J, xq = hex27_jacobians(mesh.geom_nodes, hex27_ref_coords(mesh.geom_nodes), nodes)
    else:
        J, xq = hex8_jacobians(X, nodes)
    F = np.asarray(func(xq.reshape(-1, 3))).reshape(xq.shape)
    Jt = np.einsum("enij,nj->eni", J, tang)
    loc = np.einsum("eni,eni->en", Jt, F)  # local dof values
    out = np.zeros(space.ndof)
    mask = None
    if dofs is not None:
        mask = np.zeros(space.ndof, dtype=bool)
        mask[dofs] = True
    if space.elem_R is not None:
        # x_local = R x_global  ->  x_global = R^-1 x_local
        glob = np.linalg.solve(space.elem_R, loc[:, :, None])[:, :, 0]
    else:
        glob = loc * space.elem_sign
    d = space.elem_dofs
    if mask is None:
        out[d.reshape(-1)] = glob.reshape(-1)
    else:
        sel = mask[d]
        out[d[sel]] = glob[sel]
    return out
