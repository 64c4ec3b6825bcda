"""Discrete gradient (H1 -> ND) and discrete curl (ND -> RT) on hexes -- ORACLE (test infrastructure; only tests/,
smoke() and bench.py's cpu_baseline leg may import this).

Restates what mfem::DiscreteLinearOperator + GradientInterpolator / CurlInterpolator assemble for the reference's
CurlAuxSolver (/root/reference/src/auxsolvers/curl_aux.cpp:13-30) and ScalarPotentialSource::BuildGrad
(/root/reference/src/sources/scalar_potential_source.cpp:89-97) [MFEM-upstream: fem/bilininteg.hpp GradientInterpolator /
CurlInterpolator -> FiniteElement::ProjectGrad / ProjectCurl]: the local matrix applies the RANGE element's dof
functionals to the gradient / curl of every DOMAIN basis function.  On tensor elements with MFEM's nodal bases these
functionals are point values of the pulled-back field, so the local matrices do not depend on the geometry.

Route taken here (deliberately different from the kernel's 1-D shortcut): full 3-D evaluation of every basis function's
gradient / curl at every range node.  The reference's tests hold no vectors for these operators (searched:
test/, examples/): PARITY UNPINNED against MFEM; pinned instead by exactness on polynomial fields and curl(grad) = 0
(tests/test_interp.py).  Everything is in TENSOR order (x fastest, component-major) -- MFEM's local orderings enter in
plugged-in use through TensorBasisElement::GetDofMap(), not here."""
import numpy as np

from .nd_hex import NDHex, tensor_dims
from .quadrature import gauss_legendre, gauss_lobatto, lagrange_eval


def h1_nodes(p):
    cp = gauss_lobatto(p + 1)
    return np.array([[cp[i], cp[j], cp[k]] for k in range(p + 1) for j in range(p + 1) for i in range(p + 1)])


def nd_nodes(p):
    """(nodes (nd,3), component (nd,)) of the tensor-ordered ND dofs"""
    cp, op = gauss_lobatto(p + 1), gauss_legendre(p)[0]
    nodes, comp = [], []
    for c in range(3):
        n = tensor_dims(p, c)
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    idx = (i, j, k)
                    nodes.append([op[idx[d]] if d == c else cp[idx[d]] for d in range(3)])
                    comp.append(c)
    return np.array(nodes), np.array(comp)


def rt_dims(p, c):
    n = [p, p, p]
    n[c] = p + 1
    return n


def rt_nodes(p):
    """RT_FECollection(p-1) hex element: component c closed (p+1 points) along c, open (p) along the others"""
    cp, op = gauss_lobatto(p + 1), gauss_legendre(p)[0]
    nodes, comp = [], []
    for c in range(3):
        n = rt_dims(p, c)
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    idx = (i, j, k)
                    nodes.append([cp[idx[d]] if d == c else op[idx[d]] for d in range(3)])
                    comp.append(c)
    return np.array(nodes), np.array(comp)


def local_grad_matrix(p):
    """G[t_nd, j_h1] = e_c . grad(phi_j)(node_t), phi_j the H1 tensor basis"""
    cp = gauss_lobatto(p + 1)
    nodes, comp = nd_nodes(p)
    tabs = [lagrange_eval(cp, nodes[:, d]) for d in range(3)]       # (B, G) per direction at all ND nodes
    n1 = p + 1
    G = np.zeros((len(nodes), n1 ** 3))
    for k in range(n1):
        for j in range(n1):
            for i in range(n1):
                idx = (i, j, k)
                col = i + n1 * (j + n1 * k)
                for c in range(3):
                    f = np.ones(len(nodes))
                    for d in range(3):
                        f = f * (tabs[d][1][:, idx[d]] if d == c else tabs[d][0][:, idx[d]])
                    G[comp == c, col] = f[comp == c]
    return G


def local_curl_matrix(p):
    """C[t_rt, t_nd] = e_c . curl(N_t_nd)(node_t_rt), N the ND tensor basis (unsigned)"""
    fe = NDHex(p)
    nodes, comp = rt_nodes(p)
    _, Cl = fe.shape_and_curl(nodes)                               # (n_rt, nd local, 3), MFEM local order with signs
    Ct = Cl[:, fe.loc, :] * np.asarray(fe.sgn)[None, :, None]      # tensor function t = sgn[t] * local function loc[t]
    return Ct[np.arange(len(nodes)), :, comp]


def apply(kind, p, in_dofs, out_dofs, n_out, x, tet=False):
    """y = Interp x with signed (-1-k) element dofs (hexes: tensor order; tets: MFEM's local order, order 1); every
    element overwrites its output dofs"""
    if tet:
        M = tet_local_grad_matrix() if kind == "grad" else tet_local_curl_matrix()
    else:
        M = local_grad_matrix(p) if kind == "grad" else local_curl_matrix(p)
    ii = np.where(in_dofs >= 0, in_dofs, -1 - in_dofs)
    si = np.where(in_dofs >= 0, 1.0, -1.0)
    oi = np.where(out_dofs >= 0, out_dofs, -1 - out_dofs)
    so = np.where(out_dofs >= 0, 1.0, -1.0)
    y = np.zeros(n_out)
    ye = (M @ (si * x[ii]).T).T * so
    y[oi.ravel()] = ye.ravel()
    return y


# ---- tets, order 1 (ND1 -> RT0, H1 P1 -> ND1) ---------------------------------------------------------------------
# [MFEM-upstream] RT_TetrahedronElement::nk: dof_f(v) = v(face centroid) . nk_f on the reference tet, nk = twice the
# area-weighted outward normal; faces in MFEM's order (1,2,3), (0,3,2), (0,1,3), (0,2,1)
TET_NK = np.array([[1.0, 1.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0], [0.0, 0.0, -1.0]])


def tet_local_curl_matrix():
    """C[f, e] = nk_f . curl(N_e)(centroid of face f), N the oracle's ND1 tet basis (nd_tet.py)"""
    from .nd_tet import NDTet, TET_FACES, TET_VERTS
    fe = NDTet(1)
    cen = np.array([TET_VERTS[list(f)].mean(axis=0) for f in TET_FACES])
    _, C = fe.shape_and_curl(cen)                                  # (4 points, 6 dofs, 3)
    return np.einsum("fd,fed->fe", TET_NK, C)


def tet_local_grad_matrix():
    """G[e, v] = tk_e . grad(lambda_v) with tk_e = x_b - x_a (the ND1 dof functional applied to the P1 hat functions)"""
    from .nd_tet import TET_EDGES, TET_VERTS
    gl = np.array([[-1.0, -1.0, -1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    return np.array([[(TET_VERTS[b] - TET_VERTS[a]) @ gl[v] for v in range(4)] for a, b in TET_EDGES])
