"""Prototype (numpy) of the modal/Legendre affine-element operator used by the fast hex kernels,
checked against the oracle's quadrature element matrices.  Development aid, not product code."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mesh as om, space as osp, assemble as oa
from oracle.nd_hex import NDHex, tensor_dims, tensor_index
from oracle.quadrature import gauss_legendre, gauss_lobatto, lagrange_eval


def legendre01(n, x):
    """orthonormal shifted Legendre L_0..L_{n-1} on [0,1] and derivatives at x"""
    t = 2 * x - 1
    P = np.zeros((n, len(x))); dP = np.zeros((n, len(x)))
    P[0] = 1
    if n > 1:
        P[1] = t; dP[1] = 1
    for k in range(1, n - 1):
        P[k + 1] = ((2 * k + 1) * t * P[k] - k * P[k - 1]) / (k + 1)
        dP[k + 1] = ((2 * k + 1) * (P[k] + t * dP[k]) - k * dP[k - 1]) / (k + 1)
    s = np.sqrt(2 * np.arange(n) + 1)[:, None]
    return s * P, 2 * s * dP


def tables(p):
    xq, wq = gauss_legendre(p + 3)
    L, dL = legendre01(p + 1, xq)
    fe = NDHex(p)
    Bc, _ = lagrange_eval(fe.cp, xq)  # [q][j]
    Bo, _ = lagrange_eval(fe.op, xq)
    Tc = np.einsum("q,mq,qj->mj", wq, L, Bc)          # (p+1)x(p+1)
    To = np.einsum("q,mq,qi->mi", wq, L[:p], Bo)      # p x p
    Dm = np.einsum("q,mq,nq->mn", wq, L[:p], dL)      # p x (p+1): L_n' = sum_m Dm[m,n] L_m
    return Tc, To, Dm


def modal_apply(p, GK, GM, X):
    """X: list of 3 nodal coefficient arrays, comp c shape (n0,n1,n2) indexed [i0,i1,i2] (axis order x,y,z)."""
    Tc, To, Dm = tables(p)
    def T(c, a): return To if a == c else Tc
    def fwd(c, A):
        A = np.einsum("mi,ijk->mjk", T(c, 0), A)
        A = np.einsum("mj,ijk->imk", T(c, 1), A)
        return np.einsum("mk,ijk->ijm", T(c, 2), A)
    def bwd(c, A):
        A = np.einsum("mi,mjk->ijk", T(c, 0), A)
        A = np.einsum("mj,imk->ijk", T(c, 1), A)
        return np.einsum("mk,ijm->ijk", T(c, 2), A)
    def dax(A, a):   # derivative along axis a (closed p+1 -> p modes)
        return np.moveaxis(np.tensordot(Dm, np.moveaxis(A, a, 0), axes=(1, 0)), 0, a)
    def daxT(A, a):
        return np.moveaxis(np.tensordot(Dm.T, np.moveaxis(A, a, 0), axes=(1, 0)), 0, a)
    def restrict(A, shape):  # common index set: crop / zero-pad to shape
        out = np.zeros(shape)
        sl = tuple(slice(0, min(s, t)) for s, t in zip(A.shape, shape))
        out[sl] = A[sl]
        return out
    xh = [fwd(c, X[c]) for c in range(3)]
    yh = [np.zeros_like(x) for x in xh]
    for c in range(3):
        for d in range(3):
            yh[c] += GM[c, d] * restrict(xh[d], xh[c].shape)
    r = []
    for c in range(3):
        c1, c2 = (c + 1) % 3, (c + 2) % 3
        r.append(dax(xh[c2], c1) - dax(xh[c1], c2))
    for c in range(3):
        c1, c2 = (c + 1) % 3, (c + 2) % 3
        w = sum(GK[c, d] * restrict(r[d], r[c].shape) for d in range(3))
        yh[c2] += daxT(w, c1)
        yh[c1] -= daxT(w, c2)
    return [bwd(c, yh[c]) for c in range(3)]


def check(p):
    rng = np.random.default_rng(p)
    mesh = om.cartesian_hex(1, 1, 1)
    A = rng.standard_normal((3, 3)) * 0.3 + np.eye(3)
    if np.linalg.det(A) < 0: A[:, 0] *= -1
    mesh.verts[:] = mesh.verts @ A.T + rng.standard_normal(3)
    space = osp.NDSpace(mesh, p)
    alpha, beta = np.array([0.7]), np.array([1.9])
    (_, _, Ae), = list(oa.element_matrices(space, alpha, beta, q_override=p + 1))
    Ae = Ae[0]
    fe = space.fe
    # tensor ordering <- local: x_tensor[t] = sgn[t] x_local[loc[t]]
    X8 = mesh.verts[mesh.elems[0]]
    J = np.stack([X8[1] - X8[0], X8[3] - X8[0], X8[4] - X8[0]], axis=1)  # columns = edge vectors
    det = np.linalg.det(J)
    GK = alpha[0] * J.T @ J / det
    Ji = np.linalg.inv(J)
    GM = beta[0] * det * Ji @ Ji.T
    xl = rng.standard_normal(fe.nd)
    y_ref = Ae @ xl
    xt = fe.sgn * xl[fe.loc]
    X, o = [], 0
    for c in range(3):
        n = tensor_dims(p, c)
        sz = n[0] * n[1] * n[2]
        X.append(xt[o:o + sz].reshape(n[2], n[1], n[0]).transpose(2, 1, 0)); o += sz
    Y = modal_apply(p, GK, GM, X)
    yt = np.concatenate([Y[c].transpose(2, 1, 0).ravel() for c in range(3)])
    yl = np.zeros(fe.nd); yl[fe.loc] = fe.sgn * yt
    print("p=%d rel err %.2e" % (p, np.linalg.norm(yl - y_ref) / np.linalg.norm(y_ref)))


if __name__ == "__main__":
    for p in (1, 2, 3, 4):
        check(p)
    Tc, To, Dm = tables(2)
    np.set_printoptions(precision=17, linewidth=200)
    print(Tc, To, Dm, sep="\n")
