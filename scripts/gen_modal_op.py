"""Development aid for the next element kernels: derives, for order p, the affine-element operator in the orthonormal
Legendre (modal) basis as a SPARSE bilinear form in the geometry factors,

    Yh = sum_{a<=b} GM_ab * Mhat^{ab} Xh  +  sum_{a<=b} GK_ab * Khat^{ab} Xh ,

(GK = alpha J^T J / det, GM = beta adj adj^T / det, symmetric 3x3), checks the whole chain nodal -> modal -> operator ->
nodal against the oracle's quadrature element matrices on a random affine element, and prints the non-zeros as C
statements (the hand-written order-2 kernel in csrc/hex_patch.cu is this for p = 2, in a rotated per-component frame).

    python scripts/gen_modal_op.py 1        # order 1: 12 dofs, prints the operator; all orders 1..3 are verified"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import mesh as om, space as osp, assemble as oa  # noqa: E402
from oracle.nd_hex import tensor_dims  # noqa: E402
from proto_modal import tables  # noqa: E402


def modal_sizes(p):
    return [tuple(p if a == c else p + 1 for a in range(3)) for c in range(3)]   # modes of component c along x, y, z


def modal_operator(p, GK, GM):
    """dense matrix of the modal operator (modal coefficients of all three components, component-major, x fastest)"""
    _, _, Dm = tables(p)
    sz = modal_sizes(p)
    off = np.cumsum([0] + [int(np.prod(s)) for s in sz])
    n = off[-1]

    def unpack(v):
        return [v[off[c]:off[c + 1]].reshape(sz[c][2], sz[c][1], sz[c][0]).transpose(2, 1, 0) for c in range(3)]

    def pack(X):
        return np.concatenate([X[c].transpose(2, 1, 0).ravel() for c in range(3)])

    def dax(A, a):
        return np.moveaxis(np.tensordot(Dm, np.moveaxis(A, a, 0), axes=(1, 0)), 0, a)

    def daxT(A, a):
        return np.moveaxis(np.tensordot(Dm.T, np.moveaxis(A, a, 0), axes=(1, 0)), 0, a)

    def restrict(A, shape):
        out = np.zeros(shape)
        sl = tuple(slice(0, min(s, t)) for s, t in zip(A.shape, shape))
        out[sl] = A[sl]
        return out

    def apply(v):
        xh = unpack(v)
        yh = [np.zeros_like(x) for x in xh]
        for c in range(3):
            for d in range(3):
                yh[c] += GM[c, d] * restrict(xh[d], xh[c].shape)
        r = [dax(xh[(c + 2) % 3], (c + 1) % 3) - dax(xh[(c + 1) % 3], (c + 2) % 3) for c in range(3)]
        for c in range(3):
            c1, c2 = (c + 1) % 3, (c + 2) % 3
            w = sum(GK[c, d] * restrict(r[d], r[c].shape) for d in range(3))
            yh[c2] += daxT(w, c1)
            yh[c1] -= daxT(w, c2)
        return pack(yh)

    return np.stack([apply(e) for e in np.eye(n)], axis=1)


def transforms(p):
    """dense nodal -> modal matrix T (tensor-ordered nodal dofs of all three components -> modal coefficients)"""
    Tc, To, _ = tables(p)
    blocks = []
    for c in range(3):
        t = [To if a == c else Tc for a in range(3)]
        blocks.append(np.kron(t[2], np.kron(t[1], t[0])))      # x fastest
    n = sum(b.shape[0] for b in blocks)
    T = np.zeros((n, n))
    o = 0
    for b in blocks:
        T[o:o + b.shape[0], o:o + b.shape[1]] = b
        o += b.shape[0]
    return T


def sym_basis():
    out = []
    for a in range(3):
        for b in range(a, 3):
            E = np.zeros((3, 3))
            E[a, b] = E[b, a] = 1.0
            out.append(((a, b), E))
    return out


def check(p):
    rng = np.random.default_rng(10 + p)
    mesh = om.cartesian_hex(1, 1, 1)
    A = rng.standard_normal((3, 3)) * 0.3 + np.eye(3)
    if np.linalg.det(A) < 0:
        A[:, 0] *= -1
    mesh.verts[:] = mesh.verts @ A.T + rng.standard_normal(3)
    space = osp.NDSpace(mesh, p)
    alpha, beta = np.array([0.7]), np.array([1.9])
    (_, _, Ae), = list(oa.element_matrices(space, alpha, beta, q_override=p + 1))
    fe = space.fe
    X8 = mesh.verts[mesh.elems[0]]
    J = np.stack([X8[1] - X8[0], X8[3] - X8[0], X8[4] - X8[0]], axis=1)
    det = np.linalg.det(J)
    GK = alpha[0] * J.T @ J / det
    Ji = np.linalg.inv(J)
    GM = beta[0] * det * Ji @ Ji.T
    T = transforms(p)
    At = T.T @ modal_operator(p, GK, GM) @ T                  # tensor-ordered nodal dofs
    S = np.zeros((fe.nd, fe.nd))                              # x_tensor[t] = sgn[t] x_local[loc[t]]
    S[np.arange(fe.nd), fe.loc] = fe.sgn
    err = np.abs(S.T @ At @ S - Ae[0]).max() / np.abs(Ae[0]).max()
    # the operator is bilinear in (GK, GM): sum of the basis contributions reproduces it
    lin = sum(GM[a, b] * modal_operator(p, np.zeros((3, 3)), E) + GK[a, b] * modal_operator(p, E, np.zeros((3, 3))) for (a, b), E in sym_basis())
    err2 = np.abs(lin - modal_operator(p, GK, GM)).max()
    nnz = sum(int((np.abs(modal_operator(p, np.zeros((3, 3)), E)) > 1e-14).sum() + (np.abs(modal_operator(p, E, np.zeros((3, 3)))) > 1e-14).sum())
              for _, E in sym_basis())
    print("p=%d: %3d dofs, modal chain vs oracle element matrix: %.1e; bilinearity %.1e; %d non-zero (factor, i, j) terms" % (p, fe.nd, err, err2, nnz))
    return err


def emit(p):
    names = {(0, 0): "00", (0, 1): "01", (0, 2): "02", (1, 1): "11", (1, 2): "12", (2, 2): "22"}
    print("// order %d modal operator: yh[i] += factor * coefficient * xh[j]   (modal coefficients: component-major, x fastest)" % p)
    for kind in ("M", "K"):
        for (a, b), E in sym_basis():
            Z = np.zeros((3, 3))
            H = modal_operator(p, Z, E) if kind == "M" else modal_operator(p, E, Z)
            for i, j in zip(*np.nonzero(np.abs(H) > 1e-14)):
                print("yh[%2d] = fma(G%s%s * %+.17g, xh[%2d], yh[%2d]);" % (i, kind, names[(a, b)], H[i, j], j, i))


if __name__ == "__main__":
    worst = max(check(p) for p in (1, 2, 3))
    assert worst < 1e-12
    if len(sys.argv) > 1:
        emit(int(sys.argv[1]))
