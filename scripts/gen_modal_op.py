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


def _snap(values, rel=1e-13):
    """magnitudes that differ only by rounding noise become one constant (the mean of the cluster)"""
    mags = sorted({abs(float(v)) for v in values if abs(v) > 1e-14})
    rep, cluster = {}, []
    for m in mags + [None]:
        if m is not None and (not cluster or m - cluster[0] <= rel * m):
            cluster.append(m)
            continue
        mean = float(np.mean(cluster))
        if abs(mean - round(mean)) <= rel * mean:
            mean = float(round(mean))
        for c in cluster:
            rep[c] = mean
        cluster = [m] if m is not None else []
    return rep


def emit_header(p, path):
    """Complete element operator of order p as one host/device function (tensor-ordered nodal dofs in and out):
    forward transform, sparse modal operator, backward transform -- straight-line code, every constant a literal.
    The geometry factors are multiplied once by scale and by each distinct coefficient magnitude, so that every
    non-zero of the modal operator costs one FMA."""
    T = transforms(p)
    n = T.shape[0]
    names = {(0, 0): 0, (0, 1): 1, (0, 2): 2, (1, 1): 3, (1, 2): 4, (2, 2): 5}
    Z = np.zeros((3, 3))
    ops = [(kind, names[ab], modal_operator(p, Z, E) if kind == "M" else modal_operator(p, E, Z)) for kind in ("M", "K") for ab, E in sym_basis()]
    tsnap = _snap(T.ravel())
    hsnap = _snap(np.concatenate([H.ravel() for _, _, H in ops]))
    tname = {m: "t%d" % i for i, m in enumerate(sorted(set(tsnap.values())))}
    L = []
    L.append("// GENERATED by scripts/gen_modal_op.py %d --header -- do not edit.  Order-%d Nedelec hex, affine element:" % (p, p))
    L.append("// y = scale * (K(alpha) + M(beta)) x on the %d tensor-ordered local dofs (component-major, x fastest), applied in the" % n)
    L.append("// orthonormal Legendre basis; GK[6] = alpha J^T J / det, GM[6] = beta adj adj^T / det (packed 00 01 02 11 12 22).")
    L.append("// Verified against the oracle's quadrature element matrices by the generator and by tests/test_modal_generated.py.")
    L.append("#pragma once")
    L.append("#ifndef HCURL_HD\n#ifdef __CUDACC__\n#define HCURL_HD __host__ __device__ __forceinline__\n#else\n#define HCURL_HD inline\n#endif\n#endif")
    L.append("HCURL_HD void hcurl_modal_p%d_apply(const double *x, const double *GK, const double *GM, const double scale, double *y) {" % p)
    for m, nm in sorted(tname.items(), key=lambda kv: kv[1]):
        L.append("  const double %s = %.17g;" % (nm, m))
    fac = {}
    for kind, s, H in ops:
        for v in np.abs(H[np.abs(H) > 1e-14]):
            fac.setdefault((kind, s, hsnap[float(v)]), None)
    for i, key in enumerate(sorted(fac)):
        kind, s, m = key
        fac[key] = "g%d" % i
        L.append("  const double g%d = %sG%s[%d];" % (i, "scale * " if m == 1.0 else "%.17g * scale * " % m, kind, s))
    L.append("  double xh[%d], yh[%d];" % (n, n))

    def lin(dst, coef, src):
        terms = [(j, coef[j]) for j in range(n) if abs(coef[j]) > 1e-14]
        expr = None
        for j, c in terms:
            t = "%s%s * %s[%d]" % ("-" if c < 0 else "", tname[tsnap[abs(float(c))]], src, j)
            expr = t if expr is None else "%s + %s" % (expr, t)
        L.append("  %s = %s;" % (dst, expr or "0.0"))

    for i in range(n):
        lin("xh[%d]" % i, T[i], "x")
    for i in range(n):
        expr = None
        for kind, s, H in ops:
            for j in np.nonzero(np.abs(H[i]) > 1e-14)[0]:
                g = fac[(kind, s, hsnap[abs(float(H[i, j]))])]
                t = "%s%s * xh[%d]" % ("-" if H[i, j] < 0 else "", g, j)
                expr = t if expr is None else "%s + %s" % (expr, t)
        L.append("  yh[%d] = %s;" % (i, expr or "0.0"))
    for j in range(n):
        lin("y[%d]" % j, T[:, j], "yh")
    L.append("}")
    with open(path, "w") as f:
        f.write("\n".join(L) + "\n")


if __name__ == "__main__":
    worst = max(check(p) for p in (1, 2, 3))
    assert worst < 1e-12
    if len(sys.argv) > 2 and sys.argv[2] == "--header":
        emit_header(int(sys.argv[1]), sys.argv[3])
    elif len(sys.argv) > 1:
        emit(int(sys.argv[1]))
