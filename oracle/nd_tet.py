"""Nedelec (first kind) tetrahedral element of order p (oracle; test
infrastructure).

Restates MFEM's ``ND_TetrahedronElement`` ("ND_3D_P<p>" on simplices,
/root/reference/examples/team7/Main.cpp:163) [MFEM-upstream; SURVEY.md
Appendix A.1]: the basis is the dual basis of the dof functionals

    dof_m(v) = t_m . v(x_m)

over the space ND_p = P_{p-1}^3 (+) {q in hom. P_p^3 : q.x = 0}.  Because a
dual basis depends only on the space and the functionals, any raw basis of
ND_p may be inverted; monomials are used here.  Functionals (MFEM order):
p per edge (Gauss-Legendre points along the edge, tangent = edge vector),
p(p-1) per face (open face nodes, two tangents f1-f0, f2-f0), then
p(p-1)(p-2)/2 interior nodes with the three unit tangents.
"""
import itertools

import numpy as np

from .quadrature import gauss_legendre

TET_VERTS = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)
TET_EDGES = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
TET_FACES = [(1, 2, 3), (0, 3, 2), (0, 1, 3), (0, 2, 1)]


def nd_tet_ndof(p):
    return p * (p + 2) * (p + 3) // 2


def face_nodes(p):
    """Face-node index pairs (i, j) in MFEM loop order (j outer, i inner):
    barycentric weights on (f0, f1, f2) proportional to
    (fop[p-2-i-j], fop[i], fop[j])."""
    return [(i, j) for j in range(p - 1) for i in range(p - 1 - j)]


class _VecPoly:
    """Vector polynomial = list of (coef, (a,b,c), component)."""

    def __init__(self, terms):
        self.terms = terms

    def eval(self, X):
        out = np.zeros((len(X), 3))
        for cf, (a, b, c), d in self.terms:
            out[:, d] += cf * X[:, 0] ** a * X[:, 1] ** b * X[:, 2] ** c
        return out

    def curl(self, X):
        out = np.zeros((len(X), 3))

        def dmono(e, axis):
            e = list(e)
            k = e[axis]
            if k == 0:
                return 0.0
            e[axis] -= 1
            return k * X[:, 0] ** e[0] * X[:, 1] ** e[1] * X[:, 2] ** e[2]

        for cf, e, d in self.terms:
            # curl(f e_d): component (d+1) gets +df/dx_{d+2}; component (d+2) gets -df/dx_{d+1}
            d1, d2 = (d + 1) % 3, (d + 2) % 3
            out[:, d1] += cf * dmono(e, d2)
            out[:, d2] -= cf * dmono(e, d1)
        return out


def _raw_basis(p):
    fs = []
    for deg in range(p):
        for a, b, c in itertools.product(range(deg + 1), repeat=3):
            if a + b + c == deg:
                for d in range(3):
                    fs.append(_VecPoly([(1.0, (a, b, c), d)]))
    for a, b, c in itertools.product(range(p), repeat=3):
        if a + b + c != p - 1:
            continue
        # m*(y,-x,0), m*(z,0,-x)
        fs.append(_VecPoly([(1.0, (a, b + 1, c), 0), (-1.0, (a + 1, b, c), 1)]))
        fs.append(_VecPoly([(1.0, (a, b, c + 1), 0), (-1.0, (a + 1, b, c), 2)]))
        if a == 0:  # m*(0,z,-y)
            fs.append(_VecPoly([(1.0, (a, b, c + 1), 1), (-1.0, (a, b + 1, c), 2)]))
    return fs


class NDTet:
    def __init__(self, p):
        self.p = p
        self.nd = nd_tet_ndof(p)
        eop = gauss_legendre(p)[0]
        fop = gauss_legendre(p - 1)[0] if p > 1 else None
        iop = gauss_legendre(p - 2)[0] if p > 2 else None
        nodes, tang = [], []
        for a, b in TET_EDGES:
            for i in range(p):
                nodes.append(TET_VERTS[a] + eop[i] * (TET_VERTS[b] - TET_VERTS[a]))
                tang.append(TET_VERTS[b] - TET_VERTS[a])
        for f0, f1, f2 in TET_FACES:
            for i, j in face_nodes(p):
                w = np.array([fop[p - 2 - i - j], fop[i], fop[j]])
                w = w / w.sum()
                x = w[0] * TET_VERTS[f0] + w[1] * TET_VERTS[f1] + w[2] * TET_VERTS[f2]
                nodes.append(x)
                tang.append(TET_VERTS[f1] - TET_VERTS[f0])
                nodes.append(x)
                tang.append(TET_VERTS[f2] - TET_VERTS[f0])
        for k in range(p - 2):
            for j in range(p - 2 - k):
                for i in range(p - 2 - j - k):
                    w = iop[i] + iop[j] + iop[k] + iop[p - 3 - i - j - k]
                    x = np.array([iop[i], iop[j], iop[k]]) / w
                    for d in range(3):
                        nodes.append(x)
                        tang.append(np.eye(3)[d])
        self.nodes = np.array(nodes)
        self.tang = np.array(tang)
        assert len(self.nodes) == self.nd
        self.raw = _raw_basis(p)
        assert len(self.raw) == self.nd
        V = np.zeros((self.nd, self.nd))  # V[m, n] = dof_m(raw_n)
        for n, f in enumerate(self.raw):
            V[:, n] = np.einsum("md,md->m", f.eval(self.nodes), self.tang)
        self.Vinv = np.linalg.inv(V)  # shape_m = sum_n Vinv[n, m] raw_n

    def shape_and_curl(self, pts):
        pts = np.atleast_2d(pts)
        R = np.stack([f.eval(pts) for f in self.raw], 1)  # (nq, n, 3)
        Rc = np.stack([f.curl(pts) for f in self.raw], 1)
        V = np.einsum("qnd,nm->qmd", R, self.Vinv)
        C = np.einsum("qnd,nm->qmd", Rc, self.Vinv)
        return V, C

    def dof_nodes(self):
        return self.nodes, self.tang
