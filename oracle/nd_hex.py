"""Nedelec (first kind) hexahedral element of order p (oracle; test
infrastructure).

Restates MFEM's ``ND_HexahedronElement`` as instantiated by
``FiniteElementCollection::New("ND_3D_P<p>")``
(/root/reference/src/problem_builders/problem_builder_base.cpp:110-111)
[MFEM-upstream; SURVEY.md Appendix A.1]:

  * closed 1D basis: Lagrange on p+1 Gauss-Lobatto points,
    open 1D basis: Lagrange on p Gauss-Legendre points;
  * x-block basis  l^o_i(x) l^c_j(y) l^c_k(z)  (y-, z-blocks cyclic),
    nd = 3 p (p+1)^2;
  * local ordering [edge | face | interior] with sign flips where the local
    tangent opposes the entity's own parametrisation (MFEM's ``dof_map`` with
    negative entries).

The dof_map is DERIVED from the entity parametrisations (reference edges and
faces below) rather than transcribed, so the derivation doubles as a check.
"""
import numpy as np

from .quadrature import gauss_legendre, gauss_lobatto, lagrange_eval

# MFEM reference hexahedron: vertices, edges, faces
HEX_VERTS = np.array(
    [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
    dtype=np.int64,
)
HEX_EDGES = np.array(
    [[0, 1], [1, 2], [3, 2], [0, 3], [4, 5], [5, 6], [7, 6], [4, 7], [0, 4], [1, 5], [2, 6], [3, 7]],
    dtype=np.int64,
)
HEX_FACES = np.array(
    [[3, 2, 1, 0], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]],
    dtype=np.int64,
)


def nd_hex_ndof(p):
    return 3 * p * (p + 1) ** 2


def tensor_index(p, c, i, j, k):
    """Index in the [x-block | y-block | z-block] tensor ordering."""
    n = [p + 1, p + 1, p + 1]
    n[c] = p
    return c * p * (p + 1) ** 2 + i + n[0] * (j + n[1] * k)


def tensor_dims(p, c):
    n = [p + 1, p + 1, p + 1]
    n[c] = p
    return n


def _axis_and_dir(va, vb):
    d = HEX_VERTS[vb] - HEX_VERTS[va]
    ax = int(np.nonzero(d)[0][0])
    assert np.count_nonzero(d) == 1
    return ax, int(d[ax])


def face_dof_layout(p):
    """Face-local dofs of an order-p ND quad face, in MFEM order:
    s-directed (j=1..p-1 closed in t, i=0..p-1 open in s), then t-directed
    (j=0..p-1 open in t, i=1..p-1 closed in s).  Returns list of
    (dir, open_index, closed_index) with dir 0 = s, 1 = t."""
    out = []
    for j in range(1, p):
        for i in range(p):
            out.append((0, i, j))
    for j in range(p):
        for i in range(1, p):
            out.append((1, j, i))
    return out


def dof_map(p):
    """Return (local_index, sign) arrays over the tensor ordering:
    local dof `loc[t]` (MFEM [edge|face|interior] order) carries tensor basis
    function t multiplied by `sgn[t]`."""
    nd = nd_hex_ndof(p)
    loc = -np.ones(nd, dtype=np.int64)
    sgn = np.ones(nd, dtype=np.int64)
    o = 0
    # edges: p dofs each, open index runs along the edge from va to vb
    for va, vb in HEX_EDGES:
        ax, d = _axis_and_dir(va, vb)
        base = HEX_VERTS[va] * p  # closed indices of the two fixed coordinates
        for m in range(p):
            idx = [int(base[0]), int(base[1]), int(base[2])]
            idx[ax] = m if d > 0 else p - 1 - m
            t = tensor_index(p, ax, *idx)
            assert loc[t] < 0
            loc[t] = o
            sgn[t] = d
            o += 1
    # faces
    for fv in HEX_FACES:
        v0, v1, v3 = fv[0], fv[1], fv[3]
        ax_s, d_s = _axis_and_dir(v0, v1)
        ax_t, d_t = _axis_and_dir(v0, v3)
        ax_n = 3 - ax_s - ax_t
        fixed = int(HEX_VERTS[v0][ax_n]) * p
        for direction, io, ic in face_dof_layout(p):
            idx = [0, 0, 0]
            idx[ax_n] = fixed
            if direction == 0:  # tangent along s: open in s, closed in t
                idx[ax_s] = io if d_s > 0 else p - 1 - io
                idx[ax_t] = ic if d_t > 0 else p - ic
                comp, sg = ax_s, d_s
            else:  # tangent along t: open in t, closed in s
                idx[ax_t] = io if d_t > 0 else p - 1 - io
                idx[ax_s] = ic if d_s > 0 else p - ic
                comp, sg = ax_t, d_t
            t = tensor_index(p, comp, *idx)
            assert loc[t] < 0
            loc[t] = o
            sgn[t] = sg
            o += 1
    # interior: x-, y-, z-components, i fastest
    for c in range(3):
        lo = [1, 1, 1]
        hi = [p, p, p]
        lo[c] = 0
        for k in range(lo[2], hi[2]):
            for j in range(lo[1], hi[1]):
                for i in range(lo[0], hi[0]):
                    t = tensor_index(p, c, i, j, k)
                    assert loc[t] < 0
                    loc[t] = o
                    o += 1
    assert o == nd and (loc >= 0).all()
    return loc, sgn


class NDHex:
    """Reference-element tables for order p."""

    def __init__(self, p):
        self.p = p
        self.nd = nd_hex_ndof(p)
        self.cp = gauss_lobatto(p + 1)
        self.op = gauss_legendre(p)[0]
        self.loc, self.sgn = dof_map(p)

    def tables_1d(self, xq):
        Bc, Gc = lagrange_eval(self.cp, xq)
        Bo, _ = lagrange_eval(self.op, xq)
        return Bo, Bc, Gc

    def shape_and_curl(self, pts):
        """Reference shapes V[q, dof, 3] and curls C[q, dof, 3] in MFEM local
        dof order, at arbitrary reference points pts (nq,3)."""
        p, nd = self.p, self.nd
        pts = np.atleast_2d(pts)
        nq = len(pts)
        tabs = [self.tables_1d(pts[:, d]) for d in range(3)]  # per dim: Bo,Bc,Gc at q
        V = np.zeros((nq, nd, 3))
        C = np.zeros((nq, nd, 3))
        for c in range(3):
            n = tensor_dims(p, c)
            d1, d2 = (c + 1) % 3, (c + 2) % 3
            for k in range(n[2]):
                for j in range(n[1]):
                    for i in range(n[0]):
                        idx = (i, j, k)
                        t = tensor_index(p, c, i, j, k)
                        l, s = self.loc[t], self.sgn[t]
                        fo = tabs[c][0][:, idx[c]]
                        f1, g1 = tabs[d1][1][:, idx[d1]], tabs[d1][2][:, idx[d1]]
                        f2, g2 = tabs[d2][1][:, idx[d2]], tabs[d2][2][:, idx[d2]]
                        V[:, l, c] = s * fo * f1 * f2
                        # curl of (phi e_c): e_{d1}-component = +d(phi)/d x_{d2},
                        # e_{d2}-component = -d(phi)/d x_{d1}   (c,d1,d2 cyclic)
                        C[:, l, d1] = s * fo * f1 * g2
                        C[:, l, d2] = -s * fo * g1 * f2
        return V, C

    def dof_nodes(self):
        """Nodes (nd,3) and unit tangents (nd,3) of the dof functionals in MFEM
        local order: dof_l(v) = t_l . v(node_l) on the reference element."""
        p, nd = self.p, self.nd
        nodes = np.zeros((nd, 3))
        tang = np.zeros((nd, 3))
        for c in range(3):
            n = tensor_dims(p, c)
            for k in range(n[2]):
                for j in range(n[1]):
                    for i in range(n[0]):
                        idx = (i, j, k)
                        t = tensor_index(p, c, i, j, k)
                        l, s = self.loc[t], self.sgn[t]
                        for d in range(3):
                            nodes[l, d] = self.op[idx[d]] if d == c else self.cp[idx[d]]
                        tang[l, c] = s
        return nodes, tang
