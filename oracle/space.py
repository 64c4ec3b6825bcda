"""Global Nedelec DOF numbering (oracle; test infrastructure).

Loop-based restatement of what ``ParFiniteElementSpace`` + ``ND_FECollection``
hand to the assembly on the reference path
(/root/reference/src/problem_builders/problem_builder_base.cpp:98-127)
[MFEM-upstream, SURVEY.md Appendix A.2]:

    [edge dofs: p per edge | face dofs | interior dofs, element-major]

Edges are oriented from the lower to the higher global vertex index; an
element seeing an edge the other way gets the reversed order and sign -1.
Face dofs are defined in the frame of the face's *base* vertex order (the
local vertex order of the first element that touches the face).  An element
seeing the face in another orientation gets a permutation with signs (quads)
or a node permutation with 2x2 tangent transformations (triangles, p >= 2 --
MFEM's ``ND_DofTransformation``).

The element restriction is returned as a dense matrix per element:
    x_local = R_e @ x_global[dofs_e]
which for hexes is diagonal +-1 (stored as a sign vector).
"""
import numpy as np

from .mesh import Topology
from .nd_hex import NDHex, face_dof_layout
from .nd_tet import NDTet, face_nodes

_QUAD_CORNERS = [(0, 0), (1, 0), (1, 1), (0, 1)]


def quad_face_perm(p, local_verts, base_verts):
    """For a quad face seen with vertex order `local_verts`, return
    (perm, sign): local face dof m equals sign[m] * (base face dof perm[m])."""
    pos = [base_verts.index(v) for v in local_verts]
    P0 = np.array(_QUAD_CORNERS[pos[0]])
    es = np.array(_QUAD_CORNERS[pos[1]]) - P0
    et = np.array(_QUAD_CORNERS[pos[3]]) - P0
    layout = face_dof_layout(p)
    index = {d: m for m, d in enumerate(layout)}
    perm, sign = [], []

    def map_idx(vec, io_s, ic_t, tangent_is_s):
        # returns base (dir, open_index, closed_index, sign) for a local dof
        # with open index along the tangent and closed index along the other axis
        tan = es if tangent_is_s else et
        oth = et if tangent_is_s else es
        bdir = 0 if tan[0] != 0 else 1
        sg = int(tan[bdir])
        io = io_s if sg > 0 else p - 1 - io_s
        odir = 1 - bdir
        so = int(oth[odir])
        ic = ic_t if so > 0 else p - ic_t
        return (bdir, io, ic), sg

    for direction, io, ic in layout:
        key, sg = map_idx(None, io, ic, direction == 0)
        perm.append(index[key])
        sign.append(sg)
    return perm, sign


def tri_face_transform(p, local_verts, base_verts):
    """For a triangular face: returns (nd_face x nd_face) matrix R with
    x_local_face = R @ x_base_face."""
    nodes = face_nodes(p)
    nf = 2 * len(nodes)
    R = np.zeros((nf, nf))
    pos = [base_verts.index(v) for v in local_verts]  # local vertex k is base vertex pos[k]
    node_index = {ij: n for n, ij in enumerate(nodes)}
    # tangents in terms of base tangents T1 = F1-F0, T2 = F2-F0
    bt = {0: np.array([0, 0]), 1: np.array([1, 0]), 2: np.array([0, 1])}
    t1 = bt[pos[1]] - bt[pos[0]]
    t2 = bt[pos[2]] - bt[pos[0]]
    for n, (i, j) in enumerate(nodes):
        wl = [p - 2 - i - j, i, j]  # open-point indices on local vertices 0,1,2
        wb = [0, 0, 0]
        for k in range(3):
            wb[pos[k]] = wl[k]
        nb = node_index[(wb[1], wb[2])]
        R[2 * n, 2 * nb:2 * nb + 2] = t1
        R[2 * n + 1, 2 * nb:2 * nb + 2] = t2
    return R


class NDSpace:
    def __init__(self, mesh, p, topo=None):
        self.mesh, self.p = mesh, p
        self.topo = topo or Topology(mesh)
        topo = self.topo
        hexa = mesh.kind == "hex"
        self.fe = NDHex(p) if hexa else NDTet(p)
        nd = self.fe.nd
        ned = p
        nfd = 2 * p * (p - 1) if hexa else p * (p - 1)
        nint = 3 * p * (p - 1) ** 2 if hexa else 3 * (p * (p - 1) * (p - 2) // 6)
        self.ned, self.nfd, self.nint = ned, nfd, nint
        self.face_off = topo.n_edges * ned
        self.int_off = self.face_off + topo.n_faces * nfd
        self.ndof = self.int_off + mesh.n_elem * nint
        ne = mesh.n_elem
        self.elem_dofs = np.zeros((ne, nd), dtype=np.int64)
        self.elem_sign = np.ones((ne, nd), dtype=np.int64)  # hex / tet-edge signs
        self.elem_R = None if hexa else np.zeros((ne, nd, nd))
        for e in range(ne):
            ev = mesh.elems[e]
            o = 0
            Rdiag = np.ones(nd)
            blocks = []
            for le, (a, b) in enumerate(topo.ledges):
                ge = topo.el_edges[e, le]
                fwd = ev[a] < ev[b]
                for m in range(ned):
                    self.elem_dofs[e, o] = ge * ned + (m if fwd else ned - 1 - m)
                    self.elem_sign[e, o] = 1 if fwd else -1
                    o += 1
            for lf, fv in enumerate(topo.lfaces):
                if nfd == 0:
                    break
                gf = topo.el_faces[e, lf]
                lv = [int(ev[m]) for m in fv]
                bv = list(topo.face_verts[gf])
                base = self.face_off + gf * nfd
                if hexa:
                    perm, sign = quad_face_perm(p, lv, bv)
                    for m in range(nfd):
                        self.elem_dofs[e, o] = base + perm[m]
                        self.elem_sign[e, o] = sign[m]
                        o += 1
                else:
                    R = tri_face_transform(p, lv, bv)
                    blocks.append((o, R))
                    for m in range(nfd):
                        self.elem_dofs[e, o] = base + m
                        o += 1
            for m in range(nint):
                self.elem_dofs[e, o] = self.int_off + e * nint + m
                o += 1
            assert o == nd
            if not hexa:
                Re = np.diag(self.elem_sign[e].astype(np.float64))
                for o0, R in blocks:
                    Re[o0:o0 + nfd, o0:o0 + nfd] = R
                self.elem_R[e] = Re

    def signed_elem_dofs(self):
        """MFEM's GetElementDofs encoding: k >= 0 plain, -1-k sign-flipped."""
        return np.where(self.elem_sign > 0, self.elem_dofs, -1 - self.elem_dofs)

    def boundary_dofs(self, face_filter=None):
        """DOFs on boundary faces (and their edges).  face_filter(face_index)
        optionally restricts the set (essential attributes)."""
        topo, p = self.topo, self.p
        dofs = set()
        fe_edges = {}
        # map face -> its edges through any adjacent element
        for f in topo.boundary_faces():
            if face_filter is not None and not face_filter(f):
                continue
            e = topo.face_elems[f][0]
            lf = list(topo.el_faces[e]).index(f)
            fv = [int(self.mesh.elems[e][m]) for m in topo.lfaces[lf]]
            nv = len(fv)
            for m in range(nv):
                a, b = fv[m], fv[(m + 1) % nv]
                ge = topo.edge_id[(min(a, b), max(a, b))]
                for k in range(self.ned):
                    dofs.add(ge * self.ned + k)
            for k in range(self.nfd):
                dofs.add(self.face_off + f * self.nfd + k)
        return np.array(sorted(dofs), dtype=np.int64)
