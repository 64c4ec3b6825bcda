"""Mesh ingest and Nedelec DOF-map generation (product code, vectorised numpy).

Stands in for what Hephaestus obtains from MFEM before the hot path starts --
``mfem::Mesh(file,1,1)`` + ``ParMesh`` (/root/reference/examples/team7/Main.cpp:154-155)
and ``ParFiniteElementSpace`` over an ``ND_3D_P<p>`` collection
(/root/reference/src/problem_builders/problem_builder_base.cpp:98-127) -- so that the
operator can be driven stand-alone (benchmarks, tests) at 10^7..10^9 DOFs.  In
plugged-in use the same arrays come from MFEM through adaptor/heph_gpu_operator.hpp.

DOF layout: [p per edge | face dofs | interior dofs, element-major]; edges and
faces numbered by first appearance over (element, local entity); edge
orientation low->high global vertex; quad-face orientation as signed
permutations; triangle faces by base-vertex frames (the library derives the
2x2 DofTransformation).  SURVEY.md Appendix A.2.
"""
import numpy as np

HEX_VERTS = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
HEX_EDGES = np.array([[0, 1], [1, 2], [3, 2], [0, 3], [4, 5], [5, 6], [7, 6], [4, 7], [0, 4], [1, 5], [2, 6], [3, 7]])
HEX_FACES = np.array([[3, 2, 1, 0], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]])
TET_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]])
TET_FACES = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]])


class Mesh:
    """kind 'hex' | 'tet'; verts (nv,3); elems (ne,8|4) corner ids (MFEM order);
    attr (ne,); geom_nodes optional (ne,27,3) in x-fastest tensor order."""

    def __init__(self, kind, verts, elems, attr=None, geom_nodes=None, bdr=None, bdr_attr=None):
        self.kind = kind
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.elems = np.ascontiguousarray(elems, dtype=np.int64)
        self.attr = np.ones(len(self.elems), np.int64) if attr is None else np.ascontiguousarray(attr, dtype=np.int64)
        self.geom_nodes = geom_nodes
        self.bdr = bdr
        self.bdr_attr = bdr_attr

    @property
    def n_elem(self):
        return len(self.elems)

    def elem_xyz(self):
        """Element node coordinates as the C ABI wants them."""
        if self.geom_nodes is not None:
            return np.ascontiguousarray(self.geom_nodes)
        return np.ascontiguousarray(self.verts[self.elems])


def cartesian_hex_mesh(nx, ny, nz, sx=1.0, sy=1.0, sz=1.0):
    """[0,sx]x[0,sy]x[0,sz] cut into nx*ny*nz hexes, stored as a general
    unstructured mesh (explicit connectivity/coordinates); attribute 1 for
    elements with x-centre < sx/2, else 2 (benchmark input S1)."""
    gx, gy, gz = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    vid = (gx + (nx + 1) * (gy + (ny + 1) * gz)).astype(np.int64)
    verts = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    verts[vid.ravel(), 0] = sx * gx.ravel() / nx
    verts[vid.ravel(), 1] = sy * gy.ravel() / ny
    verts[vid.ravel(), 2] = sz * gz.ravel() / nz
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")  # element order: x fastest
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    elems = np.stack([vid[i + dx, j + dy, k + dz] for dx, dy, dz in HEX_VERTS], axis=1)
    attr = np.where((i + 0.5) / nx < 0.5, 1, 2)
    return Mesh("hex", verts, elems, attr)


def perturb_interior(mesh, frac=0.2, seed=777):
    """Interior vertices displaced by U(-frac h, frac h), h = shortest edge."""
    v = mesh.verts.copy()
    lo, hi = v.min(0), v.max(0)
    interior = np.all((v > lo + 1e-12) & (v < hi - 1e-12), axis=1)
    edges = HEX_EDGES if mesh.kind == "hex" else TET_EDGES
    ev = mesh.verts[mesh.elems[:, edges]]
    h = np.linalg.norm(ev[:, :, 0] - ev[:, :, 1], axis=-1).min()
    d = np.random.default_rng(seed).uniform(-frac * h, frac * h, size=v.shape)
    v[interior] += d[interior]
    return Mesh(mesh.kind, v, mesh.elems, mesh.attr, None, mesh.bdr, mesh.bdr_attr)


_KUHN = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]])


def hex_to_tets(mesh):
    """Kuhn split: 6 positively oriented tets per hex around the 0-6 diagonal."""
    t = mesh.elems[:, _KUHN].reshape(-1, 4)
    X = mesh.verts[t]
    det = np.linalg.det(X[:, 1:] - X[:, :1])
    neg = det < 0
    t[neg, 2], t[neg, 3] = t[neg, 3].copy(), t[neg, 2].copy()
    return Mesh("tet", mesh.verts, t, np.repeat(mesh.attr, 6))


def refine_tets(mesh):
    """Uniform (red) refinement: each tet -> 8 (4 corner tets + octahedron cut
    along its shortest diagonal)."""
    e = mesh.elems
    pairs = e[:, TET_EDGES]
    key = np.sort(pairs, axis=2).reshape(-1, 2)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    nv = len(mesh.verts)
    mid = 0.5 * (mesh.verts[uniq[:, 0]] + mesh.verts[uniq[:, 1]])
    verts = np.concatenate([mesh.verts, mid])
    m = (nv + inv.reshape(-1, 6)).astype(np.int64)  # midpoints of edges 01 02 03 12 13 23
    v0, v1, v2, v3 = e.T
    m01, m02, m03, m12, m13, m23 = m.T
    corner = [np.stack(c, 1) for c in ([v0, m01, m02, m03], [m01, v1, m12, m13], [m02, m12, v2, m23], [m03, m13, m23, v3])]
    X = verts
    d = [np.linalg.norm(X[a] - X[b], axis=1) for a, b in ((m01, m23), (m02, m13), (m03, m12))]
    choice = np.argmin(np.stack(d, 1), axis=1)
    octs = {
        0: ([m01, m23, m02, m03], [m01, m23, m03, m13], [m01, m23, m13, m12], [m01, m23, m12, m02]),
        1: ([m02, m13, m01, m03], [m02, m13, m03, m23], [m02, m13, m23, m12], [m02, m13, m12, m01]),
        2: ([m03, m12, m01, m02], [m03, m12, m02, m23], [m03, m12, m23, m13], [m03, m12, m13, m01]),
    }
    inner = []
    for q in range(4):
        cand = np.stack([np.stack(octs[c][q], 1) for c in range(3)], 0)  # (3, ne, 4)
        inner.append(cand[choice, np.arange(len(e))])
    allt = np.stack(corner + inner, 1).reshape(-1, 4)
    Xt = verts[allt]
    neg = np.linalg.det(Xt[:, 1:] - Xt[:, :1]) < 0
    allt[neg, 2], allt[neg, 3] = allt[neg, 3].copy(), allt[neg, 2].copy()
    return Mesh("tet", verts, allt, np.repeat(mesh.attr, 8))


def read_mfem_mesh(path):
    """MFEM mesh v1.0 text (straight tets / hexes)."""
    toks = []
    with open(path) as f:
        for ln in f:
            ln = ln.split("#")[0].strip()
            if ln:
                toks.append(ln)
    assert toks[0].startswith("MFEM mesh v1.0")
    nv_of = {4: 4, 5: 8, 2: 3, 3: 4}
    i = 1
    elems, attr, bdr, battr, verts, kind = [], [], [], [], None, None
    while i < len(toks):
        t = toks[i]
        if t == "dimension":
            i += 2
        elif t in ("elements", "boundary"):
            n = int(toks[i + 1])
            for r in toks[i + 2:i + 2 + n]:
                v = [int(s) for s in r.split()]
                if t == "elements":
                    kind = {4: "tet", 5: "hex"}[v[1]]
                    attr.append(v[0])
                    elems.append(v[2:2 + nv_of[v[1]]])
                else:
                    battr.append(v[0])
                    bdr.append(v[2:2 + nv_of[v[1]]])
            i += 2 + n
        elif t == "vertices":
            n = int(toks[i + 1])
            verts = np.array([[float(s) for s in r.split()] for r in toks[i + 3:i + 3 + n]])
            i += 3 + n
        else:
            i += 1
    return Mesh(kind, verts, elems, attr, None, np.array(bdr, dtype=np.int64), np.array(battr, dtype=np.int64))


def _hex27_to_tensor(X27):
    """Reorder 27 geometry nodes (any file order) into x-fastest tensor order on
    {0,1/2,1}^3 by matching each node to the trilinear image of the candidates."""
    c = np.array([[i, j, k] for k in (0, .5, 1) for j in (0, .5, 1) for i in (0, .5, 1)], dtype=np.float64)
    N = np.ones((27, 8))
    for v in range(8):
        for d in range(3):
            N[:, v] *= c[:, d] if HEX_VERTS[v, d] else 1 - c[:, d]
    img = N @ X27[0, :8]  # (27,3) images for the first element
    order = np.array([np.argmin(np.linalg.norm(X27[0] - img[t], axis=1)) for t in range(27)])
    assert len(set(order.tolist())) == 27
    return X27[:, order]


def read_exodus(path):
    """ExodusII (NetCDF classic) TETRA4 / HEX8 / HEX27 reader; attribute = block
    id, side sets -> boundary elements with attribute = side-set id."""
    from scipy.io import netcdf_file

    f = netcdf_file(path, "r", mmap=False)
    var = f.variables
    if "coord" in var:
        xyz = np.array(var["coord"][:]).T.copy()
    else:
        xyz = np.stack([np.array(var["coord" + a][:]) for a in "xyz"], 1)
    ids = np.array(var["eb_prop1"][:])
    elems, attr, kind, n27 = [], [], None, []
    for b, bid in enumerate(ids):
        cv = var["connect%d" % (b + 1)]
        conn = np.array(cv[:]) - 1
        et = cv.elem_type.decode() if isinstance(cv.elem_type, bytes) else str(cv.elem_type)
        et = et.upper()
        if et.startswith("TET") and conn.shape[1] == 4:
            k = "tet"
            elems.append(conn)
        elif et.startswith("HEX") and conn.shape[1] in (8, 27):
            k = "hex"
            elems.append(conn[:, :8])
            if conn.shape[1] == 27:
                n27.append(conn)
        else:
            raise ValueError("unsupported Exodus block %s with %d nodes" % (et, conn.shape[1]))
        assert kind in (None, k), "mixed element types"
        kind = k
        attr.append(np.full(len(conn), bid))
    elems, attr = np.concatenate(elems), np.concatenate(attr)
    geom = _hex27_to_tensor(xyz[np.concatenate(n27)]) if n27 else None
    bdr, battr = [], []
    if "ss_prop1" in var:
        sides = ([[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [0, 3, 2, 1], [4, 5, 6, 7]] if kind == "hex"
                 else [[0, 1, 3], [1, 2, 3], [0, 3, 2], [0, 2, 1]])
        sides = np.array(sides)
        for s, sid in enumerate(np.array(var["ss_prop1"][:])):
            el = np.array(var["elem_ss%d" % (s + 1)][:]) - 1
            sd = np.array(var["side_ss%d" % (s + 1)][:]) - 1
            bdr.append(elems[el[:, None], sides[sd]])
            battr.append(np.full(len(el), sid))
    f.close()
    return Mesh(kind, xyz, elems, attr, geom, np.concatenate(bdr) if bdr else None, np.concatenate(battr) if bdr else None)


# ----------------------------------------------------------------------------
def _first_appearance_ids(cols):
    """Group equal rows of the integer columns; number groups by first
    appearance.  Returns (ids, n_groups, first_row_of_group)."""
    n = len(cols[0])
    order = np.lexsort(cols[::-1])
    new = np.ones(n, dtype=bool)
    diff = np.zeros(n - 1, dtype=bool)
    for c in cols:
        cs = c[order]
        diff |= cs[1:] != cs[:-1]
    new[1:] = diff
    grp = np.cumsum(new) - 1
    first = order[new]  # lexsort is stable: first row of a group is its earliest
    ng = len(first)
    by_first = np.argsort(first, kind="stable")
    rank = np.empty(ng, dtype=np.int64)
    rank[by_first] = np.arange(ng)
    ids = np.empty(n, dtype=np.int64)
    ids[order] = rank[grp]
    return ids, ng, first[by_first]


def _quad_orientation_tables(p):
    """perm/sign tables [8][nfd]: local face dof m == sign * base face dof perm
    for orientation code o = 2*pos0 + flip, where pos0 = index of the local
    first vertex in the base list and flip = 1 when the local direction of
    travel is opposite to the base's."""
    corners = np.array([[0, 0], [1, 0], [1, 1], [0, 1]])
    layout = [(0, i, j) for j in range(1, p) for i in range(p)] + [(1, j, i) for j in range(p) for i in range(1, p)]
    index = {d: m for m, d in enumerate(layout)}
    nfd = len(layout)
    perm = np.zeros((8, max(nfd, 1)), dtype=np.int64)
    sign = np.ones((8, max(nfd, 1)), dtype=np.int64)
    for pos0 in range(4):
        for flip in range(2):
            step = -1 if flip else 1
            pos = [(pos0 + step * k) % 4 for k in range(4)]
            P0 = corners[pos[0]]
            es, et = corners[pos[1]] - P0, corners[pos[3]] - P0
            for m, (direction, io, ic) in enumerate(layout):
                tan, oth = (es, et) if direction == 0 else (et, es)
                bdir = 0 if tan[0] != 0 else 1
                sg = int(tan[bdir])
                so = int(oth[1 - bdir])
                key = (bdir, io if sg > 0 else p - 1 - io, ic if so > 0 else p - ic)
                perm[2 * pos0 + flip, m] = index[key]
                sign[2 * pos0 + flip, m] = sg
    return perm, sign


class NDSpace:
    """Nedelec space of order p on a Mesh (serial numbering)."""

    def __init__(self, mesh, p):
        self.mesh, self.p = mesh, p
        hexa = mesh.kind == "hex"
        E = mesh.elems
        ne = len(E)
        ledges = HEX_EDGES if hexa else TET_EDGES
        lfaces = HEX_FACES if hexa else TET_FACES
        nle, nlf, nfv = len(ledges), len(lfaces), lfaces.shape[1]
        self.nd = 3 * p * (p + 1) ** 2 if hexa else p * (p + 2) * (p + 3) // 2
        ned = p
        nfd = 2 * p * (p - 1) if hexa else p * (p - 1)
        nint = 3 * p * (p - 1) ** 2 if hexa else p * (p - 1) * (p - 2) // 2
        # ---- edges
        ea, eb = E[:, ledges[:, 0]], E[:, ledges[:, 1]]  # (ne, nle)
        lo, hi = np.minimum(ea, eb).ravel(), np.maximum(ea, eb).ravel()
        eid, n_edges, efirst = _first_appearance_ids([lo, hi])
        eid = eid.reshape(ne, nle)
        fwd = ea < eb
        self.n_edges = n_edges
        self.edge_verts = np.stack([lo[efirst], hi[efirst]], 1)
        # ---- faces
        fv = E[:, lfaces]  # (ne, nlf, nfv)
        fs = np.sort(fv.reshape(-1, nfv), axis=1)
        fid, n_faces, ffirst = _first_appearance_ids([fs[:, 0], fs[:, 1], fs[:, 2]])
        fid = fid.reshape(ne, nlf)
        self.n_faces = n_faces
        self.face_verts = fv.reshape(-1, nfv)[ffirst]  # base vertex order
        self.face_count = np.bincount(fid.ravel(), minlength=n_faces)
        self.el_edges, self.el_faces = eid, fid
        self.ned, self.nfd, self.nint = ned, nfd, nint
        self.face_off = n_edges * ned
        self.int_off = self.face_off + n_faces * nfd
        self.ndof = self.int_off + ne * nint
        # ---- element dofs (signed, MFEM local order)
        dofs = np.zeros((ne, self.nd), dtype=np.int64)
        sgn = np.ones((ne, self.nd), dtype=np.int64)
        m = np.arange(ned)
        ed = np.where(fwd[:, :, None], m[None, None, :], ned - 1 - m[None, None, :]) + eid[:, :, None] * ned
        dofs[:, :nle * ned] = ed.reshape(ne, -1)
        sgn[:, :nle * ned] = np.repeat(np.where(fwd, 1, -1), ned, axis=1)
        o = nle * ned
        base = self.face_verts[fid]  # (ne, nlf, nfv)
        if nfd:
            if hexa:
                perm, sg = _quad_orientation_tables(p)
                pos0 = np.argmax(base == fv[:, :, :1], axis=2)  # index of local v0 in base
                nxt = np.take_along_axis(base, ((pos0 + 1) % 4)[:, :, None], axis=2)[:, :, 0]
                flip = (nxt != fv[:, :, 1]).astype(np.int64)
                code = 2 * pos0 + flip
                fd = self.face_off + fid[:, :, None] * nfd + perm[code]
                dofs[:, o:o + nlf * nfd] = fd.reshape(ne, -1)
                sgn[:, o:o + nlf * nfd] = sg[code].reshape(ne, -1)
            else:
                fd = self.face_off + fid[:, :, None] * nfd + np.arange(nfd)[None, None, :]
                dofs[:, o:o + nlf * nfd] = fd.reshape(ne, -1)
        o += nlf * nfd
        if nint:
            dofs[:, o:] = self.int_off + np.arange(ne)[:, None] * nint + np.arange(nint)[None, :]
        self.elem_dofs, self.elem_sign = dofs, sgn
        self.elem_face_base = None if hexa else np.ascontiguousarray(base, dtype=np.int32)

    def signed_elem_dofs(self):
        """FiniteElementSpace::GetElementDofs encoding (k, or -1-k if flipped)."""
        return np.where(self.elem_sign > 0, self.elem_dofs, -1 - self.elem_dofs).astype(np.int32)

    def boundary_dofs(self, select=None):
        """DOFs on boundary faces (count == 1) and on their edges.  `select`
        optionally maps boundary-face centroids (n,3) -> bool mask."""
        mesh = self.mesh
        hexa = mesh.kind == "hex"
        lfaces = HEX_FACES if hexa else TET_FACES
        ledges = HEX_EDGES if hexa else TET_EDGES
        isb = self.face_count[self.el_faces] == 1  # (ne, nlf)
        e_idx, lf_idx = np.nonzero(isb)
        if select is not None:
            cen = mesh.verts[mesh.elems[e_idx[:, None], lfaces[lf_idx]]].mean(axis=1)
            keep = np.asarray(select(cen), dtype=bool)
            e_idx, lf_idx = e_idx[keep], lf_idx[keep]
        out = []
        if self.nfd:
            f = self.el_faces[e_idx, lf_idx]
            out.append((self.face_off + f[:, None] * self.nfd + np.arange(self.nfd)[None]).ravel())
        # local edges of each local face
        f2e = np.array([[le for le, (a, b) in enumerate(ledges) if a in fv and b in fv] for fv in lfaces])
        ge = self.el_edges[e_idx[:, None], f2e[lf_idx]]
        out.append((ge.ravel()[:, None] * self.ned + np.arange(self.ned)[None]).ravel())
        return np.unique(np.concatenate(out)).astype(np.int32)

    def sideset_dofs(self, attrs):
        """DOFs on the boundary elements of the mesh file whose attribute is in
        `attrs` (BCMap::GetEssentialBdrMarkers -> GetEssentialTrueDofs,
        /root/reference/src/boundary_conditions/boundary_conditions.cpp:32-51)."""
        mesh = self.mesh
        sel = np.isin(mesh.bdr_attr, np.asarray(attrs))
        b = np.asarray(mesh.bdr[sel], dtype=np.int64)   # the vertex-triple keys below overflow 32 bits
        nfv = b.shape[1]
        out = []
        # edges of the boundary polygons
        lo_all = np.minimum(self.edge_verts[:, 0], self.edge_verts[:, 1])
        nv = int(max(self.mesh.elems.max(), b.max())) + 1
        ekey = self.edge_verts[:, 0] * nv + self.edge_verts[:, 1]
        order = np.argsort(ekey)
        for k in range(nfv):
            a, c = b[:, k], b[:, (k + 1) % nfv]
            key = np.minimum(a, c) * nv + np.maximum(a, c)
            pos = np.searchsorted(ekey[order], key)
            ge = order[pos]
            assert np.all(ekey[ge] == key)
            out.append((ge[:, None] * self.ned + np.arange(self.ned)[None]).ravel())
        if self.nfd:
            fs = np.sort(self.face_verts, axis=1)
            fkey = (fs[:, 0] * nv + fs[:, 1]) * nv + fs[:, 2]
            forder = np.argsort(fkey)
            bs = np.sort(b, axis=1)
            key = (bs[:, 0] * nv + bs[:, 1]) * nv + bs[:, 2]
            gf = forder[np.searchsorted(fkey[forder], key)]
            assert np.all(fkey[gf] == key)
            out.append((self.face_off + gf[:, None] * self.nfd + np.arange(self.nfd)[None]).ravel())
        del lo_all
        return np.unique(np.concatenate(out)).astype(np.int32)


class NativeNDSpace:
    """ND space on a hex or tet mesh generated by the library's host-side C++ code (csrc/ndspace.cpp,
    hcurl_host_ndspace_*): same numbering, signs, boundary dofs and per-face base vertex order as NDSpace
    (tests/test_ndspace.py checks them bit for bit), about 7x faster at 2 M hexes.  Offers what GpuNDSpace / the
    bench need."""

    def __init__(self, mesh, p):
        from . import _lib as L
        self.mesh, self.p = mesh, p
        lib = L.lib()
        hexa = mesh.kind == "hex"
        ev = np.ascontiguousarray(mesh.elems, dtype=np.int32)
        create = lib.hcurl_host_ndspace_create_hex if hexa else lib.hcurl_host_ndspace_create_tet
        h = L.check_ptr(create(p, len(ev), L.ptr(ev)), "hcurl_host_ndspace_create")
        try:
            self.nd = int(lib.hcurl_host_ndspace_nd(h))
            self.ndof = int(lib.hcurl_host_ndspace_ndof(h))
            self.n_edges, self.n_faces = int(lib.hcurl_host_ndspace_n_edges(h)), int(lib.hcurl_host_ndspace_n_faces(h))
            self._dofs = np.empty((len(ev), self.nd), np.int32)
            L.check(lib.hcurl_host_ndspace_elem_dofs(h, L.ptr(self._dofs)), "hcurl_host_ndspace_elem_dofs")
            nb = int(lib.hcurl_host_ndspace_boundary_dofs(h, None))
            self._bdr = np.empty(nb, np.int32)
            lib.hcurl_host_ndspace_boundary_dofs(h, L.ptr(self._bdr))
            self.elem_face_base = None
            if not hexa:
                self.elem_face_base = np.empty((len(ev), 4, 3), np.int32)
                L.check(lib.hcurl_host_ndspace_elem_face_base(h, L.ptr(self.elem_face_base)), "hcurl_host_ndspace_elem_face_base")
        finally:
            lib.hcurl_host_ndspace_destroy(h)

    def signed_elem_dofs(self):
        return self._dofs

    def boundary_dofs(self):
        return self._bdr


NativeHexNDSpace = NativeNDSpace


def interpolate_affine_hex(space, func, chunk=1 << 17):
    """Nodal ND interpolant (GridFunction::ProjectCoefficient) of a vector field on a mesh of parallelepiped hexes, from
    the signed element -> dof map (MFEM convention: -1-k = flipped) and the dof functionals dof = (J t) . F(x(node))
    (hcurl_host_hex_dof_nodes).  Chunked over elements so that it also runs at the bench size (2 M hexes, 50 M dofs).
    Checked against the oracle's own projection in tests/test_zy_fullsize_properties.py."""
    from . import _lib as L
    mesh = space.mesh
    nd = 3 * space.p * (space.p + 1) ** 2
    nodes, tang = np.zeros((nd, 3)), np.zeros((nd, 3))
    L.check(L.lib().hcurl_host_hex_dof_nodes(space.p, L.ptr(nodes), L.ptr(tang)), "hcurl_host_hex_dof_nodes")
    sd = space.signed_elem_dofs()
    out = np.zeros(space.ndof)
    for e0 in range(0, mesh.n_elem, chunk):
        X = mesh.verts[mesh.elems[e0:e0 + chunk]]  # (ne, 8, 3), MFEM vertex order
        v0 = X[:, 0]
        J = np.stack([X[:, 1] - v0, X[:, 3] - v0, X[:, 4] - v0], axis=2)  # columns = edge vectors
        ne = len(X)  # (ne*3, 3) @ (3, nd): one GEMM instead of ne small ones
        xq = v0[:, None, :] + (J.reshape(ne * 3, 3) @ nodes.T).reshape(ne, 3, -1).transpose(0, 2, 1)
        Jt = (J.reshape(ne * 3, 3) @ tang.T).reshape(ne, 3, -1).transpose(0, 2, 1)
        F = np.asarray(func(xq.reshape(-1, 3))).reshape(xq.shape)
        loc = np.einsum("eni,eni->en", Jt, F)
        s = sd[e0:e0 + chunk]
        flip = s < 0
        out[np.where(flip, -1 - s, s).reshape(-1)] = np.where(flip, -loc, loc).reshape(-1)
    return out
