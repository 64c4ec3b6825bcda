"""Meshes and mesh topology (oracle; test infrastructure).

Plain-loop restatement of what ``mfem::Mesh`` provides to the FE space on the
reference path (``mfem::Mesh(file,1,1)`` / ``ParMesh`` in
/root/reference/examples/team7/Main.cpp:154-155) [MFEM-upstream, SURVEY.md
Appendix A.2]: element->vertex tables, edges and faces numbered by first
appearance while looping elements and their local entities, boundary faces.

Element kinds: "hex" (8 corner vertices, optional 27 geometry nodes) and
"tet" (4 vertices).
"""
import numpy as np

from .nd_hex import HEX_EDGES, HEX_FACES

TET_EDGES = np.array([[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]], dtype=np.int64)
TET_FACES = np.array([[1, 2, 3], [0, 3, 2], [0, 1, 3], [0, 2, 1]], dtype=np.int64)


class Mesh:
    def __init__(self, kind, verts, elems, attr=None, geom_nodes=None, bdr=None, bdr_attr=None):
        self.kind = kind
        self.verts = np.ascontiguousarray(verts, dtype=np.float64)
        self.elems = np.ascontiguousarray(elems, dtype=np.int64)
        self.attr = np.ones(len(self.elems), dtype=np.int64) if attr is None else np.asarray(attr, dtype=np.int64)
        # optional higher-order geometry: (n_elem, n_geom_nodes, 3)
        self.geom_nodes = geom_nodes
        # boundary elements given by the file (vertex lists) with attributes
        self.bdr = bdr
        self.bdr_attr = bdr_attr

    @property
    def n_elem(self):
        return len(self.elems)

    def elem_coords(self):
        """(n_elem, nv, 3) corner coordinates."""
        return self.verts[self.elems]


def cartesian_hex(nx, ny, nz, sx=1.0, sy=1.0, sz=1.0):
    """Structured hex mesh stored as a general unstructured mesh; vertices and
    elements x-fastest, element corners in MFEM hex order.  Attribute 1 for
    x_centre < sx/2, else 2 (input S1)."""
    verts = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    v = 0
    for k in range(nz + 1):
        for j in range(ny + 1):
            for i in range(nx + 1):
                verts[v] = (sx * i / nx, sy * j / ny, sz * k / nz)
                v += 1

    def vid(i, j, k):
        return i + (nx + 1) * (j + (ny + 1) * k)

    elems = []
    attr = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                elems.append(
                    [vid(i, j, k), vid(i + 1, j, k), vid(i + 1, j + 1, k), vid(i, j + 1, k),
                     vid(i, j, k + 1), vid(i + 1, j, k + 1), vid(i + 1, j + 1, k + 1), vid(i, j + 1, k + 1)]
                )
                attr.append(1 if (i + 0.5) / nx < 0.5 else 2)
    return Mesh("hex", verts, elems, attr)


def perturb_interior(mesh, frac=0.2, seed=777):
    """Displace interior vertices by U(-frac*h, frac*h) per coordinate, h = the
    smallest edge length; boundary vertices (on the bounding box) stay."""
    v = mesh.verts.copy()
    lo, hi = v.min(0), v.max(0)
    interior = np.all((v > lo + 1e-12) & (v < hi - 1e-12), axis=1)
    if mesh.kind == "hex":
        edges = HEX_EDGES
    else:
        edges = TET_EDGES
    ev = mesh.verts[mesh.elems[:, edges]]
    h = np.linalg.norm(ev[:, :, 0] - ev[:, :, 1], axis=-1).min()
    rng = np.random.default_rng(seed)
    d = rng.uniform(-frac * h, frac * h, size=v.shape)
    v[interior] += d[interior]
    return Mesh(mesh.kind, v, mesh.elems, mesh.attr, None, mesh.bdr, mesh.bdr_attr)


# Kuhn / Freudenthal split of the reference cube into 6 tets around the
# diagonal 0-6 (hex corner numbering as above)
_KUHN = [[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]]


def hex_to_tets(mesh):
    elems = []
    attr = []
    for e, a in zip(mesh.elems, mesh.attr):
        for t in _KUHN:
            tv = [e[t[0]], e[t[1]], e[t[2]], e[t[3]]]
            X = mesh.verts[tv]
            if np.linalg.det(X[1:] - X[0]) < 0:
                tv[2], tv[3] = tv[3], tv[2]
            elems.append(tv)
            attr.append(a)
    return Mesh("tet", mesh.verts, elems, attr)


def read_mfem_mesh(path):
    """MFEM mesh v1.0 text format, straight tets or hexes
    (/root/reference/data/beam-tet.mesh)."""
    with open(path) as f:
        lines = [ln.split("#")[0].strip() for ln in f]
    lines = [ln for ln in lines if ln]
    it = iter(lines)
    assert next(it).startswith("MFEM mesh v1.0")
    dim = None
    elems, attr, bdr, battr, verts = [], [], [], [], None
    geom_nv = {4: 4, 5: 8, 2: 3, 3: 4}  # tet, hex, triangle, quad
    kind = None
    for ln in it:
        if ln == "dimension":
            dim = int(next(it))
        elif ln == "elements":
            n = int(next(it))
            for _ in range(n):
                t = [int(s) for s in next(it).split()]
                kind = {4: "tet", 5: "hex"}[t[1]]
                attr.append(t[0])
                elems.append(t[2:2 + geom_nv[t[1]]])
        elif ln == "boundary":
            n = int(next(it))
            for _ in range(n):
                t = [int(s) for s in next(it).split()]
                battr.append(t[0])
                bdr.append(t[2:2 + geom_nv[t[1]]])
        elif ln == "vertices":
            n = int(next(it))
            vdim = int(next(it))
            verts = np.array([[float(s) for s in next(it).split()] for _ in range(n)])
            assert vdim == 3
    assert dim == 3
    return Mesh(kind, verts, elems, attr, None, np.array(bdr, dtype=np.int64), np.array(battr, dtype=np.int64))


def read_exodus(path):
    """ExodusII (NetCDF classic) reader for TETRA4 / HEX8 / HEX27 blocks; the
    files under /root/reference/data/*.g,*.gen.  Element attribute = block id;
    side sets become boundary elements with attribute = side-set id."""
    from scipy.io import netcdf_file

    f = netcdf_file(path, "r", mmap=False)
    var = f.variables
    if "coord" in var:
        xyz = np.array(var["coord"][:]).T.copy()
    else:
        xyz = np.stack([np.array(var["coordx"][:]), np.array(var["coordy"][:]), np.array(var["coordz"][:])], 1)
    blk_ids = np.array(var["eb_prop1"][:])
    elems, attr, kind, nodes27 = [], [], None, []
    for b, bid in enumerate(blk_ids):
        conn = np.array(var["connect%d" % (b + 1)][:]) - 1
        et = var["connect%d" % (b + 1)].elem_type
        et = et.decode() if isinstance(et, bytes) else str(et)
        et = et.upper()
        nn = conn.shape[1]
        if et.startswith("TET") and nn == 4:
            k = "tet"
            elems.append(conn)
        elif et.startswith("HEX") and nn == 8:
            k = "hex"
            elems.append(conn)
        elif et.startswith("HEX") and nn == 27:
            k = "hex"
            elems.append(conn[:, :8])
            nodes27.append(conn)
        else:
            raise ValueError("unsupported block type %s/%d" % (et, nn))
        assert kind in (None, k)
        kind = k
        attr.append(np.full(len(conn), bid))
    elems = np.concatenate(elems)
    attr = np.concatenate(attr)
    geom = None
    if nodes27:
        geom = xyz[np.concatenate(nodes27)]  # Exodus HEX27 node order
    # side sets -> boundary faces
    bdr, battr = [], []
    if "ss_prop1" in var:
        ss_ids = np.array(var["ss_prop1"][:])
        # Exodus side numbering -> local vertex lists (outward)
        hex_sides = [[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [0, 4, 7, 3], [0, 3, 2, 1], [4, 5, 6, 7]]
        tet_sides = [[0, 1, 3], [1, 2, 3], [0, 3, 2], [0, 2, 1]]
        sides = hex_sides if kind == "hex" else tet_sides
        for s, sid in enumerate(ss_ids):
            el = np.array(var["elem_ss%d" % (s + 1)][:]) - 1
            sd = np.array(var["side_ss%d" % (s + 1)][:]) - 1
            for e, q in zip(el, sd):
                bdr.append([elems[e][m] for m in sides[q]])
                battr.append(sid)
    f.close()
    # drop unused nodes (HEX27 mid nodes) from the corner-vertex numbering? keep all: harmless
    return Mesh(kind, xyz, elems, attr, geom, np.array(bdr, dtype=np.int64) if bdr else None,
                np.array(battr, dtype=np.int64) if bdr else None)


class Topology:
    """Edges / faces numbered by first appearance; orientation data."""

    def __init__(self, mesh):
        self.mesh = mesh
        hexa = mesh.kind == "hex"
        ledges = HEX_EDGES if hexa else TET_EDGES
        lfaces = HEX_FACES if hexa else TET_FACES
        self.ledges, self.lfaces = ledges, lfaces
        edge_id = {}
        face_id = {}
        self.face_verts = []  # base vertex order of each face (first appearance)
        self.face_elems = []
        self.el_edges = np.zeros((mesh.n_elem, len(ledges)), dtype=np.int64)
        self.el_faces = np.zeros((mesh.n_elem, len(lfaces)), dtype=np.int64)
        for e, ev in enumerate(mesh.elems):
            for le, (a, b) in enumerate(ledges):
                key = (min(ev[a], ev[b]), max(ev[a], ev[b]))
                if key not in edge_id:
                    edge_id[key] = len(edge_id)
                self.el_edges[e, le] = edge_id[key]
        for e, ev in enumerate(mesh.elems):
            for lf, fv in enumerate(lfaces):
                gv = tuple(int(ev[m]) for m in fv)
                key = tuple(sorted(gv))
                if key not in face_id:
                    face_id[key] = len(face_id)
                    self.face_verts.append(gv)
                    self.face_elems.append([e])
                else:
                    self.face_elems[face_id[key]].append(e)
                self.el_faces[e, lf] = face_id[key]
        self.n_edges = len(edge_id)
        self.n_faces = len(face_id)
        self.edge_id = edge_id
        self.face_id = face_id
        self.edge_verts = np.zeros((self.n_edges, 2), dtype=np.int64)
        for (a, b), i in edge_id.items():
            self.edge_verts[i] = (a, b)

    def boundary_faces(self):
        return [f for f in range(self.n_faces) if len(self.face_elems[f]) == 1]
