"""Domain decomposition across ranks (one per GPU): the synthetic structured hex cube (BrickProblem, nothing global is
materialised) and any mesh under a host-supplied element partition (GeneralProblem, at the end of the file).

Stands in for what Hephaestus gets from ``mfem::ParMesh(MPI_COMM_WORLD, mesh)``
(/root/reference/examples/team7/Main.cpp:155) + ``ParFiniteElementSpace``: element partition,
local (L-vector) dof numbering, true-dof ownership (lowest rank of the sharing group owns, MFEM's
convention) and the shared-dof lists with relative orientations that ``hcurl_space_create`` takes
(include/hcurl_b200.h).  In plugged-in use those arrays come from MFEM's GroupTopology through
adaptor/heph_gpu_operator.hpp; this module only has to be self-consistent.

Each rank builds ONLY its own brick (n^3 elements) -- nothing global is ever materialised, so the
1e9-dof weak-scaling runs fit.  Interface entities are matched through global vertex ids, which
are computable locally on a structured grid.
"""
import numpy as np

from . import fespace


def brick_dims(world):
    """(px, py, pz) with px*py*pz == world, as cubic as possible, px >= py >= pz."""
    best = None
    for pz in range(1, world + 1):
        if world % pz:
            continue
        for py in range(pz, world // pz + 1):
            if (world // pz) % py:
                continue
            px = world // (pz * py)
            if px < py:
                continue
            cand = (px - pz, px, py, pz)
            if best is None or cand < best:
                best = cand
    return best[1:]


class LocalSpace:
    """One rank's share: arrays in the layout of hcurl_space_create."""
    pass


def _hash_unit(gid, salt):
    """Deterministic U(0,1) from global ids (same on every rank), splitmix64."""
    with np.errstate(over="ignore"):  # arithmetic modulo 2^64 is the point
        z = gid.astype(np.uint64) + np.uint64((salt * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)


class BrickProblem:
    """rank's brick of the (px n) x (py n) x (pz n) cube [0,px] x [0,py] x [0,pz] (unit bricks)."""

    def __init__(self, rank, world, n, p, perturbed=False, dims=None):
        self.rank, self.world, self.n, self.p = rank, world, n, p
        self.dims = px, py, pz = dims or brick_dims(world)
        assert px * py * pz == world
        self.brick = bx, by, bz = rank % px, (rank // px) % py, rank // (px * py)
        N = np.array([px * n, py * n, pz * n])
        self.N = N
        mesh = fespace.cartesian_hex_mesh(n, n, n)
        # global grid coordinates / ids of the local vertices (local order is monotone in global id)
        g = np.rint(mesh.verts * n).astype(np.int64) + np.array([bx, by, bz]) * n
        self.vert_grid = g
        self.vert_gid = g[:, 0] + (N[0] + 1) * (g[:, 1] + (N[1] + 1) * g[:, 2])
        verts = g / float(n)
        if perturbed:
            interior = np.all((g > 0) & (g < N), axis=1)
            h = 1.0 / n
            d = np.stack([_hash_unit(self.vert_gid, 777 + k) for k in range(3)], 1) * 0.4 * h - 0.2 * h
            verts = verts + d * interior[:, None]
        # attribute by global x-centre, as in the serial benchmark mesh
        cx = verts[mesh.elems][:, :, 0].mean(axis=1)
        attr = np.where(cx < 0.5 * px, 1, 2)
        self.mesh = fespace.Mesh("hex", verts, mesh.elems, attr)
        self.space = sp = fespace.NDSpace(self.mesh, p)
        self._build_sharing()

    # ---- which bricks contain a grid point --------------------------------------------------
    def _owners_1d(self, c, axis):
        """ranks-along-axis containing coordinate c (vector): (lo, hi) brick indices inclusive."""
        n, P = self.n, self.dims[axis]
        lo = np.maximum((c - 1) // n, 0)          # brick whose closed range [b n, (b+1) n] contains c
        lo = np.where(c % n == 0, np.maximum(c // n - 1, 0), c // n)
        hi = np.minimum(c // n, P - 1)
        return lo, hi

    def _entity_ranks(self, vg):
        """vg: (m, k, 3) grid coords of the k vertices of m entities -> list of rank sets
        as (m, 8) array padded with -1 (bricks containing ALL the vertices)."""
        px, py, pz = self.dims
        los, his = [], []
        for a in range(3):
            lo, hi = self._owners_1d(vg[:, :, a], a)
            los.append(lo.max(axis=1))
            his.append(hi.min(axis=1))
        m = len(vg)
        out = -np.ones((m, 8), dtype=np.int64)
        cnt = np.zeros(m, dtype=np.int64)
        for dz in range(2):
            for dy in range(2):
                for dx in range(2):
                    b = [los[0] + dx, los[1] + dy, los[2] + dz]
                    ok = (b[0] <= his[0]) & (b[1] <= his[1]) & (b[2] <= his[2])
                    r = b[0] + px * (b[1] + py * b[2])
                    idx = np.nonzero(ok)[0]
                    out[idx, cnt[idx]] = r[idx]
                    cnt[idx] += 1
        return out

    def _face_base_gids(self, fg, side_rank_is_low):
        """Base vertex order (global ids) of interface faces as the rank on the low / high side of
        the face numbers them: the face is local face +a of the low-side element and local face -a
        of the high-side element; NDSpace takes the base order from that element's local face."""
        # fg: (m,4,3) grid coords of the face corners (any order).  Normal axis a: constant coord.
        m = len(fg)
        N = self.N
        out = np.zeros((m, 4), dtype=np.int64)
        if m == 0:
            return out
        const = np.stack([np.all(fg[:, :, a] == fg[:, :1, a], axis=1) for a in range(3)], 1)
        ax = np.argmax(const, axis=1)
        lo = fg.min(axis=1)  # lower corner of the face
        e0 = lo.copy()
        e0[np.arange(m), ax] -= np.asarray(side_rank_is_low, dtype=np.int64)
        corners = e0[:, None, :] + fespace.HEX_VERTS[None]          # (m, 8, 3) in MFEM corner order
        want = lo[np.arange(m), ax]
        for lf in fespace.HEX_FACES:
            c = corners[:, lf]                                        # (m, 4, 3)
            inplane = np.all(c[np.arange(m), :, ax] == want[:, None], axis=1)
            gidf = c[:, :, 0] + (N[0] + 1) * (c[:, :, 1] + (N[1] + 1) * c[:, :, 2])
            out[inplane] = gidf[inplane]
        return out

    def _build_sharing(self):
        sp, p, me = self.space, self.p, self.rank
        g, gid = self.vert_grid, self.vert_gid
        ned, nfd = sp.ned, sp.nfd
        # ---- shared edges / faces and their rank sets
        eg = g[sp.edge_verts]                                  # (n_edges, 2, 3)
        er = self._entity_ranks(eg)
        e_sh = np.nonzero((er >= 0).sum(axis=1) > 1)[0]
        fg = g[sp.face_verts]                                  # (n_faces, 4, 3)
        fr = self._entity_ranks(fg)
        f_sh = np.nonzero((fr >= 0).sum(axis=1) > 1)[0]
        # global keys for a rank-independent ordering
        ekey = np.sort(gid[sp.edge_verts[e_sh]], axis=1)
        fkey = np.sort(gid[sp.face_verts[f_sh]], axis=1)
        e_own = np.where(er[e_sh] >= 0, er[e_sh], 1 << 30).min(axis=1)
        f_own = np.where(fr[f_sh] >= 0, fr[f_sh], 1 << 30).min(axis=1)
        # ---- ownership of local dofs
        owned = np.ones(sp.ndof, dtype=bool)
        for k in range(ned):
            owned[e_sh[e_own != me] * ned + k] = False
        for k in range(nfd):
            owned[sp.face_off + f_sh[f_own != me] * nfd + k] = False
        ltdof = -np.ones(sp.ndof, dtype=np.int64)
        ltdof[owned] = np.arange(int(owned.sum()))
        # ---- face dof correspondence ghost <-> owner (signed permutation by relative orientation)
        if nfd:
            perm, sign = fespace._quad_orientation_tables(p)
        nbrs = sorted(set(er[e_sh][er[e_sh] >= 0].tolist() + fr[f_sh][fr[f_sh] >= 0].tolist()) - {me})
        own_off, own_ldof, ghost_off, ghost_ldof, ghost_sign = [0], [], [0], [], []
        for r in nbrs:
            # edges shared with r, in global-key order
            has_r = (er[e_sh] == r).any(axis=1)
            for owner_is_me, acc in ((True, own_ldof), (False, ghost_ldof)):
                sel = np.nonzero(has_r & (e_own == (me if owner_is_me else r)))[0]
                order = np.lexsort((ekey[sel, 1], ekey[sel, 0]))
                ed = e_sh[sel][order]
                # local vertex order is monotone in the global id: edge orientation (low -> high) agrees
                acc.append((ed[:, None] * ned + np.arange(ned)[None]).ravel())
                if not owner_is_me:
                    ghost_sign.append(np.ones(len(ed) * ned, dtype=np.int8))
            if nfd:
                has_r = (fr[f_sh] == r).any(axis=1)
                for owner_is_me in (True, False):
                    sel = np.nonzero(has_r & (f_own == (me if owner_is_me else r)))[0]
                    order = np.lexsort(tuple(fkey[sel, k] for k in (3, 2, 1, 0)))
                    fd = f_sh[sel][order]
                    base = sp.face_off + fd[:, None] * nfd
                    if owner_is_me:
                        own_ldof.append((base + np.arange(nfd)[None]).ravel())
                        continue
                    # my base order (global ids) vs the owner's base order of the same face
                    mine = gid[sp.face_verts[fd]]
                    lo_side = self._low_side_rank(fg[fd]) == r     # owner r is on the low side?
                    theirs = self._face_base_gids(fg[fd], lo_side)
                    pos0 = np.argmax(theirs == mine[:, :1], axis=1)
                    nxt = theirs[np.arange(len(fd)), (pos0 + 1) % 4]
                    flip = (nxt != mine[:, 1]).astype(np.int64)
                    code = 2 * pos0 + flip
                    # my dof m == sign[code][m] * owner dof perm[code][m]; list my dofs in owner order
                    inv = np.argsort(perm[code], axis=1)          # owner dof k  ->  my dof m
                    ghost_ldof.append((base + inv).ravel())
                    ghost_sign.append(np.take_along_axis(sign[code], inv, axis=1).astype(np.int8).ravel())
            own_off.append(sum(len(a) for a in own_ldof))
            ghost_off.append(sum(len(a) for a in ghost_ldof))
        L = LocalSpace()
        L.elem_xyz = self.mesh.elem_xyz()
        L.elem_ldof = sp.signed_elem_dofs()
        L.elem_verts, L.elem_face_base = None, None
        L.n_ldof = sp.ndof
        L.ldof_ltdof = ltdof.astype(np.int32)
        L.n_tdof = int(owned.sum())
        L.nbr_rank = np.array(nbrs, dtype=np.int32)
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        L.own_off, L.own_ldof = np.array(own_off, np.int32), cat(own_ldof, np.int32)
        L.ghost_off, L.ghost_ldof = np.array(ghost_off, np.int32), cat(ghost_ldof, np.int32)
        L.ghost_sign = cat(ghost_sign, np.int8)
        self.local = L
        # essential true dofs: boundary of the GLOBAL cube, owned by this rank
        N = self.N
        def on_bdr(cen):
            gc = cen * self.n
            return np.any((np.abs(gc) < 1e-6) | (np.abs(gc - N[None, :]) < 1e-6), axis=1)
        # use unperturbed grid centroids of the local boundary faces
        bmesh = fespace.Mesh("hex", self.vert_grid / float(self.n), self.mesh.elems, self.mesh.attr)
        bsp = sp if not np.any(self.mesh.verts != bmesh.verts) else None
        ess_l = self._boundary_ldofs(on_bdr)
        self.ess_ldofs = ess_l
        self.ess_tdofs = ltdof[ess_l][ltdof[ess_l] >= 0].astype(np.int32)
        del bsp

    def _low_side_rank(self, fg):
        """rank of the brick on the low side of each interface face (fg: (m,4,3) grid coords)."""
        px, py, pz = self.dims
        n = self.n
        const = np.stack([np.all(fg[:, :, a] == fg[:, :1, a], axis=1) for a in range(3)], 1)
        ax = np.argmax(const, axis=1)
        lo = fg.min(axis=1).copy()
        lo[np.arange(len(fg)), ax] -= 1
        b = lo // n
        return b[:, 0] + px * (b[:, 1] + py * b[:, 2])

    def _boundary_ldofs(self, on_bdr):
        """local dofs on faces of the global cube boundary (faces with one local element AND lying
        on the global boundary), with their edges."""
        sp = self.space
        gridmesh_verts = self.vert_grid / float(self.n)
        lfaces = fespace.HEX_FACES
        isb = sp.face_count[sp.el_faces] == 1
        e_idx, lf_idx = np.nonzero(isb)
        cen = gridmesh_verts[self.mesh.elems[e_idx[:, None], lfaces[lf_idx]]].mean(axis=1)
        keep = on_bdr(cen)
        e_idx, lf_idx = e_idx[keep], lf_idx[keep]
        out = []
        if sp.nfd:
            f = sp.el_faces[e_idx, lf_idx]
            out.append((sp.face_off + f[:, None] * sp.nfd + np.arange(sp.nfd)[None]).ravel())
        f2e = np.array([[le for le, (a, b) in enumerate(fespace.HEX_EDGES) if a in fv and b in fv] for fv in lfaces])
        ge = sp.el_edges[e_idx[:, None], f2e[lf_idx]]
        out.append((ge.ravel()[:, None] * sp.ned + np.arange(sp.ned)[None]).ravel())
        return np.unique(np.concatenate(out)).astype(np.int64) if out else np.zeros(0, np.int64)


def brick_problem(rank, world, n, p, perturbed=False, dims=None):
    return BrickProblem(rank, world, n, p, perturbed, dims)


# ---- host-side reference of the exchange semantics (tests; mirrors csrc/comm.cu) --------------
def to_ext(local, x_t):
    """T-vector -> extended vector [owned | ghosts] (ghosts zero)."""
    n_ghost = int(local.ghost_off[-1])
    return np.concatenate([x_t, np.zeros(n_ghost)])


def ext_index(local):
    """extended-vector index of every local dof."""
    l2e = local.ldof_ltdof.astype(np.int64).copy()
    l2e[local.ghost_ldof] = local.n_tdof + np.arange(len(local.ghost_ldof))
    assert np.all(l2e >= 0)
    return l2e


# ---- general (unstructured) meshes: host-supplied element partition -------------------------------------------------
class GeneralProblem:
    """One rank's share of ANY mesh under a host-supplied element -> rank array ("METIS / host-supplied
    partition array for unstructured" -- what mfem::ParMesh(comm, mesh, partitioning) + ParFiniteElementSpace produce
    from METIS output, /root/reference/examples/team7/Main.cpp:155).  Built from the serial space, so every rank uses
    the serial space's dof orientation and face-dof ordering: a shared dof is the same functional on all ranks
    (ghost_sign = +1), the lists of a neighbour pair are matched by ascending serial dof number, and the owner of a
    shared dof is the lowest rank that holds it (MFEM's convention).  Suited to real meshes (team7: 57 905 tets); the
    synthetic 1e9-dof cubes use BrickProblem, which never materialises anything global.

    attributes: elems (serial element ids of this rank), l2g (local dof -> serial dof), local (LocalSpace),
    ess_tdofs (true-dof indices of the given serial essential dofs that this rank owns), space (the serial space)."""

    def __init__(self, space, elem_rank, rank, ess_dofs=None):
        mesh = space.mesh
        elem_rank = np.asarray(elem_rank)
        assert elem_rank.shape == (mesh.n_elem,)
        sd = np.asarray(space.signed_elem_dofs()).astype(np.int64)          # serial dof, -1-k = flipped
        gd = np.where(sd >= 0, sd, -1 - sd)
        # lowest rank holding each serial dof, and who else holds it
        world = int(elem_rank.max()) + 1
        owner = np.full(space.ndof, world, dtype=np.int64)
        np.minimum.at(owner, gd.ravel(), np.repeat(elem_rank, gd.shape[1]))
        self.rank, self.space = rank, space
        self.elems = np.nonzero(elem_rank == rank)[0]
        mine = gd[self.elems]
        self.l2g = np.unique(mine)                                          # ascending serial numbers
        g2l = -np.ones(space.ndof, dtype=np.int64)
        g2l[self.l2g] = np.arange(len(self.l2g))
        loc = g2l[mine]
        L = LocalSpace()
        L.elem_xyz = mesh.elem_xyz()[self.elems]
        L.elem_ldof = np.where(sd[self.elems] >= 0, loc, -1 - loc).astype(np.int32)
        L.elem_verts = mesh.elems[self.elems].astype(np.int32) if mesh.kind == "tet" else None
        fb = getattr(space, "elem_face_base", None)
        L.elem_face_base = np.asarray(fb)[self.elems].astype(np.int32) if (mesh.kind == "tet" and fb is not None and space.p >= 2) else None
        L.n_ldof = len(self.l2g)
        owned = owner[self.l2g] == rank
        ltdof = -np.ones(L.n_ldof, dtype=np.int64)
        ltdof[owned] = np.arange(int(owned.sum()))
        L.ldof_ltdof = ltdof.astype(np.int32)
        L.n_tdof = int(owned.sum())
        # which other ranks hold my dofs
        held = {}                                                           # rank q -> serial dofs q holds that I hold too
        for q in range(world):
            if q == rank:
                continue
            hq = np.zeros(space.ndof, dtype=bool)
            hq[gd[elem_rank == q].ravel()] = True
            common = self.l2g[hq[self.l2g]]
            if len(common):
                held[q] = common
        own_off, own_ldof, ghost_off, ghost_ldof = [0], [], [0], []
        nbrs = []
        for q in sorted(held):
            common = held[q]                                                # ascending serial numbers on both sides
            mine_owned = common[owner[common] == rank]
            theirs = common[owner[common] == q]
            if len(mine_owned) == 0 and len(theirs) == 0:
                continue                                                    # shared only through dofs a third rank owns
            nbrs.append(q)
            own_ldof.append(g2l[mine_owned])
            ghost_ldof.append(g2l[theirs])
            own_off.append(own_off[-1] + len(mine_owned))
            ghost_off.append(ghost_off[-1] + len(theirs))
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        L.nbr_rank = np.array(nbrs, dtype=np.int32)
        L.own_off, L.own_ldof = np.array(own_off, np.int32), cat(own_ldof, np.int32)
        L.ghost_off, L.ghost_ldof = np.array(ghost_off, np.int32), cat(ghost_ldof, np.int32)
        L.ghost_sign = np.ones(len(L.ghost_ldof), dtype=np.int8)
        # every dof I do not own must arrive from exactly one neighbour (its owner)
        assert len(L.ghost_ldof) == int((~owned).sum()) and len(np.unique(L.ghost_ldof)) == len(L.ghost_ldof)
        self.local = L
        if ess_dofs is not None:
            e = np.asarray(ess_dofs, dtype=np.int64)
            e = e[g2l[e] >= 0]
            t = ltdof[g2l[e]]
            self.ess_tdofs = np.sort(t[t >= 0]).astype(np.int32)
        else:
            self.ess_tdofs = np.zeros(0, np.int32)


def general_problem(space, elem_rank, rank, ess_dofs=None):
    return GeneralProblem(space, elem_rank, rank, ess_dofs)


def partition_by_coordinate(mesh, world):
    """A simple host-side element partition for tests and examples (recursive coordinate bisection of the element
    centroids into `world` parts of near-equal size); production use passes METIS output instead."""
    cen = mesh.verts[mesh.elems].mean(axis=1)
    part = np.zeros(mesh.n_elem, dtype=np.int64)

    def split(idx, lo, hi):
        if hi - lo == 1:
            part[idx] = lo
            return
        mid = (lo + hi) // 2
        ax = int(np.argmax(cen[idx].max(axis=0) - cen[idx].min(axis=0)))
        order = idx[np.argsort(cen[idx, ax], kind="stable")]
        k = len(order) * (mid - lo) // (hi - lo)
        split(order[:k], lo, mid)
        split(order[k:], mid, hi)
    split(np.arange(mesh.n_elem), 0, world)
    return part
