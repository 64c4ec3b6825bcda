"""Host-side patch decomposition (hephaestus_b200/csrc/patch.cpp) -- runs without a GPU.

The patch lists must reproduce exactly the signed gather x_e = G x and scatter-add y = G^T y_e
of the element -> dof map (FiniteElementSpace::GetElementDofs semantics) for any mesh."""
import ctypes as C

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace

from helpers import star_mesh

NEG = 0x80000000


def build(space, fe, slot_code, centroid, stride=None, unit=8, n_units=None, exchange=None):
    lib = L.lib()
    ne, nd = slot_code.shape
    stride = stride or fe + 1
    h = L.check_ptr(lib.hcurl_host_patches_create(nd, ne, fe, stride, unit, n_units or space.ndof, L.ptr(centroid), L.ptr(slot_code),
                                                  None if exchange is None else L.ptr(exchange)), "patches")
    SWIZZLE["n_boundary"] = int(lib.hcurl_host_patches_boundary_count(h))
    npatch = lib.hcurl_host_patches_count(h)
    nw = lib.hcurl_host_patches_blob_words(h)
    perm = np.zeros(npatch * fe, np.int32)
    off = np.zeros(npatch + 1, np.uint32)
    blob = np.zeros(nw, np.uint32)
    L.check(lib.hcurl_host_patches_copy(h, L.ptr(perm), L.ptr(off), L.ptr(blob)), "copy")
    assert lib.hcurl_host_patches_max_blob_words(h) == int(np.diff(off.astype(np.int64)).max()) * 4
    assert lib.hcurl_host_patches_max_dofs(h) == max(int(blob[int(o) * 4]) for o in off[:-1])
    part = np.zeros(2 * npatch, np.uint32)
    md, ml, mt = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    L.check(lib.hcurl_host_patches_parts(h, L.ptr(part), C.byref(md), C.byref(ml), C.byref(mt)), "parts")
    part = part.astype(np.int64).reshape(npatch, 2)
    assert np.all(off[:-1] < part[:, 0]) and np.all(part[:, 0] <= part[:, 1]) and np.all(part[:, 1] < off[1:])
    assert md.value == int((part[:, 0] - off[:-1].astype(np.int64)).max()) * 4 and ml.value == int((part[:, 1] - part[:, 0]).max()) * 4
    assert mt.value == int((off[1:].astype(np.int64) - part[:, 1]).max()) * 4
    SWIZZLE["part"] = part
    own_off = np.zeros(npatch + 1, np.uint32)
    n_own = lib.hcurl_host_patches_owned(h, L.ptr(own_off), None)
    own_idx = np.zeros(max(n_own, 1), np.uint32)
    assert lib.hcurl_host_patches_owned(h, None, L.ptr(own_idx)) == n_own == int(own_off[-1])
    SWIZZLE["owned"] = (own_off.astype(np.int64), own_idx.astype(np.int64))
    nb = 128 // unit
    row, grp = np.zeros(nd, np.uint8), np.zeros((fe + nb - 1) // nb, np.uint8)
    L.check(lib.hcurl_host_patches_swizzle(h, L.ptr(row), L.ptr(grp)), "swizzle")
    assert row.max(initial=0) < nb and grp.max(initial=0) < nb
    SWIZZLE.update(row=row.astype(np.int64), grp=grp.astype(np.int64), nb=nb)
    lib.hcurl_host_patches_destroy(h)
    return perm, off.astype(np.int64), blob, stride


SHARED, OWNED = 0x80000000, 0x10000000


HDR = 32


def decode_patch(k, off, blob, nd=None, fe=None):
    """Blob of patch k (layout: hephaestus_b200/csrc/patch.cpp) ->
    (unit ids (U,), [(first u, unit ids, None for private units | owned-by-this-patch mask for shared ones, E byte offsets
     (n, W))] per class -- private with 1 / 2 / 3..4 slots, then shared with 1 / 2 / 3..4 slots --, overflow words,
     dependency patches
     [, lidx (nd, fe), number of real elements -- when nd and fe are given])."""
    B = blob[off[k] * 4:off[k + 1] * 4]
    U, n_real, nov = [int(v) for v in B[:3]]
    assert B[3] == 0 and B[17] == 0
    n = [int(v) for v in B[4:10]]
    oL = [int(v) for v in B[10:16]]
    oov, wd, wl = int(B[16]), int(B[18]), int(B[19])
    assert np.all(B[20:HDR] == 0) and U == sum(n)
    pd0, pt0 = SWIZZLE["part"][k]
    assert wd == (pd0 - off[k]) * 4 and wl == (pt0 - pd0) * 4 and wd >= HDR + U
    assert all(o % 4 == 0 for o in oL + [oov]) and len(B) % 4 == 0
    word = B[HDR:HDR + U].astype(np.int64)
    assert np.all(word & 0x60000000 == 0)                        # essential flags are set by the operator
    assert np.all(((word & OWNED) == 0) | ((word & SHARED) != 0))   # only shared units have an owner flag
    dof, multi, owned = word & 0x0FFFFFFF, (word & SHARED) != 0, (word & OWNED) != 0
    Lw = B[wd:wd + wl]
    sh = 16 * np.arange(4, dtype=np.uint64)
    groups, u0 = [], 0
    for c in range(6):
        w, cnt, o = c % 3, n[c], oL[c]
        if w == 0:
            ent = Lw[o:o + (cnt + 1) // 2].view(np.uint16)[:cnt].astype(np.int64)[:, None]
        elif w == 1:
            ent = (Lw[o:o + cnt, None].astype(np.uint64) >> sh[None, :2]).astype(np.int64) & 0xFFFF
        else:
            ent = (Lw[o:o + 2 * cnt].view(np.uint64)[:, None] >> sh[None]).astype(np.int64) & 0xFFFF
        groups.append((u0, dof[u0:u0 + cnt], None if c < 3 else owned[u0:u0 + cnt], ent))
        assert np.all(multi[u0:u0 + cnt] == (c >= 3))
        u0 += cnt
    ov = Lw[oov:oov + nov].astype(np.int64)
    oo, oi = SWIZZLE["owned"]
    deps = oi[oo[k]:oo[k + 1]]                 # the shared units this patch owns (zeroed ahead of the sums by its CTA)
    assert np.all(np.diff(deps) > 0) and np.array_equal(deps, np.sort(dof[owned]))
    if nd is None:
        return dof, groups, ov, deps
    T = B[wd + wl:]
    lidx = T[:(nd * fe + 1) // 2].view(np.uint16)[:nd * fe].astype(np.int64).reshape(nd, fe)
    assert (nd * fe + 1) // 2 <= len(T)
    return dof, groups, ov, deps, lidx, n_real


SWIZZLE = {}   # bank swizzle of the most recent build(): row[s], grp[j // nb]


def e_index(s, j, st):
    """E-vector position (in units) of slot s of patch position j (swizzled inside aligned runs of nb positions)."""
    return s * st + (j ^ SWIZZLE["row"][s] ^ SWIZZLE["grp"][j // SWIZZLE["nb"]])


def make(kind, n, p, perturb):
    if kind == "star":
        m = star_mesh(*n)
    else:
        m = fespace.cartesian_hex_mesh(*n)
        if perturb:
            m = fespace.perturb_interior(m)
    if kind == "tet":
        m = fespace.hex_to_tets(m)
    sp = fespace.NDSpace(m, p)
    sd = sp.signed_elem_dofs().astype(np.int64)
    idx = np.where(sd >= 0, sd, -1 - sd)
    code = (idx | np.where(sd < 0, NEG, 0)).astype(np.uint32)
    cen = np.ascontiguousarray(m.verts[m.elems].mean(axis=1))
    return m, sp, np.ascontiguousarray(code), cen, idx, np.where(sd < 0, -1.0, 1.0)


@pytest.mark.parametrize("kind,n,p,fe", [("hex", (4, 4, 4), 2, 64), ("hex", (5, 3, 7), 2, 32), ("hex", (8, 8, 4), 2, 128),
                                         ("hex", (3, 3, 3), 1, 32), ("hex", (3, 2, 2), 3, 32), ("tet", (3, 3, 2), 2, 64),
                                         ("star", (5,), 2, 32), ("star", (2,), 1, 64)])
@pytest.mark.parametrize("perturb", [False, True])
def test_patch_lists_reproduce_gather_and_scatter(kind, n, p, fe, perturb):
    m, sp, code, cen, idx, sgn = make(kind, n, p, perturb)
    ne, nd = code.shape
    perm, off, blob, st = build(sp, fe, code, cen)
    npatch = len(off) - 1
    assert npatch == (ne + fe - 1) // fe
    n_ovf = sum(len(decode_patch(k, off, blob)[2]) for k in range(npatch))
    if kind in ("star", "tet"):
        assert n_ovf > 0  # these meshes have dofs with more than four slots in a patch
    elif kind == "hex":
        assert n_ovf == 0
    # perm is a permutation of the elements (+ padding)
    real = perm[perm >= 0]
    assert np.array_equal(np.sort(real), np.arange(ne))
    assert np.all(perm[ne:] == -1)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(sp.ndof)
    ye = rng.standard_normal((ne, nd))
    y_ref = np.zeros(sp.ndof)
    np.add.at(y_ref, idx.ravel(), (sgn * ye).ravel())
    y = np.full(sp.ndof, 1e30)                  # nothing is cleared: the owner's store comes first
    touched = np.zeros(sp.ndof, np.int64)
    filled_by = np.full(sp.ndof, -1)
    elem_patch = np.empty(ne, np.int64)
    elem_patch[perm[perm >= 0]] = np.nonzero(perm >= 0)[0] // fe
    pmin = np.full(sp.ndof, 1 << 40)
    pmax = np.full(sp.ndof, -1)
    np.minimum.at(pmin, idx.ravel(), np.repeat(elem_patch, nd))
    np.maximum.at(pmax, idx.ravel(), np.repeat(elem_patch, nd))
    sl = np.arange(nd)
    for k in range(npatch):
        pd_all, groups, ov, deps, lidx, n_real_blob = decode_patch(k, off, blob, nd, fe)
        el = perm[k * fe:(k + 1) * fe]
        n_real = int((el >= 0).sum())
        assert n_real == n_real_blob
        # element side: pull x through lidx (position in the patch dof list | sign bit 15)
        Xs = x[pd_all]
        neg_l = (lidx & 0x8000) != 0
        pos_l = lidx & 0x1FFF
        assert np.all(lidx & 0x6000 == 0)                       # 8-byte units: only the sign flag
        assert np.all(pos_l[:, :n_real] < len(pd_all))
        for j in range(n_real):
            xe = np.where(neg_l[:, j], -Xs[pos_l[:, j]], Xs[pos_l[:, j]])
            assert np.array_equal(xe, sgn[el[j]] * x[idx[el[j]]])
        # element results, signed on the element side, in the paired E layout
        Ey = np.full(nd * st, np.nan)
        for j in range(n_real):
            Ey[e_index(sl, j, st)] = np.where(neg_l[:, j], -ye[el[j]], ye[el[j]])
        n_slots = 0
        for u0, pd, own, ent in groups:
            # ascending global index up to the bank-spreading permutation inside aligned runs of 16 positions
            run = (u0 + np.arange(len(pd))) // 16
            srt = np.sort(pd)
            assert np.all(np.diff(srt) > 0)
            for r in np.unique(run):
                assert np.array_equal(np.sort(pd[run == r]), srt[run == r])
            assert np.array_equal(pd, pd_all[u0:u0 + len(pd)])
            W = ent.shape[1]
            used = ent != 0xFFFF
            if W < 4:
                assert used.all()
            else:
                assert used[:, :3].all()                         # only the 4th entry may be unused
            assert np.all(ent[used] % 8 == 0)
            n_slots += int(used.sum())
            # private units are stored; the owner of a shared unit is the lowest-numbered patch touching it: its CTA zeroes
            # the unit ahead of time, every patch (the owner included) adds once the zeroing watermark has passed
            if own is None:
                assert np.all(pmin[pd] == k) and np.all(pmax[pd] == k) and np.all(touched[pd] == 0)
            else:
                assert np.array_equal(own, pmin[pd] == k)
                y[pd[own]] = 0.0
                filled_by[pd[own]] = k
                assert np.all(filled_by[pd] == pmin[pd])

            for u in range(len(pd)):
                v = np.sum(Ey[ent[u][used[u]] // 8])
                if own is None:
                    y[pd[u]] = v
                else:
                    y[pd[u]] += v
                touched[pd[u]] += 1
        assert n_slots + len(ov) == n_real * nd
        for w in ov:                                             # overflow slots: units with more than four slots are classed shared
            u = w >> 16
            assert groups[5][0] <= u < groups[5][0] + len(groups[5][1])
            y[pd_all[u]] += Ey[(w & 0xFFFF) // 8]
    assert np.all(touched >= 1)
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-13)


def test_structured_cube_patches_are_blocks():
    """On a power-of-two structured cube the Morton patches are 4x4x4 element blocks."""
    m, sp, code, cen, idx, sgn = make("hex", (8, 8, 8), 2, False)
    perm, off, blob, st = build(sp, 64, code, cen)
    for k in range(len(off) - 1):
        c = cen[perm[k * 64:(k + 1) * 64]]
        ext = c.max(0) - c.min(0)
        np.testing.assert_allclose(ext, 3.0 / 8, atol=1e-12)
        # 4x4x4 block of order-2 hexes: 1944 unique dofs, 1176 of them interior to the block
        # (here blocks touch the domain boundary, which only adds private dofs)
        pd, groups, ov, deps = decode_patch(k, off, blob)
        assert len(pd) == 1944
        assert sum(len(g[1]) for g in groups[:3]) >= 1176


def test_bad_arguments_are_reported():
    lib = L.lib()
    code = np.array([[5, 1]], dtype=np.uint32)
    cen = np.zeros((1, 3))
    assert not lib.hcurl_host_patches_create(2, 1, 32, 33, 8, 3, L.ptr(cen), L.ptr(code), None)
    assert b"out of range" in lib.hcurl_last_error()
    assert not lib.hcurl_host_patches_create(2, 1, 20000, 20000, 8, 8, L.ptr(cen), L.ptr(code), None)
    assert b"too large" in lib.hcurl_last_error()
    assert not lib.hcurl_host_patches_create(2, 1, 32, 33, 12, 8, L.ptr(cen), L.ptr(code), None)
    assert b"unit size" in lib.hcurl_last_error()


SWAP, NEG1, NEG0 = 0x20000000, 0x40000000, 0x80000000


@pytest.mark.parametrize("n,fe,padded", [((4, 4, 4), 64, True), ((5, 3, 4), 32, True), ((5, 3, 4), 32, False)])
def test_pair_units_with_orientation_flags(n, fe, padded):
    """16-byte units: the dof side moves aligned pairs of doubles; each element slot pair carries a
    swap flag and one sign per entry (hex_patch.cu builds these from the order-2 map; here random
    flags exercise every combination)."""
    m = fespace.cartesian_hex_mesh(*n)
    sp = fespace.NDSpace(m, 2)
    sd = sp.signed_elem_dofs().astype(np.int64)
    ne = sd.shape[0]
    npair, n_units = 27, sp.ndof // 2
    rng = np.random.default_rng(5)
    # any map from element slot pairs to global pairs will do for the builder: reuse the even dofs
    idx = np.where(sd >= 0, sd, -1 - sd)[:, ::2] // 2
    flags = rng.integers(0, 8, size=idx.shape).astype(np.uint32) << 29
    code = np.ascontiguousarray((idx.astype(np.uint32) | flags).astype(np.uint32))
    cen = np.ascontiguousarray(m.verts[m.elems].mean(axis=1))
    perm, off, blob, st = build(sp, fe, code, cen, stride=None if padded else fe, unit=16, n_units=n_units)   # any slot stride >= elements per patch
    x = rng.standard_normal((n_units, 2))
    ye = rng.standard_normal((ne, npair, 2))
    sw, n1, n0 = (flags & SWAP) != 0, (flags & NEG1) != 0, (flags & NEG0) != 0
    s0, s1 = np.where(n0, -1.0, 1.0), np.where(n1, -1.0, 1.0)
    # reference: local = (s0 * g[swap], s1 * g[!swap]);  scatter is the transpose
    g = x[idx]
    xe_ref = np.stack([s0 * np.where(sw, g[..., 1], g[..., 0]), s1 * np.where(sw, g[..., 0], g[..., 1])], axis=-1)
    contrib = np.stack([np.where(sw, s1 * ye[..., 1], s0 * ye[..., 0]), np.where(sw, s0 * ye[..., 0], s1 * ye[..., 1])], axis=-1)
    y_ref = np.zeros((n_units, 2))
    np.add.at(y_ref, idx.ravel(), contrib.reshape(-1, 2))
    y = np.full((n_units, 2), 1e30)
    sl = np.arange(npair)
    for k in range(len(off) - 1):
        pd_all, groups, ov, deps, lidx, n_real = decode_patch(k, off, blob, npair, fe)
        el = perm[k * fe:(k + 1) * fe]
        Xs = x[pd_all]
        pos = lidx & 0x1FFF
        lsw, ln1, ln0 = (lidx & 0x2000) != 0, (lidx & 0x4000) != 0, (lidx & 0x8000) != 0
        Ey = np.full((npair * st, 2), np.nan)
        for j in range(n_real):
            gj = Xs[pos[:, j]]
            a0 = np.where(ln0[:, j], -1.0, 1.0) * np.where(lsw[:, j], gj[:, 1], gj[:, 0])
            a1 = np.where(ln1[:, j], -1.0, 1.0) * np.where(lsw[:, j], gj[:, 0], gj[:, 1])
            assert np.array_equal(np.stack([a0, a1], -1), xe_ref[el[j]])
            b0 = np.where(ln0[:, j], -1.0, 1.0) * ye[el[j], :, 0]
            b1 = np.where(ln1[:, j], -1.0, 1.0) * ye[el[j], :, 1]
            Ey[e_index(sl, j, st)] = np.where(lsw[:, j, None], np.stack([b1, b0], -1), np.stack([b0, b1], -1))
        for u0, pd, own, ent in groups:
            used = ent != 0xFFFF
            assert np.all(ent[used] % 16 == 0)
            if own is not None:
                y[pd[own]] = 0.0
            for u in range(len(pd)):
                v = Ey[ent[u][used[u]] // 16].sum(axis=0)
                y[pd[u]] = v if own is None else y[pd[u]] + v
        for w in ov:
            y[pd_all[w >> 16]] += Ey[(w & 0xFFFF) // 16]
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-13)


def test_exchange_patches_come_first():
    """Patches touching a flagged (exchange) unit are placed first; the decomposition stays a valid one."""
    m, sp, code, cen, idx, sgn = make("hex", (16, 4, 4), 2, False)
    ne, nd = code.shape
    flags = np.zeros(sp.ndof, np.uint8)
    on_plane = np.nonzero(np.isclose(m.verts[m.elems].max(axis=1)[:, 0], 1.0))[0]      # elements touching the face x = 1
    flags[np.unique(idx[on_plane])] = 1
    perm, off, blob, st = build(sp, 32, code, cen, exchange=flags)
    nb = SWIZZLE["n_boundary"]
    npatch = len(off) - 1
    assert 0 < nb < npatch
    real = perm[perm >= 0]
    assert np.array_equal(np.sort(real), np.arange(ne))
    for k in range(npatch):
        el = perm[k * 32:(k + 1) * 32]
        touches = bool(flags[idx[el[el >= 0]]].any())
        assert touches == (k < nb), (k, nb, touches)
