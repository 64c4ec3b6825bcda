"""Host-side patch decomposition (hephaestus_b200/csrc/patch.cpp) -- runs without a GPU.

The patch lists must reproduce exactly the signed gather x_e = G x and scatter-add y = G^T y_e
of the element -> dof map (FiniteElementSpace::GetElementDofs semantics) for any mesh."""
import ctypes as C

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace

NEG = 0x80000000


def build(space, fe, slot_code, centroid):
    lib = L.lib()
    ne, nd = slot_code.shape
    h = L.check_ptr(lib.hcurl_host_patches_create(nd, ne, fe, space.ndof, L.ptr(centroid), L.ptr(slot_code)), "patches")
    npatch = lib.hcurl_host_patches_count(h)
    npd = lib.hcurl_host_patches_ndof_entries(h)
    perm = np.zeros(npatch * fe, np.int32)
    off = np.zeros(npatch + 1, np.uint32)
    npriv = np.zeros(npatch, np.uint32)
    dof = np.zeros(npd, np.uint32)
    start = np.zeros(npd + npatch, np.uint16)
    slots = np.zeros(npatch * fe * nd, np.uint16)
    L.check(lib.hcurl_host_patches_copy(h, L.ptr(perm), L.ptr(off), L.ptr(npriv), L.ptr(dof), L.ptr(start), L.ptr(slots)), "copy")
    lib.hcurl_host_patches_destroy(h)
    return perm, off, npriv, dof, start, slots


def make(kind, n, p, perturb):
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
                                         ("hex", (3, 3, 3), 1, 32), ("hex", (3, 2, 2), 3, 32), ("tet", (3, 3, 2), 2, 64)])
@pytest.mark.parametrize("perturb", [False, True])
def test_patch_lists_reproduce_gather_and_scatter(kind, n, p, fe, perturb):
    m, sp, code, cen, idx, sgn = make(kind, n, p, perturb)
    ne, nd = code.shape
    perm, off, npriv, dof, start, slots = build(sp, fe, code, cen)
    npatch = len(npriv)
    assert npatch == (ne + fe - 1) // fe
    # perm is a permutation of the elements (+ padding)
    real = perm[perm >= 0]
    assert np.array_equal(np.sort(real), np.arange(ne))
    assert np.all(perm[ne:] == -1)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(sp.ndof)
    ye = rng.standard_normal((ne, nd))
    y_ref = np.zeros(sp.ndof)
    np.add.at(y_ref, idx.ravel(), (sgn * ye).ravel())
    y = np.zeros(sp.ndof)
    touched = np.zeros(sp.ndof, np.int64)
    for k in range(npatch):
        o, U = int(off[k]), int(off[k + 1] - off[k])
        st = start[o + k:o + k + U + 1].astype(np.int64)
        sl = slots[k * fe * nd:(k + 1) * fe * nd]
        pd = dof[o:o + U].astype(np.int64)
        assert st[0] == 0 and np.all(np.diff(st) >= 1)
        el = perm[k * fe:(k + 1) * fe]
        n_real = int((el >= 0).sum())
        assert st[-1] == n_real * nd
        # classes sorted ascending, private first
        assert np.all(np.diff(pd[:npriv[k]]) > 0) and np.all(np.diff(pd[npriv[k]:]) > 0)
        E = np.full(nd * fe, np.nan)
        Ey = np.zeros(nd * fe)
        for j in range(n_real):
            Ey[np.arange(nd) * fe + j] = ye[el[j]]
        for u in range(U):
            ent = sl[st[u]:st[u + 1]].astype(np.int64)
            pos, neg = ent & 0x7FFF, (ent & 0x8000) != 0
            E[pos] = np.where(neg, -x[pd[u]], x[pd[u]])
            v = np.sum(np.where(neg, -Ey[pos], Ey[pos]))
            if u < npriv[k]:
                assert touched[pd[u]] == 0
                y[pd[u]] = v
            else:
                y[pd[u]] += v
            touched[pd[u]] += 1
        for j in range(n_real):
            xe = E[np.arange(nd) * fe + j]
            assert np.array_equal(xe, sgn[el[j]] * x[idx[el[j]]])
    assert np.all(touched >= 1)
    np.testing.assert_allclose(y, y_ref, rtol=0, atol=1e-13)
    # every dof listed as private really belongs to one patch only
    elem_patch = np.empty(ne, np.int64)
    elem_patch[perm[perm >= 0]] = np.nonzero(perm >= 0)[0] // fe
    pmin = np.full(sp.ndof, 1 << 40)
    pmax = np.full(sp.ndof, -1)
    np.minimum.at(pmin, idx.ravel(), np.repeat(elem_patch, nd))
    np.maximum.at(pmax, idx.ravel(), np.repeat(elem_patch, nd))
    for k in range(npatch):
        o = int(off[k])
        pd = dof[o:o + int(off[k + 1] - off[k])].astype(np.int64)
        assert np.all(pmin[pd[:npriv[k]]] == pmax[pd[:npriv[k]]])
        assert np.all(pmin[pd[npriv[k]:]] != pmax[pd[npriv[k]:]])


def test_structured_cube_patches_are_blocks():
    """On a power-of-two structured cube the Morton patches are 4x4x4 element blocks."""
    m, sp, code, cen, idx, sgn = make("hex", (8, 8, 8), 2, False)
    perm, off, npriv, dof, start, slots = build(sp, 64, code, cen)
    for k in range(len(npriv)):
        c = cen[perm[k * 64:(k + 1) * 64]]
        ext = c.max(0) - c.min(0)
        np.testing.assert_allclose(ext, 3.0 / 8, atol=1e-12)
    # 4x4x4 block of order-2 hexes: 1944 unique dofs, 1176 of them interior to the block
    # (here blocks touch the domain boundary, which only adds private dofs)
    assert np.all(np.diff(off) == 1944)
    assert npriv.min() >= 1176


def test_bad_arguments_are_reported():
    lib = L.lib()
    code = np.array([[5, 1]], dtype=np.uint32)
    cen = np.zeros((1, 3))
    assert not lib.hcurl_host_patches_create(2, 1, 32, 3, L.ptr(cen), L.ptr(code))
    assert b"out of range" in lib.hcurl_last_error()
    assert not lib.hcurl_host_patches_create(2, 1, 20000, 8, L.ptr(cen), L.ptr(code))
    assert b"too large" in lib.hcurl_last_error()
