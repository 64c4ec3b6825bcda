"""Order-2 patch kernel as an Env program (hephaestus_b200/csrc/hex_patch_p2_body.h: plain and split-blob variants),
checked WITHOUT a GPU on the host environment of tests/emu/host_env.h against the oracle's assembled operator -- the
same scheme as tests/test_patch_p1_emu.py.  Besides pinning this restatement of the headline kernel, the split-blob
variant (three bulk copies per patch with different lifetimes, four mbarriers) gets its first execution here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import space as osp

from helpers import coefficients
from test_patch_p1_emu import affine_pair, geometry_factors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ESS0, ESS1, SHARED, IDX = 0x20000000, 0x40000000, 0x80000000, 0x1FFFFFFF
FE = 32


@pytest.fixture(scope="session")
def emu2(tmp_path_factory, lib_built):
    path = os.environ.get("HCURL_EMU2_LIB")
    if not path:
        path = str(tmp_path_factory.mktemp("emu2") / "libp2emu.so")
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(ROOT, "hephaestus_b200", "csrc"),
               "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "patch_p2_emu.cpp"),
               "-o", path]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.emu_patch_p2_apply.restype = C.c_int
    lib.emu_patch_p2_apply.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P, P, C.c_double,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P, P, C.c_int, P, P]
    return lib


VARIANT = {"half": 2, "single": 3, "single64": 4, "single_general": 5}


def build_patches(sp, mesh, split):
    """split: False = plain blobs, True = split-blob layout, "half" / "single" / "single64" = plain blobs for the
    two-threads-per-component program / the one-thread-per-element program with 32- / 64-element patches"""
    lib = L.lib()
    variant, split = VARIANT.get(split, 0), split is True
    FE = 64 if variant == 4 else 32
    sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
    ne = sd.shape[0]
    assert sd.shape[1] == 54 and sp.ndof % 2 == 0
    code = np.zeros((ne, 27), np.uint32)
    L.check(lib.hcurl_host_hex_p2_pair_codes(ne, L.ptr(sd), L.ptr(code)), "pair codes")
    if variant in (3, 4):  # the reversed-line kernels: MFEM-convention maps flag a slot pair with all three bits or none
        assert np.all(np.isin(code >> 29, (0, 7)))
    cen = np.ascontiguousarray(mesh.verts[mesh.elems].mean(axis=1))
    h = L.check_ptr(lib.hcurl_host_patches_create(27, ne, FE, FE if split else FE + 1, 16, sp.ndof // 2, L.ptr(cen), L.ptr(code), None), "patches")
    npatch = lib.hcurl_host_patches_count(h)
    perm, off = np.zeros(npatch * FE, np.int32), np.zeros(npatch + 1, np.uint32)
    blob = np.zeros(lib.hcurl_host_patches_blob_words(h), np.uint32)
    L.check(lib.hcurl_host_patches_copy(h, L.ptr(perm), L.ptr(off), L.ptr(blob)), "copy")
    row, grp = np.zeros(27, np.uint8), np.zeros((FE + 7) // 8, np.uint8)
    L.check(lib.hcurl_host_patches_swizzle(h, L.ptr(row), L.ptr(grp)), "swizzle")
    part = np.zeros(2 * npatch, np.uint32)
    md, ml = C.c_int64(0), C.c_int64(0)
    L.check(lib.hcurl_host_patches_parts(h, L.ptr(part), C.byref(md), C.byref(ml)), "parts")
    al = lambda w: (int(w) * 4 + 15) & ~15                                            # noqa: E731
    caps = dict(blob=al(lib.hcurl_host_patches_max_blob_words(h)), xs=int(lib.hcurl_host_patches_max_dofs(h)) * 16, d=al(md.value), l=al(ml.value))
    lib.hcurl_host_patches_destroy(h)
    T = np.zeros(7)
    L.check(lib.hcurl_host_modal_tables_p2(L.ptr(T)), "tables")
    return dict(perm=perm, off=off, blob=blob, row=row, grp=grp, part=part, caps=caps, T=T, split=variant or int(split), FE=FE)


def run_emu(emu2, P, edata, x, y, scale=1.0, accumulate=False, grid=2, dot=False, blob=None, k_begin=0, k_end=None):
    d = np.zeros(1)
    blob = P["blob"] if blob is None else blob
    npatch = len(P["off"]) - 1 if k_end is None else k_end
    c = P["caps"]
    rc = emu2.emu_patch_p2_apply(int(P["split"]), int(accumulate), npatch, k_begin, grid, L.ptr(P["off"]), L.ptr(P["part"]), L.ptr(blob),
                                 L.ptr(edata), L.ptr(x), L.ptr(y), scale, c["blob"], c["xs"], c["d"], c["l"], L.ptr(P["row"]), L.ptr(P["grp"]),
                                 len(P["grp"]), L.ptr(P["T"]), L.ptr(d) if dot else None)
    assert rc == 0, "emulation protocol error flags 0x%x" % rc
    return d[0]


def shared_mask(off, blob, ndof):
    sh = np.zeros(ndof // 2, bool)
    for k in range(len(off) - 1):
        B = blob[int(off[k]) * 4:int(off[k + 1]) * 4]
        w = B[16:16 + int(B[0])]
        sh[(w & IDX)[(w & SHARED) != 0]] = True
    return np.repeat(sh, 2)


# (6,6,5) = 180 elements -> 6 patches: with grid 1 / 2 the double-buffered pipeline wraps around; scrambled vertex
# numbering makes pairs swapped / sign-flipped (the non-plain path), the unscrambled cube exercises the plain path
@pytest.mark.parametrize("split", [False, True, "half", "single", "single64", "single_general"])
@pytest.mark.parametrize("n,grid,scramble", [((6, 6, 5), 1, True), ((6, 6, 5), 2, False), ((4, 4, 4), 3, True), ((3, 2, 2), 1, True)])
def test_emulated_p2_kernel_matches_oracle(emu2, split, n, grid, scramble):
    m, mo = affine_pair("hex", n)
    if scramble:
        perm = np.random.default_rng(5).permutation(len(m.verts))
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        m = fespace.Mesh("hex", m.verts[perm], inv[m.elems], m.attr)
        mo = type(mo)("hex", mo.verts[perm], inv[mo.elems], mo.attr)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    assert np.array_equal(sp.signed_elem_dofs(), np.where(so.elem_sign > 0, so.elem_dofs, -1 - so.elem_dofs))
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    P = build_patches(sp, m, split)
    edata = geometry_factors(m, P["perm"], P["FE"], alpha, beta)
    rng = np.random.default_rng(8)
    x = rng.standard_normal(sp.ndof)
    sh = shared_mask(P["off"], P["blob"], sp.ndof)
    y = np.where(sh, 0.0, 1e30)
    dot = run_emu(emu2, P, edata, x, y, grid=grid, dot=True)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    y2 = rng.standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu2, P, edata, x, y2, scale=-0.37, accumulate=True, grid=grid)
    assert np.abs(y2 - (y2_0 - 0.37 * ref)).max() <= 1e-12 * np.abs(ref).max()
    npatch = len(P["off"]) - 1
    if npatch >= 3:
        y3 = np.where(sh, 0.0, 1e30)
        run_emu(emu2, P, edata, x, y3, grid=grid, k_begin=npatch // 2)
        run_emu(emu2, P, edata, x, y3, grid=grid, k_begin=0, k_end=npatch // 2)
        assert np.abs(y3 - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("split", [False, True, "half", "single", "single64", "single_general"])
def test_emulated_p2_kernel_overflow_slots(emu2, split):
    """star meshes (six rhombic prisms around an edge, stacked): pairs with more than four element slots in one patch go
    through the overflow list and are RED-added"""
    m, mo = affine_pair("star", (6,))
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    if sp.ndof % 2:
        pytest.skip("odd number of dofs: no pair structure")
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    P = build_patches(sp, m, split)
    n_ovf = sum(int(P["blob"][int(o) * 4 + 7]) for o in P["off"][:-1])
    assert n_ovf > 0
    edata = geometry_factors(m, P["perm"], P["FE"], alpha, beta)
    x = np.random.default_rng(4).standard_normal(sp.ndof)
    sh = shared_mask(P["off"], P["blob"], sp.ndof)
    y = np.where(sh, 0.0, 1e30)
    dot = run_emu(emu2, P, edata, x, y, grid=2, dot=True)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)


@pytest.mark.parametrize("split", [False, True, "half", "single", "single64", "single_general"])
def test_emulated_p2_kernel_essential_dofs(emu2, split):
    m, mo = affine_pair("hex", (5, 4, 3), seed=9)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    alpha, beta = coefficients(mo, seed=5)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    # an arbitrary essential set: single dofs of a pair, both, none
    ess = np.sort(np.random.default_rng(11).choice(sp.ndof, sp.ndof // 4, replace=False))
    flag = np.zeros(sp.ndof, bool)
    flag[ess] = True
    P = build_patches(sp, m, split)
    blob_ess = P["blob"].copy()
    for k in range(len(P["off"]) - 1):
        o = int(P["off"][k]) * 4
        U = int(P["blob"][o])
        w = blob_ess[o + 16:o + 16 + U]
        q = (w & IDX).astype(np.int64)
        w[flag[2 * q]] |= ESS0
        w[flag[2 * q + 1]] |= ESS1
    edata = geometry_factors(m, P["perm"], P["FE"], alpha, beta)
    x = np.random.default_rng(2).standard_normal(sp.ndof)
    sh = shared_mask(P["off"], P["blob"], sp.ndof)
    y = np.where(sh, 0.0, 1e30)
    y[ess] = 777.0
    dot = run_emu(emu2, P, edata, x, y, grid=2, dot=True, blob=blob_ess)
    xz = x.copy()
    xz[ess] = 0.0
    ref = A @ xz
    assert np.abs(y[~flag] - ref[~flag]).max() <= 1e-12 * np.abs(ref).max()
    assert np.all(y[ess] == 777.0)
    assert abs(dot - xz @ ref) <= 1e-12 * abs(xz @ ref)
