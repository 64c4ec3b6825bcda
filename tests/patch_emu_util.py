"""Shared helpers of the host-emulation tests of the patch pipeline (tests/test_patch_p1_emu.py, test_patch_p2_emu.py):
build the emulation library (tests/emu/patch_emu.cpp = hephaestus_b200/csrc/patch_pipeline.h on host threads), build
patch blobs through the library's host API, decode their headers, run the emulated kernel."""
import ctypes as C
import os
import subprocess

import numpy as np

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle import mesh as om

from helpers import star_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = 32
ESS0, ESS1, SHARED, IDX = 0x20000000, 0x40000000, 0x80000000, 0x0FFFFFFF


def build_emu_lib(tmp_path_factory):
    path = os.environ.get("HCURL_EMU_LIB")
    if not path:
        path = str(tmp_path_factory.mktemp("emu") / "libpatchemu.so")
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(ROOT, "hephaestus_b200", "csrc"),
               "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "patch_emu.cpp"), "-o", path]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.emu_patch_apply.restype = C.c_int
    lib.emu_patch_apply.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P, P, C.c_double,
                                    C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, P, P, C.c_int, P, P, P, P, C.c_uint32, C.c_int, P]
    return lib


def affine_pair(kind, n, seed=4):
    """(product mesh, oracle mesh): every element a parallelepiped, element numbering scrambled"""
    if kind == "star":
        m, mo = star_mesh(*n), star_mesh(*n, cls=om.Mesh)
    else:
        m, mo = fespace.cartesian_hex_mesh(*n, 1.0, 0.7, 1.3), om.cartesian_hex(*n, 1.0, 0.7, 1.3)
    rng = np.random.default_rng(seed)
    A = np.eye(3) + 0.25 * rng.standard_normal((3, 3))
    if np.linalg.det(A) < 0:
        A[:, 0] *= -1
    perm = rng.permutation(m.n_elem)
    out = []
    for mesh, cls in ((m, fespace.Mesh), (mo, om.Mesh)):
        out.append(cls("hex", mesh.verts @ A.T, mesh.elems[perm], np.arange(m.n_elem) % 3 + 1))
    return out


def geometry_factors(mesh, perm, fe, alpha, beta):
    """[n_patch][12][FE]: alpha J^T J / det and beta adj adj^T / det, packed 00 01 02 11 12 22 (patch_edata_kernel)"""
    X = mesh.verts[mesh.elems]                                   # MFEM corner order
    J = np.stack([X[:, 1] - X[:, 0], X[:, 3] - X[:, 0], X[:, 4] - X[:, 0]], axis=2)   # columns = edge vectors
    det = np.linalg.det(J)
    GK = alpha[:, None, None] * np.einsum("eki,ekj->eij", J, J) / det[:, None, None]
    Ji = np.linalg.inv(J)
    GM = beta[:, None, None] * det[:, None, None] * np.einsum("eik,ejk->eij", Ji, Ji)
    iu = ([0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2])
    per_elem = np.concatenate([GK[:, iu[0], iu[1]], GM[:, iu[0], iu[1]]], axis=1)      # (ne, 12)
    npatch = len(perm) // fe
    out = np.zeros((npatch, 12, fe))
    pe = perm.reshape(npatch, fe)
    for k in range(npatch):
        real = pe[k] >= 0
        out[k][:, real] = per_elem[pe[k][real]].T
    return out


def build_patches(sp, mesh, order, fe):
    """patch blobs of an order-1 (8-byte units) or order-2 (16-byte pair units) hex space through the host API"""
    lib = L.lib()
    sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
    ne = sd.shape[0]
    if order == 2:
        assert sd.shape[1] == 54 and sp.ndof % 2 == 0
        nu, unit, n_units = 27, 16, sp.ndof // 2
        code = np.zeros((ne, nu), np.uint32)
        L.check(lib.hcurl_host_hex_p2_pair_codes(ne, L.ptr(sd), L.ptr(code)), "pair codes")
    else:
        assert sd.shape[1] == 12
        nu, unit, n_units = 12, 8, sp.ndof
        code = np.zeros((ne, nu), np.uint32)
        L.check(lib.hcurl_host_hex_slot_codes(1, ne, L.ptr(sd), L.ptr(code)), "slot codes")
    cen = np.ascontiguousarray(mesh.verts[mesh.elems].mean(axis=1))
    h = L.check_ptr(lib.hcurl_host_patches_create(nu, ne, fe, fe + 1, unit, n_units, L.ptr(cen), L.ptr(code), None), "patches")
    npatch = lib.hcurl_host_patches_count(h)
    perm, off = np.zeros(npatch * fe, np.int32), np.zeros(npatch + 1, np.uint32)
    blob = np.zeros(lib.hcurl_host_patches_blob_words(h), np.uint32)
    L.check(lib.hcurl_host_patches_copy(h, L.ptr(perm), L.ptr(off), L.ptr(blob)), "copy")
    nb = 128 // unit
    row, grp = np.zeros(27, np.uint8), np.zeros(16, np.uint8)
    L.check(lib.hcurl_host_patches_swizzle(h, L.ptr(row), L.ptr(grp)), "swizzle")
    part = np.zeros(2 * npatch, np.uint32)
    md, ml, mt = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    L.check(lib.hcurl_host_patches_parts(h, L.ptr(part), C.byref(md), C.byref(ml), C.byref(mt)), "parts")
    own_off = np.zeros(npatch + 1, np.uint32)
    n_own = lib.hcurl_host_patches_owned(h, L.ptr(own_off), None)
    own_idx = np.zeros(max(n_own, 1), np.uint32)
    lib.hcurl_host_patches_owned(h, None, L.ptr(own_idx))
    caps = dict(d=md.value * 4, l=ml.value * 4, t=mt.value * 4, xs=(int(lib.hcurl_host_patches_max_dofs(h)) * unit + 15) & ~15)
    lib.hcurl_host_patches_destroy(h)
    T = np.zeros(7)
    L.check(lib.hcurl_host_modal_tables_p2(L.ptr(T)), "tables")
    general = bool(order == 2 and not np.all(np.isin(code >> 29, (0, 7))))
    return dict(order=order, FE=fe, perm=perm, off=off, part=part, blob=blob, row=row, grp=grp, n_grp=(fe + nb - 1) // nb, caps=caps, T=T,
                general=general, own_off=own_off, own_idx=own_idx, count=np.zeros(npatch // 8 + 4, np.uint32), arrivals=0, unit=unit)


def header(P, k, blob=None):
    blob = P["blob"] if blob is None else blob
    o = int(P["off"][k]) * 4
    return blob[o:o + HDR]


def unit_words(P, k, blob=None):
    """(view of the unit words of patch k, number of private units)"""
    blob = P["blob"] if blob is None else blob
    o = int(P["off"][k]) * 4
    B = blob[o:o + HDR]
    return blob[o + HDR:o + HDR + int(B[0])], int(B[4] + B[5] + B[6])


def flag_essential(P, flag):
    """copy of the blobs with the essential flags on the unit words (what patch_flag_kernel / patch_flag8_kernel do)"""
    b = P["blob"].copy()
    for k in range(len(P["off"]) - 1):
        w, _ = unit_words(P, k, b)
        q = (w & IDX).astype(np.int64)
        if P["order"] == 2:
            w[flag[2 * q]] |= ESS0
            w[flag[2 * q + 1]] |= ESS1
        else:
            w[flag[q]] |= ESS0
        if np.any(w & (ESS0 | ESS1)):
            b[int(P["off"][k]) * 4 + 20] = 1      # header word 20: the patch has essential dofs
    return b


def run_emu(emu, P, edata, x, y, scale=1.0, accumulate=False, grid=2, dot=False, blob=None, k_begin=0, k_end=None, general=None,
            ess_identity=True, stream=False):
    """stream = True: the warp-specialised program (patch_stream.h); `grid` then counts streams of 32 + 64 threads"""
    d = np.zeros(1)
    blob = P["blob"] if blob is None else blob
    npatch = len(P["off"]) - 1 if k_end is None else k_end
    c = P["caps"]
    g = P["general"] if general is None else general
    if not accumulate:   # every arrival counter receives one arrival per CTA of the launch (what the launcher accounts for)
        P["arrivals"] += grid if stream else min(grid, npatch - k_begin)   # (idle streams arrive too)
    rc = emu.emu_patch_apply(P["order"], -P["FE"] if stream else P["FE"], int(g), int(accumulate), npatch, k_begin, grid, L.ptr(P["off"]), L.ptr(P["part"]),
                             L.ptr(blob), L.ptr(edata), L.ptr(x), L.ptr(y), scale, c["d"], c["l"], c["t"], c["xs"], L.ptr(P["row"]),
                             L.ptr(P["grp"]), P["n_grp"], L.ptr(P["T"]), L.ptr(P["own_off"]), L.ptr(P["own_idx"]), L.ptr(P["count"]), P["arrivals"], int(ess_identity),
                             L.ptr(d) if dot else None)
    assert rc == 0, "emulation protocol error flags 0x%x" % rc
    return d[0]
