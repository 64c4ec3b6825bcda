"""Order-1 patch kernel, checked WITHOUT a GPU: the per-thread program of the CUDA kernel
(hephaestus_b200/csrc/hex_patch_p1_body.h) is compiled for the host with the emulation environment of
tests/emu/patch_p1_emu.cpp (one host thread per CUDA thread, std::barrier = __syncthreads, asynchronous copies poisoned
at issue and completed at their wait) and compared with the oracle's assembled operator on the same seeded inputs.

What this pins: the blob consumption, the pull through the per-slot table with sign flips, the generated modal element
operator, the result-buffer swizzle, the per-dof sums with private stores / shared REDs / overflow slots, essential
flags, accumulation, the fused dot product, the software pipeline (two blob buffers, gather of the next patch, mbarrier
phases) and -- under scripts/tsan_emu_check.sh -- the barrier placement.  What it cannot pin is the device environment
(hex_patch_p1.cu: PTX of the bulk copy, cp.async, red): GPU parity tests in test_gpu_apply.py, opt-in until run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import mesh as om
from oracle import space as osp

from helpers import coefficients, star_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ESS0 = 0x20000000
IDX = 0x1FFFFFFF


@pytest.fixture(scope="session")
def emu(tmp_path_factory, lib_built):
    path = os.environ.get("HCURL_EMU_LIB")
    if not path:
        path = str(tmp_path_factory.mktemp("emu") / "libp1emu.so")
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-I" + os.path.join(ROOT, "hephaestus_b200", "csrc"),
               "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", "patch_p1_emu.cpp"), "-o", path]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.emu_patch_p1_apply.restype = C.c_int
    lib.emu_patch_p1_apply.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, P, P, P, P, C.c_double, C.c_uint32, C.c_uint32,
                                       P, P, C.c_int, P]
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


def build_patches(sp, mesh, fe):
    lib = L.lib()
    sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
    ne, nd = sd.shape
    assert nd == 12
    code = np.zeros((ne, nd), np.uint32)
    L.check(lib.hcurl_host_hex_slot_codes(1, ne, L.ptr(sd), L.ptr(code)), "slot codes")
    cen = np.ascontiguousarray(mesh.verts[mesh.elems].mean(axis=1))
    h = L.check_ptr(lib.hcurl_host_patches_create(nd, ne, fe, fe + 1, 8, sp.ndof, L.ptr(cen), L.ptr(code), None), "patches")
    npatch = lib.hcurl_host_patches_count(h)
    perm = np.zeros(npatch * fe, np.int32)
    off = np.zeros(npatch + 1, np.uint32)
    blob = np.zeros(lib.hcurl_host_patches_blob_words(h), np.uint32)
    L.check(lib.hcurl_host_patches_copy(h, L.ptr(perm), L.ptr(off), L.ptr(blob)), "copy")
    row, grp = np.zeros(nd, np.uint8), np.zeros((fe + 15) // 16, np.uint8)
    L.check(lib.hcurl_host_patches_swizzle(h, L.ptr(row), L.ptr(grp)), "swizzle")
    caps = (int(lib.hcurl_host_patches_max_blob_words(h)) * 4 + 15) & ~15, (int(lib.hcurl_host_patches_max_dofs(h)) * 8 + 15) & ~15
    lib.hcurl_host_patches_destroy(h)
    return perm, off, blob, row, grp, caps


def run_emu(emu, fe, P, edata, x, y, scale=1.0, accumulate=False, grid=3, dot=False, blob=None, k_begin=0, k_end=None):
    perm, off, blob0, row, grp, caps = P
    blob = blob0 if blob is None else blob
    d = np.zeros(1)
    npatch = len(off) - 1 if k_end is None else k_end
    rc = emu.emu_patch_p1_apply(fe, int(accumulate), npatch, k_begin, grid, L.ptr(off), L.ptr(blob), L.ptr(edata), L.ptr(x), L.ptr(y), scale,
                                caps[0], caps[1], L.ptr(row), L.ptr(grp), len(grp), L.ptr(d) if dot else None)
    assert rc == 0, "emulation protocol error flags 0x%x" % rc
    return d[0]


def shared_mask(off, blob, ndof):
    """dofs the kernel RED-adds (must be zero before a non-accumulating apply), as hex_patch_zero_shared clears them"""
    sh = np.zeros(ndof, bool)
    for k in range(len(off) - 1):
        B = blob[int(off[k]) * 4:int(off[k + 1]) * 4]
        w = B[16:16 + int(B[0])]
        sh[(w & IDX)[(w & 0x80000000) != 0]] = True
    return sh


# (the 6x6x6 / 8x8x5 cases give one CTA 7 / 4 / 5 patches: the double-buffered blob pipeline wraps around)
CASES = [("hex", (6, 6, 6), 32, 1), ("hex", (6, 6, 6), 32, 2), ("hex", (8, 8, 5), 64, 1), ("hex", (4, 4, 4), 32, 3), ("hex", (5, 3, 7), 32, 2), ("hex", (6, 5, 4), 64, 1), ("hex", (8, 4, 4), 128, 1), ("hex", (3, 3, 3), 128, 4),
         ("star", (4,), 32, 1), ("star", (7,), 32, 2)]


@pytest.mark.parametrize("kind,n,fe,grid", CASES)
def test_emulated_kernel_matches_oracle(emu, kind, n, fe, grid):
    m, mo = affine_pair(kind, n)
    sp, so = fespace.NDSpace(m, 1), osp.NDSpace(mo, 1)
    assert np.array_equal(sp.signed_elem_dofs(), np.where(so.elem_sign > 0, so.elem_dofs, -1 - so.elem_dofs))   # same dof map on both sides
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=2).tocsr()
    P = build_patches(sp, m, fe)
    perm, off, blob = P[0], P[1], P[2]
    n_ovf = sum(int(blob[int(o) * 4 + 7]) for o in off[:-1])
    assert (n_ovf > 0) == (kind == "star")                # the star meshes exercise the overflow slots (edge valence 6)
    edata = geometry_factors(m, perm, fe, alpha, beta)
    rng = np.random.default_rng(8)
    x = rng.standard_normal(sp.ndof)
    # --- plain apply: private dofs are stored (y may hold garbage there), shared ones must start from zero
    sh = shared_mask(off, blob, sp.ndof)
    y = np.where(sh, 0.0, 1e30)
    dot = run_emu(emu, fe, P, edata, x, y, scale=1.0, grid=grid, dot=True)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    # --- scaled, accumulating apply
    y2 = rng.standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu, fe, P, edata, x, y2, scale=-0.37, accumulate=True, grid=grid)
    assert np.abs(y2 - (y2_0 - 0.37 * ref)).max() <= 1e-12 * np.abs(ref).max()
    # --- ranged launches (the multi-GPU overlap splits the patch range) add up to the full apply
    if len(off) - 1 >= 3:
        y3 = np.where(sh, 0.0, 1e30)
        mid = (len(off) - 1) // 2
        run_emu(emu, fe, P, edata, x, y3, grid=grid, k_begin=mid)
        run_emu(emu, fe, P, edata, x, y3, grid=grid, k_begin=0, k_end=mid)
        assert np.abs(y3 - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("kind,n,fe,grid", [("hex", (4, 4, 4), 32, 2), ("hex", (5, 4, 3), 64, 1), ("star", (5,), 32, 2)])
def test_emulated_kernel_essential_dofs(emu, kind, n, fe, grid):
    """essential flags on the dof words (patch_flag8_kernel): flagged dofs enter as zeros and their rows are not written"""
    m, mo = affine_pair(kind, n, seed=9)
    sp, so = fespace.NDSpace(m, 1), osp.NDSpace(mo, 1)
    alpha, beta = coefficients(mo, seed=5)
    A = oa.assemble(so, alpha, beta, q_override=2).tocsr()
    ess = np.asarray(sp.boundary_dofs(), dtype=np.int64)
    assert 0 < len(ess) < sp.ndof
    P = build_patches(sp, m, fe)
    perm, off, blob = P[0], P[1], P[2]
    flag = np.zeros(sp.ndof, bool)
    flag[ess] = True
    blob_ess = blob.copy()
    for k in range(len(off) - 1):
        o = int(off[k]) * 4
        U = int(blob[o])
        w = blob_ess[o + 16:o + 16 + U]
        w[flag[w & IDX]] |= ESS0
    edata = geometry_factors(m, perm, fe, alpha, beta)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(sp.ndof)
    sh = shared_mask(off, blob, sp.ndof)
    y = np.where(sh, 0.0, 1e30)
    y[ess] = 777.0                                      # must stay untouched
    dot = run_emu(emu, fe, P, edata, x, y, grid=grid, dot=True, blob=blob_ess)
    xz = x.copy()
    xz[ess] = 0.0
    ref = A @ xz
    free = ~flag
    assert np.abs(y[free] - ref[free]).max() <= 1e-12 * np.abs(ref).max()
    assert np.all(y[ess] == 777.0)
    assert abs(dot - xz @ ref) <= 1e-12 * abs(xz @ ref)


def test_generated_header_is_current(tmp_path):
    """the committed gen/hcurl_modal_p1.h is what scripts/gen_modal_op.py emits"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_modal_op", os.path.join(ROOT, "scripts", "gen_modal_op.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    assert g.check(1) < 1e-12
    out = str(tmp_path / "p1.h")
    g.emit_header(1, out)
    assert open(out).read() == open(os.path.join(ROOT, "hephaestus_b200", "csrc", "gen", "hcurl_modal_p1.h")).read()
