"""Order-2 patch kernel (hephaestus_b200/csrc/patch_pipeline.h + hex_patch_p2_elem.h: one thread per element, pair
units), checked WITHOUT a GPU on the host environment of tests/emu/ against the oracle's assembled operator -- the same
scheme as tests/test_patch_p1_emu.py.  Variants: 32- / 64-element patches, reversed-line and general orientation
handling."""
import numpy as np
import pytest

from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import space as osp

from helpers import coefficients
from patch_emu_util import affine_pair, build_emu_lib, build_patches, flag_essential, geometry_factors, header, run_emu


@pytest.fixture(scope="session")
def emu(tmp_path_factory, lib_built):
    return build_emu_lib(tmp_path_factory)


VARIANTS = [(32, False), (64, False), (32, True), (-32, False), (-32, True), (-64, False), (-64, True), (-128, False)]   # (elements per patch, general orientation handling); negative: the warp-specialised stream program (patch_stream.h)


# (6,6,5) = 180 elements -> 6 patches of 32: with grid 1 / 2 the pipeline wraps around; scrambled vertex numbering makes
# element slot pairs reversed (MFEM-convention maps flag a slot pair with all three orientation bits or none)
@pytest.mark.parametrize("fe,general", VARIANTS)
@pytest.mark.parametrize("n,grid,scramble", [((6, 6, 5), 1, True), ((6, 6, 5), 2, False), ((4, 4, 4), 3, True), ((3, 2, 2), 1, True), ((10, 8, 7), 1, True), ((10, 8, 7), 2, False)])
def test_emulated_p2_kernel_matches_oracle(emu, fe, general, n, grid, scramble):
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
    stream, fe = fe < 0, abs(fe)
    if stream:
        grid = {1: 1, 2: 3, 3: 8}[grid]   # streams; 8 > the 2 patches of the (4, 4, 4) case: idle streams
    P = build_patches(sp, m, 2, fe)
    assert not P["general"]
    edata = geometry_factors(m, P["perm"], fe, alpha, beta)
    rng = np.random.default_rng(8)
    x = rng.standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)   # nothing is cleared before an apply
    dot = run_emu(emu, P, edata, x, y, grid=grid, dot=True, general=general, stream=stream)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    y2 = rng.standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu, P, edata, x, y2, scale=-0.37, accumulate=True, grid=grid, general=general, stream=stream)
    assert np.abs(y2 - (y2_0 - 0.37 * ref)).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("fe,general", VARIANTS)
def test_emulated_p2_kernel_overflow_slots(emu, fe, general):
    """star meshes (six rhombic prisms around an edge, stacked): pairs with more than four element slots in one patch go
    through the overflow list and are RED-added after the patch's own store"""
    m, mo = affine_pair("star", (6,))
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    if sp.ndof % 2:
        pytest.skip("odd number of dofs: no pair structure")
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    stream, fe = fe < 0, abs(fe)
    P = build_patches(sp, m, 2, fe)
    assert sum(int(header(P, k)[2]) for k in range(len(P["off"]) - 1)) > 0
    edata = geometry_factors(m, P["perm"], fe, alpha, beta)
    x = np.random.default_rng(4).standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)
    dot = run_emu(emu, P, edata, x, y, grid=2, dot=True, general=general or P["general"], stream=stream)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)


@pytest.mark.parametrize("fe,general", VARIANTS)
def test_emulated_p2_kernel_essential_dofs(emu, fe, general):
    m, mo = affine_pair("hex", (5, 4, 3), seed=9)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    alpha, beta = coefficients(mo, seed=5)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    # an arbitrary essential set: single dofs of a pair, both, none
    ess = np.sort(np.random.default_rng(11).choice(sp.ndof, sp.ndof // 4, replace=False))
    flag = np.zeros(sp.ndof, bool)
    flag[ess] = True
    stream, fe = fe < 0, abs(fe)
    P = build_patches(sp, m, 2, fe)
    blob_ess = flag_essential(P, flag)
    edata = geometry_factors(m, P["perm"], fe, alpha, beta)
    x = np.random.default_rng(2).standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)
    dot = run_emu(emu, P, edata, x, y, grid=2, dot=True, blob=blob_ess, general=general, stream=stream)
    ref = oa.constrained_apply(A, ess, x)
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(y[ess], x[ess])
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    # accumulating constrained apply: free dofs accumulate, essential rows are still the identity
    y2 = np.random.default_rng(3).standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu, P, edata, x, y2, scale=2.0, accumulate=True, grid=2, blob=blob_ess, general=general, stream=stream)
    xz = x.copy()
    xz[ess] = 0.0
    exp = y2_0 + 2.0 * (A @ xz)
    exp[ess] = x[ess]
    assert np.abs(y2 - exp).max() <= 1e-12 * np.abs(ref).max()
    # several ranks: essential rows are left to the caller (zeros)
    y = np.full(sp.ndof, 1e30)
    run_emu(emu, P, edata, x, y, grid=2, blob=blob_ess, general=general, ess_identity=False, stream=stream)
    ref0 = ref.copy()
    ref0[ess] = 0.0
    assert np.abs(y - ref0).max() <= 1e-12 * np.abs(ref).max() and np.all(y[ess] == 0.0)


def test_general_orientation_flags(emu):
    """random swap / sign flags on a consistent map (flip the flags of an element slot pair together with the local
    input / output it stands for): the general kernel is exercised on combinations MFEM's maps never produce"""
    m, mo = affine_pair("hex", (4, 3, 3))
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    P = build_patches(sp, m, 2, 32)
    # sign-flip one global dof of a pair everywhere: x' = S x, y' = S y with the flags NEG toggled for that dof in every
    # slot referencing it -- equivalent to A' = S A S; emulate by toggling table bits and comparing with S A S x
    rng = np.random.default_rng(1)
    flipped = rng.random(sp.ndof) < 0.3
    S = np.where(flipped, -1.0, 1.0)
    blob = P["blob"].copy()
    for k in range(len(P["off"]) - 1):
        o, t0 = int(P["off"][k]) * 4, int(P["part"][2 * k + 1]) * 4
        B = blob[o:o + 32]
        U = int(B[0])
        q = (blob[o + 32:o + 32 + U] & 0x0FFFFFFF).astype(np.int64)
        lidx = blob[t0:t0 + (27 * 32 + 1) // 2].view(np.uint16)
        nreal = int(B[1])
        tab = lidx[:27 * 32].reshape(27, 32)
        for j in range(nreal):
            for r in range(27):
                l = int(tab[r, j])
                pos, sw = l & 0x1FFF, bool(l & 0x2000)
                g0, g1 = 2 * q[pos], 2 * q[pos] + 1          # local slot 0 reads g[swap], slot 1 reads g[!swap]
                first, second = (g1, g0) if sw else (g0, g1)
                if flipped[first]:
                    l ^= 0x8000
                if flipped[second]:
                    l ^= 0x4000
                tab[r, j] = l
    edata = geometry_factors(m, P["perm"], 32, alpha, beta)
    x = rng.standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)
    run_emu(emu, P, edata, x, y, grid=2, blob=blob, general=True)
    ref = S * (A @ (S * x))
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("grid", [1, 2])
def test_emulated_p2_stream128_zeroing_batches(emu, grid):
    """the default form of the stream program (128-element patches, 4 element + 8 sum + 4 gather warps) over MANY pipeline
    iterations per stream: 2304 elements = 18 patches, i.e. 18 / 9 iterations on 1 / 2 streams -- the batch arrivals of the
    zeroing ahead (8 iterations per batch), the watermark waits of the gather group, the double-buffered slot lists and the
    four-deep ring of unit lists all wrap around; reversed lines from a scrambled vertex numbering"""
    m, mo = affine_pair("hex", (16, 12, 12))
    perm = np.random.default_rng(11).permutation(len(m.verts))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    m = fespace.Mesh("hex", m.verts[perm], inv[m.elems], m.attr)
    mo = type(mo)("hex", mo.verts[perm], inv[mo.elems], mo.attr)
    sp, so = fespace.NDSpace(m, 2), osp.NDSpace(mo, 2)
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=3).tocsr()
    P = build_patches(sp, m, 2, 128)
    assert len(P["off"]) - 1 == 18
    edata = geometry_factors(m, P["perm"], 128, alpha, beta)
    x = np.random.default_rng(12).standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)   # nothing is cleared before an apply
    dot = run_emu(emu, P, edata, x, y, grid=grid, dot=True, general=False, stream=True)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    y2 = np.random.default_rng(13).standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu, P, edata, x, y2, scale=0.61, accumulate=True, grid=grid, general=False, stream=True)
    assert np.abs(y2 - (y2_0 + 0.61 * ref)).max() <= 1e-12 * np.abs(ref).max()
