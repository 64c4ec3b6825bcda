"""Order-1 patch kernel, checked WITHOUT a GPU: the pipeline program of the CUDA kernel
(hephaestus_b200/csrc/patch_pipeline.h + hex_patch_p1_elem.h) is compiled for the host with the emulation environment
of tests/emu/ (one host thread per CUDA thread, all CTAs concurrent, std::barrier = __syncthreads, asynchronous copies
poisoned at issue and completed at their wait, release / acquire atomics for the per-patch flags) and compared with the
oracle's assembled operator on the same seeded inputs.

What this pins: the blob consumption (three parts, four mbarriers), the pull through the per-slot table with sign flips,
the generated modal element operator, the result-buffer swizzle, the per-unit sums -- owner stores, the other patches'
REDs after the owner's flag, overflow slots -- with NOTHING cleared beforehand, the identity on essential dofs,
accumulation, the fused dot product, the software pipeline and -- under scripts/tsan_emu_check.sh -- the barrier and
flag placement.  What it cannot pin is the device environment (hex_patch_kernels.cu: PTX of the bulk copy, cp.async,
red, st.release / ld.acquire): GPU parity tests in test_gpu_apply.py."""
import os

import numpy as np
import pytest

from hephaestus_b200 import fespace
from oracle import assemble as oa
from oracle import space as osp

from helpers import coefficients
from patch_emu_util import ROOT, affine_pair, build_emu_lib, build_patches, flag_essential, geometry_factors, header, run_emu


@pytest.fixture(scope="session")
def emu(tmp_path_factory, lib_built):
    return build_emu_lib(tmp_path_factory)


# (the 6x6x6 / 8x8x5 cases give one CTA 7 / 4 / 5 patches: the double-buffered pipeline wraps around; 10x8x8 with 32
# elements per patch gives one CTA 20 / 10 / 7 iterations: three / two batches of the zeroing watermark)
CASES = [("hex", (10, 8, 8), 32, 1), ("hex", (10, 8, 8), 32, 2), ("hex", (10, 8, 8), 32, 3), ("hex", (6, 6, 6), 32, 1), ("hex", (6, 6, 6), 32, 2), ("hex", (8, 8, 5), 64, 1), ("hex", (4, 4, 4), 32, 3), ("hex", (5, 3, 7), 32, 2), ("hex", (6, 5, 4), 64, 1), ("hex", (8, 4, 4), 128, 1), ("hex", (3, 3, 3), 128, 4),
         ("star", (4,), 32, 1), ("star", (7,), 32, 2)]


@pytest.mark.parametrize("kind,n,fe,grid", CASES)
def test_emulated_kernel_matches_oracle(emu, kind, n, fe, grid):
    m, mo = affine_pair(kind, n)
    sp, so = fespace.NDSpace(m, 1), osp.NDSpace(mo, 1)
    assert np.array_equal(sp.signed_elem_dofs(), np.where(so.elem_sign > 0, so.elem_dofs, -1 - so.elem_dofs))   # same dof map on both sides
    alpha, beta = coefficients(mo)
    A = oa.assemble(so, alpha, beta, q_override=2).tocsr()
    P = build_patches(sp, m, 1, fe)
    npatch = len(P["off"]) - 1
    n_ovf = sum(int(header(P, k)[2]) for k in range(npatch))
    assert (n_ovf > 0) == (kind == "star")                # the star meshes exercise the overflow slots (edge valence 6)
    edata = geometry_factors(m, P["perm"], fe, alpha, beta)
    rng = np.random.default_rng(8)
    x = rng.standard_normal(sp.ndof)
    # --- plain apply: y holds garbage everywhere -- every dof is stored by its owner patch before anyone adds to it
    y = np.full(sp.ndof, 1e30)
    dot = run_emu(emu, P, edata, x, y, scale=1.0, grid=grid, dot=True)
    ref = A @ x
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    # --- scaled, accumulating apply
    y2 = rng.standard_normal(sp.ndof)
    y2_0 = y2.copy()
    run_emu(emu, P, edata, x, y2, scale=-0.37, accumulate=True, grid=grid)
    assert np.abs(y2 - (y2_0 - 0.37 * ref)).max() <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("kind,n,fe,grid", [("hex", (4, 4, 4), 32, 2), ("hex", (5, 4, 3), 64, 1), ("star", (5,), 32, 2)])
def test_emulated_kernel_essential_dofs(emu, kind, n, fe, grid):
    """essential flags on the unit words (patch_flag8_kernel): flagged dofs enter as zeros and get y = x (the identity
    rows of FormLinearSystem), x_ess^2 enters the fused dot; with ess_identity = 0 (several ranks) they get 0"""
    m, mo = affine_pair(kind, n, seed=9)
    sp, so = fespace.NDSpace(m, 1), osp.NDSpace(mo, 1)
    alpha, beta = coefficients(mo, seed=5)
    A = oa.assemble(so, alpha, beta, q_override=2).tocsr()
    ess = np.asarray(sp.boundary_dofs(), dtype=np.int64)
    assert 0 < len(ess) < sp.ndof
    P = build_patches(sp, m, 1, fe)
    flag = np.zeros(sp.ndof, bool)
    flag[ess] = True
    blob_ess = flag_essential(P, flag)
    edata = geometry_factors(m, P["perm"], fe, alpha, beta)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(sp.ndof)
    y = np.full(sp.ndof, 1e30)
    dot = run_emu(emu, P, edata, x, y, grid=grid, dot=True, blob=blob_ess)
    ref = oa.constrained_apply(A, ess, x)
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.array_equal(y[ess], x[ess])
    assert abs(dot - x @ ref) <= 1e-12 * abs(x @ ref)
    y = np.full(sp.ndof, 1e30)
    run_emu(emu, P, edata, x, y, grid=grid, blob=blob_ess, ess_identity=False)
    ref0 = ref.copy()
    ref0[ess] = 0.0
    assert np.abs(y - ref0).max() <= 1e-12 * np.abs(ref).max() and np.all(y[ess] == 0.0)


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
