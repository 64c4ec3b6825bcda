"""Reads the per-phase cycle counters of a -DHCURL_PHASE_PROFILE build (scripts/phase_profile.sh) after a few applies
of the bench workload.  Diagnostic only: the marks perturb the kernel (clock reads + shared-memory updates by thread 0)."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hephaestus_b200 import _lib as L, fespace, operators  # noqa: E402

NAMES = ["top: loads of extents / zero indices", "geometry issue + wait table", "pull + forward transforms", "sync 1",
         "wait D(i+1)", "gather issue", "arrive + zero stores", "element operator", "backward transforms + store to EY",
         "sync 2 + request T + wait lists", "sums: private units (stores)", "watermark wait", "sums: shared units (REDs)",
         "gather wait", "sync 3 + requests"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=128)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--applies", type=int, default=5)
    a = ap.parse_args()
    lib = L.lib()
    f = C.CDLL(L.LIB_PATH).hcurl_debug_phase_cycles
    f.restype, f.argtypes = C.c_int, [C.c_void_p, C.c_int]
    ctx = operators.Context(0)
    mesh = fespace.cartesian_hex_mesh(a.n, a.n, a.n)
    sp = fespace.NativeNDSpace(mesh, a.order)
    gsp = operators.GpuNDSpace(ctx, sp)
    op = operators.HephGpuOperator(ctx, gsp, a.order + 1)
    ne = mesh.elems.shape[0]
    op.SetCoefficients(np.full(ne, 0.05), np.ones(ne))
    op.SetEssentialTrueDofs(sp.boundary_dofs())
    x = np.random.default_rng(1).standard_normal(sp.ndof)
    dx, dy = ctx.alloc(sp.ndof).upload(x), ctx.alloc(sp.ndof)
    for _ in range(3):
        op.Mult(dx, dy)
    ctx.synchronize()
    assert f(None, 1) == 0
    ctx.mark(0)
    for _ in range(a.applies):
        op.Mult(dx, dy)
    ctx.mark(1)
    ctx.synchronize()
    ms = ctx.elapsed_ms() / a.applies
    out = (C.c_ulonglong * 32)()
    assert f(out, 0) == 0
    v = np.array(list(out), dtype=np.float64)
    if os.environ.get("HCURL_PATCH_STREAM") == "1":
        names = {0: "E: end of iteration -> top", 1: "E: geometry issue + wait T", 2: "E: wait XS_FULL", 3: "E: pull + fwd + signal", 4: "E: operator",
                 5: "E: extents + wait EY_FREE", 6: "E: bwd + store + warp sync", 7: "E: signal + request T",
                 8: "S: end of iteration -> top", 14: "S: extents of the next requests", 15: "S: wait D, L", 16: "S: wait EY_FULL", 17: "S: sums private",
                 20: "S: group sync + arrive + wait EY_FREE + requests",
                 9: "G: top (zero indices)", 21: "G: gather wait", 22: "G: group sync + signal XS_FULL + arrive", 10: "G: wait D(i+1)",
                 11: "G: wait XS_FREE", 12: "G: gather issue", 13: "S: zero stores", 18: "G: wait L + watermark", 23: "G: wait EY_FULL",
                 19: "G: sums shared"}
        groups = [("element warps", [0, 1, 2, 3, 4, 5, 6, 7]), ("sum group", [8, 14, 15, 13, 16, 17, 20]), ("gather group", [9, 21, 22, 10, 11, 12, 18, 23, 19])]
        fe = int(os.environ.get("HCURL_PATCH_FE", "128"))
        its = a.n ** 3 // fe * a.applies
        print("apply %.4f ms (instrumented); cycles per pipeline iteration (%d-element patch) and stream" % (ms, fe))
        for gname, ids in groups:
            tot = sum(v[i] for i in ids)
            print(" %s: %.0f cycles" % (gname, tot / its))
            for i in ids:
                print("  %2d %-42s %8.0f cycles  %5.1f %%" % (i, names[i], v[i] / its, 100 * v[i] / tot))
        return
    iters = v[15]
    tot = v[:15].sum()
    print("apply %.4f ms (instrumented), %d pipeline iterations per apply, %.0f cycles per iteration" % (ms, iters / a.applies, tot / iters))
    for i in range(15):
        print("  %2d %-42s %8.0f cycles  %5.1f %%" % (i, NAMES[i], v[i] / iters, 100 * v[i] / tot))


if __name__ == "__main__":
    main()
