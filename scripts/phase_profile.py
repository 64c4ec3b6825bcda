"""Per-phase cycle profile of the order-2 patch kernel (thread 0 of every CTA, clock64):
where a CTA's lifetime goes -- blob wait, gather, transforms, operator, scatter.
    python scripts/phase_profile.py [n]     (GPU; HCURL_PATCH_FE selects the patch size)"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hephaestus_b200 import _lib as L, fespace, operators  # noqa: E402

lib = L.lib()
lib.hcurl_debug_patch_profile.restype = C.c_int
lib.hcurl_debug_patch_profile.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = operators.Context(0)
mesh = fespace.cartesian_hex_mesh(n, n, n)
space = fespace.NDSpace(mesh, 2)
gsp = operators.GpuNDSpace(ctx, space)
op = operators.HephGpuOperator(ctx, gsp, 3)
op.SetCoefficients(np.ones(mesh.n_elem), np.ones(mesh.n_elem))
op.SetEssentialTrueDofs(space.boundary_dofs())
dx = ctx.alloc(space.ndof).upload(np.random.default_rng(0).standard_normal(space.ndof))
dy = ctx.alloc(space.ndof)
for _ in range(3):
    op.Mult(dx, dy)
ctx.synchronize()
op.kernel_stats_reset()
for _ in range(10):
    op.Mult(dx, dy)
ms, cnt = op.kernel_stats_ms()
print("FE=%s n=%d: kernel %.4f ms (no profiling)" % (os.environ.get("HCURL_PATCH_FE", "32"), n, ms))
lib.hcurl_debug_patch_profile(ctx.h, 1, None)
reps = 5
op.kernel_stats_reset()
for _ in range(reps):
    op.Mult(dx, dy)
ms2, _ = op.kernel_stats_ms()
out = (C.c_ulonglong * 8)()
lib.hcurl_debug_patch_profile(ctx.h, 0, out)
tot = float(sum(out[:6]))
names = ["prologue (per CTA)", "-", "pull + fwd", "op + bwd + next gather + Y store", "-", "dof sums + stores + gather wait"]
print("  with profiling %.4f ms; cycles of thread 0 per patch (share):" % ms2)
npatch = (mesh.n_elem + int(os.environ.get("HCURL_PATCH_FE", "32")) - 1) // int(os.environ.get("HCURL_PATCH_FE", "32"))
for i, nm in enumerate(names):
    print("  %-24s %9.0f  %5.1f%%" % (nm, out[i] / reps / npatch, 100.0 * out[i] / tot))
