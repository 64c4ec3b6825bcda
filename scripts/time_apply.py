"""Times device-resident applies of the bench workload WITHOUT any correctness check -- for timing-only diagnostic builds
(scripts/variant_build.sh with HCURL_TIMING_* switches, which compute wrong results on purpose).  usage:
HCURL_LIB_PATH=build/diag/libhcurl_X.so python scripts/time_apply.py [--n 128] [--applies 100]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from hephaestus_b200 import fespace, operators  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--applies", type=int, default=100)
a = ap.parse_args()
ctx = operators.Context(0)
mesh = fespace.cartesian_hex_mesh(a.n, a.n, a.n)
sp = fespace.NativeNDSpace(mesh, 2)
op = operators.HephGpuOperator(ctx, operators.GpuNDSpace(ctx, sp), 3)
ne = mesh.elems.shape[0]
op.SetCoefficients(np.where(mesh.attr == 1, 0.05, 0.1), np.where(mesh.attr == 1, 1.0, 1e3))
op.SetEssentialTrueDofs(sp.boundary_dofs())
dx, dy = ctx.alloc(sp.ndof).upload(np.random.default_rng(1).standard_normal(sp.ndof)), ctx.alloc(sp.ndof)
for _ in range(5):
    op.Mult(dx, dy)
ctx.synchronize()
ctx.mark(0)
for _ in range(a.applies):
    op.Mult(dx, dy)
ctx.mark(1)
ctx.synchronize()
print("%s: %.4f ms per apply (kernel %d)" % (os.environ.get("HCURL_LIB_PATH", "default library"), ctx.elapsed_ms() / a.applies, op.LastKernel()))
