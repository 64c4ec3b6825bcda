"""Hardware micro-benchmarks that back the kernel design: FP64 FMA / DMMA
peaks and the gather + FP64-atomic-scatter floor of the order-2 hex apply."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hephaestus_b200 import _lib as L, fespace, operators  # noqa: E402

lib = L.lib()
lib.hcurl_debug_gather_scatter_ms.restype = C.c_double
lib.hcurl_debug_gather_scatter_ms.argtypes = [C.c_void_p, C.c_int, C.c_int]
lib.hcurl_debug_dmma_tflops.restype = C.c_double
lib.hcurl_debug_dmma_tflops.argtypes = [C.c_void_p, C.c_int]
ctx = operators.Context(0)
print("DFMA peak (api)      %.2f TFLOP/s" % ctx.fp64_peak_tflops())
for mode, name in ((1, "DFMA only"), (0, "DMMA only"), (2, "DMMA+DFMA co-issued (sum)")):
    print("%-28s %.2f TFLOP/s" % (name, lib.hcurl_debug_dmma_tflops(ctx.h, mode)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mesh = fespace.cartesian_hex_mesh(n, n, n)
space = fespace.NDSpace(mesh, 2)
gsp = operators.GpuNDSpace(ctx, space)
op = operators.HephGpuOperator(ctx, gsp, 3)
op.SetCoefficients(np.ones(mesh.n_elem), np.ones(mesh.n_elem))
for mode, name in ((0, "map + x gather only"), (1, "gather + RED.F64 scatter")):
    ms = lib.hcurl_debug_gather_scatter_ms(op.h, mode, 5)
    print("%-28s %.3f ms  (%.1f GDOF/s equivalent, n=%d, %d DOFs)" % (name, ms, space.ndof / ms / 1e6, n, space.ndof))
