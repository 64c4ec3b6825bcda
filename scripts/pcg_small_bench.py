#!/usr/bin/env python
"""Jacobi-PCG at the reference's own problem sizes (data/team7.g: 68 547 dofs at order 1, 369 962 at order 2): iterations
per second of hcurl_pcg_solve with the Krylov loop as one CUDA graph (default) or driven by the host (HCURL_PCG_GRAPH=0),
and of the CPU arm (oracle/csr_kernels.c: assembled CSR + PCG, all host cores) at the same size.  One JSON line per case.

    python scripts/pcg_small_bench.py [--no-cpu]          (run it twice, with and without HCURL_PCG_GRAPH=0)"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hephaestus_b200 import fespace, operators  # noqa: E402
from helpers import coef_by_attr, golden_mesh  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--iters", type=int, default=2000)
    a = ap.parse_args()
    ctx = operators.Context(0)
    m, mo = golden_mesh("team7")
    dt = float(np.float32(0.001))
    for p in (1, 2):
        sp = fespace.NativeNDSpace(m, p)
        gsp = operators.GpuNDSpace(ctx, sp)
        op = operators.HephGpuOperator(ctx, gsp, 0)
        alpha, beta = dt * coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
        op.SetCoefficients(alpha, beta)
        op.SetEssentialTrueDofs(sp.boundary_dofs())
        b = np.random.default_rng(3).standard_normal(sp.ndof)
        b[sp.boundary_dofs()] = 0.0
        solver = operators.HephGpuPCG(ctx, tol=1e-30, abstol=0.0, max_iter=a.iters)   # run to the iteration limit
        solver.SetOperator(op)
        db, dx = ctx.alloc(sp.ndof).upload(b), ctx.alloc(sp.ndof)
        solver.Mult(db, dx)     # builds the graph / warms up
        best = 1e30
        for _ in range(3):
            solver.Mult(db, dx)
            best = min(best, solver.last_loop_ms())
        line = {"mesh": "team7.g", "order": p, "dofs": sp.ndof, "elements": int(m.n_elem), "iters": solver.num_iterations,
                "graph_loop": solver.last_used_graph(), "gpu_iters_per_s": solver.num_iterations / (best * 1e-3), "gpu_us_per_iter": best * 1e3 / solver.num_iterations}
        if not a.no_cpu:
            from oracle import csr as ocsr
            from oracle import assemble as oa, space as osp
            so = osp.NDSpace(mo, p)
            A = oa.assemble(so, alpha, beta).tocsr()
            Ac, _ = oa.eliminate(A, so.boundary_dofs())
            ocsr.lib().csr_set_threads(os.cpu_count() or 1)
            co = ocsr.CsrOperator(Ac)
            dinv = 1.0 / co.diagonal()
            co.pcg(b, dinv, rtol=1e-30, atol=0.0, max_iter=10)
            n_it = 300
            t0 = time.perf_counter()
            _, it_c, _, _ = co.pcg(b, dinv, rtol=1e-30, atol=0.0, max_iter=n_it)
            t = time.perf_counter() - t0
            line.update(cpu_iters_per_s=it_c / t, cpu_cores=ocsr.lib().csr_num_threads())
        print(json.dumps(line), flush=True)
        solver.close()
        op.close()


if __name__ == "__main__":
    main()
