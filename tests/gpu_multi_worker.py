"""Worker of tests/test_gpu_multi.py: one rank (one GPU) of an N-rank run.  The product path
(partition -> hcurl_space_create with shared-dof lists -> NCCL exchange inside hcurl_op_apply /
hcurl_pcg_solve) against the serial oracle.  The NCCL id travels through a file."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hephaestus_b200 import operators, partition  # noqa: E402
from oracle import assemble as oa, mesh as om, pcg as opcg, space as osp  # noqa: E402
from helpers import dof_descriptors, match_dofs  # noqa: E402
from test_partition import coef, serial_problem  # noqa: E402


def general_main(ctx, rank, world, name, p):
    """a reference mesh (committed fixture) under a host-supplied element partition (partition.GeneralProblem): apply and
    Jacobi-PCG against the serial oracle; local dofs keep the serial orientation, so no sign bookkeeping here"""
    from hephaestus_b200 import fespace
    from helpers import coef_by_attr, golden_mesh
    m, mo = golden_mesh(name)
    sp, so = fespace.NDSpace(m, p), osp.NDSpace(mo, p)
    part = partition.partition_by_coordinate(m, world)
    ess = sp.boundary_dofs()
    gp = partition.general_problem(sp, part, rank, ess)
    L = gp.local
    gsp = operators.GpuNDSpace(ctx, sp, L)
    op = operators.HephGpuOperator(ctx, gsp, 0)
    a, b = 0.05 * coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
    op.SetCoefficients(a[gp.elems], b[gp.elems])
    op.SetEssentialTrueDofs(gp.ess_tdofs)
    A = oa.assemble(so, a, b)
    own = np.nonzero(L.ldof_ltdof >= 0)[0]
    own = own[np.argsort(L.ldof_ltdof[own])]
    g = gp.l2g[own]
    x_s = np.random.default_rng(5).standard_normal(so.ndof)
    y_s = oa.constrained_apply(A, ess, x_s)
    err = np.linalg.norm(op.Mult(x_s[g]) - y_s[g]) / np.linalg.norm(y_s)
    assert err < 1e-12, ("apply", rank, err)
    Ac, _ = oa.eliminate(A, ess)
    b_s = Ac @ np.random.default_rng(54321).standard_normal(so.ndof)
    solver = operators.HephGpuPCG(ctx, tol=1e-8, max_iter=5000)
    solver.SetOperator(op)
    xs = solver.Mult(b_s[g])
    xo, it_o, _, conv_o, _ = opcg.pcg(Ac, b_s, diag=Ac.diagonal(), rtol=solver.tol, atol=solver.abstol, max_iter=5000)
    assert solver.converged and conv_o
    assert abs(solver.num_iterations - it_o) <= max(2, it_o // 100), (solver.num_iterations, it_o)
    serr = np.linalg.norm(xs - xo[g]) / np.linalg.norm(xo)
    assert serr < 1e-5, ("solve", rank, serr)
    print("OK rank %d (%s, %d of %d elements) apply err %.2e iters %d/%d solve err %.2e" % (rank, name, len(gp.elems), m.n_elem, err,
                                                                                           solver.num_iterations, it_o, serr))


def main():
    rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    general = sys.argv[4] == "general"
    if not general:
        n, p, perturbed = int(sys.argv[4]), int(sys.argv[5]), sys.argv[6] == "1"
    idf = os.path.join(tmp, "nccl_id")
    if rank == 0:
        uid = operators.Context.unique_id()
        with open(idf + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(idf + ".tmp", idf)
    else:
        t0 = time.time()
        while not os.path.exists(idf):
            time.sleep(0.05)
            assert time.time() - t0 < 120
        uid = open(idf, "rb").read()
    ctx = operators.Context(rank, uid, rank, world)
    if general:
        return general_main(ctx, rank, world, sys.argv[5], int(sys.argv[6]))
    bp = partition.brick_problem(rank, world, n, p, perturbed)
    L = bp.local
    gsp = operators.GpuNDSpace(ctx, bp.space, L)
    op = operators.HephGpuOperator(ctx, gsp, p + 1 if not perturbed else 0)
    a, b = coef(bp.mesh.attr)
    op.SetCoefficients(a, b)
    op.SetEssentialTrueDofs(bp.ess_tdofs)
    # serial oracle
    m, sp, mo, so = serial_problem(partition.brick_dims(world), n, p, perturbed)
    A = oa.assemble(so, *coef(mo.attr), q_override=None if perturbed else p + 1)
    ess = sp.boundary_dofs()
    lmo = om.Mesh("hex", bp.mesh.verts, bp.mesh.elems, bp.mesh.attr)
    idx, sgn = match_dofs(*dof_descriptors(lmo, osp.NDSpace(lmo, p)), *dof_descriptors(mo, so))
    own = np.nonzero(L.ldof_ltdof >= 0)[0]
    own = own[np.argsort(L.ldof_ltdof[own])]
    x_s = np.random.default_rng(5).standard_normal(so.ndof)
    y_s = oa.constrained_apply(A, ess, x_s)
    y = op.Mult(sgn[own] * x_s[idx[own]])
    err = np.linalg.norm(y - sgn[own] * y_s[idx[own]]) / np.linalg.norm(y_s)
    assert err < 1e-12, ("apply", rank, err)
    # PCG: same iteration count / solution as the serial oracle solve
    Ac, _ = oa.eliminate(A, ess)
    b_s = Ac @ np.random.default_rng(54321).standard_normal(so.ndof)
    solver = operators.HephGpuPCG(ctx, tol=1e-9, max_iter=3000, print_level=int(os.environ.get("HCURL_TEST_PCG_PRINT", "0")))
    solver.SetOperator(op)
    xs = solver.Mult(sgn[own] * b_s[idx[own]])
    xo, it_o, _, conv_o, _ = opcg.pcg(Ac, b_s, diag=Ac.diagonal(), rtol=solver.tol, atol=solver.abstol, max_iter=3000)
    assert solver.converged and conv_o
    assert abs(solver.num_iterations - it_o) <= 2, (solver.num_iterations, it_o)
    serr = np.linalg.norm(xs - sgn[own] * xo[idx[own]]) / np.linalg.norm(xo)
    assert serr < 1e-6, ("solve", rank, serr)
    print("OK rank %d apply err %.2e iters %d/%d solve err %.2e" % (rank, err, solver.num_iterations, it_o, serr))


if __name__ == "__main__":
    main()
