"""Worker of tests/test_partition.py::test_gloo_world2_exchange: one rank of a world_size-2 run on
CPU (gloo).  Applies the rank-local operator (oracle matrix, the checker) with the product's shared-dof
lists driving real send/recv, and compares with the serial oracle apply."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hephaestus_b200 import partition  # noqa: E402
from oracle import assemble as oa, mesh as om, space as osp  # noqa: E402
from helpers import dof_descriptors, match_dofs  # noqa: E402
from test_partition import coef, serial_problem  # noqa: E402


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, p = 3, 2
    bp = partition.brick_problem(rank, world, n, p, True)
    L = bp.local
    lmo = om.Mesh("hex", bp.mesh.verts, bp.mesh.elems, bp.mesh.attr)
    lso = osp.NDSpace(lmo, p)
    A_loc = oa.assemble(lso, *coef(lmo.attr))
    m, sp, mo, so = serial_problem(partition.brick_dims(world), n, p, True)
    A = oa.assemble(so, *coef(mo.attr))
    idx, sgn = match_dofs(*dof_descriptors(lmo, lso), *dof_descriptors(mo, so))
    x_s = np.random.default_rng(5).standard_normal(so.ndof)
    y_s = A @ x_s
    own = np.nonzero(L.ldof_ltdof >= 0)[0]
    l2e = partition.ext_index(L)
    ext = np.zeros(L.n_tdof + int(L.ghost_off[-1]))
    ext[L.ldof_ltdof[own]] = sgn[own] * x_s[idx[own]]

    def exchange(v, reduce):
        reqs, recvs = [], []
        for k, q in enumerate(L.nbr_rank):
            o0, o1, g0, g1 = L.own_off[k], L.own_off[k + 1], L.ghost_off[k], L.ghost_off[k + 1]
            if not reduce:   # owner -> ghost
                if o1 > o0:
                    reqs.append(dist.isend(torch.from_numpy(v[l2e[L.own_ldof[o0:o1]]].copy()), int(q)))
                if g1 > g0:
                    buf = torch.zeros(g1 - g0, dtype=torch.float64)
                    reqs.append(dist.irecv(buf, int(q)))
                    recvs.append((buf, k))
            else:            # ghost -> owner (sum)
                if g1 > g0:
                    reqs.append(dist.isend(torch.from_numpy(v[L.n_tdof + g0:L.n_tdof + g1] * L.ghost_sign[g0:g1]), int(q)))
                if o1 > o0:
                    buf = torch.zeros(o1 - o0, dtype=torch.float64)
                    reqs.append(dist.irecv(buf, int(q)))
                    recvs.append((buf, k))
        for r in reqs:
            r.wait()
        for buf, k in recvs:
            if not reduce:
                g0, g1 = L.ghost_off[k], L.ghost_off[k + 1]
                v[L.n_tdof + g0:L.n_tdof + g1] = buf.numpy() * L.ghost_sign[g0:g1]
            else:
                np.add.at(v, l2e[L.own_ldof[L.own_off[k]:L.own_off[k + 1]]], buf.numpy())

    exchange(ext, False)
    yl = A_loc @ ext[l2e]
    ye = np.zeros_like(ext)
    np.add.at(ye, l2e, yl)
    exchange(ye, True)
    got = ye[L.ldof_ltdof[own]]
    ref = sgn[own] * y_s[idx[own]]
    err = np.linalg.norm(got - ref) / np.linalg.norm(y_s)
    # global dot product over true dofs == serial dot (all-reduce of the PCG scalars)
    t = torch.tensor([float(ye[:L.n_tdof] @ ext[:L.n_tdof])], dtype=torch.float64)
    dist.all_reduce(t)
    derr = abs(t.item() - x_s @ y_s) / abs(x_s @ y_s)
    dist.destroy_process_group()
    assert err < 1e-12 and derr < 1e-12, (err, derr)
    print("OK rank %d err %.2e dot err %.2e" % (rank, err, derr))


if __name__ == "__main__":
    main()
