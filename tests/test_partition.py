"""Multi-rank layout (hephaestus_b200/partition.py): ownership, shared-dof lists and relative
orientations must make the rank-local applies + exchange equal the serial operator (SURVEY.md 8e,
A.7 vi).  CPU only: the exchange of csrc/comm.cu is emulated in numpy (all ranks in one process) and
run for real over torch.distributed/gloo with world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest

from hephaestus_b200 import fespace, partition
from oracle import assemble as oa
from oracle import mesh as om
from oracle import space as osp

from helpers import dof_descriptors, match_dofs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def serial_problem(dims, n, p, perturbed):
    px, py, pz = dims
    m = fespace.cartesian_hex_mesh(px * n, py * n, pz * n, px, py, pz)
    if perturbed:
        N = np.array([px * n, py * n, pz * n])
        g = np.rint(m.verts * n).astype(np.int64)
        gid = g[:, 0] + (N[0] + 1) * (g[:, 1] + (N[1] + 1) * g[:, 2])
        interior = np.all((g > 0) & (g < N), axis=1)
        h = 1.0 / n
        d = np.stack([partition._hash_unit(gid, 777 + k) for k in range(3)], 1) * 0.4 * h - 0.2 * h
        m = fespace.Mesh("hex", m.verts + d * interior[:, None], m.elems, m.attr)
    mo = om.Mesh("hex", m.verts, m.elems, m.attr)
    return m, fespace.NDSpace(m, p), mo, osp.NDSpace(mo, p)


def coef(attr):
    return np.where(attr == 1, 0.05, 0.1), np.where(attr == 1, 1.0, 1.0e3)


def emulate_apply(bps, mats, x_t):
    """x_t: list of T-vectors.  Returns list of T-vectors y = A x with the broadcast / local apply /
    reduce sequence of csrc/api.cu:op_apply_ext."""
    W = len(bps)
    l2e = [partition.ext_index(b.local) for b in bps]
    ext = [partition.to_ext(b.local, x) for b, x in zip(bps, x_t)]
    # broadcast owner -> ghosts
    for r in range(W):
        L = bps[r].local
        for k, q in enumerate(L.nbr_rank):
            Lq = bps[q].local
            kq = list(Lq.nbr_rank).index(r)
            send = ext[q][l2e[q][Lq.own_ldof[Lq.own_off[kq]:Lq.own_off[kq + 1]]]]
            g0, g1 = L.ghost_off[k], L.ghost_off[k + 1]
            assert len(send) == g1 - g0
            ext[r][L.n_tdof + g0:L.n_tdof + g1] = send * L.ghost_sign[g0:g1]
    # local apply on L-vectors
    yext = []
    for r in range(W):
        xl = ext[r][l2e[r]]
        yl = mats[r] @ xl
        ye = np.zeros_like(ext[r])
        np.add.at(ye, l2e[r], yl)
        yext.append(ye)
    # reduce ghosts -> owner
    for r in range(W):
        L = bps[r].local
        for k, q in enumerate(L.nbr_rank):
            Lq = bps[q].local
            kq = list(Lq.nbr_rank).index(r)
            g0, g1 = L.ghost_off[k], L.ghost_off[k + 1]
            part = yext[r][L.n_tdof + g0:L.n_tdof + g1] * L.ghost_sign[g0:g1]
            np.add.at(yext[q], l2e[q][Lq.own_ldof[Lq.own_off[kq]:Lq.own_off[kq + 1]]], part)
    return [ye[:b.local.n_tdof] for ye, b in zip(yext, bps)]


@pytest.mark.parametrize("world,n,p", [(2, 2, 1), (2, 3, 2), (4, 2, 2), (8, 2, 2), (8, 2, 1), (2, 2, 3), (6, 2, 2)])
@pytest.mark.parametrize("perturbed", [False, True])
def test_partitioned_apply_equals_serial(world, n, p, perturbed):
    dims = partition.brick_dims(world)
    m, sp, mo, so = serial_problem(dims, n, p, perturbed)
    a, b = coef(mo.attr)
    A = oa.assemble(so, a, b)
    Ps, Ds = dof_descriptors(mo, so)
    rng = np.random.default_rng(5)
    x_s = rng.standard_normal(so.ndof)
    y_s = A @ x_s
    bps, mats, maps = [], [], []
    for r in range(world):
        bp = partition.brick_problem(r, world, n, p, perturbed)
        lmo = om.Mesh("hex", bp.mesh.verts, bp.mesh.elems, bp.mesh.attr)
        lso = osp.NDSpace(lmo, p)
        assert np.array_equal(lso.signed_elem_dofs(), bp.space.signed_elem_dofs())
        la, lb = coef(lmo.attr)
        mats.append(oa.assemble(lso, la, lb))
        P, D = dof_descriptors(lmo, lso)
        maps.append(match_dofs(P, D, Ps, Ds))      # local dof -> (serial dof, sign)
        bps.append(bp)
    # every serial dof is owned exactly once
    owned_serial = np.concatenate([maps[r][0][bps[r].local.ldof_ltdof >= 0] for r in range(world)])
    assert np.array_equal(np.sort(owned_serial), np.arange(so.ndof))
    # owner = lowest rank holding the dof
    holder_min = np.full(so.ndof, world)
    for r in range(world):
        np.minimum.at(holder_min, maps[r][0], r)
    for r in range(world):
        own = bps[r].local.ldof_ltdof >= 0
        assert np.all(holder_min[maps[r][0][own]] == r)
    # T-vectors of the serial x
    x_t = []
    for r in range(world):
        L = bps[r].local
        own = np.nonzero(L.ldof_ltdof >= 0)[0]
        xt = np.zeros(L.n_tdof)
        xt[L.ldof_ltdof[own]] = maps[r][1][own] * x_s[maps[r][0][own]]
        x_t.append(xt)
    y_t = emulate_apply(bps, mats, x_t)
    for r in range(world):
        L = bps[r].local
        own = np.nonzero(L.ldof_ltdof >= 0)[0]
        ref = maps[r][1][own] * y_s[maps[r][0][own]]
        got = y_t[r][L.ldof_ltdof[own]]
        assert np.linalg.norm(got - ref) <= 1e-12 * np.linalg.norm(y_s), (r, np.abs(got - ref).max())
    # essential dofs: union over ranks == boundary dofs of the serial space
    ess_serial = np.concatenate([maps[r][0][np.nonzero(bps[r].local.ldof_ltdof >= 0)[0][np.argsort(
        bps[r].local.ldof_ltdof[bps[r].local.ldof_ltdof >= 0])][bps[r].ess_tdofs]] for r in range(world)])
    assert np.array_equal(np.sort(ess_serial), np.sort(sp.boundary_dofs()))


def test_gloo_world2_exchange(tmp_path):
    """Two real processes over torch.distributed (gloo): broadcast + reduce with send/recv in the
    order csrc/comm.cu issues them."""
    script = os.path.join(ROOT, "tests", "gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(tmp_path)], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert all("OK" in o for o in outs), outs


# ---- general meshes with a host-supplied element partition (partition.GeneralProblem) -----------------------------------
def _local_matrix(so, a, b, gp):
    """the rank's share of the assembled operator in its L-vector numbering (serial dof orientation)"""
    import scipy.sparse as sps
    n = so.ndof
    acc = sps.csr_matrix((n, n))
    mine = np.zeros(so.mesh.n_elem, bool)
    mine[gp.elems] = True
    for e0, e1, A in oa.element_matrices(so, a, b):
        G = oa.to_global(so, e0, e1, A)
        sel = mine[e0:e1]
        if not sel.any():
            continue
        d = so.elem_dofs[e0:e1][sel]
        nd = d.shape[1]
        acc = acc + sps.coo_matrix((G[sel].ravel(), (np.repeat(d[:, :, None], nd, 2).ravel(), np.repeat(d[:, None, :], nd, 1).ravel())),
                                   shape=(n, n)).tocsr()
    return acc[gp.l2g][:, gp.l2g]


@pytest.mark.parametrize("kind,n,p,world", [("tet", (3, 2, 2), 1, 3), ("tet", (2, 2, 2), 2, 4), ("hex", (4, 3, 3), 2, 3), ("hex", (3, 3, 2), 1, 5),
                                            ("tet", (2, 2, 1), 3, 2)])
def test_general_partition_apply_equals_serial(kind, n, p, world):
    from helpers import make_pair
    m, sp, mo, so = make_pair(kind, n, p, True)
    a, b = coef(np.arange(mo.n_elem) % 2 + 1)
    A = oa.assemble(so, a, b)
    part = partition.partition_by_coordinate(m, world)
    assert len(np.unique(part)) == world and np.bincount(part).max() - np.bincount(part).min() <= 1
    ess = sp.boundary_dofs()
    gps = [partition.general_problem(sp, part, r, ess) for r in range(world)]
    mats = [_local_matrix(so, a, b, g) for g in gps]
    # every serial dof owned exactly once, by the lowest rank holding it
    owned = np.concatenate([g.l2g[g.local.ldof_ltdof >= 0] for g in gps])
    assert np.array_equal(np.sort(owned), np.arange(sp.ndof))
    x_s = np.random.default_rng(5).standard_normal(sp.ndof)
    y_s = A @ x_s
    x_t = []
    for g in gps:
        own = g.local.ldof_ltdof >= 0
        xt = np.zeros(g.local.n_tdof)
        xt[g.local.ldof_ltdof[own]] = x_s[g.l2g[own]]
        x_t.append(xt)
    y_t = emulate_apply(gps, mats, x_t)
    for g, yt in zip(gps, y_t):
        own = g.local.ldof_ltdof >= 0
        assert np.linalg.norm(yt[g.local.ldof_ltdof[own]] - y_s[g.l2g[own]]) <= 1e-12 * np.linalg.norm(y_s)
    # essential dofs: the union over ranks is the serial set
    ess_union = np.concatenate([g.l2g[np.nonzero(g.local.ldof_ltdof >= 0)[0][np.argsort(g.local.ldof_ltdof[g.local.ldof_ltdof >= 0])][g.ess_tdofs]]
                                for g in gps])
    assert np.array_equal(np.sort(ess_union), np.sort(ess))
    # local maps only reference local dofs and keep the serial orientation
    for g in gps:
        sd = np.asarray(sp.signed_elem_dofs())[g.elems]
        ld = g.local.elem_ldof
        assert np.array_equal(ld < 0, sd < 0)
        assert np.array_equal(g.l2g[np.where(ld >= 0, ld, -1 - ld)], np.where(sd >= 0, sd, -1 - sd))


@pytest.mark.parametrize("kind,n,p,world", [("tet", (3, 2, 2), 2, 4), ("hex", (4, 3, 3), 2, 3), ("hex", (3, 3, 2), 1, 5)])
def test_cpp_general_partition_equals_python(kind, n, p, world, lib_built):
    """csrc/partition.cpp (hcurl_host_partition_*) builds the arrays of partition.GeneralProblem"""
    import ctypes as C
    from hephaestus_b200 import _lib as L
    from helpers import make_pair
    lib = L.lib()
    m, sp, mo, so = make_pair(kind, n, p, True)
    part = partition.partition_by_coordinate(m, world)
    sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
    er = np.ascontiguousarray(part, dtype=np.int32)
    for r in range(world):
        g = partition.general_problem(sp, part, r)
        h = L.check_ptr(lib.hcurl_host_partition_create(len(sd), sd.shape[1], sp.ndof, L.ptr(sd), L.ptr(er), r), "partition")
        ne, nl, nt, nn, no, ng = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int(), C.c_int32(), C.c_int32()
        L.check(lib.hcurl_host_partition_sizes(h, C.byref(ne), C.byref(nl), C.byref(nt), C.byref(nn), C.byref(no), C.byref(ng)), "sizes")
        Lc = g.local
        assert (ne.value, nl.value, nt.value, nn.value, no.value, ng.value) == (len(g.elems), Lc.n_ldof, Lc.n_tdof, len(Lc.nbr_rank),
                                                                                   len(Lc.own_ldof), len(Lc.ghost_ldof))
        elems, l2g = np.zeros(ne.value, np.int64), np.zeros(nl.value, np.int32)
        eld, lt = np.zeros((ne.value, sd.shape[1]), np.int32), np.zeros(nl.value, np.int32)
        nbr, oo, go = np.zeros(nn.value, np.int32), np.zeros(nn.value + 1, np.int32), np.zeros(nn.value + 1, np.int32)
        ol, gl = np.zeros(max(no.value, 1), np.int32), np.zeros(max(ng.value, 1), np.int32)
        L.check(lib.hcurl_host_partition_copy(h, L.ptr(elems), L.ptr(l2g), L.ptr(eld), L.ptr(lt), L.ptr(nbr), L.ptr(oo), L.ptr(ol), L.ptr(go), L.ptr(gl)), "copy")
        lib.hcurl_host_partition_destroy(h)
        assert np.array_equal(elems, g.elems) and np.array_equal(l2g, g.l2g) and np.array_equal(eld, Lc.elem_ldof)
        assert np.array_equal(lt, Lc.ldof_ltdof) and np.array_equal(nbr, Lc.nbr_rank)
        assert np.array_equal(oo, Lc.own_off) and np.array_equal(go, Lc.ghost_off)
        assert np.array_equal(ol[:no.value], Lc.own_ldof) and np.array_equal(gl[:ng.value], Lc.ghost_ldof)
    assert lib.hcurl_host_partition_create(len(sd), sd.shape[1], 5, L.ptr(sd), L.ptr(er), 0) is None      # dof index out of range
    assert b"out of range" in lib.hcurl_last_error()
