"""Discrete gradient (H1 -> ND) and curl (ND -> RT) interpolators (reference: CurlAuxSolver,
/root/reference/src/auxsolvers/curl_aux.cpp:13-30, ScalarPotentialSource::BuildGrad,
/root/reference/src/sources/scalar_potential_source.cpp:89-97).

CPU part: the kernel's per-thread program (hephaestus_b200/csrc/interp_body.h, compiled for the host) against the
oracle's full 3-D evaluation (oracle/interp.py), exactness on polynomial fields, curl(grad) = 0, sign flips, map
encoding.  GPU part (opt-in, HCURL_TEST_INTERP=1: the kernel was written after the round-1 GPU budget was spent):
the same comparisons through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from oracle import interp as oi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GRAD, CURL = 0, 1


@pytest.fixture(scope="session")
def emu(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("emu") / "libinterpemu.so")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "hephaestus_b200", "csrc"),
           os.path.join(ROOT, "tests", "emu", "interp_emu.cpp"), "-o", path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.emu_interp_apply.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, P, P, P, P, P, C.c_int, P]
    return lib


# ---- structured grids with consistent (unsigned) global numberings, tensor-ordered element dofs -------------------------
def grid_dofs(kind, p, n):
    """element dofs [ne][nd] (tensor order) and dof count of the H1 / ND / RT space on an n[0] x n[1] x n[2] grid.
    Along a direction a component has either p points per element (open, not shared) or p+1 (closed, end points shared)."""
    def axis(e, npts):                      # global 1-D indices of an element's points along one direction
        return e * p + np.arange(npts)      # closed: p+1 points, last shared with the next element; open: p points
    comps = {"h1": [(p + 1, p + 1, p + 1)], "nd": [tuple(p if d == c else p + 1 for d in range(3)) for c in range(3)],
             "rt": [tuple(p + 1 if d == c else p for d in range(3)) for c in range(3)]}[kind]
    out, off = [], 0
    blocks = []
    for dims in comps:
        gdim = [n[d] * p + (1 if dims[d] == p + 1 else 0) for d in range(3)]
        rows = []
        for ez in range(n[2]):
            for ey in range(n[1]):
                for ex in range(n[0]):
                    ix, iy, iz = axis(ex, dims[0]), axis(ey, dims[1]), axis(ez, dims[2])
                    g = off + ix[None, None, :] + gdim[0] * (iy[None, :, None] + gdim[1] * iz[:, None, None])
                    rows.append(g.ravel())
        blocks.append(np.array(rows))
        off += gdim[0] * gdim[1] * gdim[2]
    return np.concatenate(blocks, axis=1).astype(np.int64), off


def grid_nodes(kind, p, n, h):
    """physical node of every global dof and its component"""
    nodes_ref, comp = {"h1": (oi.h1_nodes(p), np.zeros((p + 1) ** 3, int)), "nd": oi.nd_nodes(p), "rt": oi.rt_nodes(p)}[kind]
    dofs, nd = grid_dofs(kind, p, n)
    X, Cc = np.zeros((nd, 3)), np.zeros(nd, int)
    e = 0
    for ez in range(n[2]):
        for ey in range(n[1]):
            for ex in range(n[0]):
                X[dofs[e]] = (np.array([ex, ey, ez]) + nodes_ref) * h
                Cc[dofs[e]] = comp
                e += 1
    return X, Cc


def signed(dofs, sigma):
    return np.where(sigma[dofs] > 0, dofs, -1 - dofs).astype(np.int32)


def run_emu(emu, kind, p, in_dofs, out_dofs, n_in, n_out, x, tet=False):
    lib = L.lib()
    ne = in_dofs.shape[0]
    im, om = np.zeros(in_dofs.shape, np.uint32), np.zeros(out_dofs.shape, np.uint32)
    if tet:   # the encoding does not depend on the element type: the same bits, written here for the 4 / 6 / 4 local dofs
        ii, oo = np.where(in_dofs >= 0, in_dofs, -1 - in_dofs), np.where(out_dofs >= 0, out_dofs, -1 - out_dofs)
        first = np.zeros(n_out, bool)
        skip = np.zeros(oo.size, bool)
        for k, d in enumerate(oo.ravel()):
            skip[k] = first[d]
            first[d] = True
        im[:] = (ii | np.where(in_dofs < 0, 0x80000000, 0)).astype(np.uint32)
        om[:] = (oo | np.where(out_dofs < 0, 0x80000000, 0) | np.where(skip.reshape(oo.shape), 0x40000000, 0)).astype(np.uint32)
        M = np.zeros(24)
        L.check(lib.hcurl_host_interp_tet_matrix(kind, L.ptr(M)), "tet matrix")
        y = np.full(n_out, np.nan)
        emu.emu_interp_apply(kind, 1, ne, in_dofs.shape[1], out_dofs.shape[1], L.ptr(im), L.ptr(om), None, L.ptr(x), L.ptr(y), 1, L.ptr(M))
        return y
    L.check(lib.hcurl_host_interp_encode(kind, p, ne, L.ptr(in_dofs), L.ptr(out_dofs), n_in, n_out, L.ptr(im), L.ptr(om)), "encode")
    # every output dof has exactly one writer
    writers = np.bincount((om & 0x3FFFFFFF)[(om & 0x40000000) == 0].ravel(), minlength=n_out)
    assert np.all(writers == 1)
    D = np.zeros(p * (p + 1))
    L.check(lib.hcurl_host_interp_matrix(p, L.ptr(D)), "matrix")
    y = np.full(n_out, np.nan)
    emu.emu_interp_apply(kind, p, ne, in_dofs.shape[1], out_dofs.shape[1], L.ptr(im), L.ptr(om), L.ptr(D), L.ptr(x), L.ptr(y), 0, None)
    return y


@pytest.mark.parametrize("p", [1, 2, 3, 4])
def test_local_matrices_match_oracle(emu, p, lib_built):
    """one element, identity maps: the kernel program reproduces the oracle's local matrices column by column"""
    for kind, M in ((GRAD, oi.local_grad_matrix(p)), (CURL, oi.local_curl_matrix(p))):
        n_out, n_in = M.shape
        ind, outd = np.arange(n_in, dtype=np.int32)[None], np.arange(n_out, dtype=np.int32)[None]
        got = np.stack([run_emu(emu, kind, p, ind, outd, n_in, n_out, e) for e in np.eye(n_in)], axis=1)
        assert np.abs(got - M).max() < 1e-12 * np.abs(M).max()


@pytest.mark.parametrize("p,n", [(1, (3, 2, 4)), (2, (2, 3, 2)), (3, (2, 2, 1)), (4, (1, 2, 1))])
def test_grid_apply_signs_exactness_and_complex_property(emu, p, n, lib_built):
    h = np.array([0.7, 1.1, 0.4])
    rng = np.random.default_rng(p)
    dh, nh = grid_dofs("h1", p, n)
    dn, nn = grid_dofs("nd", p, n)
    dr, nr = grid_dofs("rt", p, n)
    sn, sr = rng.choice([-1.0, 1.0], nn), rng.choice([-1.0, 1.0], nr)      # orientation of the global ND / RT dofs
    dh_s, dn_s, dr_s = dh.astype(np.int32), signed(dn, sn), signed(dr, sr)
    # --- gradient of a polynomial of degree p in every variable is reproduced exactly
    cx, cy, cz = (rng.standard_normal(p + 1) for _ in range(3))
    P = np.polynomial.polynomial
    Xh, _ = grid_nodes("h1", p, n, h)
    phi = P.polyval(Xh[:, 0], cx) * P.polyval(Xh[:, 1], cy) * P.polyval(Xh[:, 2], cz)
    Xn, Cn = grid_nodes("nd", p, n, h)
    val = [P.polyval(Xn[:, d], c) for d, c in enumerate((cx, cy, cz))]
    der = [P.polyval(Xn[:, d], P.polyder(c)) for d, c in enumerate((cx, cy, cz))]
    grad = np.stack([der[0] * val[1] * val[2], val[0] * der[1] * val[2], val[0] * val[1] * der[2]], axis=1)
    expect = sn * h[Cn] * grad[np.arange(nn), Cn]                            # covariant pull-back: h_c * u_c
    g = run_emu(emu, GRAD, p, dh_s, dn_s, nh, nn, phi)
    assert np.abs(g - expect).max() < 1e-11 * np.abs(expect).max()
    assert np.abs(g - oi.apply("grad", p, dh_s, dn_s, nn, phi)).max() < 1e-12 * np.abs(expect).max()
    # --- curl of a field of the ND space: u_c = x_c^(p-1) * (product of degree-p polynomials in the other variables) ...
    a = rng.standard_normal((3, 3, p + 1))
    for c in range(3):
        a[c, c, p] = 0.0                                                     # degree p-1 along the own direction

    def comp(c, X, dvar=None):
        f = np.ones(len(X))
        for d in range(3):
            co = P.polyder(a[c, d]) if d == dvar else a[c, d]
            f = f * P.polyval(X[:, d], co)
        return f
    u = sn * h[Cn] * np.choose(Cn, [comp(0, Xn), comp(1, Xn), comp(2, Xn)])
    Xr, Cr = grid_nodes("rt", p, n, h)
    curl = np.stack([comp(2, Xr, 1) - comp(1, Xr, 2), comp(0, Xr, 2) - comp(2, Xr, 0), comp(1, Xr, 0) - comp(0, Xr, 1)], axis=1)
    area = np.array([h[1] * h[2], h[0] * h[2], h[0] * h[1]])
    expect = sr * area[Cr] * curl[np.arange(nr), Cr]                         # contravariant pull-back: h_a h_b * w_c
    w = run_emu(emu, CURL, p, dn_s, dr_s, nn, nr, u)
    assert np.abs(w - expect).max() < 1e-10 * np.abs(expect).max()
    assert np.abs(w - oi.apply("curl", p, dn_s, dr_s, nr, u)).max() < 1e-12 * np.abs(expect).max()
    # --- curl(grad phi) = 0 for arbitrary H1 vectors
    phi = rng.standard_normal(nh)
    g = run_emu(emu, GRAD, p, dh_s, dn_s, nh, nn, phi)
    assert np.abs(run_emu(emu, CURL, p, dn_s, dr_s, nn, nr, g)).max() < 1e-11 * np.abs(g).max()


# ---- tets, order 1 (the TEAM7 configuration: ND1 -> RT0, H1 P1 -> ND1) -------------------------------------------------
def tet_spaces(n=(2, 2, 2)):
    """(vertices, H1 dofs [ne][4], signed ND1 dofs [ne][6], signed RT0 dofs [ne][4], counts): ND1 from the product's
    MFEM-convention generator; RT0 numbered by face, oriented outward of the first element that holds the face"""
    from helpers import make_pair
    from oracle import mesh as om
    m, sp, mo, so = make_pair("tet", n, 1, True)
    topo = om.Topology(mo)
    first = np.array([fe[0] for fe in topo.face_elems])
    e_idx = np.arange(mo.n_elem)[:, None]
    rt = np.where(first[topo.el_faces] == e_idx, topo.el_faces, -1 - topo.el_faces).astype(np.int32)
    return mo, mo.elems.astype(np.int32), np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32), rt, (len(mo.verts), sp.ndof, topo.n_faces)


def test_tet_order1_local_matrices_and_exactness(emu, lib_built):
    lib = L.lib()
    for kind, M in ((GRAD, oi.tet_local_grad_matrix()), (CURL, oi.tet_local_curl_matrix())):
        got = np.zeros(24)
        L.check(lib.hcurl_host_interp_tet_matrix(kind, L.ptr(got)), "tet matrix")
        assert np.array_equal(got.reshape(M.shape), M)          # closed form (incidence, x 2 for the curl) == basis evaluation
    mo, h1, nd, rt, (nv, nn, nf) = tet_spaces()
    rng = np.random.default_rng(6)
    X = mo.verts
    # gradient: edge dof = phi(x_b) - phi(x_a) in the global edge orientation; checked through curl(grad) = 0 and,
    # for a linear phi, against the ND1 dofs of its (constant) gradient: dof = g . (x_b - x_a)
    g = rng.standard_normal(3)
    phi = X @ g + 0.3
    u = run_emu(emu, GRAD, 1, h1, nd, nv, nn, phi, tet=True)
    ii = np.where(nd >= 0, nd, -1 - nd)
    sg = np.where(nd >= 0, 1.0, -1.0)
    from oracle.nd_tet import TET_EDGES, TET_FACES
    ev = X[mo.elems]                                                             # (ne, 4, 3)
    loc = np.stack([(ev[:, b] - ev[:, a]) @ g for a, b in TET_EDGES], axis=1)    # local-orientation dofs of grad phi
    expect = np.zeros(nn)
    expect[ii.ravel()] = (sg * loc).ravel()
    assert np.abs(u - expect).max() < 1e-13 * np.abs(expect).max()
    assert np.abs(u - oi.apply("grad", 1, h1, nd, nn, phi, tet=True)).max() < 1e-13
    # curl of u = a + b x X (in ND1): RT0 dof_f = nk_f . (det J J^-1 (2 b)), J = [x1-x0, x2-x0, x3-x0]
    a0, b0 = rng.standard_normal(3), rng.standard_normal(3)
    mid = np.stack([0.5 * (ev[:, a] + ev[:, b]) for a, b in TET_EDGES], axis=1)
    uloc = np.stack([np.einsum("ed,ed->e", a0 + np.cross(b0, mid[:, k]), ev[:, b] - ev[:, a]) for k, (a, b) in enumerate(TET_EDGES)], axis=1)
    uu = np.zeros(nn)
    uu[ii.ravel()] = (sg * uloc).ravel()
    w = run_emu(emu, CURL, 1, nd, rt, nn, nf, uu, tet=True)
    J = np.stack([ev[:, 1] - ev[:, 0], ev[:, 2] - ev[:, 0], ev[:, 3] - ev[:, 0]], axis=2)
    det = np.linalg.det(J)
    assert np.all(det > 0)
    what = det[:, None] * np.linalg.solve(J, np.tile(2 * b0, (len(J), 1))[..., None])[..., 0]   # contravariant pull-back
    floc = what @ oi.TET_NK.T                                                                    # (ne, 4)
    fi, fs = np.where(rt >= 0, rt, -1 - rt), np.where(rt >= 0, 1.0, -1.0)
    expect = np.zeros(nf)
    expect[fi.ravel()] = (fs * floc).ravel()
    both = np.zeros(nf)
    np.add.at(both, fi.ravel(), np.abs(fs * floc - expect[fi]).ravel())
    assert both.max() < 1e-12 * np.abs(expect).max()             # the two elements of an interior face agree
    assert np.abs(w - expect).max() < 1e-12 * np.abs(expect).max()
    # and = 2 x the flux of curl u through the face (nk is twice the area-weighted normal)
    f0 = TET_FACES[0]
    nA = 0.5 * np.cross(ev[:, f0[1]] - ev[:, f0[0]], ev[:, f0[2]] - ev[:, f0[0]])
    assert np.abs(2 * np.abs(nA @ (2 * b0)) - np.abs(floc[:, 0])).max() < 1e-12 * np.abs(floc).max()
    assert np.abs(w - oi.apply("curl", 1, nd, rt, nf, uu, tet=True)).max() < 1e-12 * np.abs(expect).max()
    # curl(grad) = 0
    gphi = run_emu(emu, GRAD, 1, h1, nd, nv, nn, rng.standard_normal(nv), tet=True)
    assert np.abs(run_emu(emu, CURL, 1, nd, rt, nn, nf, gphi, tet=True)).max() < 1e-13 * np.abs(gphi).max()


def test_bad_arguments_are_reported(lib_built):
    lib = L.lib()
    D = np.zeros(30)
    assert lib.hcurl_host_interp_matrix(5, L.ptr(D)) != 0 and b"bad arguments" in lib.hcurl_last_error()
    ind, outd = np.array([[0, 1, 2, 3, 4, 5, 6, 9]], np.int32), np.arange(12, dtype=np.int32)[None]
    im, om = np.zeros((1, 8), np.uint32), np.zeros((1, 12), np.uint32)
    assert lib.hcurl_host_interp_encode(GRAD, 1, 1, L.ptr(ind), L.ptr(outd), 8, 12, L.ptr(im), L.ptr(om)) != 0
    assert b"out of range" in lib.hcurl_last_error()
    assert lib.hcurl_host_interp_encode(7, 1, 1, L.ptr(ind), L.ptr(outd), 10, 12, L.ptr(im), L.ptr(om)) != 0


# ---- GPU (opt-in until run on a B200) -------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("p,n", [(1, (9, 7, 5)), (2, (5, 4, 3)), (3, (3, 2, 2)), (4, (2, 2, 1))])
def test_gpu_interpolators_match_oracle(ctx, p, n):
    from hephaestus_b200 import operators
    rng = np.random.default_rng(10 + p)
    dh, nh = grid_dofs("h1", p, n)
    dn, nn = grid_dofs("nd", p, n)
    dr, nr = grid_dofs("rt", p, n)
    sn, sr = rng.choice([-1.0, 1.0], nn), rng.choice([-1.0, 1.0], nr)
    dh_s, dn_s, dr_s = dh.astype(np.int32), signed(dn, sn), signed(dr, sr)
    G = operators.HephGpuInterpolator(ctx, GRAD, p, dh_s, dn_s, nh, nn)
    Cu = operators.HephGpuInterpolator(ctx, CURL, p, dn_s, dr_s, nn, nr)
    phi, u = rng.standard_normal(nh), rng.standard_normal(nn)
    g = G.Mult(phi)
    ref = oi.apply("grad", p, dh_s, dn_s, nn, phi)
    assert np.abs(g - ref).max() <= 1e-12 * np.abs(ref).max()
    w = Cu.Mult(u)
    ref = oi.apply("curl", p, dn_s, dr_s, nr, u)
    assert np.abs(w - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(Cu.Mult(g)).max() <= 1e-11 * np.abs(g).max()
    # device-resident path
    dg, dw = ctx.alloc(nn).upload(g), ctx.alloc(nr)
    Cu.Mult(dg, dw)
    ctx.synchronize()
    assert np.abs(dw.download()).max() <= 1e-11 * np.abs(g).max()


@pytest.mark.gpu
def test_gpu_tet_order1_interpolators_match_oracle(ctx):
    from hephaestus_b200 import operators
    mo, h1, nd, rt, (nv, nn, nf) = tet_spaces((3, 3, 2))
    rng = np.random.default_rng(21)
    G = operators.HephGpuInterpolator(ctx, GRAD, 1, h1, nd, nv, nn, geom=L.TET)
    Cu = operators.HephGpuInterpolator(ctx, CURL, 1, nd, rt, nn, nf, geom=L.TET)
    phi, u = rng.standard_normal(nv), rng.standard_normal(nn)
    g = G.Mult(phi)
    assert np.abs(g - oi.apply("grad", 1, h1, nd, nn, phi, tet=True)).max() <= 1e-13 * np.abs(g).max()
    w = Cu.Mult(u)
    ref = oi.apply("curl", 1, nd, rt, nf, u, tet=True)
    assert np.abs(w - ref).max() <= 1e-12 * np.abs(ref).max()
    assert np.abs(Cu.Mult(g)).max() <= 1e-12 * np.abs(g).max()
