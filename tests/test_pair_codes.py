"""Dof pairs of an order-2 hex space (hcurl_host_hex_p2_pair_codes, csrc/patch.cpp: the 16-byte units of the patch
kernel) against an independent restatement with the oracle's tables; runs without a GPU."""
import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle.nd_hex import NDHex

from helpers import make_pair, star_mesh

SWAP, NEG1, NEG0, IDX = 0x20000000, 0x40000000, 0x80000000, 0x1FFFFFFF


def rot_slot_to_tensor(p, slot):
    """slot s of the per-component rotated layout (component c; indices along axes c, c+1, c+2, the open one fastest)
    -> index in the standard tensor order (x fastest inside each component)"""
    p1, nc = p + 1, p * (p + 1) ** 2
    c, t = divmod(slot, nc)
    ir = [t % p, (t // p) % p1, t // (p * p1)]
    idx = [0, 0, 0]
    for k in range(3):
        idx[(c + k) % 3] = ir[k]
    n0 = p if c == 0 else p1
    n1 = p if c == 1 else p1
    return c * nc + idx[0] + n0 * (idx[1] + n1 * idx[2])


def pair_codes(sd):
    lib = L.lib()
    sd = np.ascontiguousarray(sd, dtype=np.int32)
    codes = np.zeros((len(sd), 27), np.uint32)
    rc = lib.hcurl_host_hex_p2_pair_codes(len(sd), L.ptr(sd), L.ptr(codes))
    return rc, codes


@pytest.mark.parametrize("kind", ["cube", "scrambled", "perturbed", "star"])
def test_pair_codes_reproduce_the_map(kind):
    if kind == "star":
        mesh = star_mesh(4)
        sp = fespace.NDSpace(mesh, 2)
    else:
        mesh, sp, _, _ = make_pair("hex", (4, 3, 5), 2, kind == "perturbed", scramble=kind == "scrambled")
    sd = sp.signed_elem_dofs().astype(np.int64)
    rc, codes = pair_codes(sd)
    assert rc == 0
    fe = NDHex(2)
    tens = np.array([rot_slot_to_tensor(2, s) for s in range(54)])
    dof = np.where(sd >= 0, sd, -1 - sd)[:, fe.loc[tens]]                 # [e][rotated slot]
    neg = ((sd < 0)[:, fe.loc[tens]]) != (fe.sgn[tens] < 0)[None, :]
    q = (codes & IDX).astype(np.int64)
    sw = (codes & SWAP) != 0
    for h, flag in ((0, NEG0), (1, NEG1)):
        want = 2 * q + np.where(sw, 1 - h, h)
        assert np.array_equal(dof[:, h::2], want)
        assert np.array_equal(neg[:, h::2], (codes & flag) != 0)
    if kind == "scrambled":
        assert sw.any() and (~sw).any() and ((codes & NEG0) != 0).any()
    if kind == "cube":
        # lexicographic vertex numbering: the pairs of the third component never carry a flag (the kernel's warp-uniform
        # fast path), those of the first do on (almost) every element (faces seen in a rotated frame)
        flagged = (codes & (SWAP | NEG0 | NEG1)) != 0
        assert not flagged[:, 18:].any() and flagged[:, :9].any()


def test_unpaired_numbering_is_rejected():
    mesh, sp, _, _ = make_pair("hex", (3, 3, 3), 2, False)
    sd = sp.signed_elem_dofs().astype(np.int64)
    perm = np.random.default_rng(0).permutation(sp.ndof)
    k = np.where(sd >= 0, sd, -1 - sd)
    sd2 = np.where(sd >= 0, perm[k], -1 - perm[k])
    rc, _ = pair_codes(sd2)
    assert rc == 1 and b"aligned dof pairs" in L.lib().hcurl_last_error()
