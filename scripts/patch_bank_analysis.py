"""Offline model of the shared-memory wavefronts of the patch kernel's two irregular phases, from the host-built patch
blobs (no GPU needed), for the layout the order-2 kernel actually uses: 16-byte units (dof pairs), 8 lanes per
128-byte wavefront.
  pull  element side reads Xs[lidx[s][el]] (lanes = consecutive elements; equal addresses broadcast)
  sums  dof side reads the result buffer at the byte offsets of each pair's slots (lanes = consecutive pairs of a
        width class, one read per slot rank)
wavefronts of an access = max number of distinct 16-byte addresses falling in one of the 8 bank groups."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hephaestus_b200 import fespace
from test_host_patches import build, decode_patch

NB = 8


def quarter_wavefronts(addr):
    a = np.unique(addr)
    return np.bincount(a % NB, minlength=NB).max() if len(a) else 0


def analyse(n=16, fe=32, nsample=24, padded=True):
    m = fespace.cartesian_hex_mesh(n, n, n)
    sp = fespace.NDSpace(m, 2)
    sd = sp.signed_elem_dofs().astype(np.int64)
    idx = np.where(sd >= 0, sd, -1 - sd)
    code = np.ascontiguousarray((np.sort(idx, axis=1)[:, ::2] // 2).astype(np.uint32))  # the 27 dof pairs of every element
    cen = np.ascontiguousarray(m.verts[m.elems].mean(axis=1))
    perm, off, blob, st = build(sp, fe, code, cen, stride=None if padded else fe, unit=16, n_units=sp.ndof // 2)
    npatch = len(off) - 1
    rng = np.random.default_rng(0)
    pull = ideal_pull = sums = ideal_sums = 0
    ks = rng.choice(npatch, min(nsample, npatch), replace=False)
    for k in ks:
        dof, groups, ov, lidx, n_real = decode_patch(k, off, blob, 27, fe)
        pos = lidx & 0x1FFF
        for s in range(27):
            for h0 in range(0, n_real, NB):
                pull += quarter_wavefronts(pos[s, h0:min(n_real, h0 + NB)]); ideal_pull += 1
        for u0, pd, priv, ent in groups:
            for h0 in range(0, len(ent), NB):
                for kk in range(ent.shape[1]):
                    a = ent[h0:h0 + NB, kk]; a = a[a != 0xFFFF] // 16
                    if len(a):
                        sums += quarter_wavefronts(a); ideal_sums += 1
    ns = len(ks)
    print("n=%d fe=%d %s rows: wavefronts per patch: pull %.0f (ideal %.0f, x%.2f)   dof sums %.0f (ideal %.0f, x%.2f)" % (
        n, fe, "padded" if padded else "unpadded", pull / ns, ideal_pull / ns, pull / ideal_pull, sums / ns, ideal_sums / ns, sums / ideal_sums))
    return pull / ns, sums / ns


if __name__ == "__main__":
    for fe, padded in ((32, True), (32, False), (64, True)):
        analyse(16, fe, padded=padded)
