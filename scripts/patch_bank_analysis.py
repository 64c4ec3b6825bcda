"""Offline model of the shared-memory wavefronts of the patch kernel's two irregular phases, from the
host-built patch blobs (no GPU needed):
  pull  element side reads Xs[lidx[s][el]] (lanes = consecutive el; equal addresses broadcast)
  sums  dof side reads E at the byte offsets of each dof's slots (lanes = consecutive dofs of a width class)
8-byte accesses are served per half-warp; wavefronts = max distinct addresses falling in one of 16 banks."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_host_patches import build, decode_patch, make


def half_wavefronts(addr):
    """addr: word (8-byte) addresses of one half-warp (distinct or equal) -> wavefronts"""
    a = np.unique(addr)
    return np.bincount(a % 16, minlength=16).max() if len(a) else 0


def analyse(n=16, fe=64, nsample=24, nd=54):
    m, sp, code, cen, idx, sgn = make("hex", (n, n, n), 2, False)
    perm, off, blob, st = build(sp, fe, code, cen)
    npatch = len(off) - 1
    rng = np.random.default_rng(0)
    pull = ideal_pull = sums = ideal_sums = 0
    ks = rng.choice(npatch, min(nsample, npatch), replace=False)
    for k in ks:
        dof, groups, ov, lidx, n_real = decode_patch(k, off, blob, nd, fe)
        pos = lidx & 0x1FFF
        for s in range(nd):
            for h0 in range(0, n_real, 16):
                pull += half_wavefronts(pos[s, h0:min(n_real, h0 + 16)]); ideal_pull += 1
        # the kernel walks each width class separately
        for w in range(3):
            ent = groups[w][3]
            for h0 in range(0, len(ent), 16):
                for kk in range(ent.shape[1]):
                    a = ent[h0:h0 + 16, kk]; a = a[a != 0xFFFF] // 8
                    if len(a):
                        sums += half_wavefronts(a); ideal_sums += 1
    ns = len(ks)
    print("n=%d fe=%d: wavefronts per patch: pull %.0f (ideal %.0f, x%.2f)   dof sums %.0f (ideal %.0f, x%.2f)" % (
        n, fe, pull / ns, ideal_pull / ns, pull / ideal_pull, sums / ns, ideal_sums / ns, sums / ideal_sums))
    return pull / ns, sums / ns


if __name__ == "__main__":
    for fe in (32, 64):
        analyse(16, fe)
