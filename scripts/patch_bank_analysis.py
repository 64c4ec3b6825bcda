"""Offline model of the shared-memory wavefronts / global sectors of the per-dof gather (A) and
scatter (C) phases of the patch kernel, from the host-built patch blobs (no GPU needed)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from hephaestus_b200 import fespace
from test_host_patches import build, decode_patch, make

def analyse(n=16, fe=64, nsample=40):
    m, sp, code, cen, idx, sgn = make("hex", (n, n, n), 2, False)
    # slot order used by the kernel does not matter for the statistics of random scatter, but use the real one
    perm, off, blob, st = build(sp, fe, code, cen)
    npatch = len(off) - 1
    rng = np.random.default_rng(0)
    tot = dict(wf=0, ideal=0, sectors=0, reqs=0)
    for k in rng.choice(npatch, min(nsample, npatch), replace=False):
        dof, groups, ov = decode_patch(k, off, blob)
        U = len(dof)
        ent = np.full((U, 4), -1, np.int64)
        for priv, u0, pd, e in groups:
            ent[u0:u0 + len(pd), :e.shape[1]] = np.where(e == 0xFFFF, -1, e & 0x7FFF)
        for w0 in range(0, U, 32):
            lanes = np.arange(w0, min(U, w0 + 32))
            tot["sectors"] += len(np.unique(dof[lanes] // 4)); tot["reqs"] += 1
            for kk in range(4):
                a = ent[lanes, kk]
                if (a >= 0).sum() == 0: continue
                for h in (lanes[:16], lanes[16:]):
                    aa = ent[h, kk]; aa = aa[aa >= 0]
                    if len(aa) == 0: continue
                    tot["wf"] += np.bincount(aa % 16, minlength=16).max()
                    tot["ideal"] += 1
    ns = min(nsample, npatch)
    print("n=%d fe=%d: per patch: slot wavefronts %.0f (ideal %.0f, x%.2f) per phase; gather sectors/request %.1f (%d requests)" % (
        n, fe, tot["wf"] / ns, tot["ideal"] / ns, tot["wf"] / tot["ideal"], tot["sectors"] / tot["reqs"], tot["reqs"] / ns))

if __name__ == "__main__":
    for fe in (32, 64):
        analyse(16, fe)
