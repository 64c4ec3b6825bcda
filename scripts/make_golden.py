#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the reference's mesh files.

    python scripts/make_golden.py            (needs /root/reference/data; build container only)

The reference holds NO operator-level golden vectors for this path; what it does
hold are the meshes its integration tests run on.  They cannot travel to the GPU box (no
/root/reference there), so the arrays the FE space is built from are extracted here with the
oracle's readers and committed as compressed fixtures:

  beam_tet.npz          data/beam-tet.mesh       48 tets   -- TestAForm / TestHFormSource MMS mesh
                                                             (test/integration/test_aform_source.cpp:81)
  cylinder_hex_q2.npz   data/cylinder-hex-q2.gen 252 HEX27 -- TestHFormRod / TestEBFormRod / TestComplexAFormRod
                                                             (test/integration/test_hform_rod.cpp:40)
  team7.npz             data/team7.g             57 905 tets -- examples/team7/Main.cpp:154, TestComplexTeam7
  team4.npz             data/team4.g             127 074 tets -- BASELINE configs[2] (closed_coil / team4 order-2 A-formulation), with
                                                             team4_samples.npz: samples of the oracle's apply at order 2 (808 658 dofs)
  irises.npz            data/irises.g            8 346 HEX8 -- TestComplexIrisWaveguide, the reference's known-answer test
                                                             for the hex path (test/integration/test_complex_iris_wg.cpp:90,191-195;
                                                             `python scripts/make_golden.py irises` writes only this one and
                                                             irises_vectors.npz)

plus `oracle_vectors.npz`: y = A x of the ORACLE on those meshes for fixed seeded inputs.  These are
regression vectors of the restatement (they pin the oracle against accidental change and let the
GPU tests run without re-assembling team7); they are NOT outputs of MFEM -- see DESIGN.md
"parity pinning".
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import assemble as oa  # noqa: E402
from oracle import mesh as om  # noqa: E402
from oracle import space as osp  # noqa: E402

DATA = "/root/reference/data"
OUT = os.path.join(ROOT, "tests", "golden")


def save_mesh(name, m):
    # only the vertices the elements use (HEX27 files carry mid nodes in the vertex table)
    d = {"kind": np.array(m.kind), "verts": m.verts, "elems": m.elems.astype(np.int32), "attr": m.attr.astype(np.int32)}
    if m.geom_nodes is not None:
        d["geom_nodes"] = m.geom_nodes
    if m.bdr is not None:
        d["bdr"] = m.bdr.astype(np.int32)
        d["bdr_attr"] = m.bdr_attr.astype(np.int32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def coef_by_attr(attr, lo, hi):
    """deterministic per-attribute coefficient values in [lo, hi]"""
    ids = np.unique(attr)
    vals = {a: lo + (hi - lo) * ((7 * i + 3) % 11) / 10.0 for i, a in enumerate(ids)}
    return np.array([vals[a] for a in attr])


def irises():
    """mesh fixture + regression vectors of the oracle for the volume operator of the iris-waveguide KAT
    (K(1/mu0) + M(-omega^2 eps0), ND1) -- the KAT itself is recomputed by tests/test_oracle_iris_kat.py"""
    from oracle import iris_kat
    m = om.read_exodus(os.path.join(DATA, "irises.g"))
    save_mesh("irises", m)
    s = osp.NDSpace(m, 1)
    om2 = (2 * np.pi * iris_kat.FREQ) ** 2
    A = oa.assemble(s, 1.0 / iris_kat.MU0, -om2 * iris_kat.EPS0)
    x = np.random.default_rng(12345).standard_normal(s.ndof)
    np.savez_compressed(os.path.join(OUT, "irises_vectors.npz"), irises_p1_y=A @ x, irises_p1_diag=A.diagonal(), irises_p1_ndof=np.array(s.ndof))
    print("irises", s.ndof, "dofs", A.nnz, "nnz")


def team4():
    """BASELINE.json configs[2]: data/team4.g (127 074 tets), order-2 Nedelec = 808 658 dofs.  The oracle needs two minutes
    for this space and matrix, and its output vector would be 6.5 MB, so the fixture holds the mesh and SAMPLES of the
    oracle's y = (K(alpha) + M(beta)) x and diagonal for the seeded x of the other fixtures: 4096 entries at seeded
    positions plus x.y and |y|^2 (`python scripts/make_golden.py team4`)."""
    m = om.read_exodus(os.path.join(DATA, "team4.g"))
    save_mesh("team4", m)
    s = osp.NDSpace(m, 2)
    alpha, beta = coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
    A = oa.assemble(s, alpha, beta)
    x = np.random.default_rng(12345).standard_normal(s.ndof)
    y = A @ x
    idx = np.sort(np.random.default_rng(777).choice(s.ndof, 4096, replace=False))
    np.savez_compressed(os.path.join(OUT, "team4_samples.npz"), ndof=np.array(s.ndof), idx=idx, y=y[idx], diag=A.diagonal()[idx],
                        xdoty=np.array(x @ y), ynorm2=np.array(y @ y), ymax=np.array(np.abs(y).max()))
    print("team4", s.ndof, "dofs", A.nnz, "nnz")


def main():
    os.makedirs(OUT, exist_ok=True)
    if sys.argv[1:] == ["irises"]:
        return irises()
    if sys.argv[1:] == ["team4"]:
        return team4()
    irises()
    meshes = {
        "beam_tet": om.read_mfem_mesh(os.path.join(DATA, "beam-tet.mesh")),
        "cylinder_hex_q2": om.read_exodus(os.path.join(DATA, "cylinder-hex-q2.gen")),
        "team7": om.read_exodus(os.path.join(DATA, "team7.g")),
    }
    for k, m in meshes.items():
        save_mesh(k, m)
    vec = {}
    for name, orders in (("beam_tet", (1, 2, 3)), ("cylinder_hex_q2", (1, 2, 3)), ("team7", (1,))):
        m = meshes[name]
        alpha, beta = coef_by_attr(m.attr, 0.5, 2.0), coef_by_attr(m.attr, 1.0, 1.0e3)
        for p in orders:
            s = osp.NDSpace(m, p)
            A = oa.assemble(s, alpha, beta)
            x = np.random.default_rng(12345).standard_normal(s.ndof)
            vec["%s_p%d_y" % (name, p)] = A @ x
            vec["%s_p%d_diag" % (name, p)] = A.diagonal()
            vec["%s_p%d_ndof" % (name, p)] = np.array(s.ndof)
            print(name, p, s.ndof, "dofs", A.nnz, "nnz")
    np.savez_compressed(os.path.join(OUT, "oracle_vectors.npz"), **vec)


if __name__ == "__main__":
    main()
