"""examples/standalone_solve.c -- the C ABI used from plain C99 (ExodusII reader -> ND space -> operator -> PCG) --
compiles warning-free, reads a mesh and builds the space on the host, and without a GPU fails loudly at context
creation (no CPU fallback)."""
import os
import subprocess

import numpy as np
import pytest

from hephaestus_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_two_hex_exodus(path):
    from scipy.io import netcdf_file
    f = netcdf_file(path, "w", version=2)
    xyz = np.array([[x, y, z] for z in (0, 1) for y in (0, 1) for x in (0, 1, 2)], dtype=float)
    for nm, n in (("num_nodes", len(xyz)), ("num_el_blk", 1), ("num_el_in_blk1", 2), ("num_nod_per_el1", 8), ("num_side_sets", 1),
                  ("num_side_ss1", 1), ("num_dim", 3)):
        f.createDimension(nm, n)
    v = f.createVariable("coord", "d", ("num_dim", "num_nodes")); v[:] = xyz.T
    v = f.createVariable("eb_prop1", "i", ("num_el_blk",)); v[:] = [1]
    c = f.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1"))
    c[:] = [[1, 2, 5, 4, 7, 8, 11, 10], [2, 3, 6, 5, 8, 9, 12, 11]]
    c.elem_type = "HEX8"
    v = f.createVariable("ss_prop1", "i", ("num_side_sets",)); v[:] = [5]
    v = f.createVariable("elem_ss1", "i", ("num_side_ss1",)); v[:] = [1]
    v = f.createVariable("side_ss1", "i", ("num_side_ss1",)); v[:] = [4]
    f.close()


def test_standalone_c_driver(tmp_path, lib_built):
    exe = str(tmp_path / "standalone_solve")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "standalone_solve.c"), "-L" + os.path.dirname(L.LIB_PATH), "-lhcurl_b200", "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    mesh = str(tmp_path / "two_hex.e")
    write_two_hex_exodus(mesh)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.dirname(L.LIB_PATH) + ":/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([exe, mesh, "2", "5"], capture_output=True, text=True, env=env, timeout=300)
    # two hexes sharing a face: 20 edges, 11 faces -> 2*20 + 4*11 + 6*2 = 96 ND2 dofs; side set 5 = one face: 4 + 4*2 dofs
    assert "2 hexes, ND2: 96 dofs (12 essential)" in r.stdout, r.stdout + r.stderr
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert "converged 1" in r.stdout
    else:
        assert r.returncode == 1 and "no CUDA device" in r.stderr
