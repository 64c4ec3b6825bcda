"""Host-side C++ ExodusII reader (hephaestus_b200/csrc/exodus.cpp) against the Python reader (scipy's netCDF) on the
reference's own mesh files, and on a small CDF-1 file written here (runs without a GPU).  The reference's data
directory only exists in the build container; those cases are skipped elsewhere."""
import ctypes as C
import os

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace

REF_DATA = "/root/reference/data"


def cpp_read(path):
    lib = L.lib()
    h = L.check_ptr(lib.hcurl_host_exodus_open(path.encode()), "exodus_open")
    nn, ne, ns = C.c_int64(), C.c_int64(), C.c_int64()
    geom, npe = C.c_int(), C.c_int()
    L.check(lib.hcurl_host_exodus_counts(h, C.byref(nn), C.byref(ne), C.byref(geom), C.byref(npe), C.byref(ns)), "counts")
    nfv = 4 if geom.value == L.HEX else 3
    xyz = np.zeros((nn.value, 3))
    conn = np.zeros((ne.value, npe.value), np.int32)
    attr = np.zeros(ne.value, np.int32)
    bdr = np.zeros((ns.value, nfv), np.int32)
    battr = np.zeros(ns.value, np.int32)
    L.check(lib.hcurl_host_exodus_read(h, L.ptr(xyz), L.ptr(conn), L.ptr(attr), L.ptr(bdr) if ns.value else None,
                                       L.ptr(battr) if ns.value else None), "read")
    lib.hcurl_host_exodus_close(h)
    return dict(kind="hex" if geom.value == L.HEX else "tet", xyz=xyz, conn=conn, attr=attr, bdr=bdr, battr=battr)


@pytest.mark.parametrize("name", ["team7.g", "cylinder-hex-q2.gen", "team4.g", "irises.g", "inhomogeneous_beam.g", "solenoid.g"])
def test_reference_files_match_python_reader(name):
    path = os.path.join(REF_DATA, name)
    if not os.path.exists(path):
        pytest.skip("reference data not present")
    try:
        ref = fespace.read_exodus(path)
    except Exception as exc:   # element types the Python reader does not handle either
        lib = L.lib()
        assert not lib.hcurl_host_exodus_open(path.encode()), "C++ reader accepted what the Python reader rejects: %s" % exc
        return
    got = cpp_read(path)
    assert got["kind"] == ref.kind
    assert np.array_equal(got["xyz"], ref.verts)
    nc = 8 if ref.kind == "hex" else 4
    assert np.array_equal(got["conn"][:, :nc], ref.elems)
    assert np.array_equal(got["attr"], ref.attr)
    if ref.bdr is not None:
        assert np.array_equal(got["bdr"], ref.bdr) and np.array_equal(got["battr"], ref.bdr_attr)
    else:
        assert len(got["bdr"]) == 0
    if ref.geom_nodes is not None:       # HEX27: the 27 geometry nodes per element, in the Python reader's tensor order
        assert got["conn"].shape[1] == 27
        assert np.array_equal(fespace._hex27_to_tensor(got["xyz"][got["conn"]]), ref.geom_nodes)


def test_small_cdf1_file(tmp_path):
    """A two-block TETRA4 mesh with one side set written with scipy (CDF-1, 32-bit offsets)."""
    from scipy.io import netcdf_file
    path = str(tmp_path / "two_tets.e")
    f = netcdf_file(path, "w", version=1)
    for nm, n in (("num_nodes", 5), ("num_el_blk", 2), ("num_el_in_blk1", 1), ("num_nod_per_el1", 4), ("num_el_in_blk2", 1),
                  ("num_nod_per_el2", 4), ("num_side_sets", 1), ("num_side_ss1", 2), ("num_dim", 3)):
        f.createDimension(nm, n)
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1.0]])
    for d, a in enumerate("xyz"):
        v = f.createVariable("coord" + a, "d", ("num_nodes",))
        v[:] = xyz[:, d]
    v = f.createVariable("eb_prop1", "i", ("num_el_blk",)); v[:] = [7, 9]
    c1 = f.createVariable("connect1", "i", ("num_el_in_blk1", "num_nod_per_el1")); c1[:] = [[1, 2, 3, 4]]; c1.elem_type = "TETRA4"
    c2 = f.createVariable("connect2", "i", ("num_el_in_blk2", "num_nod_per_el2")); c2[:] = [[2, 3, 4, 5]]; c2.elem_type = "tetra"
    v = f.createVariable("ss_prop1", "i", ("num_side_sets",)); v[:] = [3]
    v = f.createVariable("elem_ss1", "i", ("num_side_ss1",)); v[:] = [1, 2]
    v = f.createVariable("side_ss1", "i", ("num_side_ss1",)); v[:] = [4, 1]
    f.close()
    got = cpp_read(path)
    ref = fespace.read_exodus(path)
    assert got["kind"] == "tet" and np.array_equal(got["xyz"], xyz)
    assert np.array_equal(got["conn"], [[0, 1, 2, 3], [1, 2, 3, 4]]) and np.array_equal(got["attr"], [7, 9])
    assert np.array_equal(got["bdr"], ref.bdr) and np.array_equal(got["battr"], [3, 3])


def test_rejects_non_netcdf(tmp_path):
    p = tmp_path / "x.g"
    p.write_bytes(b"not a mesh")
    lib = L.lib()
    assert not lib.hcurl_host_exodus_open(str(p).encode())
    assert b"not a NetCDF" in lib.hcurl_last_error()
    assert not lib.hcurl_host_exodus_open(b"/nonexistent/file.g")
    assert b"cannot open" in lib.hcurl_last_error()
