"""Host-side C++ ND hex space (hephaestus_b200/csrc/ndspace.cpp) against the Python generator and the oracle:
bit-identical signed element -> dof maps, counts and boundary dofs (runs without a GPU)."""
import os

import numpy as np
import pytest

from hephaestus_b200 import _lib as L
from hephaestus_b200 import fespace
from oracle import mesh as om, space as osp

from helpers import star_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cpp_space(mesh, p):
    lib = L.lib()
    ev = np.ascontiguousarray(mesh.elems, dtype=np.int32)
    create = lib.hcurl_host_ndspace_create_hex if mesh.kind == "hex" else lib.hcurl_host_ndspace_create_tet
    h = L.check_ptr(create(p, len(ev), L.ptr(ev)), "ndspace")
    nd, ndof = lib.hcurl_host_ndspace_nd(h), lib.hcurl_host_ndspace_ndof(h)
    dofs = np.zeros((len(ev), nd), np.int32)
    L.check(lib.hcurl_host_ndspace_elem_dofs(h, L.ptr(dofs)), "elem_dofs")
    nb = lib.hcurl_host_ndspace_boundary_dofs(h, None)
    bd = np.zeros(nb, np.int32)
    assert lib.hcurl_host_ndspace_boundary_dofs(h, L.ptr(bd)) == nb
    fb = np.zeros((len(ev), 6, 4) if mesh.kind == "hex" else (len(ev), 4, 3), np.int32)
    L.check(lib.hcurl_host_ndspace_elem_face_base(h, L.ptr(fb)), "face_base")
    out = dict(nd=nd, ndof=ndof, n_edges=lib.hcurl_host_ndspace_n_edges(h), n_faces=lib.hcurl_host_ndspace_n_faces(h), dofs=dofs, bdr=bd,
               face_base=fb)
    lib.hcurl_host_ndspace_destroy(h)
    return out


def scrambled(mesh, seed=5):
    perm = np.random.default_rng(seed).permutation(len(mesh.verts))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return fespace.Mesh("hex", mesh.verts[perm], inv[mesh.elems], mesh.attr)


@pytest.mark.parametrize("p", [1, 2, 3, 4])
@pytest.mark.parametrize("kind", ["cube", "scrambled", "star"])
def test_matches_python_generator_and_oracle(kind, p):
    mesh = star_mesh(3) if kind == "star" else fespace.cartesian_hex_mesh(4, 3, 5)
    if kind == "scrambled":
        mesh = scrambled(mesh)     # random vertex numbering: every edge / face orientation code occurs
    ref = fespace.NDSpace(mesh, p)
    got = cpp_space(mesh, p)
    assert (got["nd"], got["ndof"], got["n_edges"], got["n_faces"]) == (ref.nd, ref.ndof, ref.n_edges, ref.n_faces)
    assert np.array_equal(got["dofs"], ref.signed_elem_dofs())
    assert np.array_equal(got["bdr"], ref.boundary_dofs())
    omesh = om.Mesh("hex", mesh.verts, mesh.elems, mesh.attr)
    assert np.array_equal(got["dofs"], osp.NDSpace(omesh, p).signed_elem_dofs())
    if kind == "scrambled":
        assert (got["dofs"] < 0).any() and (got["dofs"] >= 0).any()


@pytest.mark.parametrize("p", [1, 2, 3])
def test_reference_mesh_cylinder_hex(p):
    """The reference's hex mesh data/cylinder-hex-q2.gen (golden fixture of its connectivity): 6 882 ND2 dofs
    (SURVEY.md App. C), same map as the Python generator."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "cylinder_hex_q2.npz"))
    mesh = fespace.Mesh("hex", z["verts"], z["elems"], z["attr"])
    got = cpp_space(mesh, p)
    ref = fespace.NDSpace(mesh, p)
    if p == 2:
        assert got["ndof"] == 6882
    assert got["ndof"] == ref.ndof
    assert np.array_equal(got["dofs"], ref.signed_elem_dofs())
    assert np.array_equal(got["bdr"], ref.boundary_dofs())


def test_bad_arguments():
    lib = L.lib()
    ev = np.zeros((1, 8), np.int32)
    assert not lib.hcurl_host_ndspace_create_hex(7, 1, L.ptr(ev))
    assert b"bad arguments" in lib.hcurl_last_error()
    ev[0, 3] = -2
    assert not lib.hcurl_host_ndspace_create_hex(1, 1, L.ptr(ev))
    assert b"negative" in lib.hcurl_last_error()


@pytest.mark.parametrize("kind", ["hex", "tet"])
def test_python_wrapper_is_a_drop_in(kind):
    mesh = scrambled(fespace.cartesian_hex_mesh(3, 4, 2))
    if kind == "tet":
        mesh = fespace.hex_to_tets(mesh)
    a, b = fespace.NativeNDSpace(mesh, 2), fespace.NDSpace(mesh, 2)
    assert (a.nd, a.ndof, a.p) == (b.nd, b.ndof, b.p)
    assert np.array_equal(a.signed_elem_dofs(), b.signed_elem_dofs())
    assert np.array_equal(a.boundary_dofs(), b.boundary_dofs())
    if kind == "tet":
        assert np.array_equal(a.elem_face_base, b.elem_face_base) and a.elem_face_base.dtype == np.int32
    else:
        assert a.elem_face_base is None


@pytest.mark.parametrize("p", [1, 2, 3])
@pytest.mark.parametrize("src", ["kuhn", "beam_tet", "team7"])
def test_tets_match_python_generator(src, p):
    """Tetrahedra: same edge / face enumeration, dofs and per-face base vertex order (what the triangle-face
    DofTransformation needs) as the Python generator, on a Kuhn-split cube and on the reference's tet meshes."""
    if src == "kuhn":
        mesh = fespace.hex_to_tets(fespace.cartesian_hex_mesh(3, 2, 4))
    else:
        if src == "team7" and p > 1:
            pytest.skip("one order is enough for the large mesh")
        z = np.load(os.path.join(ROOT, "tests", "golden", src + ".npz"))
        mesh = fespace.Mesh("tet", z["verts"], z["elems"], z["attr"])
    ref = fespace.NDSpace(mesh, p)
    got = cpp_space(mesh, p)
    assert (got["nd"], got["ndof"], got["n_edges"], got["n_faces"]) == (ref.nd, ref.ndof, ref.n_edges, ref.n_faces)
    assert np.array_equal(got["dofs"], ref.signed_elem_dofs())
    assert np.array_equal(got["bdr"], ref.boundary_dofs())
    assert np.array_equal(got["face_base"], ref.elem_face_base)
    if src == "team7":
        assert got["ndof"] == 68547


@pytest.mark.parametrize("name,p", [("team7", 1), ("cylinder_hex_q2", 2), ("cylinder_hex_q2", 1), ("beam_tet", 2)])
def test_sideset_dofs_match_python(name, p):
    """Essential dofs of a set of boundary attributes (side sets of the reference's meshes) as the Python generator's
    sideset_dofs() selects them."""
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    if "bdr" not in z.files or z["bdr"].ndim != 2 or len(z["bdr"]) == 0:
        pytest.skip("fixture has no boundary elements")
    kind = str(z["kind"])
    mesh = fespace.Mesh(kind, z["verts"], z["elems"], z["attr"], None, z["bdr"], z["bdr_attr"])
    ref = fespace.NDSpace(mesh, p)
    attrs = np.unique(z["bdr_attr"])[:2]
    want = ref.sideset_dofs(attrs)
    faces = np.ascontiguousarray(z["bdr"][np.isin(z["bdr_attr"], attrs)], dtype=np.int32)
    lib = L.lib()
    ev = np.ascontiguousarray(mesh.elems, dtype=np.int32)
    create = lib.hcurl_host_ndspace_create_hex if kind == "hex" else lib.hcurl_host_ndspace_create_tet
    h = L.check_ptr(create(p, len(ev), L.ptr(ev)), "ndspace")
    n = lib.hcurl_host_ndspace_face_dofs(h, len(faces), L.ptr(faces), None)
    assert n == len(want)
    got = np.zeros(n, np.int32)
    assert lib.hcurl_host_ndspace_face_dofs(h, len(faces), L.ptr(faces), L.ptr(got)) == n
    assert np.array_equal(got, want)
    bad = faces[:1].copy()
    bad[0, 0] = int(ev.max())  # most likely not a face of the mesh
    if lib.hcurl_host_ndspace_face_dofs(h, 1, L.ptr(bad), None) < 0:
        assert b"not in the mesh" in lib.hcurl_last_error()
    lib.hcurl_host_ndspace_destroy(h)
