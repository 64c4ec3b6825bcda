"""Drives every host-only entry point of the library (patch builder, pair codes, ND space generator, ExodusII reader)\nthrough an AddressSanitizer / UBSan build of those sources: scripts/asan_host_check.sh builds the library and runs this."""
import ctypes as C, numpy as np, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hephaestus_b200 import fespace
lib = C.CDLL(os.environ.get('HCURL_ASAN_LIB', '/tmp/hcurl_asan/libhost_asan.so'))
P = C.c_void_p
def ptr(a): return a.ctypes.data_as(P)
lib.hcurl_host_patches_create.restype = P
lib.hcurl_host_patches_create.argtypes = [C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int64, P, P, P]
lib.hcurl_host_patches_destroy.argtypes = [P]
lib.hcurl_host_patches_count.restype = C.c_int64; lib.hcurl_host_patches_count.argtypes=[P]
lib.hcurl_host_patches_blob_words.restype = C.c_int64; lib.hcurl_host_patches_blob_words.argtypes=[P]
lib.hcurl_host_patches_copy.argtypes=[P,P,P,P]
lib.hcurl_host_hex_p2_pair_codes.argtypes=[C.c_int64,P,P]
for f in ("hcurl_host_ndspace_create_hex","hcurl_host_ndspace_create_tet"):
    getattr(lib,f).restype=P; getattr(lib,f).argtypes=[C.c_int,C.c_int64,P]
lib.hcurl_host_ndspace_destroy.argtypes=[P]
lib.hcurl_host_ndspace_nd.argtypes=[P]
lib.hcurl_host_ndspace_ndof.restype=C.c_int64; lib.hcurl_host_ndspace_ndof.argtypes=[P]
lib.hcurl_host_ndspace_elem_dofs.argtypes=[P,P]
lib.hcurl_host_ndspace_boundary_dofs.restype=C.c_int64; lib.hcurl_host_ndspace_boundary_dofs.argtypes=[P,P]
lib.hcurl_host_ndspace_face_dofs.restype=C.c_int64; lib.hcurl_host_ndspace_face_dofs.argtypes=[P,C.c_int64,P,P]
lib.hcurl_host_ndspace_elem_face_base.argtypes=[P,P]
lib.hcurl_host_exodus_open.restype=P; lib.hcurl_host_exodus_open.argtypes=[C.c_char_p]
lib.hcurl_host_exodus_close.argtypes=[P]
lib.hcurl_host_exodus_counts.argtypes=[P,P,P,P,P,P]
lib.hcurl_host_exodus_read.argtypes=[P,P,P,P,P,P]
# --- ND spaces
for kind in ("hex","tet"):
    for p in (1,2,3):
        m = fespace.cartesian_hex_mesh(5,4,3)
        if kind=="tet": m = fespace.hex_to_tets(m)
        ev = np.ascontiguousarray(m.elems, dtype=np.int32)
        h = (lib.hcurl_host_ndspace_create_hex if kind=="hex" else lib.hcurl_host_ndspace_create_tet)(p, len(ev), ptr(ev))
        nd = lib.hcurl_host_ndspace_nd(h); ndof = lib.hcurl_host_ndspace_ndof(h)
        d = np.zeros((len(ev), nd), np.int32); lib.hcurl_host_ndspace_elem_dofs(h, ptr(d))
        nb = lib.hcurl_host_ndspace_boundary_dofs(h, None); b = np.zeros(nb, np.int32); lib.hcurl_host_ndspace_boundary_dofs(h, ptr(b))
        fb = np.zeros((len(ev), 6 if kind=="hex" else 4, 4 if kind=="hex" else 3), np.int32); lib.hcurl_host_ndspace_elem_face_base(h, ptr(fb))
        faces = np.ascontiguousarray(fb[:7, 0], dtype=np.int32)
        n = lib.hcurl_host_ndspace_face_dofs(h, len(faces), ptr(faces), None); o = np.zeros(max(n,1), np.int32); lib.hcurl_host_ndspace_face_dofs(h, len(faces), ptr(faces), ptr(o))
        lib.hcurl_host_ndspace_destroy(h)
        ref = fespace.NDSpace(m, p)
        assert np.array_equal(d, ref.signed_elem_dofs()) and np.array_equal(b, ref.boundary_dofs())
# --- pair codes + patches (pairs, exchange flags) and scalar units
m = fespace.cartesian_hex_mesh(9,8,6); sp = fespace.NDSpace(m,2)
sd = np.ascontiguousarray(sp.signed_elem_dofs(), dtype=np.int32)
codes = np.zeros((len(sd),27), np.uint32); assert lib.hcurl_host_hex_p2_pair_codes(len(sd), ptr(sd), ptr(codes)) == 0
cen = np.ascontiguousarray(m.verts[m.elems].mean(axis=1))
flags = np.zeros(sp.ndof//2, np.uint8); flags[::17] = 1
for fe, stride in ((32,33),(32,32),(64,65),(128,129)):
    h = lib.hcurl_host_patches_create(27, len(sd), fe, stride, 16, sp.ndof//2, ptr(cen), ptr(codes), ptr(flags))
    assert h
    npatch = lib.hcurl_host_patches_count(h); nw = lib.hcurl_host_patches_blob_words(h)
    perm = np.zeros(npatch*fe, np.int32); off = np.zeros(npatch+1, np.uint32); blob = np.zeros(nw, np.uint32)
    lib.hcurl_host_patches_copy(h, ptr(perm), ptr(off), ptr(blob)); lib.hcurl_host_patches_destroy(h)
idx = np.where(sd>=0, sd, -1-sd); code8 = np.ascontiguousarray((idx | np.where(sd<0, 0x80000000, 0)).astype(np.uint32))
h = lib.hcurl_host_patches_create(54, len(sd), 32, 33, 8, sp.ndof, ptr(cen), ptr(code8), None); assert h; lib.hcurl_host_patches_destroy(h)
# --- split-blob parts, tensor-ordered slot codes (order-1 patch path, 8-byte units)
lib.hcurl_host_patches_parts.argtypes = [P, P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
lib.hcurl_host_hex_slot_codes.argtypes = [C.c_int, C.c_int64, P, P]
h = lib.hcurl_host_patches_create(27, len(sd), 32, 32, 16, sp.ndof//2, ptr(cen), ptr(codes), None); assert h
part = np.zeros(2*lib.hcurl_host_patches_count(h), np.uint32); md, ml = C.c_int64(), C.c_int64()
assert lib.hcurl_host_patches_parts(h, ptr(part), C.byref(md), C.byref(ml)) == 0 and md.value > 0 and ml.value > 0
lib.hcurl_host_patches_destroy(h)
sp1 = fespace.NDSpace(m,1); sd1 = np.ascontiguousarray(sp1.signed_elem_dofs(), dtype=np.int32)
c1 = np.zeros(sd1.shape, np.uint32); assert lib.hcurl_host_hex_slot_codes(1, len(sd1), ptr(sd1), ptr(c1)) == 0
for fe in (32, 128, 256):
    h = lib.hcurl_host_patches_create(12, len(sd1), fe, fe+1, 8, sp1.ndof, ptr(cen), ptr(c1), None); assert h; lib.hcurl_host_patches_destroy(h)
# --- rank-local spaces under an element partition
lib.hcurl_host_partition_create.restype = P; lib.hcurl_host_partition_create.argtypes = [C.c_int64, C.c_int, C.c_int64, P, P, C.c_int]
lib.hcurl_host_partition_destroy.argtypes = [P]; lib.hcurl_host_partition_copy.argtypes = [P]*10
lib.hcurl_host_partition_sizes.argtypes = [P] + [P]*6
er = np.ascontiguousarray((np.arange(len(sd)) * 7) % 3, dtype=np.int32)
for r in range(3):
    h = lib.hcurl_host_partition_create(len(sd), 54, sp.ndof, ptr(sd), ptr(er), r); assert h
    ne, nl, nt, nn, no, ng = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int(), C.c_int32(), C.c_int32()
    lib.hcurl_host_partition_sizes(h, C.byref(ne), C.byref(nl), C.byref(nt), C.byref(nn), C.byref(no), C.byref(ng))
    bufs = [np.zeros(ne.value, np.int64), np.zeros(nl.value, np.int32), np.zeros(ne.value*54, np.int32), np.zeros(nl.value, np.int32), np.zeros(nn.value, np.int32),
            np.zeros(nn.value+1, np.int32), np.zeros(max(no.value,1), np.int32), np.zeros(nn.value+1, np.int32), np.zeros(max(ng.value,1), np.int32)]
    lib.hcurl_host_partition_copy(h, *[ptr(b) for b in bufs]); lib.hcurl_host_partition_destroy(h)
# --- exodus
for f in ("team7.g","cylinder-hex-q2.gen","irises.g"):
    path = "/root/reference/data/"+f
    if not os.path.exists(path): continue
    h = lib.hcurl_host_exodus_open(path.encode()); assert h
    nn, ne, ns = C.c_int64(), C.c_int64(), C.c_int64(); g, npe = C.c_int(), C.c_int()
    lib.hcurl_host_exodus_counts(h, C.byref(nn), C.byref(ne), C.byref(g), C.byref(npe), C.byref(ns))
    xyz = np.zeros((nn.value,3)); conn = np.zeros((ne.value,npe.value), np.int32); attr = np.zeros(ne.value, np.int32)
    nfv = 4 if g.value==0 else 3
    bdr = np.zeros((max(ns.value,1), nfv), np.int32); ba = np.zeros(max(ns.value,1), np.int32)
    assert lib.hcurl_host_exodus_read(h, ptr(xyz), ptr(conn), ptr(attr), ptr(bdr), ptr(ba)) == 0
    lib.hcurl_host_exodus_close(h)
open('/tmp/hcurl_asan/garbage.g','wb').write(b'CDF\x02' + os.urandom(300))
h = lib.hcurl_host_exodus_open(b'/tmp/hcurl_asan/garbage.g'); assert not h
print("sanitizer drive: all host entry points exercised OK")
