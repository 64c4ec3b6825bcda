"""Shared builders for the parity tests: the same seeded inputs go to the CUDA
path (through the C ABI) and to the CPU oracle."""
import numpy as np

from hephaestus_b200 import fespace
from oracle import mesh as om
from oracle import space as osp


def make_pair(kind, n, p, perturb=True, dims=(1.0, 0.7, 1.3), scramble=False):
    """(product mesh/space, oracle mesh/space) built independently."""
    nx, ny, nz = (n, n, n) if np.isscalar(n) else n
    m = fespace.cartesian_hex_mesh(nx, ny, nz, *dims)
    mo = om.cartesian_hex(nx, ny, nz, *dims)
    if perturb:
        m, mo = fespace.perturb_interior(m), om.perturb_interior(mo)
    if kind == "tet":
        m, mo = fespace.hex_to_tets(m), om.hex_to_tets(mo)
    if scramble:
        perm = np.random.default_rng(5).permutation(len(m.verts))
        inv = np.empty_like(perm)
        inv[perm] = np.arange(len(perm))
        m = fespace.Mesh(kind, m.verts[perm], inv[m.elems], m.attr)
        mo = om.Mesh(kind, mo.verts[perm], inv[mo.elems], mo.attr)
    return m, fespace.NDSpace(m, p), mo, osp.NDSpace(mo, p)


def coefficients(mesh, seed=3):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.5, 2.0, mesh.n_elem), rng.uniform(0.5, 2.0, mesh.n_elem)


def rel_err(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def star_mesh(layers=3, cls=None):
    """Six 60-degree rhombic prisms around a central edge (all elements affine, the central
    edges have valence 6 > ELL width), stacked in z."""
    ang = np.pi / 3 * np.arange(6)
    ring1 = np.stack([np.cos(ang), np.sin(ang)], 1)
    ring2 = ring1 + np.roll(ring1, -1, axis=0)
    pts2 = np.concatenate([[[0.0, 0.0]], ring1, ring2])
    nl = len(pts2)
    verts = np.concatenate([np.column_stack([pts2, np.full(nl, float(z))]) for z in range(layers + 1)])
    elems = []
    for z in range(layers):
        for k in range(6):
            q = [0, 1 + k, 7 + k, 1 + (k + 1) % 6]  # counter-clockwise rhombus
            elems.append([z * nl + v for v in q] + [(z + 1) * nl + v for v in q])
    return (cls or fespace.Mesh)("hex", verts, np.array(elems))




# ---- geometric identification of dofs (independent of any numbering convention) ---------------
def dof_descriptors(omesh, ospace):
    """Physical node (ndof,3) and oriented tangent (ndof,3) of every global dof functional of an
    oracle hex NDSpace: dof d acts as v -> T_d . v(X_d).  Two spaces on the same geometry describe the
    same dof when the nodes coincide; the relative sign of the basis functions is sign(T . T')."""
    from oracle.assemble import hex8_jacobians
    fe = ospace.fe
    nodes, tang = fe.dof_nodes()                      # reference nodes / unit tangents, MFEM local order
    X = omesh.elem_coords()
    J, xq = hex8_jacobians(X, nodes)                 # (ne, nd, 3, 3), (ne, nd, 3)
    T = np.einsum("eqij,qj->eqi", J, tang) * ospace.elem_sign[:, :, None]
    P = np.zeros((ospace.ndof, 3))
    D = np.zeros((ospace.ndof, 3))
    P[ospace.elem_dofs.ravel()] = xq.reshape(-1, 3)
    D[ospace.elem_dofs.ravel()] = T.reshape(-1, 3)
    return P, D


def match_dofs(P_a, D_a, P_b, D_b, decimals=9):
    """For every dof of space a: (index in b, relative sign).  Nodes must coincide."""
    key = lambda P: [tuple(r) for r in np.round(P, decimals) + 0.0]  # noqa: E731
    where = {k: i for i, k in enumerate(key(P_b))}
    idx = np.array([where[k] for k in key(P_a)])
    sgn = np.sign(np.einsum("ij,ij->i", D_a, D_b[idx]))
    assert np.all(sgn != 0)
    return idx, sgn
