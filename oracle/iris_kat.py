"""Known-answer test of the reference for the HEX path: TestComplexIrisWaveguide (ORACLE; test infrastructure).

/root/reference/test/integration/test_complex_iris_wg.cpp: ND1 hexes on data/irises.g, ComplexEFormulation at 9.3 GHz,
PEC walls (side set 1), rectangular-waveguide TE10 port conditions on side sets 2 (input) and 3 (output), direct solve,
and the two numbers the test requires (:191-195):

    max |E_real| = 4896.771 +- 0.001,   max |E_imag| = 5357.650 +- 0.001     (GridFunction::ComputeMaxError vs 0)

Restated pieces, with where they come from:
  * operator  A = K(1/mu0) + M(-omega^2 eps0) + i [ M(omega sigma = 0) + (k_z / mu0) M_port ]
    (/root/reference/src/formulations/ComplexMaxwell/complex_maxwell_formulation.cpp:181-191, RegisterCoefficients :67-123;
    HERMITIAN convention, complex_maxwell_formulation.hpp:82) -- K, M are the oracle's element matrices (assemble.py);
  * port terms (/root/reference/src/boundary_conditions/Robin/rwte10_port_rbc.cpp:6-90): boundary VectorFEMassIntegrator
    with coefficient k_z / mu0 on both ports (imaginary part), VectorFEBoundaryTangentLFIntegrator of
    u_re = -2 k_z Im(E10) / mu0, u_im = 2 k_z Re(E10) / mu0 on the input port, E10 as coded there (including its
    component placement E = (e e_hat[1], e e_hat[2], e e_hat[0]));
  * [MFEM-upstream] VectorFEBoundaryTangentLFIntegrator: L(v) = int (n x f) . v, rule of order 2*order + 0 (2 x 2 Gauss
    on ND1 quad faces); boundary mass: exact on planar rectangular faces; ComputeMaxError: maximum over the elements'
    Gauss points of order 2*order + 3 (3 x 3 x 3) of the Euclidean norm of the field.
The traces of the 3-D element basis are used for the face integrals (the tangential trace of an ND hex function is the ND
quad function of the face), so no separate face element is restated."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import assemble as oa
from . import space as osp
from .quadrature import gauss_legendre

EPS0 = 8.8541878176e-12
MU0 = 4.0e-7 * np.pi
FREQ = 9.3e9
A1 = np.array([0.0, 22.86e-3, 0.0])   # port length vector
A2 = np.array([0.0, 0.0, 10.16e-3])   # port width vector

# reference coordinates of the hex faces: (fixed axis, fixed value), in HEX_FACES order
FACE_PLANE = [(2, 0.0), (1, 0.0), (0, 1.0), (1, 1.0), (0, 0.0), (2, 1.0)]


class Port:
    """RWTE10PortRBC's constants (rwte10_port_rbc.cpp:6-30)"""

    def __init__(self):
        self.omega = 2 * np.pi * FREQ
        a3 = np.cross(A1, A2)
        a2xa3 = np.cross(A2, a3)
        v = A1 @ a2xa3
        kc = np.pi / np.linalg.norm(A1)
        k0 = self.omega * np.sqrt(EPS0 * MU0)
        self.kz = np.sqrt(k0 * k0 - kc * kc)              # _k.imag()
        self.k_a = a2xa3 * np.pi / v
        self.k_c = a3 * self.kz / np.linalg.norm(a3)

    def e10(self, X):
        """RWTE10 (:49-62), complex (n, 3)"""
        e_hat = np.cross(self.k_c, self.k_a)
        e_hat /= np.linalg.norm(e_hat)
        e0 = np.sqrt(2 * self.omega * MU0 / (np.linalg.norm(A1) * np.linalg.norm(A2) * self.kz))
        e_mag = e0 * np.sin(X @ self.k_a) * np.exp(-1j * (X @ self.k_c))
        return np.stack([e_mag * e_hat[1], e_mag * e_hat[2], e_mag * e_hat[0]], axis=1)

    def u_real(self, X):
        return -2 * self.kz * self.e10(X).imag / MU0

    def u_imag(self, X):
        return 2 * self.kz * self.e10(X).real / MU0


def boundary_face_elements(mesh, topo):
    """for every boundary quad of mesh.bdr: (element, local face)"""
    out = []
    for fv in mesh.bdr:
        f = topo.face_id[tuple(sorted(int(v) for v in fv))]
        (e,) = topo.face_elems[f]
        out.append((e, list(topo.el_faces[e]).index(f)))
    return np.array(out)


def face_terms(space, elems, lface, npts, coef=None, func=None):
    """Per boundary face (all on local face `lface` of their element): the tangential mass matrix
    coef * int u_t . v_t dS and / or the load vector int (n x f) . v dS, in the element's LOCAL dofs, by an
    npts x npts Gauss rule; n is the outward unit normal."""
    fe = space.fe
    ax, val = FACE_PLANE[lface]
    free = [d for d in range(3) if d != ax]
    g, w = gauss_legendre(npts)
    pts = np.zeros((npts * npts, 3))
    wts = np.zeros(npts * npts)
    for i in range(npts):
        for j in range(npts):
            pts[i * npts + j, ax] = val
            pts[i * npts + j, free[0]] = g[i]
            pts[i * npts + j, free[1]] = g[j]
            wts[i * npts + j] = w[i] * w[j]
    V, _ = fe.shape_and_curl(pts)
    X = space.mesh.elem_coords()[elems]
    J, xq = oa.hex8_jacobians(X, pts)                                # (nf, nq, 3, 3)
    nvec = np.cross(J[..., :, free[0]], J[..., :, free[1]])          # x_s x x_t, |.| = dS / (ds dt)
    dS = np.linalg.norm(nvec, axis=-1)
    n = nvec / dS[..., None]
    outward = 1.0 if val == 1.0 else -1.0                            # d x / d(ref axis) points out of the element on the +1 face
    sgn = np.sign(np.einsum("fqi,fqi->fq", n, J[..., :, ax])) * outward
    n = n * sgn[..., None]
    U = np.einsum("fqji,qnj->fqni", np.linalg.inv(J), V)             # J^-T v_hat
    Ut = U - np.einsum("fqni,fqi->fqn", U, n)[..., None] * n[:, :, None, :]
    M = b = None
    if coef is not None:
        M = coef * np.einsum("q,fq,fqni,fqmi->fnm", wts, dS, Ut, Ut)
    if func is not None:
        F = np.asarray(func(xq.reshape(-1, 3))).reshape(xq.shape)
        b = np.einsum("q,fq,fqi,fqni->fn", wts, dS, np.cross(n, F), U)
    return M, b


def solve(mesh, lf_points=None, p=1):
    """Returns (max |E_real|, max |E_imag|, space, x) -- the two numbers of test_complex_iris_wg.cpp:191-195 for p = 1
    (the reference's ND_3D_P1); p > 1 runs the same problem at a higher order with MFEM's order-dependent rules (a
    consistency check of the higher-order hex operator on a real mesh: the answer must stay close)."""
    lf_points = lf_points if lf_points is not None else (2 * p) // 2 + 1     # order 2p + 0
    space = osp.NDSpace(mesh, p)
    topo = space.topo
    port = Port()
    n = space.ndof
    Ar = oa.assemble(space, 1.0 / MU0, -port.omega ** 2 * EPS0)
    fel = boundary_face_elements(mesh, topo)
    Ai = sp.csr_matrix((n, n))
    br, bi = np.zeros(n), np.zeros(n)
    for attr in (2, 3):
        for lf in range(6):
            sel = (mesh.bdr_attr == attr) & (fel[:, 1] == lf)
            if not sel.any():
                continue
            el = fel[sel, 0]
            M, vr = face_terms(space, el, lf, p + 1, coef=port.kz / MU0, func=None)   # exact on planar rectangular faces
            s = space.elem_sign[el].astype(float)
            d = space.elem_dofs[el]
            G = M * s[:, :, None] * s[:, None, :]
            nd = d.shape[1]
            Ai = Ai + sp.coo_matrix((G.ravel(), (np.repeat(d[:, :, None], nd, 2).ravel(), np.repeat(d[:, None, :], nd, 1).ravel())),
                                    shape=(n, n)).tocsr()
            if attr == 2:
                _, vr = face_terms(space, el, lf, lf_points, func=port.u_real)
                _, vi = face_terms(space, el, lf, lf_points, func=port.u_imag)
                np.add.at(br, d.ravel(), (vr * s).ravel())
                np.add.at(bi, d.ravel(), (vi * s).ravel())
    pec = {topo.face_id[tuple(sorted(int(v) for v in fv))] for fv, a in zip(mesh.bdr, mesh.bdr_attr) if a == 1}
    ess = space.boundary_dofs(lambda f: f in pec)
    keep = np.ones(n, bool)
    keep[ess] = False
    idx = np.nonzero(keep)[0]
    Z = (Ar + 1j * Ai).tocsr()[idx][:, idx].tocsc()
    x = np.zeros(n, complex)
    x[idx] = spla.splu(Z).solve((br + 1j * bi)[idx])
    # ComputeMaxError against zero: Gauss points of order 2*p + 3 (p = 1: 5 -> 3 per direction)
    g, _ = gauss_legendre((2 * p + 3) // 2 + 1)
    pts = np.array([[a, b, c] for c in g for b in g for a in g])
    V, _ = space.fe.shape_and_curl(pts)
    J, _ = oa.hex8_jacobians(mesh.elem_coords(), pts)
    xe = x[space.elem_dofs] * space.elem_sign
    E = np.einsum("eqji,en,qnj->eqi", np.linalg.inv(J), xe, V)
    return float(np.linalg.norm(E.real, axis=-1).max()), float(np.linalg.norm(E.imag, axis=-1).max()), space, x
