"""Quadrature rules and 1D nodal bases (oracle; test infrastructure).

Restates what MFEM's ``IntegrationRules`` / ``Poly_1D`` provide to the
integrators named in /root/reference/src/kernels/curl_curl_kernel.cpp:23 and
vector_fe_mass_kernel.cpp:23 [MFEM-upstream, SURVEY.md Appendix A.1/A.3]:

  * Gauss-Legendre points/weights on [0,1] (open basis nodes, hex rules);
  * Gauss-Lobatto points on [0,1] (closed basis nodes);
  * Lagrange basis + derivative evaluation on arbitrary nodes;
  * symmetric Gauss rules on the reference tetrahedron.
"""
import numpy as np


def gauss_legendre(n):
    """n-point Gauss-Legendre rule on [0,1] (exact to degree 2n-1)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto(n):
    """n-point Gauss-Lobatto points on [0,1] (n >= 2)."""
    if n < 2:
        raise ValueError("Gauss-Lobatto needs n >= 2")
    if n == 2:
        return np.array([0.0, 1.0])
    # interior points: roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    r = np.sort(np.polynomial.legendre.legroots(dc).real)
    # Newton polish for full double accuracy
    for _ in range(3):
        f = np.polynomial.legendre.legval(r, dc)
        df = np.polynomial.legendre.legval(r, np.polynomial.legendre.legder(dc))
        r = r - f / df
    x = np.concatenate([[-1.0], r, [1.0]])
    x = 0.5 * (x - x[::-1])  # enforce symmetry
    return 0.5 * (x + 1.0)


def lagrange_eval(nodes, x):
    """Values B[q, i] = l_i(x_q) and derivatives G[q, i] = l_i'(x_q)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(nodes)
    B = np.ones((len(x), n))
    G = np.zeros((len(x), n))
    for i in range(n):
        denom = 1.0
        for m in range(n):
            if m != i:
                denom *= nodes[i] - nodes[m]
        for m in range(n):
            if m == i:
                continue
            B[:, i] *= x - nodes[m]
            term = np.ones(len(x))
            for l in range(n):
                if l != i and l != m:
                    term *= x - nodes[l]
            G[:, i] += term
        B[:, i] /= denom
        G[:, i] /= denom
    return B, G


def hex_rule_points(order):
    """MFEM's choice for tensor elements: Gauss-Legendre with
    n = (order|1)/2 + 1 points per dimension [MFEM-upstream IntegrationRules]."""
    real_order = order | 1
    return real_order // 2 + 1


# ---------------------------------------------------------------------------
# Tetrahedron rules.  On straight tets with per-element constant coefficients
# any rule exact to the integrand degree reproduces MFEM's element matrix to
# round-off, so the oracle uses a collapsed
# (Duffy / Gauss-Jacobi) rule that is exact to a requested degree.
# ---------------------------------------------------------------------------
def _gauss_jacobi01(n, alpha):
    """n-point Gauss-Jacobi rule on [0,1] with weight (1-x)^alpha."""
    from scipy.special import roots_jacobi

    x, w = roots_jacobi(n, alpha, 0.0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)


def tet_rule(degree):
    """Points (nq,3) / weights (nq,) on the reference tet, exact to `degree`.
    Weights sum to 1/6."""
    n = degree // 2 + 1
    xa, wa = _gauss_jacobi01(n, 0)
    xb, wb = _gauss_jacobi01(n, 1)
    xc, wc = _gauss_jacobi01(n, 2)
    pts = []
    wts = []
    for k in range(n):
        for j in range(n):
            for i in range(n):
                z = xc[k]
                y = xb[j] * (1 - z)
                x = xa[i] * (1 - y - z)
                pts.append((x, y, z))
                wts.append(wa[i] * wb[j] * wc[k])
    return np.array(pts), np.array(wts)
