"""FIAT 2019.1.0 default quadrature schemes, restated (oracle -- test infrastructure only).

SURVEY.md section 3.4 lists which scheme FFC picks for which degree; the constants below are
the published Strang-Fix / Dunavant / collapsed Gauss-Jacobi rules that FIAT's
``create_quadrature(triangle, degree, "default")`` returns.  Degree 4 and degree 6 rules and the
collapsed m=5 rule are pinned by the time-of-flight golden (tests/test_oracle_tof.py).
Points are returned as barycentric triples (l0, l1, l2) with l0 = 1-X-Y, l1 = X, l2 = Y on the
reference triangle (0,0),(1,0),(0,1); weights sum to 1/2.
"""
import numpy as np
from scipy.special import roots_jacobi


def _bary(xy):
    xy = np.asarray(xy, dtype=np.float64)
    return np.stack([1.0 - xy[:, 0] - xy[:, 1], xy[:, 0], xy[:, 1]], axis=1)


def collapsed_gauss_jacobi(m):
    """Collapsed (Duffy) rule with m points per direction, FIAT ordering (x outer, y inner)."""
    ptx, wx = roots_jacobi(m, 0.0, 0.0)
    pty, wy = roots_jacobi(m, 1.0, 0.0)
    pts, wts = [], []
    for xi, wxi in zip(ptx, wx):
        for yj, wyj in zip(pty, wy):
            X = 0.5 * (1.0 + xi) * (1.0 - yj) - 1.0      # on the (-1,-1),(1,-1),(-1,1) triangle
            Y = yj
            pts.append((0.5 * (X + 1.0), 0.5 * (Y + 1.0)))
            wts.append(0.5 * wxi * wyj * 0.25)
    return _bary(pts), np.asarray(wts)


def triangle_rule(degree):
    """FIAT default scheme on the triangle for polynomial degree ``degree``."""
    if degree <= 1:
        return _bary([(1.0 / 3.0, 1.0 / 3.0)]), np.array([0.5])
    if degree == 2:
        pts = [(1.0 / 6.0, 1.0 / 6.0), (1.0 / 6.0, 2.0 / 3.0), (2.0 / 3.0, 1.0 / 6.0)]
        return _bary(pts), np.full(3, 1.0 / 6.0)
    if degree == 3:
        pts = [(0.659027622374092, 0.231933368553031), (0.659027622374092, 0.109039009072877),
               (0.231933368553031, 0.659027622374092), (0.231933368553031, 0.109039009072877),
               (0.109039009072877, 0.659027622374092), (0.109039009072877, 0.231933368553031)]
        return _bary(pts), np.full(6, 1.0 / 12.0)
    if degree == 4:
        a, b = 0.816847572980459, 0.091576213509771
        c, d = 0.108103018168070, 0.445948490915965
        pts = [(a, b), (b, a), (b, b), (c, d), (d, c), (d, d)]
        w = [0.109951743655322 / 2.0] * 3 + [0.223381589678011 / 2.0] * 3
        return _bary(pts), np.asarray(w)
    if degree == 5:
        a1, a2 = 0.101286507323456, 0.797426985353087
        b1, b2 = 0.470142064105115, 0.059715871789770
        pts = [(1.0 / 3.0, 1.0 / 3.0), (a1, a1), (a2, a1), (a1, a2), (b1, b1), (b2, b1), (b1, b2)]
        w = [0.225 / 2.0] + [0.125939180544827 / 2.0] * 3 + [0.132394152788506 / 2.0] * 3
        return _bary(pts), np.asarray(w)
    if degree == 6:
        p, q = 0.873821971016996, 0.063089014491502
        s, t = 0.501426509658179, 0.249286745170910
        e, f, g = 0.636502499121399, 0.310352451033785, 0.053145049844816
        pts = [(p, q), (q, p), (q, q), (s, t), (t, s), (t, t),
               (e, f), (e, g), (f, e), (f, g), (g, e), (g, f)]
        w = [0.050844906370207 / 2.0] * 3 + [0.116786275726379 / 2.0] * 3 + [0.082851075618374 / 2.0] * 6
        return _bary(pts), np.asarray(w)
    return collapsed_gauss_jacobi((degree + 2) // 2)


def facet_rule(degree):
    """Gauss-Legendre on [0,1] with m=(degree+2)//2 points; returns (t, w), sum(w) = 1."""
    m = (degree + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (x + 1.0), 0.5 * w
