"""CPU oracle for BASELINE.json configs[0]: the 1-D electron-swarm time of flight (test infrastructure only).

Restates examples/time_of_flight_1D/fedm-tof_1d.py of the reference: IntervalMesh(4000, 0, 1e-3), P2 Lagrange unknown
u = ln n_e (:97-104), analytic Gaussian swarm as initial state and exact solution (:106-108), Gamma = -grad(D e^u) + w e^u
(:114), source f interpolated into P2 (:115), the log-representation balance equation in Cartesian form (r = 1/(2 pi),
fedm/functions.py:350-368), BDF1 first step then constant-step BDF2 (:143-171), Newton with rtol 1e-10 (:20-22) and a direct
solve.  Quadrature: no global degree is set, UFL estimates 8 (e^u u v with P2 u, v) => FIAT's 5-point Gauss-Legendre rule.

PARITY UNPINNED against FEniCS: the reference ships no fixture for this case (it is an example, not a test).  The oracle is
anchored on the case's own exact solution instead -- the discrete solution must converge to it at the expected orders
(tests/test_oracle_tof1d.py) -- and on the 2-D time-of-flight oracle, which shares the weak form and IS pinned by a FEniCS
golden file.  There is no GPU kernel for this configuration (P2 / 1-D elements are outside the hot path).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

WEZ, DE, ALPHA = 1.7e5, 0.12, 5009.51        # drift velocity, diffusion coefficient, ionisation coefficient (fedm-tof_1d.py:40-42)
X0, WIDTH = 3e-4, 4e-5                       # initial position and width of the Gaussian (:43-44)


def ln_swarm(x, t):
    """ln of the exact density (fedm-tof_1d.py:106)."""
    s = 1.0 + 4.0 * DE * t / WIDTH ** 2
    return -((x - X0 - WEZ * t) / WIDTH) ** 2 / s + ALPHA * WEZ * t - 0.5 * np.log(s)


def source(x, t):
    """f = alpha w n_exact (fedm-tof_1d.py:115)."""
    return np.exp(ln_swarm(x, t)) * WEZ * ALPHA


class Tof1D:
    def __init__(self, n_cells=4000, length=1e-3, n_quad=5):
        self.n, self.h = n_cells, length / n_cells
        self.x = np.linspace(0.0, length, 2 * n_cells + 1)           # P2 nodes: vertices and midpoints, ordered by coordinate
        self.dofs = np.stack([2 * np.arange(n_cells), 2 * np.arange(n_cells) + 1, 2 * np.arange(n_cells) + 2], axis=1)
        g, w = np.polynomial.legendre.leggauss(n_quad)
        xi, self.w = 0.5 * (g + 1.0), 0.5 * w
        # P2 basis on [0, 1], local order (left vertex, midpoint, right vertex)
        self.phi = np.stack([(1 - xi) * (1 - 2 * xi), 4 * xi * (1 - xi), xi * (2 * xi - 1)])            # (3, q)
        self.dphi = np.stack([4 * xi - 3, 4 - 8 * xi, 4 * xi - 1]) / self.h
        self.xq = self.x[self.dofs[:, 0], None] + xi[None, :] * self.h                                   # (cells, q)
        self.rows = np.repeat(self.dofs, 3, axis=1).ravel()
        self.cols = np.tile(self.dofs, (1, 3)).ravel()
        self.ndof = self.x.size

    def _local(self, vec):
        return vec[self.dofs]                                                                            # (cells, 3)

    def residual_jacobian(self, u, u_old, u_old1, f_nodal, dt, dt_old, jac=True):
        tr = dt / dt_old
        ct = (1.0 + 2.0 * tr) / ((1.0 + tr) * dt)
        hist = (-(1.0 + tr) ** 2 * u_old + tr ** 2 * u_old1) / ((1.0 + tr) * dt)
        ul, hl, fl = self._local(u), self._local(hist), self._local(f_nodal)
        uq, duq = ul @ self.phi, ul @ self.dphi                                                          # (cells, q)
        hq, fq = hl @ self.phi, fl @ self.phi
        n = np.exp(uq)
        beta = ct * uq + hq
        gamma = -DE * n * duq + WEZ * n
        W = self.w[None, :] * self.h
        s_q = (n * beta - fq) * W
        g_q = gamma * W
        Re = s_q @ self.phi.T - g_q @ self.dphi.T                                                        # (cells, 3)
        b = np.zeros(self.ndof)
        np.add.at(b, self.dofs.ravel(), Re.ravel())
        if not jac:
            return b, None
        # dR_a/du_b = int [ n (beta + ct) phi_b phi_a - ( -D n (phi_b' + u' phi_b) + w n phi_b ) phi_a' ]
        c0 = n * (beta + ct) * W
        c1 = (-DE * n * duq + WEZ * n) * W                                                               # multiplies phi_b phi_a'
        c2 = -DE * n * W                                                                                 # multiplies phi_b' phi_a'
        Je = (np.einsum("cq,aq,bq->cab", c0, self.phi, self.phi) - np.einsum("cq,aq,bq->cab", c1, self.dphi, self.phi)
              - np.einsum("cq,aq,bq->cab", c2, self.dphi, self.dphi))
        A = sp.csr_matrix((Je.ravel(), (self.rows, self.cols)), shape=(self.ndof, self.ndof))
        return b, A

    def newton(self, u, u_old, u_old1, f_nodal, dt, dt_old, rtol=1e-10, atol=1e-10, max_it=50):
        b, A = self.residual_jacobian(u, u_old, u_old1, f_nodal, dt, dt_old)
        f0 = np.linalg.norm(b)
        for it in range(1, max_it + 1):
            u = u - spla.splu(A.tocsc()).solve(b)
            b, A = self.residual_jacobian(u, u_old, u_old1, f_nodal, dt, dt_old)
            fn = np.linalg.norm(b)
            if fn < atol or fn <= rtol * f0:
                return u, it
        raise RuntimeError("Newton solver did not converge")

    def l2_relative_error(self, u, t):
        """|| e^{u_h} - n_exact ||_L2 / || n_exact ||_L2 by quadrature (the reference projects both onto P2 first, :155-157)."""
        nh = np.exp(self._local(u) @ self.phi)
        ne = np.exp(ln_swarm(self.xq, t))
        W = self.w[None, :] * self.h
        return float(np.sqrt(((nh - ne) ** 2 * W).sum() / (ne ** 2 * W).sum()))

    def run(self, T_final=3e-9, dt=1e-11, t0=0.0):
        t, dt_old = t0, 1e30
        u_new = np.log(np.exp(ln_swarm(self.x, t)) + 3.0e-16)                                            # initial guess (:121)
        u_old = ln_swarm(self.x, t)
        u_old1 = u_old.copy()
        its = []
        while abs(t - T_final) / T_final > 1e-6:
            u_old1, u_old = u_old, u_new.copy()                                                          # :144-145
            t += dt
            u_new, k = self.newton(u_new, u_old, u_old1, source(self.x, t), dt, dt_old)
            its.append(k)
            if t > t0 + dt:
                dt_old = dt
        return u_new, t, its
