"""Time-of-flight electron swarm, one log-density unknown (oracle -- test infrastructure only).

Restates tests/integrated_tests/time_of_flight/fedm_tof.py of the reference (constants :45-52,
time set-up :61-74, mesh :95, analytic solution :116-118, flux and source :120-125, form :127,
initial guess :129, SNES :131-141, loop :146-177) using the recipe of SURVEY.md Appendix F.
Pinned by the golden VTU field and the golden relative_error scalar (tests/test_oracle_tof.py).
"""
import numpy as np
import scipy.sparse.linalg as spla

from . import dual as ad
from .fem import Assembler, Geometry, mass_matrix_cartesian
from .mesh import rectangle_mesh
from .newton import newton
from .quadrature import triangle_rule

WEZ, DE, ALPHA = 1.7e5, 0.12, 5009.51
DOLFIN_EPS = 3.0e-16


def u_analytical(r, z, t):
    return np.log(np.exp(-((z - WEZ * t) ** 2 + r ** 2) / (4.0 * DE * t) + ALPHA * WEZ * t)
                  / (4 * DE * t * np.pi) ** 1.5)


def source(r, z, t):
    return (np.exp(-((z - WEZ * t) ** 2 + r ** 2) / (4.0 * DE * t) + ALPHA * WEZ * t)
            * (WEZ * ALPHA) / (8 * np.pi ** 1.5 * (DE * t) ** 1.5))


class TofProblem:
    def __init__(self, nx=40, ny=40, width=2.5e-4, height=5e-4):
        self.mesh = rectangle_mesh(0.0, 0.0, width, height, nx, ny, "right")
        self.asm = Assembler(self.mesh, 1)
        self.degree = 8                       # UFL's estimate for the summed integrand (no global setting)
        self.dt, self.dt_old = 1e-12, 1e30
        xc = self.mesh.coords[self.mesh.cells]              # (nc,3,2)
        mid = np.stack([0.5 * (xc[:, 1] + xc[:, 2]), 0.5 * (xc[:, 0] + xc[:, 2]), 0.5 * (xc[:, 0] + xc[:, 1])], axis=1)
        self.p2_nodes = np.concatenate([xc, mid], axis=1)   # (nc,6,2)
        self.f6 = None
        self.Uold = self.Uold1 = None

    def _point(self, ctx):
        dt, dt_old = self.dt, self.dt_old
        tr = dt / dt_old
        trp1, tr2p1 = 1.0 + tr, 1.0 + 2.0 * tr
        two_pi_r = 2.0 * np.pi * ctx.x[0]
        u = ctx.u(0)
        hx, hy = ctx.grad_u(0)
        n = ad.exp(u)
        u_part = (u * tr2p1 - trp1 ** 2.0 * ctx.coef("old") + tr ** 2.0 * ctx.coef("old1")) / trp1
        f = ctx.p2(self.f6)
        s = two_pi_r * (n * u_part / dt - f)
        gx = -DE * n * hx
        gy = -DE * n * hy + WEZ * n
        return [(s, (-two_pi_r * gx, -two_pi_r * gy))]

    def residual(self, x):
        Re, _ = self.asm.assemble_cells(x.reshape(-1, 1), self._point, self.degree,
                                        {"old": self.Uold, "old1": self.Uold1}, jac=False)
        return self.asm.scatter_vector(Re)

    def jacobian(self, x):
        _, Je = self.asm.assemble_cells(x.reshape(-1, 1), self._point, self.degree,
                                        {"old": self.Uold, "old1": self.Uold1}, jac=True)
        return self.asm.scatter_matrix([(Je, None)])

    def set_time(self, t):
        p = self.p2_nodes
        self.f6 = source(p[:, :, 0], p[:, :, 1], t)

    def project_exp(self, nodal_fn, degree, elem="P1"):
        """project(exp(w), V): Cartesian mass solve with the FFC-estimated quadrature degree."""
        mesh = self.mesh
        geo = Geometry(mesh)
        lam_all, w_all = triangle_rule(degree)
        b = np.zeros(mesh.num_vertices)
        for lam, w in zip(lam_all, w_all):
            val = np.exp(nodal_fn(lam))
            for a in range(3):
                np.add.at(b, mesh.cells[:, a], w * geo.detJ * val * lam[a])
        M = mass_matrix_cartesian(mesh)
        return spla.splu(M.tocsc()).solve(b), M

    def run(self, t0=2.5e-9, T_final=2.6e-9, rtol=1e-10, max_it=50):
        c = self.mesh.coords
        ua = u_analytical(c[:, 0], c[:, 1], t0)
        x_old = ua.copy()
        x_old1 = ua.copy()
        x_new = np.log(np.exp(ua) + DOLFIN_EPS)
        t = t0
        its_log = []
        while abs(t - T_final) / T_final > 1e-6:
            x_old1 = x_old.copy()
            x_old = x_new.copy()
            self.Uold, self.Uold1 = x_old, x_old1
            t += self.dt
            self.set_time(t)
            its, _ = newton(self.residual, self.jacobian, x_new, rtol=rtol, max_it=max_it)
            its_log.append(its)
            if t > t0 + 1e-12:
                self.dt_old = self.dt
        self.t_end = t
        return x_new, its_log

    def postprocess(self, x_new):
        """n_num = project(exp(u_new)) (degree 4) and the relative L2 error vs project(exp(P3 u_an)) (degree 6)."""
        cells = self.mesh.cells
        Uc = x_new[cells]
        n_num, M = self.project_exp(lambda lam: Uc @ lam, 4)
        t = self.t_end
        xc = self.mesh.coords[cells]

        def p3(lam):
            l = lam
            val = 0.0
            for i in range(3):
                xi = xc[:, i]
                val = val + 0.5 * l[i] * (3 * l[i] - 1) * (3 * l[i] - 2) * u_analytical(xi[:, 0], xi[:, 1], t)
                for j in range(3):
                    if j != i:
                        xe = (2.0 * xc[:, i] + xc[:, j]) / 3.0
                        val = val + 4.5 * l[i] * l[j] * (3 * l[i] - 1) * u_analytical(xe[:, 0], xe[:, 1], t)
            xm = (xc[:, 0] + xc[:, 1] + xc[:, 2]) / 3.0
            val = val + 27.0 * l[0] * l[1] * l[2] * u_analytical(xm[:, 0], xm[:, 1], t)
            return val

        n_ex, _ = self.project_exp(p3, 6)
        e = n_num - n_ex
        rel = np.sqrt(e @ (M @ e)) / np.sqrt(n_ex @ (M @ n_ex))
        return n_num, n_ex, float(rel)
