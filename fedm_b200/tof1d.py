"""1-D electron-swarm time of flight on the device path (BASELINE.json configs[0]).

Host-side mirror of examples/time_of_flight_1D/fedm-tof_1d.py (reference): constants :40-44, time set-up :53-58,
IntervalMesh(4000, 0, 1e-3) :87, P2 function space :98, analytic initial state / exact solution :106, source f interpolated into
P2 :115, initial guess :121, SNES parameters :20-22 (relative tolerance 1e-10, direct linear solver), time loop :143-171 (BDF1
for the first step, constant-step BDF2 afterwards), relative L2 error :155-157.  Per-step arithmetic runs in libfedmx
(fedm_b200/csrc/fedmx_p2line.cu): row-wise assembly of residual + pentadiagonal Jacobian and a banded direct solve.
"""
import ctypes as C

import numpy as np
import torch

from . import hotpath as hp

WEZ, DE, ALPHA = 1.7e5, 0.12, 5009.51
X0, WIDTH = 3e-4, 4e-5
DOLFIN_EPS = 3.0e-16


def ln_swarm(x, t):
    s = 1.0 + 4.0 * DE * t / WIDTH ** 2
    return -((x - X0 - WEZ * t) / WIDTH) ** 2 / s + ALPHA * WEZ * t - 0.5 * np.log(s)


def source(x, t):
    return np.exp(ln_swarm(x, t)) * WEZ * ALPHA


class Tof1DCase:
    def __init__(self, n_cells=4000, length=1e-3, device=0, newton_rtol=1e-10, newton_atol=1e-10, newton_max_it=50):
        self.hot = hp.HotPath(device)                     # context only (stream, error reporting, reductions): no triangle mesh
        self.n_cells, self.h = n_cells, length / n_cells
        self.x = np.linspace(0.0, length, 2 * n_cells + 1)
        self.ndof = self.x.size
        self.rtol, self.atol, self.max_it = newton_rtol, newton_atol, newton_max_it
        dev = self.hot.device
        mk = lambda: torch.zeros(self.ndof, dtype=torch.float64, device=dev)
        self.u_new, self.u_old, self.u_old1, self.f, self.hist, self.res, self.dx = mk(), mk(), mk(), mk(), mk(), mk(), mk()
        self.band = torch.zeros(5 * self.ndof, dtype=torch.float64, device=dev)
        self.newton_its = []

    def _assemble(self, dt, dt_old, jac=True):
        hot, p = self.hot, (lambda t: C.c_void_p(t.data_ptr()))
        hot._ck(hot.lib.fedmx_p2line_assemble(hot.h, self.n_cells, self.h, dt, dt_old, DE, WEZ, p(self.u_new), p(self.u_old),
                                              p(self.u_old1), p(self.f), p(self.hist), p(self.res),
                                              p(self.band) if jac else None))
        return hot.norm2(self.res, self.ndof)

    def newton(self, dt, dt_old):
        """SNES newtonls/basic with a direct linear solve; raises RuntimeError on non-convergence like dolfin."""
        hot = self.hot
        f0 = fn = self._assemble(dt, dt_old)
        if not np.isfinite(f0):
            raise RuntimeError("Newton solver did not converge: residual is not finite")
        if f0 < self.atol:
            return 0
        for it in range(1, self.max_it + 1):
            hot._ck(hot.lib.fedmx_p2line_solve(hot.h, self.ndof, C.c_void_p(self.band.data_ptr()), C.c_void_p(self.res.data_ptr()),
                                               C.c_void_p(self.dx.data_ptr())))
            self.u_new.sub_(self.dx)
            fn = self._assemble(dt, dt_old)
            if not np.isfinite(fn):
                raise RuntimeError("Newton solver did not converge: residual is not finite")
            if fn < self.atol or fn <= self.rtol * f0:
                return it
        raise RuntimeError(f"Newton solver did not converge in {self.max_it} iterations (|F| = {fn:g})")

    def run(self, T_final=3e-9, dt=1e-11, t0=0.0):
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.hot.device)
        t, dt_old = t0, 1e30
        self.u_new.copy_(to_dev(np.log(np.exp(ln_swarm(self.x, t)) + DOLFIN_EPS)))
        self.u_old.copy_(to_dev(ln_swarm(self.x, t)))
        self.u_old1.copy_(self.u_old)
        while abs(t - T_final) / T_final > 1e-6:
            self.u_old1.copy_(self.u_old)
            self.u_old.copy_(self.u_new)
            t += dt
            self.f.copy_(to_dev(source(self.x, t)))               # f = interpolate(Expression(...), V), fedm-tof_1d.py:115
            self.newton_its.append(self.newton(dt, dt_old))
            if t > t0 + dt:
                dt_old = dt
        self.t_end = t
        return self.u_new

    def l2_relative_error(self, t=None):
        """|| e^{u_h} - n_exact ||_L2 / || n_exact ||_L2 with the 5-point rule per cell (post-processing on the host)."""
        t = self.t_end if t is None else t
        u = self.u_new.cpu().numpy()
        g, w = np.polynomial.legendre.leggauss(5)
        xi, w = 0.5 * (g + 1.0), 0.5 * w
        phi = np.stack([(1 - xi) * (1 - 2 * xi), 4 * xi * (1 - xi), xi * (2 * xi - 1)])
        c = np.arange(self.n_cells)
        ul = np.stack([u[2 * c], u[2 * c + 1], u[2 * c + 2]], axis=1)
        xq = self.x[2 * c][:, None] + xi[None, :] * self.h
        nh, ne = np.exp(ul @ phi), np.exp(ln_swarm(xq, t))
        W = w[None, :] * self.h
        return float(np.sqrt(((nh - ne) ** 2 * W).sum() / (ne ** 2 * W).sum()))
