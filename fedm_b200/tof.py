"""Time-of-flight electron swarm on the device path.

Host-side mirror of tests/integrated_tests/time_of_flight/fedm_tof.py (reference): constants :45-52, time set-up
:61-66, mesh :95, analytic initial condition :116-118, initial guess :129, SNES parameters :28-30/:138-140, time loop
:146-177 and the final ``project(exp(u_new), V)`` :165.  All per-step arithmetic runs in libfedmx.
"""
import numpy as np
import torch

from . import hotpath as hp
from .mesh import rectangle

WEZ, DE, ALPHA = 1.7e5, 0.12, 5009.51
DOLFIN_EPS = 3.0e-16


def u_analytical(r, z, t):
    return np.log(np.exp(-((z - WEZ * t) ** 2 + r ** 2) / (4.0 * DE * t) + ALPHA * WEZ * t) / (4 * DE * t * np.pi) ** 1.5)


class TofCase:
    def __init__(self, nx=40, ny=40, width=2.5e-4, height=5e-4, device=0, newton_rtol=1e-10, newton_max_it=50,
                 ksp_rtol=1e-13):
        self.mesh = rectangle((0.0, 0.0), (width, height), nx, ny, "right")
        self.hot = hp.HotPath(device)
        self.hot.set_mesh(1, self.mesh.coords, self.mesh.cells)
        self.hot.set_model(hp.MODEL_TOF, [DE, WEZ, ALPHA], 1)
        self.newton_kw = dict(rtol=newton_rtol, max_it=newton_max_it, ksp_rtol=ksp_rtol, ksp_restart=30, ksp_max_it=20000)
        self.device = self.hot.device
        self.newton_its = []

    def run(self, t0=2.5e-9, T_final=2.6e-9, dt=1e-12):
        hot, c = self.hot, self.mesh.coords
        ua = u_analytical(c[:, 0], c[:, 1], t0)
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        x_old, x_old1 = to_dev(ua), to_dev(ua)
        x_new = to_dev(np.log(np.exp(ua) + DOLFIN_EPS))
        t, dt_old = t0, 1e30
        while abs(t - T_final) / T_final > 1e-6:
            x_old1.copy_(x_old)
            x_old.copy_(x_new)
            t += dt
            hot.bdf2_history(1, x_old, x_old1, dt, dt_old)
            res = hot.newton(x_new, dt, dt_old, t, **self.newton_kw)
            self.newton_its.append(res.iterations)
            if t > t0 + dt:
                dt_old = dt                      # BDF1 for the first two steps, BDF2 afterwards (fedm_tof.py:175-176)
        self.t_end, self.x_new = t, x_new
        return x_new

    def project_density(self, x):
        """n = project(exp(u), V): Cartesian P1 mass solve on the device (ProjectionModel, 6-point degree-4 rule)."""
        pr = hp.HotPath(self.device.index or 0)
        pr.set_mesh(1, self.mesh.coords, self.mesh.cells)
        pr.set_model(hp.MODEL_PROJECTION, [0.0, 1.0], 1)
        pr.fields.copy_(x.view(-1, 1))
        out = pr.new_vector()
        pr.newton(out, 1.0, 1.0, rtol=1e-13, atol=0.0, max_it=3, ksp_rtol=1e-14, ksp_restart=30, ksp_max_it=5000,
                  raise_on_failure=False)
        res = out.cpu().numpy()
        pr.close()
        return res
