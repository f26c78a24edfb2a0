"""Argon glow discharge (LMEA, 5 coupled fields) on the device path.

Host-side mirror of tests/integrated_tests/glow_discharge/fedm_gd.py (reference): input decks through the
fedm.file_io-compatible readers (:61-98), conditions (:45-56, :111-160), initial state (:244-264, :300-307), coefficient
tables and their derivatives (:83-95, :312-340), Dirichlet conditions (:345-347), SNES parameters (:38-40, :405-412) and
the time loop (:419-470).  Per step the device does: L2 projection of the reduced field, 17 table interpolations, the
BDF2 history, the Newton solve (GlowModel kernels), the error estimate and the mean-energy update.
"""
import numpy as np
import torch

from . import file_io
from . import hotpath as hp
from .mesh import rectangle
from .physical_constants import elementary_charge, epsilon_0, kB
from .timestepping import adaptive_timestep, attempt_until_accepted

GAP = WALL = 0.01
BOUNDARIES = [["line", 0.0, 0.0, 0.0, WALL], ["line", GAP, GAP, 0.0, WALL],
              ["line", 0.0, GAP, 0.0, 0.0], ["line", 0.0, GAP, WALL, WALL]]
# field columns (see fedm_b200/csrc/glow_model.cuh)
F_HIST, F_EPS_OLD, F_U3_OLD, F_D3, F_D3D, F_MU3, F_MU3D, F_MU2, F_D2, F_D1, F_EPS, F_K, F_KD, NF = 0, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 21, 28


class GlowCase:
    n_eq = 5

    def __init__(self, input_dir, nx=100, ny=100, device=0, ttol=2e-3, newton_rtol=1e-4, newton_max_it=20,
                 ksp_rtol=1e-5, precond="amg"):
        file_io.files.file_input = input_dir
        model = "4_particles"
        path = file_io.files.file_input / model
        self.N0, self.Tgas, self.U_w = 1.0 * 3.21877e22, 300.0, -250.0
        nsp, species, prop_files, stems = file_io.read_speclist(path)
        M, sign = file_io.read_particle_properties(prop_files, model)
        power, loss, gain = file_io.reaction_matrices(path, species)
        kfiles = file_io.rate_coefficient_file_names(path)
        energy_loss = file_io.read_energy_loss(path)
        mu_x, mu_y, mu_dep = file_io.read_transport_coefficients(stems, "mobility", model)
        D_x, D_y, D_dep = file_io.read_transport_coefficients(stems, "Diffusion", model)
        k_dep = file_io.read_dependences(kfiles)
        k_x, k_y = file_io.read_rate_coefficients(kfiles, k_dep)
        e = nsp - 1
        assert nsp == 4 and len(kfiles) <= 7 and mu_dep[e] == "Umean" and D_dep[e] == "Umean" and D_dep[2] == "ESR"
        N0 = self.N0
        De_diff = np.gradient(D_y[e], D_x[e]) / N0
        mue_diff = np.gradient(mu_y[e], mu_x[e]) / N0
        self.k_dep, self.R = k_dep, len(kfiles)
        self.mesh = m = rectangle((0.0, 0.0), (WALL, GAP), nx, ny, "crossed")
        self.hot = hot = hp.HotPath(device)
        dev = self.device = hot.device
        hot.set_mesh(5, m.coords, m.cells, m.exterior_facets(), m.mark_lines(BOUNDARIES))
        z = m.coords[:, 1]
        pw = np.nonzero(np.abs(z) < 3e-16)[0]
        gr = np.nonzero(np.abs(z - GAP) < 3e-16)[0]
        self.bc_nodes = np.concatenate([pw, gr]).astype(np.int32)
        hot.set_dirichlet(4, self.bc_nodes, np.concatenate([np.zeros(pw.size), np.ones(gr.size)]).astype(np.int32))
        p = np.zeros(128)
        ref = [0.3, 0.3, 5e-4, 0.3]
        p[0], p[1], p[2], p[3] = N0, elementary_charge / epsilon_0, self.R, float((1 << 1) | (1 << 2))
        p[4] = np.sqrt(8.0 * kB * self.Tgas / (np.pi * M[1]))
        p[5] = np.sqrt(8.0 * kB * self.Tgas / (np.pi * M[2]))
        p[6], p[7], p[8], p[9], p[10] = ref[1], ref[2], ref[3], 0.06, 5.0
        p[11] = 16.0 * elementary_charge / (3.0 * np.pi * M[e])
        p[12] = 1.3333
        gml = gain - loss
        for r in range(self.R):
            p[16 + 4 * r:20 + 4 * r] = power[r]
            p[48 + 4 * r:52 + 4 * r] = gml[r]
            p[80 + r] = energy_loss[r]
            p[88 + r] = 1.0 if k_dep[r] == "Umean" else 0.0
        hot.set_model(hp.MODEL_GLOW, p, NF)
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        self.tab = {"mu2": (td(mu_x[2]), td(mu_y[2])), "mu3": (td(mu_x[e]), td(mu_y[e])), "D3": (td(D_x[e]), td(D_y[e])),
                    "mu3d": (td(mu_x[e]), td(mue_diff)), "D3d": (td(D_x[e]), td(De_diff))}
        self.ktab = {}
        for r in range(self.R):
            if k_dep[r] == "Umean":
                self.ktab[r] = (td(k_x[r]), td(k_y[r]), td(np.gradient(k_y[r], k_x[r])))
            else:
                hot.fields[:, F_K + r] = float(k_y[r])           # 'const' rate coefficients: set once ('initial' status)
        hot.fields[:, F_D1] = float(D_y[1]) / N0                  # Ar* diffusion: const
        self.flat = hot.fields.view(-1)
        # reduced-field projector: project(1e21*sqrt(dot(-grad(Phi),-grad(Phi)))/N0) (fedm_gd.py:309,432)
        self.proj = pr = hp.HotPath(device)
        pr.set_mesh(1, m.coords, m.cells)
        pr.set_model(hp.MODEL_PROJECTION, [1.0, 1e21 / N0], 1)
        self.redE = pr.new_vector()
        self.precond = precond
        nv = m.num_vertices
        x0 = np.zeros((nv, 5))
        x0[:, 1:4] = np.log(1e12)
        x0[:, 0] = np.log(3.0) + np.log(1e12)        # initial Poisson problem has zero charge and zero electrode values: Phi = 0
        self.x_new, self.x_old, self.x_old1 = td(x0.ravel()), td(x0.ravel()), hot.new_vector()
        hot.fields[:, F_EPS] = 3.0
        if precond == "amg":
            hot.bc_values = np.array([0.0, 0.0])
            self._update_coefficients()
            hot.bdf2_history(4, self.x_old, self.x_old1, 1e-13, 1e30)
            hot.jacobian(self.x_new, 1e-13, 1e30)
            hot.setup_two_stage_preconditioner(4, self.bc_nodes)
        self.newton_kw = dict(rtol=newton_rtol, max_it=newton_max_it, ksp_rtol=ksp_rtol, ksp_restart=30, ksp_max_it=20000)
        self.t, self.dt, self.dt_old = 0.0, 1e-13, 1e30
        self.dt_min, self.dt_max, self.ttol = 1e-15, 1e-8, ttol
        self.max_error = [1, 1, 1]
        self.rows, self.stats = [], []

    def _col(self, c):
        return self.flat[c:]                        # contiguous view starting at column c; stride NF selects the column

    def _interp(self, x, table, col, divisor=1.0):
        self.hot.interp_table(x, table[0], table[1], self._col(col), divisor=divisor, x_stride=x[1] if isinstance(x, tuple) else 1,
                              out_stride=NF, n=self.mesh.num_vertices)

    def _update_coefficients(self):
        """fedm_gd.py:426-443: shift the mean energy, project the reduced field, interpolate every table."""
        hot, nv = self.hot, self.mesh.num_vertices
        hot.strided_copy(self.flat, NF, F_EPS, self.flat, NF, F_EPS_OLD, nv)            # mean_energy_old <- mean_energy
        hot.strided_copy(self.x_old, 5, 3, self.flat, NF, F_U3_OLD, nv)                 # ln n_e of the previous level
        pr = self.proj
        pr.strided_copy(self.x_new, 5, 4, pr.fields.view(-1), 1, 0, nv)                 # Phi
        self.redE.zero_()
        pr.newton(self.redE, 1.0, 1.0, rtol=1e-13, atol=0.0, max_it=3, ksp_rtol=1e-14, ksp_restart=30, ksp_max_it=5000,
                  raise_on_failure=False)
        eps_old = self._col(F_EPS_OLD)
        N0 = self.N0
        it = lambda x, xs, tab, col, div=1.0: hot.interp_table(x, tab[0], tab[1], self._col(col), divisor=div, x_stride=xs,
                                                               out_stride=NF, n=nv)
        it(self.redE, 1, self.tab["mu2"], F_MU2, N0)
        hot.strided_scale(self.flat, NF, F_MU2, self.flat, NF, F_D2, nv, kB * self.Tgas, elementary_charge)    # 'ESR'
        it(eps_old, NF, self.tab["mu3"], F_MU3, N0)
        it(eps_old, NF, self.tab["D3"], F_D3, N0)
        it(eps_old, NF, self.tab["mu3d"], F_MU3D)
        it(eps_old, NF, self.tab["D3d"], F_D3D)
        for r, (kx, ky, kd) in self.ktab.items():
            it(eps_old, NF, (kx, ky), F_K + r)
            it(eps_old, NF, (kx, kd), F_KD + r)

    def step(self):
        hot = self.hot
        self.x_old1.copy_(self.x_old)
        self.x_old.copy_(self.x_new)
        self._update_coefficients()
        self.t_old = self.t

        def solve_attempt(t_new, dt):
            hot.bc_values = np.array([self.U_w * (1.0 - np.exp(-t_new / 1e-9)), 0.0])      # Phi_powered.t = t (fedm/functions.py:1042-1044)
            hot.bdf2_history(4, self.x_old, self.x_old1, dt, self.dt_old)
            res = hot.newton(self.x_new, dt, self.dt_old, t_new, **self.newton_kw)
            self.stats.append((res.iterations, res.linear_iterations, res.fnorm0, res.fnorm))
            return hot.error_norm(self.x_new, self.x_old, 0)                                 # LMEA: ln w_e

        def reset():
            self.x_new.copy_(self.x_old)

        self.t, self.dt, attempts = attempt_until_accepted(
            solve_attempt, reset, self.t, self.dt, self.dt_old, self.ttol, self.dt_min, self.max_error,
            error_log=lambda row: self.rows.append(tuple(float(v) for v in row.split())))
        hot.exp_diff(self.x_new, 5, 0, 3, self.flat, NF, F_EPS, self.mesh.num_vertices)       # mean_energy = exp(ln w_e - ln n_e)
        self.dt_old = self.dt
        self.dt = adaptive_timestep(self.dt, self.max_error, self.ttol, self.dt_min, self.dt_max)
        self.max_error[2] = self.max_error[1]
        self.max_error[1] = self.max_error[0]
        return attempts

    def run(self, T_final=1e-11, t_out=1e-11):
        """Returns the state linearly interpolated to t_out (file_output semantics, fedm/file_io.py:585-587)."""
        out = None
        while self.t < T_final:
            self.step()
            if out is None and self.t_old < t_out <= self.t * (1 + 1e-12):
                xo, xn = self.x_old.cpu().numpy(), self.x_new.cpu().numpy()
                out = xo + (t_out - self.t_old) * (xn - xo) / (self.t - self.t_old)
        return out
