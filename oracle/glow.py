"""Argon glow discharge, LMEA, 5 coupled P1 fields (oracle -- test infrastructure only).

Literal restatement of tests/integrated_tests/glow_discharge/fedm_gd.py of the reference (conditions :45-56,
readers :61-98, time set-up :111-134, geometry/boundaries :139-160, IC :244-264, initial Poisson :270-307, reduced
field :309, coefficient interpolation :312-340, forms :345-388, SNES :405-412, loop :419-470) on top of
fedm/functions.py (Flux :219-237, balance equation :350-368, Poisson :401, Boundary_flux :514-522,
interpolation :621-639/:724-750, semi-implicit coefficients :768-774, Source_term :835-843,
Energy_Source_term :901-912) and fedm/file_io.py (readers :226-521).  Recipe: SURVEY.md Appendix B.
Pinned by the golden HDF5 vectors and the 6-row 'relative error.log' (tests/test_oracle_glow.py).
Unknowns per node: U = [ln w_e, ln n_Ar*, ln n_Ar+, ln n_e, Phi].
"""
import re
from pathlib import Path

import numpy as np
import scipy.sparse.linalg as spla

from . import dual as ad
from .controller import adaptive_attempts, adaptive_timestep, error_estimate
from .fem import Assembler, Geometry, apply_dirichlet_matrix, apply_dirichlet_vector, mass_matrix_cartesian
from .mesh import rectangle_mesh
from .newton import lu_solve, newton

E_CH = 1.6021766208e-19
EPS0 = 8.854187817e-12
KB = 1.38064852e-23
GAP = WALL = 0.01
BOUNDARIES = [["line", 0.0, 0.0, 0.0, WALL], ["line", GAP, GAP, 0.0, WALL],
              ["line", 0.0, GAP, 0.0, 0.0], ["line", 0.0, GAP, WALL, WALL]]


# ---- readers (fedm/file_io.py restated with plain numpy) ----------------------------------------------------------
def _decomment(path):
    out = []
    for line in Path(path).read_text().splitlines():
        line = line.split("#", 1)[0].strip()
        if line:
            out.append(line)
    return out


def _dependence(path):
    for line in Path(path).read_text().splitlines():
        if "Dependence:" in line:
            return line.split()[2]
    raise RuntimeError(path)


def _table(path):
    dep = _dependence(path)
    rows = _decomment(path)
    if dep in ("Umean", "E/N"):
        a = np.array([[float(t) for t in r.split()] for r in rows])
        return dep, a[:, 0], a[:, 1]
    if dep == "ESR":
        return dep, 0.0, 0.0
    return dep, 0.0, float(rows[0])


class GlowInputs:
    def __init__(self, path):
        path = Path(path)
        lines = [l.replace("file:", "").split() for l in _decomment(path / "speclist.cfg") if "file:" in l]
        self.species = [l[0] for l in lines]
        self.file_names = [l[1].split(".")[0] for l in lines]
        self.M, self.sign = [], []
        for l in lines:
            txt = "\n".join(_decomment(path / "species" / l[1]))
            self.M.append(float(re.search(r"Mass\s?=\s?([+-]?\d+.\d+[eE]?[-+]?\d+|0|1.0)", txt).group(1)))
            self.sign.append(float(re.search(r"Z\s+?=\s+?([+-]?\d+)", txt).group(1)))
        reac = [l.partition(" Type:")[0] for l in _decomment(path / "reacscheme.cfg")]
        loss = [r.partition(" -> ")[0].rstrip() for r in reac]
        gain = [r.partition(" -> ")[2].rstrip() for r in reac]
        L = np.array([[lo.count(s) for s in self.species] for lo in loss])
        G = np.array([[ga.count(s) for s in self.species] for ga in gain])
        self.power = L
        lg = L - G
        self.gain_minus_loss = np.where(lg < 0, -lg, 0) - np.where(lg > 0, lg, 0)
        full = "\n".join(_decomment(path / "reacscheme.cfg"))
        self.kfiles = re.findall(r"kfile: ([A-Za-z0-9_]+.[A-Za-z0-9_]+)", full)
        self.energy_loss = [float(x) for x in re.findall(r"Uin:\s?([+-]?\d+.\d+[eE]?[-+]?\d+|0|1.0)", full)]
        self.rates = [_table(path / "rate_coefficients" / k) for k in self.kfiles]
        self.mob, self.dif = [], []
        for fn in self.file_names:
            pm = path / "transport_coefficients" / (fn + "_Nb.dat")
            self.mob.append(_table(pm) if pm.is_file() else (0, 0, 0))
            self.dif.append(_table(path / "transport_coefficients" / (fn + "_ND.dat")))


class GlowProblem:
    n_eq = 5
    degree = 4                               # parameters["form_compiler"]["quadrature_degree"] = 4

    def __init__(self, input_dir, nx=100, ny=100, ttol=2e-3):
        self.inp = GlowInputs(Path(input_dir) / "4_particles")
        self.N0 = 1.0 * 3.21877e22
        self.Tgas = 300.0
        self.U_w = -250.0
        self.ttol = ttol
        self.mesh = rectangle_mesh(0.0, 0.0, WALL, GAP, nx, ny, "crossed")
        self.asm = Assembler(self.mesh, 5)
        self.markers = self.mesh.mark_boundaries(BOUNDARIES)
        z = self.mesh.coords[:, 1]
        pw = np.nonzero(np.abs(z - 0.0) < 3e-16)[0]
        gr = np.nonzero(np.abs(z - GAP) < 3e-16)[0]
        self.bc_nodes = np.concatenate([pw, gr])
        self.n_powered = pw.size
        self.bc_dofs = 5 * self.bc_nodes + 4
        inp = self.inp
        self.vth = [0.0, np.sqrt(8.0 * KB * self.Tgas / (np.pi * inp.M[1])), np.sqrt(8.0 * KB * self.Tgas / (np.pi * inp.M[2]))]
        self.ref = [0.3, 0.3, 5e-4, 0.3]
        self.gamma = 0.06
        self.we_see = 5.0
        # derivative tables (fedm_gd.py:83-95)
        self.De_diff = np.gradient(inp.dif[3][2], inp.dif[3][1]) / self.N0
        self.mue_diff = np.gradient(inp.mob[3][2], inp.mob[3][1]) / self.N0
        self.k_diff = [np.gradient(ky, kx) if dep == "Umean" else 0.0 for dep, kx, ky in inp.rates]
        self.M_cart = mass_matrix_cartesian(self.mesh)
        self.M_lu = spla.splu(self.M_cart.tocsc())
        self.geo = Geometry(self.mesh)
        self.coef = {}
        self.dt, self.dt_old, self.t = 1e-13, 1e30, 0.0

    def bc_vals(self, t):
        return np.concatenate([np.full(self.n_powered, self.U_w * (1 - np.exp(-t / 1e-9))),
                               np.zeros(self.bc_nodes.size - self.n_powered)])

    # ---- per-step coefficient update (fedm_gd.py:432-443) -----------------------------------------------------------
    def reduced_field(self, phi):
        """project(1e21*sqrt(dot(-grad(Phi),-grad(Phi)))/N0): Cartesian mass solve, cellwise-constant integrand."""
        g, cells = self.geo.g, self.mesh.cells
        P = phi[cells]
        gx = (P * g[:, :, 0]).sum(axis=1)
        gy = (P * g[:, :, 1]).sum(axis=1)
        val = 1e21 * np.sqrt(gx * gx + gy * gy) / self.N0
        b = np.zeros(self.mesh.num_vertices)
        contrib = val * self.geo.detJ / 6.0            # val * area / 3
        for a in range(3):
            np.add.at(b, cells[:, a], contrib)
        return self.M_lu.solve(b)

    def update_coefficients(self, eps_old, redE):
        inp, N0, c = self.inp, self.N0, self.coef
        n = self.mesh.num_vertices
        c["mu2"] = np.interp(redE, inp.mob[2][1], inp.mob[2][2]) / N0
        c["D2"] = KB * self.Tgas * c["mu2"] / E_CH
        c["D1"] = np.full(n, inp.dif[1][2] / N0)
        c["mu3"] = np.interp(eps_old, inp.mob[3][1], inp.mob[3][2]) / N0
        c["D3"] = np.interp(eps_old, inp.dif[3][1], inp.dif[3][2]) / N0
        c["mu3d"] = np.interp(eps_old, inp.mob[3][1], self.mue_diff)
        c["D3d"] = np.interp(eps_old, inp.dif[3][1], self.De_diff)
        for r, (dep, kx, ky) in enumerate(inp.rates):
            if dep == "Umean":
                c[f"k{r}"] = np.interp(eps_old, kx, ky)
                c[f"kd{r}"] = np.interp(eps_old, kx, self.k_diff[r])
            else:
                c[f"k{r}"] = np.full(n, ky)
        c["eps_old"] = eps_old.copy()

    # ---- integrands ---------------------------------------------------------------------------------------------------
    def _si(self, u0, u3, C):
        """Semi-implicit pieces at a point given (possibly nested-dual) fields: returns delta = eps_e - eps_old."""
        eps_e = C["eps_old"] + (ad.exp(u0) - ad.exp(u3) * C["eps_old"]) / ad.exp(C["u3_old"])
        return eps_e - C["eps_old"]

    def _rates(self, delta, n_list, get):
        inp = self.inp
        rates = []
        for r, (dep, _, _) in enumerate(inp.rates):
            k = get(f"k{r}")
            if dep == "Umean":
                k = k + get(f"kd{r}") * delta
            prod = k
            for j in range(4):
                for _ in range(int(inp.power[r, j])):
                    prod = prod * n_list[j]
            rates.append(prod)
        return rates

    def _point(self, ctx):
        dt, dt_old = self.dt, self.dt_old
        tr = dt / dt_old
        trp1, tr2p1 = 1.0 + tr, 1.0 + 2.0 * tr
        two_pi_r = 2.0 * np.pi * ctx.x[0]
        u = [ctx.u(k) for k in range(5)]
        gu = [ctx.grad_u(k) for k in range(5)]
        Ex, Ey = -gu[4][0], -gu[4][1]
        names = ["eps_old", "u3_old", "D3", "D3d", "mu3", "mu3d"]
        C = {n: ctx.coef(n) for n in names + ["mu2", "D2", "D1"] + [k for k in self.coef if k.startswith("k")]}
        delta = self._si(u[0], u[3], C)
        D3si = C["D3"] + C["D3d"] * delta
        mu3si = C["mu3"] + C["mu3d"] * delta
        n = [self.N0, ad.exp(u[1]), ad.exp(u[2]), ad.exp(u[3])]
        rates = self._rates(delta, n, lambda k: C[k])
        gml = self.inp.gain_minus_loss
        f = [sum(rates[r] * float(gml[r, s]) for r in range(len(rates)) if gml[r, s] != 0) for s in range(4)]
        f_en = 0.0
        for r, loss in enumerate(self.inp.energy_loss):
            f_en = f_en - rates[r] * loss

        # spatial gradients of composites: grad(D3si e^{u3}), grad(5/3 D3si e^{u0}), grad(D1 e^{u1})
        def comp_e(U, Cn):
            d = self._si(U[0], U[3], Cn)
            return (Cn["D3"] + Cn["D3d"] * d) * ad.exp(U[3])

        def comp_en(U, Cn):
            d = self._si(U[0], U[3], Cn)
            return (5.0 * (Cn["D3"] + Cn["D3d"] * d) / 3.0) * ad.exp(U[0])

        _, gDe = ctx.spatial_grad(comp_e, unknowns=(0, 3), coef_names=names)
        _, gDen = ctx.spatial_grad(comp_en, unknowns=(0, 3), coef_names=names)
        _, gD1 = ctx.spatial_grad(lambda U, Cn: Cn["D1"] * ad.exp(U[1]), unknowns=(1,), coef_names=["D1"])
        ge = (-gDe[0] - mu3si * Ex * n[3], -gDe[1] - mu3si * Ey * n[3])                     # Gamma_e
        w = ad.exp(u[0])
        gen = (-gDen[0] - (5.0 * mu3si / 3.0) * Ex * w, -gDen[1] - (5.0 * mu3si / 3.0) * Ey * w)   # Gamma_en
        g2 = (-C["D2"] * n[2] * gu[2][0] + C["mu2"] * Ex * n[2], -C["D2"] * n[2] * gu[2][1] + C["mu2"] * Ey * n[2])
        f_en = f_en - (ge[0] * Ex + ge[1] * Ey)

        def tder(k, dens):
            u_part = (u[k] * tr2p1 - trp1 ** 2.0 * ctx.coef(f"old{k}") + tr ** 2.0 * ctx.coef(f"old1{k}")) / trp1
            return dens * u_part / dt

        terms = [None] * 5
        terms[0] = (two_pi_r * (tder(0, w) - f_en), (-two_pi_r * gen[0], -two_pi_r * gen[1]))
        terms[1] = (two_pi_r * (tder(1, n[1]) - f[1]), (two_pi_r * gD1[0], two_pi_r * gD1[1]))   # -dot(-grad(D e^u), grad v)
        terms[2] = (two_pi_r * (tder(2, n[2]) - f[2]), (-two_pi_r * g2[0], -two_pi_r * g2[1]))
        terms[3] = (two_pi_r * (tder(3, n[3]) - f[3]), (-two_pi_r * ge[0], -two_pi_r * ge[1]))
        terms[4] = (-two_pi_r * (E_CH / EPS0) * (n[2] - n[3]), (two_pi_r * gu[4][0], two_pi_r * gu[4][1]))
        return terms

    def _facet(self, ctx, nrm):
        two_pi_r = 2.0 * np.pi * ctx.x[0]
        u = [ctx.u(k) for k in range(5)]
        gu = [ctx.grad_u(k) for k in range(5)]
        En = -(gu[4][0] * nrm[0] + gu[4][1] * nrm[1])
        C = {n: ctx.coef(n) for n in ["eps_old", "u3_old", "D3", "D3d", "mu3", "mu3d", "mu2", "D2", "eps"]}
        delta = self._si(u[0], u[3], C)
        mu3si = C["mu3"] + C["mu3d"] * delta
        n1, n2, n3, w = ad.exp(u[1]), ad.exp(u[2]), ad.exp(u[3]), ad.exp(u[0])
        g2n = -C["D2"] * n2 * (gu[2][0] * nrm[0] + gu[2][1] * nrm[1]) + C["mu2"] * En * n2
        ion = (g2n + ad.absolute(g2n)) / 2.0
        vth_e = np.sqrt(16.0 * E_CH * C["eps"] / (3.0 * np.pi * self.inp.M[3]))
        c = lambda ref: (1.0 - ref) / (1.0 + ref)
        r1 = c(self.ref[1]) * (0.5 * self.vth[1] * n1)
        r2 = c(self.ref[2]) * ((0.5 * self.vth[2] + ad.absolute(1.0 * C["mu2"] * En)) * n2)
        r3 = c(self.ref[3]) * ((0.5 * vth_e + ad.absolute(-1.0 * mu3si * En)) * n3) - 2.0 * self.gamma * ion / (1.0 + self.ref[3])
        r0 = (c(self.ref[3]) * ((0.5 * (1.3333 * vth_e) + ad.absolute(-1.0 * (5.0 * mu3si / 3.0) * En)) * w)
              - 2.0 * (self.gamma * self.we_see) * ion / (1.0 + self.ref[3]))
        return [two_pi_r * r0, two_pi_r * r1, two_pi_r * r2, two_pi_r * r3, None]

    def _coefs(self):
        c = dict(self.coef)
        for k in range(4):
            c[f"old{k}"] = self.Uold[:, k]
            c[f"old1{k}"] = self.Uold1[:, k]
        c["u3_old"] = self.Uold[:, 3]
        c["eps"] = self.eps
        return c

    def residual(self, x, apply_bc=True):
        U = x.reshape(-1, 5)
        coefs = self._coefs()
        Re, _ = self.asm.assemble_cells(U, self._point, self.degree, coefs, jac=False)
        fc, Rf, _ = self.asm.assemble_facets(U, self._facet, self.degree, self.markers, (1, 2), coefs, jac=False)
        b = self.asm.scatter_vector(Re)
        self.asm.scatter_vector(Rf, fc, out=b)
        if apply_bc:
            apply_dirichlet_vector(b, x, self.bc_dofs, self.bc_vals(self.t))
        return b

    def jacobian(self, x, apply_bc=True):
        U = x.reshape(-1, 5)
        coefs = self._coefs()
        _, Je = self.asm.assemble_cells(U, self._point, self.degree, coefs, jac=True)
        fc, _, Jf = self.asm.assemble_facets(U, self._facet, self.degree, self.markers, (1, 2), coefs, jac=True)
        A = self.asm.scatter_matrix([(Je, None), (Jf, fc)])
        if apply_bc:
            A = apply_dirichlet_matrix(A, self.bc_dofs)
        return A

    # ---- run (fedm_gd.py:419-470) -----------------------------------------------------------------------------------
    def initial_state(self):
        n = self.mesh.num_vertices
        X = np.zeros((n, 5))
        X[:, 1:4] = np.log(1e12)
        X[:, 0] = np.log(3.0) + np.log(1e12)
        return X.ravel()          # initial Poisson: zero net charge, both electrode values 0 at t = 0  =>  Phi = 0

    def run(self, T_final=1e-11, dt_min=1e-15, dt_max=1e-8, rtol=1e-4, max_it=20, t_out=1e-11, max_steps=None):
        n = self.mesh.num_vertices
        x_new = self.initial_state()
        x_old = x_new.copy()
        x_old1 = np.zeros_like(x_new)
        self.eps = np.full(n, 3.0)
        eps_old = np.full(n, 3.0)
        self.t, self.dt, self.dt_old = 0.0, 1e-13, 1e30
        max_error = [1, 1, 1]
        rows, its, out = [], [], None
        steps = 0
        while self.t < T_final:
            t_old = self.t
            x_old1 = x_old.copy()
            x_old = x_new.copy()
            self.Uold, self.Uold1 = x_old.reshape(-1, 5), x_old1.reshape(-1, 5)
            eps_old = self.eps.copy()
            redE = self.reduced_field(x_new[4::5])
            self.update_coefficients(eps_old, redE)

            def solve_fn(t_new, dt):
                self.dt, self.t = dt, t_new
                hist = []
                k, _ = newton(self.residual, self.jacobian, x_new, rtol=rtol, max_it=max_it, history=hist)
                its.append((k, hist))
                return error_estimate(x_new[0::5], x_old[0::5])          # LMEA: variable index 0 = ln w_e

            def reset_fn():
                x_new[:] = x_old

            self.t, self.dt, _ = adaptive_attempts(solve_fn, reset_fn, t_old, self.dt, self.ttol, dt_min, max_error,
                                                   log_rows=rows, dt_old=self.dt_old)
            self.eps = np.exp(x_new[0::5] - x_new[3::5])
            if out is None and t_old < t_out <= self.t * (1 + 1e-12):
                # file_output: linear interpolation in time to the output instant (fedm/file_io.py:585-587)
                out = x_old + (t_out - t_old) * (x_new - x_old) / (self.t - t_old)
            self.dt_old = self.dt
            self.dt = adaptive_timestep(self.dt, max_error, self.ttol, dt_min, dt_max)
            max_error[2] = max_error[1]
            max_error[1] = max_error[0]
            steps += 1
            if max_steps is not None and steps >= max_steps:
                break
        return x_new, rows, its, out
