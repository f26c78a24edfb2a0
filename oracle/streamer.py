"""Positive-streamer benchmark (Bagheri 2018), LFA, 3 coupled P1 fields (oracle -- test infrastructure only).

Literal restatement of tests/integrated_tests/streamer_discharge/fedm_streamer.py of the reference
(problem definition :39-114, IC :177-180, initial Poisson :194-223, coefficients :243-253, forms
:260-279, SNES set-up :302-305, time loop :312-353) on top of fedm/functions.py:219-237 (Flux),
:350-368 (BDF2 balance equation in log form), :401 (Poisson), :523-524 (Neumann drift flux).
The reference's mesh.xml is missing from the checkout, so the mesh is a synthetic dolfin-style
"right" grid (optionally graded); the assembled entries are therefore *parity unpinned* against
FEniCS and trusted through the shared code path pinned by the ToF and glow goldens.
Unknowns per node: U = [ln n_i, ln n_e, Phi].
"""
import numpy as np

from . import dual as ad
from .controller import (adaptive_attempts, adaptive_timestep, error_estimate, SolveFailed)
from .fem import Assembler, apply_dirichlet_matrix, apply_dirichlet_vector
from .mesh import rectangle_mesh
from .newton import lu_solve, newton

ELEMENTARY_CHARGE = 1.6021766208e-19      # fedm/physical_constants.py:5
EPSILON_0 = 8.854187817e-12               # fedm/physical_constants.py:7
BOX = 0.0125
U_W = 18750.0
BOUNDARIES = [["line", 0.0, 0.0, 0.0, BOX], ["line", BOX, BOX, 0.0, BOX],
              ["line", 0.0, BOX, 0.0, 0.0], ["line", 0.0, BOX, BOX, BOX]]


def streamer_mesh(n, grading_power=1.0):
    """n x n 'right' grid of [0,BOX]^2; grading_power p>1 refines towards r=0 (r = R s^p) and, in z,
    towards the seed at z = 1e-2 is *not* attempted (kept uniform) -- SURVEY.md 8d, C4."""
    if grading_power == 1.0:
        return rectangle_mesh(0.0, 0.0, BOX, BOX, n, n, "right")
    return rectangle_mesh(0.0, 0.0, BOX, BOX, n, n, "right",
                          grading=(lambda s: s ** grading_power, lambda s: s))


def initial_log_densities(coords):
    r, z = coords[:, 0], coords[:, 1]
    ui = np.log(1e13 + 5e18 * np.exp(-(r ** 2 + (z - 1e-2) ** 2) / (0.4e-3) ** 2))
    ue = np.full_like(ui, np.log(1e13))
    return ui, ue


def transport(Em):
    """mu_e, D_e, alpha as the eval'ed strings of e_Nb.dat / e_ND.dat and fedm_streamer.py:247."""
    mu = 2.3987 * Em ** (-0.26)
    D = 4.3628e-3 * Em ** (0.22)
    alpha = (1.1944e6 + 4.3666e26 * Em ** (-3.0)) * ad.exp(-2.73e7 / Em) - 340.75
    return mu, D, alpha


class StreamerProblem:
    n_eq = 3
    degree = 2                     # parameters["form_compiler"]["quadrature_degree"] = 2

    def __init__(self, mesh):
        self.mesh = mesh
        self.asm = Assembler(mesh, 3)
        self.markers = mesh.mark_boundaries(BOUNDARIES)
        z = mesh.coords[:, 1]
        eps = 3.0e-16
        cath = np.nonzero(np.abs(z - 0.0) < eps)[0]
        anod = np.nonzero(np.abs(z - BOX) < eps)[0]
        self.bc_nodes = np.concatenate([cath, anod])
        self.bc_dofs = 3 * self.bc_nodes + 2
        self.bc_vals = np.concatenate([np.zeros(cath.size), np.full(anod.size, U_W)])
        self.dt = 5e-12
        self.dt_old = 1e30
        self.Uold = None
        self.Uold1 = None

    # ---- integrands -----------------------------------------------------------------------
    def _point(self, ctx):
        dt, dt_old = self.dt, self.dt_old
        tr = dt / dt_old
        trp1 = 1.0 + tr
        tr2p1 = 1.0 + 2.0 * tr
        two_pi_r = 2.0 * np.pi * ctx.x[0]
        ui, ue = ctx.u(0), ctx.u(1)
        Hx, Hy = ctx.grad_u(1)
        Gx, Gy = ctx.grad_u(2)
        Em = ad.sqrt(Gx * Gx + Gy * Gy)
        mu, D, alpha = transport(Em)
        ni, ne = ad.exp(ui), ad.exp(ue)
        src = alpha * mu * Em * ne
        terms = []
        for k, (u, n) in enumerate(((ui, ni), (ue, ne))):
            u_part = (u * tr2p1 - trp1 ** 2.0 * ctx.coef(f"old{k}") + tr ** 2.0 * ctx.coef(f"old1{k}")) / trp1
            s = two_pi_r * (n * u_part / dt - src)
            if k == 0:
                terms.append((s, None))
            else:
                # Gamma = -D grad(e^u) + sign*mu*E*e^u, sign=-1, E=-grad(Phi)
                gam_x = -D * ne * Hx + mu * Gx * ne
                gam_y = -D * ne * Hy + mu * Gy * ne
                terms.append((s, (-two_pi_r * gam_x, -two_pi_r * gam_y)))
        f2 = (ni - ne) * (ELEMENTARY_CHARGE / EPSILON_0)
        terms.append((-two_pi_r * f2, (two_pi_r * Gx, two_pi_r * Gy)))
        return terms

    def _facet(self, ctx, n):
        two_pi_r = 2.0 * np.pi * ctx.x[0]
        Gx, Gy = ctx.grad_u(2)
        Em = ad.sqrt(Gx * Gx + Gy * Gy)
        mu, _, _ = transport(Em)
        ne = ad.exp(ctx.u(1))
        # 'Neumann': dot(sign*mu*E, n) e^u  with sign=-1, E=-G
        return [None, two_pi_r * mu * (Gx * n[0] + Gy * n[1]) * ne, None]

    def _coefs(self):
        c = {}
        for k in range(2):
            c[f"old{k}"] = self.Uold[:, k]
            c[f"old1{k}"] = self.Uold1[:, k]
        return c

    # ---- global residual / Jacobian with Dirichlet rows --------------------------------------
    def element_tensors(self, U, jac=True):
        coefs = self._coefs()
        Re, Je = self.asm.assemble_cells(U, self._point, self.degree, coefs, jac=jac)
        fc, Rf, Jf = self.asm.assemble_facets(U, self._facet, self.degree, self.markers, (1, 2), coefs, jac=jac)
        return Re, Je, fc, Rf, Jf

    def residual(self, x, apply_bc=True):
        U = x.reshape(-1, 3)
        Re, _, fc, Rf, _ = self.element_tensors(U, jac=False)
        b = self.asm.scatter_vector(Re)
        if Rf is not None:
            self.asm.scatter_vector(Rf, fc, out=b)
        if apply_bc:
            apply_dirichlet_vector(b, x, self.bc_dofs, self.bc_vals)
        return b

    def jacobian(self, x, apply_bc=True):
        U = x.reshape(-1, 3)
        _, Je, fc, _, Jf = self.element_tensors(U, jac=True)
        blocks = [(Je, None)]
        if Jf is not None:
            blocks.append((Jf, fc))
        A = self.asm.scatter_matrix(blocks)
        if apply_bc:
            A = apply_dirichlet_matrix(A, self.bc_dofs)
        return A

    # ---- set-up ---------------------------------------------------------------------------------
    def initial_potential(self, ui, ue):
        """Initial Poisson solve (fedm_streamer.py:210-223): a = 2 pi r grad u.grad v, L = 2 pi r f v."""
        asm1 = Assembler(self.mesh, 1)
        coefs = {"ui": ui, "ue": ue}

        def point(ctx):
            two_pi_r = 2.0 * np.pi * ctx.x[0]
            f = (np.exp(ctx.coef("ui")) - np.exp(ctx.coef("ue"))) * (ELEMENTARY_CHARGE / EPSILON_0)
            gx, gy = ctx.grad_u(0)
            return [(-two_pi_r * f, (two_pi_r * gx, two_pi_r * gy))]

        zero = np.zeros((self.mesh.num_vertices, 1))
        Re, Je = asm1.assemble_cells(zero, point, self.degree, coefs, jac=True)
        A = asm1.scatter_matrix([(Je, None)])
        b = -asm1.scatter_vector(Re)
        A = apply_dirichlet_matrix(A, self.bc_nodes)
        b[self.bc_nodes] = self.bc_vals
        return lu_solve(A, b)

    def initial_state(self):
        ui, ue = initial_log_densities(self.mesh.coords)
        phi = self.initial_potential(ui, ue)
        return np.stack([ui, ue, phi], axis=1).ravel()

    # ---- time loop (fedm_streamer.py:312-353) --------------------------------------------------
    def run(self, T_final=1e-10, ttol=1e-3, dt_min=1e-15, dt_max=5e-12, rtol=1e-4, max_it=20,
            linear="lu", max_steps=None, callback=None):
        x_new = self.initial_state()
        x_old = x_new.copy()
        t = 0.0
        self.dt, self.dt_old = 5e-12, 1e30
        max_error = [1, 1, 1]
        rows, newton_its = [], []
        steps = 0
        while t < T_final * (1 + 1e-6):
            self.Uold1 = x_old.reshape(-1, 3).copy()
            x_old = x_new.copy()
            self.Uold = x_old.reshape(-1, 3)

            def solve_fn(t_new, dt):
                self.dt = dt
                hist = []
                its, _ = newton(self.residual, self.jacobian, x_new, rtol=rtol, max_it=max_it,
                                linear=linear, history=hist)
                newton_its.append((its, hist))
                return error_estimate(x_new[1::3], x_old[1::3])      # LFA: var index -2 = ln n_e

            def reset_fn():
                x_new[:] = x_old

            t, self.dt, _ = adaptive_attempts(solve_fn, reset_fn, t, self.dt, ttol, dt_min, max_error,
                                              log_rows=rows, dt_old=self.dt_old)
            self.dt_old = self.dt
            self.dt = adaptive_timestep(self.dt, max_error, ttol, dt_min, dt_max)
            max_error[2] = max_error[1]
            max_error[1] = max_error[0]
            steps += 1
            if callback is not None:
                callback(steps, t, x_new)
            if max_steps is not None and steps >= max_steps:
                break
        return x_new, rows, newton_its
