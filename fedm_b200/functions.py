"""Problem-definition and time-stepping API of the hot path, with the reference's names and call signatures
(``fedm/functions.py``), on top of libfedmx.

What is here is what the per-time-step path calls: ``Problem`` (:174-202), the step-size controllers (:915-951),
``ErrorGreaterThanTTOL`` (:954), ``adaptive_solver`` (:958-1130), the coefficient interpolation routines (:531-750) and
``modify_approximation_vars`` (:15-45).  The UFL-building helpers (``Flux``, ``weak_form_*``, ``Boundary_flux``,
``Source_term`` ...) have no UFL to build here: their integrands are the compiled form families of
``fedm_b200/csrc``, selected through ``Problem(model=...)``.
"""
from pathlib import Path

import numpy as np

from . import dolfin_shim as df
from . import hotpath as hp
from .physical_constants import elementary_charge, kB, kB_eV
from .utils import comma_separated, print_rank_0


def modify_approximation_vars(approximation_type, number_of_species, particle_species, masses, charges):
    """LFA drops the first (background) species, LMEA keeps all; number of equations = species + 1."""
    known = ["LFA", "LMEA"]
    if approximation_type not in known:
        raise ValueError(f"fedm.modify_approximation_vars: The approximation type {approximation_type} is not recognised. "
                         f"Must be one of {comma_separated(known)}.")
    if approximation_type == "LFA":
        for seq in (particle_species, masses, charges):
            del seq[0]
        number_of_species -= 1
    return number_of_species, number_of_species + 1, particle_species, masses, charges


class Problem(df.NonlinearProblem):
    """``Problem(J, F, bcs)`` of the reference with the forms replaced by a compiled form family living in a HotPath.
    ``F(b, x)`` / ``J(A, x)`` keep their meaning: assemble + apply the Dirichlet conditions."""

    def __init__(self, hot, dt, dt_old, u_old, u_old1, n_hist, bc_values=lambda t: np.zeros(0), time=lambda: 0.0):
        super().__init__()
        self.hot, self.dt, self.dt_old = hot, dt, dt_old
        self.u_old, self.u_old1, self.n_hist = u_old, u_old1, n_hist
        self.bc_values, self.time = bc_values, time
        self.t = 0.0

    def _prepare(self):
        self.hot.bc_values = np.asarray(self.bc_values(self.t), dtype=np.float64)
        self.hot.bdf2_history(self.n_hist, self.u_old.vector().t, self.u_old1.vector().t, self.dt.time_step, self.dt_old.time_step)

    def F(self, b, x):
        self._prepare()
        self.hot.residual(x.t, self.dt.time_step, self.dt_old.time_step, self.t, out=b.t)

    def J(self, A, x):
        self._prepare()
        self.hot.jacobian(x.t, self.dt.time_step, self.dt_old.time_step, self.t)

    def solve_newton(self, x, **kw):
        self._prepare()
        return self.hot.newton(x.t, self.dt.time_step, self.dt_old.time_step, self.t, **kw)


# ---- step-size controllers (host libm through the C ABI: bit-exact with the reference's Python floats) -------------
def adaptive_timestep(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return hp.pid_step(dt, error, tol, dt_min, dt_max)


def adaptive_timestep_PI34(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return hp.pi34_step(dt, error, tol, dt_min, dt_max)


def adaptive_timestep_H211b(dt, dt_old, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return hp.h211b_step(dt, dt_old, error, tol, dt_min, dt_max)


class ErrorGreaterThanTTOL(Exception):
    pass


def adaptive_solver(nonlinear_solver, problem, t, dt, dt_old, u_new, u_old, var_list_new, var_list_old, assigner, error,
                    error_file, max_error, ttol, dt_min, time_dependent_arguments=None, approximation="LMEA"):
    """One accepted time step: attempt, estimate, log, accept -- or restore, shrink dt and try again (recursively, as the
    reference does); ``SystemExit`` when dt falls below dt_min.  Returns the new time."""
    print_rank_0(f"Attempting to solve the equation for t = {t} with dt = {dt.time_step}", flush=True)
    try:
        t += dt.time_step
        for arg in time_dependent_arguments or []:
            arg.t = t
        problem.t = t
        nonlinear_solver.solve(problem, u_new.vector())
        assigner.assign(var_list_new, u_new)
        if approximation in ("LMEA", "LFA"):
            pick = 0 if approximation == "LMEA" else -2
            new_vec, old_vec = var_list_new[pick].vector(), var_list_old[pick].vector()
        else:
            new_vec, old_vec = u_new.vector(), u_old.vector()
        hot = new_vec.hot
        n_entries = hot.n_owned * new_vec.n_comp
        error[0] = hot.error_norm_strided(new_vec.t, old_vec.t, 1, 0, n_entries, df.DOLFIN_EPS)
        with open(error_file, "a") as log:
            log.write(f"{error[0]:<23}  {dt_old.time_step:<23}  {dt.time_step:<23}\n")
            log.flush()
        max_error[0] = max(error)
        if error[0] >= ttol:
            raise ErrorGreaterThanTTOL
    except Exception as exc:
        t -= dt.time_step
        if isinstance(exc, ErrorGreaterThanTTOL):
            dt.time_step *= 0.5 * ttol / max_error[0]
            print_rank_0("Residual is greater than the prescribed tolerance. Reducing time-step size and repeating calculation.")
        else:
            dt.time_step *= 0.5
            print_rank_0("An exception was raised while solving. Reducing time-step size and repeating calculation.")
        if dt.time_step < dt_min:
            raise SystemExit("Minimum time-step size reached, program is terminating.")
        u_new.assign(u_old)
        assigner.assign(var_list_new, u_new)
        t = adaptive_solver(nonlinear_solver, problem, t, dt, dt_old, u_new, u_old, var_list_new, var_list_old, assigner,
                            error, error_file, max_error, ttol, dt_min, time_dependent_arguments, approximation)
    return t


# ---- coefficient interpolation (np.interp on dof vectors -> fedmx_interp_table) --------------------------------------
def _check(name, status, dependences, allowed):
    if status not in ("initial", "update"):
        raise ValueError(f"fedm.{name}: status '{status}' not recognised. Must be one of 'initial', 'update'.")
    for dep in dependences:
        if dep not in allowed:
            raise ValueError(f"fedm.{name}: dependence '{dep}' not recognised. Must be one of {comma_separated(allowed)}.")


def _table(hot, kx, ky):
    import torch
    mk = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(hot.device)
    return mk(kx), mk(ky)


def Transport_coefficient_interpolation(status, dependences, N0, Tgas, k_coeffs, kxs, kys, energy, redfield, mus=None):
    """Per-node transport coefficients: 'const' (initial only), 'Umean', 'E/N', 'ESR' (kB*Tgas*mu/e), 'Tgas'."""
    allowed = [0, "const", "Umean", "E/N", "ESR", "Tgas"]
    _check("Transport_coefficient_interpolation", status, dependences, allowed)
    if mus is None:
        if "ESR" in dependences:
            raise ValueError("fedm.Transport_coefficient_interpolation: Must provide mus (mobilities) when using ESR dependence.")
        mus = [None] * len(k_coeffs)
    if not all(len(x) == len(k_coeffs) for x in (dependences, kxs, kys, mus)):
        raise ValueError("fedm.Transport_coefficient_interpolation: The lists 'dependences', 'kxs', 'kys', 'k_coeffs', and "
                         "(optionally) 'mus' must be the same length.")
    for coeff, dep, kx, ky, mu in zip(k_coeffs, dependences, kxs, kys, mus):
        vec = coeff.vector()
        hot = vec.hot
        if dep == "const" and status == "initial":
            vec[:] = ky / N0
        elif dep in ("Umean", "E/N"):
            src = energy if dep == "Umean" else redfield
            tx, ty = _table(hot, kx, ky)
            hot.interp_table(src.vector().t, tx, ty, vec.t, divisor=N0)
        elif dep == "ESR":
            hot.strided_scale(mu.vector().t, 1, 0, vec.t, 1, 0, vec.t.numel(), kB * Tgas, elementary_charge)
        elif dep == "Tgas":
            vec[:] = np.interp(Tgas, kx, ky) / N0
        else:
            continue
        vec.update_ghost_values()


def Rate_coefficient_interpolation(status, dependences, k_coeffs, kxs, kys, energy, redfield, Te=300.0, Tgas=300.0):
    """Per-node rate coefficients: 'const' (initial only), 'Umean', 'E/N', 'Te' (table versus 2 eps/(3 kB_eV))."""
    allowed = [0, "const", "Umean", "E/N", "Te", "fun:Te,Tgas", "fun:Tgas"]
    _check("Rate_coefficient_interpolation", status, dependences, allowed)
    if not all(len(x) == len(k_coeffs) for x in (dependences, kxs, kys)):
        raise ValueError("fedm.Rate_coefficient_interpolation: The lists 'dependences', 'kxs', 'kys', and 'k_coeffs' must be the same length.")
    for coeff, dep, kx, ky in zip(k_coeffs, dependences, kxs, kys):
        vec = coeff.vector()
        hot = vec.hot
        if dep == "const" and status == "initial":
            vec[:] = ky
        elif dep in ("Umean", "E/N"):
            src = energy if dep == "Umean" else redfield
            tx, ty = _table(hot, kx, ky)
            hot.interp_table(src.vector().t, tx, ty, vec.t)
        elif dep == "Te":
            tx, ty = _table(hot, np.asarray(kx) * (3 * kB_eV) / 2, ky)        # interp(2 eps/(3 kB_eV); kx, ky) == interp(eps; kx*3kB_eV/2, ky)
            hot.interp_table(energy.vector().t, tx, ty, vec.t)
        else:
            continue
        vec.update_ghost_values()


def _no_ufl(name):
    def stub(*args, **kwargs):
        raise NotImplementedError(f"fedm_b200.functions.{name}: UFL forms are not built on the device path; the integrand is a "
                                  "compiled form family (fedm_b200.hotpath.MODEL_*), see DESIGN.md section 2")
    stub.__name__ = name
    return stub


for _n in ("Flux", "weak_form_balance_equation", "weak_form_balance_equation_log_representation", "weak_form_Poisson_equation",
           "Boundary_flux", "Source_term", "Energy_Source_term", "semi_implicit_coefficients", "Max", "Min",
           "Normal_vector", "Poisson_solver", "BoundaryGradient"):
    globals()[_n] = _no_ufl(_n)
