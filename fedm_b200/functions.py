"""Problem-definition and time-stepping API with the reference's names, arguments and behaviour (``fedm/functions.py``),
written against this repo's dolfin shim (``fedm_b200.dolfin``): weak forms are UFL-lite expressions that the form
compiler turns into CUDA row kernels; the solve is libfedmx's device Newton-Krylov.

Reference map: modify_approximation_vars :15-45, subdomains/marking :48-124, list helpers :127-171, Problem :174-202,
Max/Min :205-216, Flux :219-237, weak forms :240-401, Boundary_flux :404-528, coefficient interpolation :531-750,
semi_implicit_coefficients :753-774, Source_term :777-843, Energy_Source_term :845-912, controllers :915-951,
adaptive_solver :958-1130, Normal_vector / Poisson_solver / BoundaryGradient :1133-1208.
"""
import warnings

import numpy as np

from . import dolfin as df
from . import hotpath as _hp
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


# ---- boundaries -----------------------------------------------------------------------------------------------------------
class CircleSubDomain(df.SubDomain):
    """Part of a circle of given centre (z, r) and radius lying outside the gap (z <= 0 or z >= gap_length)."""

    def __init__(self, center_z, center_r, radius, gap_length, submesh=False, tol=1e-8):
        super().__init__()
        self._cz, self._cr, self._rad = float(center_z), float(center_r), float(radius)
        self._gap_length, self._submesh, self._tol = float(gap_length), bool(submesh), float(tol)

    def inside(self, x, on_boundary):
        r, z = x[0], x[1]
        on_circle = abs((r - self._cr) ** 2 + (z - self._cz) ** 2 - self._rad ** 2) <= self._tol
        outside_gap = z <= 0 if self._cz <= 0 else z >= self._gap_length
        return on_circle and outside_gap and (on_boundary or self._submesh)


class LineSubDomain(df.SubDomain):
    def __init__(self, r_range, z_range):
        super().__init__()
        self._r, self._z = r_range, z_range

    def inside(self, x, on_boundary):
        return df.between(x[0], self._r) and df.between(x[1], self._z) and on_boundary


def Marking_boundaries(mesh, boundaries, submesh=False, gap_length=0.01):
    """Facet markers 1..len(boundaries) for ['line', z1, z2, r1, r2] and ['circle', cz, cr, radius] entries."""
    markers = df.MeshFunction("size_t", mesh, mesh.topology().dim() - 1, 0)
    for number, entry in enumerate(boundaries, start=1):
        kind = entry[0]
        print_rank_0(kind)
        if kind == "circle":
            sub = CircleSubDomain(entry[1], entry[2], entry[3], gap_length, submesh)
        elif kind == "line":
            eps = df.DOLFIN_EPS
            sub = LineSubDomain((entry[3] - eps, entry[4] + eps), (entry[1] - eps, entry[2] + eps))
        else:
            raise ValueError(f"fedm.Marking_boundaries: Invalid boundary_type '{kind}'. Possible values are 'circle', 'line'.")
        sub.mark(markers, number)
    return markers


def Mixed_element_list(number_of_equations, element):
    return [element] * number_of_equations


def Function_space_list(number_of_equations, function_space):
    return [function_space] * number_of_equations


def Function_definition(function_space, function_type, eq_number=1):
    makers = {"TrialFunction": df.TrialFunction, "TestFunction": df.TestFunction, "Function": df.Function}
    if function_type not in makers:
        raise ValueError(f"fedm.Function_definition: Invalid function_type '{function_type}'. "
                         f"Possible values are {comma_separated(makers)}.")
    return [makers[function_type](function_space) for _ in range(eq_number)]


class Problem(df.NonlinearProblem):
    """``Problem(J, F, bcs)``.  The solver reads ``linear_form`` / ``bcs`` and runs assembly + Dirichlet rows + Newton on the
    device; ``F``/``J`` remain for code that assembles by hand."""

    def __init__(self, J, F, bcs):
        self.bilinear_form, self.linear_form, self.bcs = J, F, bcs
        df.NonlinearProblem.__init__(self)

    def F(self, b, x):
        df.assemble(self.linear_form, tensor=b)
        for bc in self.bcs:
            bc.apply(b, x)

    def J(self, A, x):
        df.assemble(self.bilinear_form, tensor=A)
        for bc in self.bcs:
            bc.apply(A)


# ---- form builders ----------------------------------------------------------------------------------------------------------
def Max(a, b):
    return (a + b + abs(a - b)) / df.Constant(2.0)


def Min(a, b):
    return (a + b - abs(a - b)) / df.Constant(2.0)


def Flux(sign, u, D, mu, E, grad_diffusion=True, logarithm_representation=True):
    """Drift-diffusion flux  -grad(D n) (or -D grad n)  + sign mu E n  with n = exp(u) in the log representation."""
    n = df.exp(u) if logarithm_representation else u
    diffusion = -df.grad(D * n) if grad_diffusion else -D * df.grad(n)
    return diffusion + sign * mu * E * n


_EQ_TYPES = ["reaction", "diffusion-reaction", "drift-diffusion-reaction"]


def weak_form_balance_equation(equation_type, dt, dt_old, dx, u, u_old, u_old1, v, f, Gamma, r=0.5 / df.pi, D=None,
                               log_representation=False):
    """Variable-step BDF2 particle balance: 2 pi [ n' v - Gamma.grad(v) - f v ] r dx with
    n' = (exp(u)|1) ((1+2w) u - (1+w)^2 u_old + w^2 u_old1) / ((1+w) dt), w = dt/dt_old."""
    if equation_type not in _EQ_TYPES:
        raise ValueError(f"fedm.weak_form_balance_equation_log_representation: The equation type {equation_type}' is not "
                         f"recognised. Must be one of {comma_separated(_EQ_TYPES)}.")
    if equation_type == "diffusion-reaction" and D is None:
        raise ValueError("fedm.weak_form_balance_equation_log_representation: When 'equation_type' is diffusion-reaction, "
                         "must also supply the diffusion coefficient 'D'.")
    w = dt / dt_old
    wp1, w2p1 = 1.0 + w, 1.0 + 2.0 * w
    weight = df.exp(u) if log_representation else 1.0
    bdf2 = (u * w2p1 - wp1 ** 2.0 * u_old + w ** 2.0 * u_old1) / wp1
    form = 2.0 * df.pi * (weight * bdf2 * v / dt) * r * dx
    if equation_type == "diffusion-reaction":
        n = df.exp(u) if log_representation else u
        form = form - 2.0 * df.pi * df.dot(-df.grad(D * n), df.grad(v)) * r * dx
    elif equation_type == "drift-diffusion-reaction":
        form = form - 2.0 * df.pi * df.dot(Gamma, df.grad(v)) * r * dx
    return form - 2.0 * df.pi * v * f * r * dx


def weak_form_balance_equation_log_representation(*args, **kwargs):
    return weak_form_balance_equation(*args, **kwargs, log_representation=True)


def weak_form_Poisson_equation(dx, u, v, f, r=0.5 / df.pi):
    return 2.0 * df.pi * r * (df.inner(df.grad(u), df.grad(v)) - f * v) * dx


def Boundary_flux(bc_type, equation_type, particle_type, sign, mu, E, normal, u, gamma, v, ds_temp, r=0.5 / df.pi, vth=0.0,
                  ref=1.0, Ion_flux=0.0):
    """Boundary integral of a balance equation: 'zero flux' -> 0, 'flux source' (thermal + drift flux with reflection,
    secondary emission for electrons), 'Neumann' (pure drift flux)."""
    bc_types, particle_types = ["zero flux", "flux source", "Neumann"], ["Heavy", "electrons"]
    if "_" in bc_type:
        warnings.warn("fedm.BoundaryFlux: bc_type should have spaces, not underscores")
        bc_type = bc_type.replace("_", " ")
    if bc_type not in bc_types:
        raise ValueError(f"fedm.Boundary_flux: boundary condition type '{bc_type}' not recognised. Must be one of {comma_separated(bc_types)}.")
    if bc_type != "zero flux" and equation_type not in _EQ_TYPES:
        raise ValueError(f"fedm.Boundary_flux: equation type '{equation_type}' not recognised. Must be one of {comma_separated(_EQ_TYPES)}.")
    if bc_type == "flux source" and equation_type == "diffusion-reaction" and particle_type not in particle_types:
        raise ValueError(f"fedm.Boundary_flux: particle type '{particle_type}' not recognised. Must be one of {comma_separated(particle_types)}.")
    if bc_type == "flux source" and equation_type != "reaction":
        out = (1.0 - ref) / (1.0 + ref)
        if equation_type == "diffusion-reaction":
            out = out * (0.5 * vth * df.exp(u))
        else:
            out = out * ((0.5 * vth + abs(sign * mu * df.dot(E, normal))) * df.exp(u))
            if particle_type == "electrons":
                out = out - 2.0 * gamma * Ion_flux / (1.0 + ref)
        return 2.0 * df.pi * out * v * r * ds_temp
    if bc_type == "Neumann" and equation_type == "drift-diffusion-reaction":
        return 2.0 * df.pi * df.dot(sign * mu * E, normal) * df.exp(u) * v * r * ds_temp
    return 0.0


def semi_implicit_coefficients(dependences, mean_energy_new, mean_energy_old, coefficients, coefficient_diffs):
    if not all(len(x) == len(dependences) for x in (coefficients, coefficient_diffs)):
        raise ValueError("fedm.semi_implicit_coefficients: The lists 'dependences', 'coefficients', and 'coefficient_diffs' must be the same length.")
    return [c + d * (mean_energy_new - mean_energy_old) if dep == "Umean" else c
            for c, d, dep in zip(coefficients, coefficient_diffs, dependences)]


def _densities(N0, u, first, last):
    return [N0] + [df.exp(u[i]) for i in range(first, last)]


def _rates(densities, p_matrix, k_coeffs):
    rates = []
    for r, k in enumerate(k_coeffs):
        prod = 1.0
        for dens, power in zip(densities, p_matrix[r]):          # same association as np.power(...).prod(axis=-1) * k
            prod = prod * dens ** int(power)
        rates.append(prod * k)
    return rates


def Source_term(coupling, approx, p_matrix, l_matrix, g_matrix, k_coeffs, N0, u):
    """f_s = sum_r (gain - loss)[r, s] k_r prod_j n_j^p[r, j], n_0 = N0 (LFA counts species from u[0], LMEA from u[1])."""
    if coupling not in ("coupled", "uncoupled"):
        raise ValueError("fedm.Source_term: coupling must be 'coupled' or 'uncoupled'.")
    if approx not in ("LFA", "LMEA"):
        raise ValueError("fedm.Source_term: approx must be 'LFA' or 'LMEA'.")
    first = 0 if coupling == "coupled" and approx == "LFA" else 1
    last = len(u) - 1 if coupling == "coupled" else len(u)
    rates = _rates(_densities(N0, u, first, last), np.asarray(p_matrix), k_coeffs)
    net = np.asarray(g_matrix) - np.asarray(l_matrix)
    out = []
    for s in range(net.shape[1]):
        total = 0.0
        for r, rate in enumerate(rates):
            if net[r, s] != 0:
                total = total + rate * int(net[r, s])
        out.append(total)
    return out


def Energy_Source_term(coupling, p_matrix, l_matrix, g_matrix, k_coeffs, u_loss, mean_energy, N0, n, Ei=0):
    """-sum_r k_r prod n^p * loss_r; sentinel losses 7e77..8e77 / 9e99..1e100 stand for (Ei - eps) / eps."""
    last = len(n) - 1 if coupling == "coupled" else len(n)
    rates = _rates(_densities(N0, n, 1, last), np.asarray(p_matrix), k_coeffs)
    total = 0.0
    for rate, loss in zip(rates, u_loss):
        if 7e77 < loss < 8e77:
            factor = Ei - mean_energy
        elif 9e99 < loss < 1e100:
            factor = mean_energy
        else:
            factor = loss
        total = total + (-rate) * factor
    return total


# ---- coefficient interpolation (np.interp on dof vectors; the CUDA backend runs fedmx_interp_table) -------------------------
def _check(name, status, dependences, allowed):
    if status not in ("initial", "update"):
        raise ValueError(f"fedm.{name}: status '{status}' not recognised. Must be one of 'initial', 'update'.")
    for dep in dependences:
        if dep not in allowed:
            raise ValueError(f"fedm.{name}: dependence '{dep}' not recognised. Must be one of {comma_separated(allowed)}.")


def _interp_into(coeff, source, kx, ky, divisor=1.0):
    dst, src = coeff.vector(), source.vector()
    if hasattr(dst.store, "interp_from"):
        dst.store.interp_from(src.store, kx, ky, divisor)          # device kernel (bit-exact with np.interp)
    else:
        dst[:] = np.interp(src[:], kx, ky) / divisor if divisor != 1.0 else np.interp(src[:], kx, ky)


def Transport_coefficient_interpolation(status, dependences, N0, Tgas, k_coeffs, kxs, kys, energy, redfield, mus=None):
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
        if dep == "const" and status == "initial":
            coeff.vector()[:] = ky / N0
        elif dep == "Umean":
            _interp_into(coeff, energy, kx, ky, N0)
        elif dep == "E/N":
            _interp_into(coeff, redfield, kx, ky, N0)
        elif dep == "ESR":
            coeff.vector()[:] = kB * Tgas * mu.vector()[:] / elementary_charge
        elif dep == "Tgas":
            coeff.vector()[:] = np.interp(Tgas, kx, ky) / N0
        else:
            continue
        coeff.vector().update_ghost_values()


def Rate_coefficient_interpolation(status, dependences, k_coeffs, kxs, kys, energy, redfield, Te=300.0, Tgas=300.0):
    allowed = [0, "const", "Umean", "E/N", "Te", "fun:Te,Tgas", "fun:Tgas"]
    _check("Rate_coefficient_interpolation", status, dependences, allowed)
    if not all(len(x) == len(k_coeffs) for x in (dependences, kxs, kys)):
        raise ValueError("fedm.Rate_coefficient_interpolation: The lists 'dependences', 'kxs', 'kys', and 'k_coeffs' must be the same length.")
    for coeff, dep, kx, ky in zip(k_coeffs, dependences, kxs, kys):
        if dep == "const" and status == "initial":
            coeff.vector()[:] = ky
        elif dep == "Te":
            coeff.vector()[:] = np.interp(2 * energy.vector()[:] / (3 * kB_eV), kx, ky)
        elif dep == "Umean":
            _interp_into(coeff, energy, kx, ky)
        elif dep == "E/N":
            _interp_into(coeff, redfield, kx, ky)
        else:
            continue
        coeff.vector().update_ghost_values()


# ---- step-size control ---------------------------------------------------------------------------------------------------------
def adaptive_timestep(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    """PID controller; evaluated by libfedmx's host routine (plain libm pow: bit-identical to the reference)."""
    return _hp.pid_step(dt, error, tol, dt_min, dt_max)


def adaptive_timestep_PI34(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return _hp.pi34_step(dt, error, tol, dt_min, dt_max)


def adaptive_timestep_H211b(dt, dt_old, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return _hp.h211b_step(dt, dt_old, error, tol, dt_min, dt_max)


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
        nonlinear_solver.solve(problem, u_new.vector())
        assigner.assign(var_list_new, u_new)
        if approximation in ("LMEA", "LFA"):
            pick = 0 if approximation == "LMEA" else -2
            new_vec, old_vec = var_list_new[pick].vector(), var_list_old[pick].vector()
        else:
            new_vec, old_vec = u_new.vector(), u_old.vector()
        error[0] = df.norm(new_vec - old_vec + df.DOLFIN_EPS) / df.norm(old_vec + df.DOLFIN_EPS)
        if df.MPI.rank(df.MPI.comm_world) == 0:          # every rank holds the same (replicated) estimate: one row per attempt
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


def Normal_vector(mesh):
    """Outward unit normal as a P1 vector field: L2 projection of the facet normal onto the boundary trace space; interior
    dofs have no equation and are set to zero through ``ident_zeros`` (fedm/functions.py:1133-1151)."""
    W = df.VectorFunctionSpace(mesh, "CG", 1)
    u, v = df.TrialFunction(W), df.TestFunction(W)
    facet_normal = df.FacetNormal(mesh)
    A = df.assemble(df.inner(u, v) * df.ds, keep_diagonal=True)
    b = df.assemble(df.inner(facet_normal, v) * df.ds)
    A.ident_zeros()
    n = df.Function(W)
    df.solve(A, n.vector(), b, "mumps")
    return n


def Poisson_solver(A, L, b, bcs, u, solver_type="mumps", preconditioner="hypre_amg"):
    """Segregated Poisson step: re-assemble the right-hand side, apply the Dirichlet conditions and solve with the
    pre-assembled matrix ``A`` (fedm/functions.py:1154-1161)."""
    b = df.assemble(L, tensor=b)
    for bc in bcs:
        bc.apply(b)
    extra = () if solver_type == "mumps" else (preconditioner,)
    df.solve(A, u.vector(), b, solver_type, *extra)


def BoundaryGradient(var, zeroDomain, source_term, ds_extract, epsilon=8.854187817e-12):
    """Flux of ``-epsilon grad(var)`` through a boundary, recovered consistently from the weak Poisson residual instead of
    differentiating the P1 field (D. Kamensky's boundary-flux trick, as adapted by fedm/functions.py:1164-1208): find
    E on the boundary with  int eps E v ds = -(residual of the Poisson equation tested with v) + consistency term on
    ds_extract[0];  all dofs inside ``zeroDomain`` (everything but the boundary of interest) are pinned to zero."""
    V = var.ufl_function_space()
    anti_bcs = [df.DirichletBC(V, df.Constant(0.0), zeroDomain())]
    n = df.FacetNormal(V.mesh())
    E, v = df.TrialFunction(V), df.TestFunction(V)
    En = df.Function(V)
    residual = epsilon * df.inner(df.grad(var), df.grad(v)) * df.dx - source_term * v * df.dx
    F = residual - epsilon * df.inner(-df.grad(var), n) * v * ds_extract[0]
    for _ in ds_extract[1:]:
        F += epsilon * E * v * df.ds
    A = df.assemble(df.lhs(F), keep_diagonal=True)
    b = df.assemble(df.rhs(F))
    for bc in anti_bcs:
        bc.apply(A)
        bc.apply(b)
    df.solve(A, En.vector(), b, "gmres", "hypre_amg")
    return En
