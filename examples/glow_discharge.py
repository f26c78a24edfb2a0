"""Low-pressure argon glow discharge (4 species, local-mean-energy approximation, semi-implicit coefficients) written
against the FEDM problem-definition API -- the multi-species / electron-energy configuration of BASELINE.json.

Cf. examples/glow_discharge/fedm-gd.py of the reference (this version uses the reference *test's* mesh, input deck and end
time, tests/integrated_tests/glow_discharge).  `from dolfin import *` resolves to the shim (compat/dolfin): the coupled
5-equation form below is lowered by the form compiler to one CUDA row kernel and solved by libfedmx's Newton-Krylov.
Run:  PYTHONPATH=compat python examples/glow_discharge.py --input tests/golden/glow_discharge_input
"""
import sys
from argparse import ArgumentParser
from pathlib import Path

import numpy as np

sys.path[:0] = [str(Path(__file__).resolve().parent.parent), str(Path(__file__).resolve().parent.parent / "compat")]

from dolfin import *                       # noqa: E402,F401,F403
from fedm.physical_constants import *      # noqa: E402,F401,F403
from fedm.file_io import *                 # noqa: E402,F401,F403
from fedm.functions import *               # noqa: E402,F401,F403


def main(input_dir=None, output_dir=None, nx=100, T_final=1e-11, linear_solver="mumps", max_steps=None):
    if input_dir is not None:
        files.file_input = Path(input_dir)
    if output_dir is not None:
        files.output_folder_path = Path(output_dir)
    parameters["form_compiler"]["quadrature_degree"] = 4

    model, approximation = "4_particles", "LMEA"
    Tgas, p0, U_w = 300.0, 1.0, -250.0
    N0 = p0 * 3.21877e22
    gap, wall = 0.01, 0.01
    path = files.file_input / model

    # ---- input deck: species, reaction scheme, look-up tables
    n_sp, species, prop_files, names = read_speclist(path)
    M, sign = read_particle_properties(prop_files, model)
    equation_type = ["reaction", "diffusion-reaction", "drift-diffusion-reaction", "drift-diffusion-reaction"]
    particle_type = ["Heavy", "Heavy", "Heavy", "electrons"]
    species_kind = ["Neutral", "Neutral", "Ion", "electrons"]
    n_ic = [N0, 1e12, 1e12, 1e12]
    grad_diff = [k == "electrons" for k in species_kind]
    P, L, G = reaction_matrices(path, species)
    k_files = rate_coefficient_file_names(path)
    energy_loss = read_energy_loss(path)
    n_re = len(k_files)
    mu_x, mu_y, mu_dep = read_transport_coefficients(names, "mobility", model)
    D_x, D_y, D_dep = read_transport_coefficients(names, "Diffusion", model)
    k_dep = read_dependences(k_files)
    k_x, k_y = read_rate_coefficients(k_files, k_dep)
    e = n_sp - 1                                                    # electrons are the last species
    dDe = np.gradient(D_y[e], D_x[e]) / N0                          # table derivatives for the semi-implicit treatment
    dmue = np.gradient(mu_y[e], mu_x[e]) / N0
    dk = [np.gradient(k_y[i], k_x[i]) if k_dep[i] == "Umean" else 0.0 for i in range(n_re)]
    n_sp, n_eq, species, M, sign = modify_approximation_vars(approximation, n_sp, species, M, sign)

    xdmf_u = output_files("xdmf", "number density", names)
    pvd_phi = output_files("pvd", "potential", ["Phi"])
    out_files, out_types = [pvd_phi[0], xdmf_u[1], xdmf_u[2], xdmf_u[3]], ["pvd", "xdmf", "xdmf", "xdmf"]

    # ---- time stepping
    t, dt_min, dt_max, ttol = 0.0, 1e-15, 1e-8, 2e-3
    dt = Expression("time_step", time_step=1e-13, degree=0)
    dt_old = Expression("time_step", time_step=1e30, degree=0)      # huge previous step: BDF2 starts as BDF1
    t_out_list = [1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5]
    t_step_list = [1e-11, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-6]
    t_out_step, t_out = t_out_list[0], t_step_list[0]
    error, max_error = [0.0] * (n_sp + 1), [1] * 3

    # ---- geometry, boundaries and wall data
    r = Expression("x[0]", degree=1)
    boundaries = [["line", 0.0, 0.0, 0.0, wall], ["line", gap, gap, 0.0, wall], ["line", 0.0, gap, 0.0, 0.0], ["line", 0.0, gap, wall, wall]]
    metal, none = [0.3, 0.3, 5e-4, 0.3], [1.0, 1.0, 1.0, 1.0]       # reflection coefficients (1 = zero flux)
    reflection = [metal, metal, none, none]
    gamma = [0.06, 0.06, 0, 0]                                     # secondary electron emission per boundary
    we_see = 5.0                                                   # mean energy of secondary electrons [eV]
    mesh = RectangleMesh(Point(0, 0), Point(wall, gap), nx, nx, "crossed")
    markers = Marking_boundaries(mesh, boundaries)
    ds = Measure("ds", domain=mesh, subdomain_data=markers)
    normal = FacetNormal(mesh)

    # ---- spaces and functions
    P1 = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    ME = FunctionSpace(mesh, MixedElement(Mixed_element_list(n_eq, P1)))
    V = FunctionSpace(mesh, P1)
    spaces = Function_space_list(n_eq, V)
    assigner, rev_assigner = FunctionAssigner(spaces, ME), FunctionAssigner(ME, spaces)
    u, v = TrialFunction(ME), TestFunctions(ME)
    u_new, u_old, u_old1 = Function(ME), Function(ME), Function(ME)
    u_phi, v_phi = TrialFunction(V), TestFunction(V)
    Phi, Phi_old, Phi_old1 = Function(V), Function(V), Function(V)
    redE, redE_old = Function(V), Function(V)
    E = -grad(u[n_eq - 1])
    u_oldV, u_old1V, u_newV = (Function_definition(V, "Function", n_sp) for _ in range(3))
    mean_energy, mean_energy_old, mean_energy_old1 = Function(V), Function(V), Function(V)
    # mean energy consistent with the new energy density and electron density, linearised around the old level
    mean_energy_e = mean_energy_old + (exp(u[0]) - exp(u[e]) * mean_energy_old) / exp(u_oldV[e])
    vth = [0] * n_sp
    for i in range(1, e):
        vth[i] = np.sqrt(8.0 * kB * Tgas / (pi * M[i]))
    vth[e] = sqrt(16.0 * elementary_charge * mean_energy / (3.0 * pi * M[e]))
    D, D_diff, mu, mu_diff = (Function_definition(V, "Function", n_sp) for _ in range(4))
    k_coef, k_coef_diff = Function_definition(V, "Function", n_re), Function_definition(V, "Function", n_re)
    epsilon = Constant(1.0) * epsilon_0

    # ---- initial state
    energy0 = interpolate(Expression("3.0", degree=1), V)
    mean_energy.assign(energy0)
    mean_energy_old.assign(energy0)
    mean_energy_old1.assign(Constant(0.0))
    rho0, rho = 0, 0
    for i in range(n_sp):
        ic = Expression("std::log(ic)", ic=n_ic[i], degree=1)
        u_newV[i].assign(ic)
        u_oldV[i].assign(ic)
        u_old1V[i].assign(Constant(0.0))
        rho0 += elementary_charge * sign[i] * exp(u_oldV[i])         # charge density of the initial state
        rho += elementary_charge * sign[i] * exp(u[i])               # ... of the unknowns (coupled Poisson equation)
    we_newV = interpolate(Expression("std::log(a) + b", a=mean_energy, b=u_oldV[e], degree=1), V)
    we_oldV = interpolate(Expression("std::log(a) + b", a=mean_energy, b=u_oldV[e], degree=1), V)
    we_old1V = interpolate(Constant(0.0), V)
    tmp = Function(V)
    for i in range(n_sp):
        tmp.assign(u_oldV[i])
        tmp.rename(names[i], str(i))
        xdmf_u[i].write_checkpoint(tmp, names[i], t * 1e6, XDMFFile.Encoding.HDF5, False)

    # ---- electrodes and the initial potential
    Phi_grounded = Constant(0.0)
    Phi_powered = Expression("U0*(1-exp(-t/1e-9))", U0=U_w, t=t, pi=pi, degree=0)

    def powered(x, on_boundary):
        return near(x[1], 0, DOLFIN_EPS) and on_boundary

    def grounded(x, on_boundary):
        return near(x[1], gap, DOLFIN_EPS) and on_boundary

    bcs0 = [DirichletBC(V, Phi_powered, powered), DirichletBC(V, Phi_grounded, grounded)]
    F0 = weak_form_Poisson_equation(dx, u_phi, v_phi, rho0 / epsilon, r)
    A0, b0 = assemble(lhs(F0)), assemble(rhs(F0))
    for bc in bcs0:
        bc.apply(A0)
        bc.apply(b0)
    solve(A0, Phi.vector(), b0, "mumps")
    Phi_old1.assign(Phi_old)
    Phi_old.assign(Phi)
    tmp.assign(Phi)
    tmp.rename("Phi", "0")
    pvd_phi[0] << (tmp, t)

    def reduced_field():
        return project(1e21 * sqrt(dot(-grad(Phi), -grad(Phi))) / N0, solver_type="mumps")

    redE.assign(reduced_field())
    redE_old.assign(redE)

    # ---- coefficients: tables -> nodal functions, semi-implicit in the mean energy
    Transport_coefficient_interpolation("initial", mu_dep, N0, Tgas, mu, mu_x, mu_y, mean_energy, redE, mu)
    Transport_coefficient_interpolation("initial", D_dep, N0, Tgas, D, D_x, D_y, mean_energy, redE, mu)
    Rate_coefficient_interpolation("initial", k_dep, k_coef, k_x, k_y, mean_energy, redE, Te=0, Tgas=0)
    k_si = semi_implicit_coefficients(k_dep, mean_energy_e, mean_energy_old, k_coef, k_coef_diff)
    mu_si = semi_implicit_coefficients(mu_dep, mean_energy_e, mean_energy_old, mu, mu_diff)
    D_si = semi_implicit_coefficients(D_dep, mean_energy_e, mean_energy_old, D, D_diff)

    def update_table_derivatives():
        w = mean_energy_old.vector()[:]
        for i in range(n_re):
            if k_dep[i] == "Umean":
                k_coef_diff[i].vector()[:] = np.interp(w, k_x[i], dk[i])
        mu_diff[e].vector()[:] = np.interp(w, mu_x[e], dmue)
        D_diff[e].vector()[:] = np.interp(w, D_x[e], dDe)

    update_table_derivatives()

    # ---- the coupled weak form: heavy species + electrons, electron energy, Poisson
    bcs = [DirichletBC(ME.sub(n_eq - 1), Phi_powered, powered), DirichletBC(ME.sub(n_eq - 1), Phi_grounded, grounded)]
    Gamma, ion_flux = [0], 0
    for i in range(1, n_sp):
        Gamma.append(Flux(sign[i], u[i], D_si[i], mu_si[i], E, grad_diffusion=grad_diff[i], logarithm_representation=True))
        if species_kind[i] == "Ion":
            ion_flux += Max(dot(Gamma[i], normal), 0)
    Gamma_en = Flux(sign[e], u[0], 5.0 * D_si[e] / 3.0, 5.0 * mu_si[e] / 3.0, E, grad_diffusion=grad_diff[e], logarithm_representation=True)
    see_energy = Expression("u_p", u_p=we_see, degree=1)
    f = Source_term("coupled", approximation, P, L, G, k_si, N0, u)
    f_en = Energy_Source_term("coupled", P, L, G, k_si, energy_loss, u[0] / u[e], N0, u)
    f_en += -dot(Flux(sign[e], u[e], D_si[e], mu_si[e], E, grad_diffusion=grad_diff[e], logarithm_representation=True), E)
    F = 0
    for i in range(1, n_sp):
        F += weak_form_balance_equation_log_representation(equation_type[i], dt, dt_old, dx, u[i], u_old[i], u_old1[i], v[i], f[i], Gamma[i], r, D_si[i])
    for b in range(len(boundaries)):
        for j in range(1, n_sp):
            F += Boundary_flux("flux source", equation_type[j], particle_type[j], sign[j], mu_si[j], E, normal, u[j], gamma[b], v[j],
                               ds(b + 1), r, vth[j], reflection[b][j], ion_flux)
    F_en = weak_form_balance_equation_log_representation(equation_type[e], dt, dt_old, dx, u[0], u_old[0], u_old1[0], v[0], f_en, Gamma_en, r)
    for b in range(len(boundaries)):
        F_en += Boundary_flux("flux source", equation_type[e], particle_type[e], sign[e], 5.0 * mu_si[e] / 3.0, E, normal, u[0],
                              gamma[b] * see_energy, v[0], ds(b + 1), r, 1.3333 * vth[e], reflection[b][e], ion_flux)
    F += F_en
    F += weak_form_Poisson_equation(dx, u[n_eq - 1], v[n_eq - 1], rho / epsilon, r)

    new_list = [we_newV, u_newV[1], u_newV[2], u_newV[3], Phi]
    old_list = [we_oldV, u_oldV[1], u_oldV[2], u_oldV[3], Phi_old]
    old1_list = [we_old1V, u_old1V[1], u_old1V[2], u_old1V[3], Phi_old1]
    out_old, out_new = [Phi_old, u_oldV[1], u_oldV[2], u_oldV[3]], [Phi, u_newV[1], u_newV[2], u_newV[3]]
    out_names = ["Phi", names[1], names[2], names[3]]
    rev_assigner.assign(u_new, new_list)
    rev_assigner.assign(u_old, old_list)
    rev_assigner.assign(u_old1, old1_list)

    F = action(F, u_new)
    problem = Problem(derivative(F, u_new, u), F, bcs)
    solver = PETScSNESSolver()
    solver.parameters["relative_tolerance"] = 1e-4
    solver.parameters["linear_solver"] = linear_solver
    solver.parameters["maximum_iterations"] = 20

    steps = 0
    while t < T_final:
        t_old = t
        u_old1.assign(u_old)
        u_old.assign(u_new)
        assigner.assign(old_list, u_old)
        redE_old.assign(redE)
        mean_energy_old1.assign(mean_energy_old)
        mean_energy_old.assign(mean_energy)
        redE.assign(reduced_field())
        Transport_coefficient_interpolation("update", mu_dep, N0, Tgas, mu, mu_x, mu_y, mean_energy_old, redE)
        Transport_coefficient_interpolation("update", D_dep, N0, Tgas, D, D_x, D_y, mean_energy_old, redE, mu)
        Rate_coefficient_interpolation("update", k_dep, k_coef, k_x, k_y, mean_energy_old, redE, Te=0, Tgas=0)
        update_table_derivatives()
        t = adaptive_solver(solver, problem, t, dt, dt_old, u_new, u_old, new_list, old_list, assigner, error, files.error_file,
                            max_error, ttol, dt_min, time_dependent_arguments=[Phi_powered], approximation=approximation)
        mean_energy.vector()[:] = np.exp(we_newV.vector()[:] - u_newV[e].vector()[:])
        t_out, t_out_step = file_output(t, t_old, t_out, t_out_step, t_out_list, t_step_list, out_types, out_files, out_names,
                                        out_new, out_old, unit="us")
        dt_old.time_step = dt.time_step
        dt.time_step = adaptive_timestep(dt.time_step, max_error, ttol, dt_min, dt_max)
        max_error[2], max_error[1] = max_error[1], max_error[0]
        steps += 1
        if max_steps is not None and steps >= max_steps:
            break
    return u_new, {"problem": problem, "solver": solver, "names": names, "new": u_newV, "mesh": mesh}


if __name__ == "__main__":
    ap = ArgumentParser(description="argon glow discharge on the FEDM API (B200 backend)")
    ap.add_argument("-i", "--input", type=Path, default=Path(__file__).resolve().parent.parent / "tests" / "golden" / "glow_discharge_input")
    ap.add_argument("-o", "--output", type=Path, default=None)
    ap.add_argument("--nx", type=int, default=100)
    ap.add_argument("--T-final", type=float, default=1e-11)
    a = ap.parse_args()
    main(input_dir=a.input, output_dir=a.output, nx=a.nx, T_final=a.T_final)
