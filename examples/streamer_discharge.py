"""Positive streamer in air (Bagheri et al. 2018 benchmark) written against the FEDM problem-definition API.

This is how a FEDM user script looks on this stack: `from dolfin import *` resolves to the shim (compat/dolfin), the
`fedm` package to the API mirror (compat/fedm), and every call below has the reference's name and signature
(cf. examples/streamer_discharge/fedm-streamer.py of the reference).  The reference ships an unstructured mesh.xml that is
not part of its checkout, so the mesh is a structured n x n grid of the same 1.25 cm x 1.25 cm domain.
Run:  PYTHONPATH=compat python examples/streamer_discharge.py --n 64 --input tests/golden/streamer_discharge_input
"""
import sys
from argparse import ArgumentParser
from pathlib import Path

sys.path[:0] = [str(Path(__file__).resolve().parent.parent), str(Path(__file__).resolve().parent.parent / "compat")]

from dolfin import *                       # noqa: E402,F401,F403
from fedm.physical_constants import *      # noqa: E402,F401,F403
from fedm.file_io import *                 # noqa: E402,F401,F403
from fedm.functions import *               # noqa: E402,F401,F403


def main(n=64, T_final=1e-10, input_dir=None, output_dir=None, linear_solver="gmres", max_steps=None, mesh_file=None):
    if input_dir is not None:
        files.file_input = Path(input_dir)
    if output_dir is not None:
        files.output_folder_path = Path(output_dir)
    parameters["form_compiler"]["quadrature_degree"] = 2

    model, approximation = "benchmark_model", "LFA"
    U_w, box = 18750.0, 0.0125
    path = files.file_input / model
    n_species, species, prop_files, species_files = read_speclist(path)
    M, sign = read_particle_properties(prop_files, model)
    equation_type = ["reaction", "drift-diffusion-reaction"]
    n_species, n_eq, species, M, sign = modify_approximation_vars(approximation, n_species, species, M, sign)

    dt_min, dt_max, ttol = 1e-15, 5e-12, 1e-3
    dt = Expression("time_step", time_step=5e-12, degree=0)
    dt_old = Expression("time_step", time_step=1e30, degree=0)      # huge previous step: BDF2 degenerates to BDF1 at start
    error, max_error = [0.0] * n_species, [1] * 3

    r = Expression("x[0]", degree=1)
    boundaries = [["line", 0.0, 0.0, 0.0, box], ["line", box, box, 0.0, box], ["line", 0.0, box, 0.0, 0.0], ["line", 0.0, box, box, box]]
    bc_type = [["zero flux", "Neumann"], ["zero flux", "Neumann"], ["zero flux", "zero flux"], ["zero flux", "zero flux"]]
    gamma = [0.0, 0.0]
    species_type = ["Ions", "electrons"]

    # the reference reads an unstructured dolfin-XML mesh (fedm-streamer.py:116); without one, a structured grid of the same box
    mesh = Mesh(str(mesh_file)) if mesh_file is not None else RectangleMesh(Point(0.0, 0.0), Point(box, box), n, n)
    facet_markers = Marking_boundaries(mesh, boundaries)
    normal = FacetNormal(mesh)
    dx = Measure("dx", domain=mesh)
    ds = Measure("ds", domain=mesh, subdomain_data=facet_markers)

    P1 = FiniteElement("Lagrange", mesh.ufl_cell(), 1)
    ME = FunctionSpace(mesh, MixedElement(Mixed_element_list(n_eq, P1)))
    V = FunctionSpace(mesh, P1)
    assigner = FunctionAssigner(Function_space_list(n_eq, V), ME)
    rev_assigner = FunctionAssigner(ME, Function_space_list(n_eq, V))
    u, v = TrialFunction(ME), TestFunctions(ME)
    u_new, u_old, u_old1 = Function(ME), Function(ME), Function(ME)
    u_newV, u_oldV = Function_definition(V, "Function", n_eq), Function_definition(V, "Function", n_eq)
    mu, D = Function_definition(V, "Function", n_eq), Function_definition(V, "Function", n_eq)
    Gamma, f = Function_definition(V, "Function", n_eq), Function_definition(V, "Function", n_eq)

    seed = "std::log(1e13+5e18*exp(-(pow(x[0], 2)+pow(x[1]-1e-2, 2))/pow(0.4e-3, 2)))"
    for fn_list in (u_oldV, u_newV):
        fn_list[0] = interpolate(Expression(seed, degree=1), V)
        fn_list[1] = interpolate(Expression("std::log(1e13)", degree=1), V)

    # initial potential from the initial space charge
    Phi, PhiV, vp = Function(V), TrialFunction(V), TestFunction(V)
    grounded = lambda x, on_boundary: near(x[1], 0) and on_boundary
    powered = lambda x, on_boundary: near(x[1], box) and on_boundary
    bcs_V = [DirichletBC(V, Constant(0.0), grounded), DirichletBC(V, Constant(U_w), powered)]
    rho0 = (exp(u_oldV[0]) - exp(u_oldV[1])) * elementary_charge / epsilon_0
    Fp = weak_form_Poisson_equation(dx, PhiV, vp, rho0, r)
    A0, b0 = assemble(lhs(Fp)), assemble(rhs(Fp))
    for bc in bcs_V:
        bc.apply(A0)
        bc.apply(b0)
    solve(A0, Phi.vector(), b0)
    u_oldV[2].assign(Phi)
    u_newV[2].assign(Phi)

    E = -grad(u[2])
    E_m = sqrt(inner(-grad(u[2]), -grad(u[2])))
    D_x, D_y, D_dep = read_transport_coefficients(species, "Diffusion", model)
    mu_x, mu_y, mu_dep = read_transport_coefficients(species, "mobility", model)
    bc = [DirichletBC(ME.sub(2), Constant(0.0), grounded), DirichletBC(ME.sub(2), Constant(U_w), powered)]
    mu[0], D[0] = mu_y[0], D_y[0]
    mu[1], D[1] = eval(mu_y[1]), eval(D_y[1])              # 'fun:E' entries: expressions in E_m from the input deck
    alpha = (1.1944e6 + 4.3666e26 * E_m ** (-3)) * exp(-2.73e7 / E_m) - 340.75
    Gamma[0] = 0.0
    Gamma[1] = Flux(sign[1], u[1], D[1], mu[1], E, grad_diffusion=False)
    f[0] = alpha * mu[1] * E_m * exp(u[1])
    f[1] = alpha * mu[1] * E_m * exp(u[1])
    for i in range(n_species):
        f[2] += sign[i] * exp(u[i]) * elementary_charge / epsilon_0

    F = 0.0
    for i in range(n_species):
        F += weak_form_balance_equation_log_representation(equation_type[i], dt, dt_old, dx, u[i], u_old[i], u_old1[i], v[i],
                                                           f[i], Gamma[i], r, D[i])
    F += weak_form_Poisson_equation(dx, u[n_eq - 1], v[n_eq - 1], f[n_eq - 1], r)
    for i in range(len(boundaries)):
        for j in range(n_species):
            F += Boundary_flux(bc_type[i][j], equation_type[j], species_type[j], sign[j], mu[j], E, normal, u[j], gamma[j], v[j],
                               ds(i + 1), r)

    variable_list_new, variable_list_old = list(u_newV), list(u_oldV)
    rev_assigner.assign(u_old, variable_list_old)
    rev_assigner.assign(u_new, variable_list_new)
    F = action(F, u_new)
    J = derivative(F, u_new, u)
    problem = Problem(J, F, bc)
    nonlinear_solver = PETScSNESSolver()
    nonlinear_solver.parameters["relative_tolerance"] = 1e-4
    nonlinear_solver.parameters["linear_solver"] = linear_solver
    nonlinear_solver.parameters["maximum_iterations"] = 20

    t, steps = 0.0, 0
    while t < T_final * (1 + 1e-6):
        u_old1.assign(u_old)
        u_old.assign(u_new)
        assigner.assign(variable_list_old, u_old)
        t = adaptive_solver(nonlinear_solver, problem, t, dt, dt_old, u_new, u_old, variable_list_new, variable_list_old, assigner,
                            error, files.error_file, max_error, ttol, dt_min, time_dependent_arguments=[], approximation=approximation)
        dt_old.time_step = dt.time_step
        dt.time_step = adaptive_timestep(dt.time_step, max_error, ttol, dt_min, dt_max)
        max_error[2], max_error[1] = max_error[1], max_error[0]
        steps += 1
        if max_steps is not None and steps >= max_steps:
            break
    main.last = {"F": F, "u_new": u_new, "bcs": bc, "problem": problem, "solver": nonlinear_solver}
    return u_new, mesh


if __name__ == "__main__":
    ap = ArgumentParser(description="FEDM streamer discharge on the B200 stack")
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--mesh", type=Path, default=None, help="dolfin-XML triangle mesh (the reference's mesh.xml); default: n x n grid")
    ap.add_argument("--input", type=Path, default=Path(__file__).resolve().parent.parent / "tests" / "golden" / "streamer_discharge_input")
    ap.add_argument("--output", type=Path, default=None)
    ap.add_argument("--t-final", type=float, default=1e-10)
    a = ap.parse_args()
    main(n=a.n, T_final=a.t_final, input_dir=a.input, output_dir=a.output, mesh_file=a.mesh)
