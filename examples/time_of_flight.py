"""Electron swarm time of flight (2-D axisymmetric, one log-density unknown) on the FEDM API; analytic solution known.
Cf. examples/time_of_flight/fedm-tof.py of the reference (this version uses the reference *test's* box, mesh and end time).
Run:  PYTHONPATH=compat python examples/time_of_flight.py
"""
import sys
from pathlib import Path

sys.path[:0] = [str(Path(__file__).resolve().parent.parent), str(Path(__file__).resolve().parent.parent / "compat")]

from dolfin import *                       # noqa: E402,F401,F403
from fedm.physical_constants import *      # noqa: E402,F401,F403
from fedm.file_io import *                 # noqa: E402,F401,F403
from fedm.functions import *               # noqa: E402,F401,F403


def main(output_dir=None, nx=40, ny=40, width=2.5e-4, height=5e-4, t0=2.5e-9, T_final=2.6e-9, dt_init=1e-12):
    if output_dir is not None:
        files.output_folder_path = Path(output_dir)
    wez, De, alpha_e = 1.7e5, 0.12, 5009.51
    dt = Expression("time_step", time_step=dt_init, degree=0)
    dt_old = Expression("time_step", time_step=1e30, degree=0)
    r = Expression("x[0]", degree=1)
    mesh = RectangleMesh(Point(0, 0), Point(width, height), nx, ny)
    V = FunctionSpace(mesh, "P", 1)
    W = VectorFunctionSpace(mesh, "P", 1)
    u, v = TrialFunction(V), TestFunction(V)
    u_old, u_old1, u_new = Function(V), Function(V), Function(V)
    t = t0
    swarm = "exp(-(pow(x[1]-w*t, 2)+pow(x[0], 2))/(4.0*D*t)+alpha*w*t)"
    kw = dict(D=De, w=wez, alpha=alpha_e, t=t, pi=pi)
    u_analytical = Expression(f"std::log({swarm}/pow(4*D*t*pi,1.5))", degree=3, **kw)
    u_old.assign(interpolate(u_analytical, V))
    u_old1.assign(interpolate(u_analytical, V))
    w = interpolate(Constant(("0", wez)), W)
    D = interpolate(Constant(De), V)
    Gamma = -grad(D * exp(u)) + w * exp(u)
    f = Expression(f"{swarm}*(w*alpha)/(8*pow(pi,1.5)*pow(D*t, 1.5))", degree=2, **kw)
    F = weak_form_balance_equation_log_representation("drift-diffusion-reaction", dt, dt_old, dx, u, u_old, u_old1, v, f, Gamma, r)
    u_new.assign(interpolate(Expression(f"std::log({swarm}/pow(4.0*D*t*pi,1.5) + DOLFIN_EPS)", degree=2, **kw), V))
    F = action(F, u_new)
    problem = Problem(derivative(F, u_new, u), F, [])
    solver = PETScSNESSolver()
    solver.parameters["relative_tolerance"] = 1e-10
    solver.parameters["linear_solver"] = "mumps"
    solver.parameters["maximum_iterations"] = 50
    while abs(t - T_final) / T_final > 1e-6:
        u_old1.assign(u_old)
        u_old.assign(u_new)
        t += dt.time_step
        f.t = t
        u_analytical.t = t
        solver.solve(problem, u_new.vector())
        if t > t0 + dt_init:
            dt_old.time_step = dt.time_step
    n_exact = project(exp(u_analytical), V, solver_type="mumps")
    n_num = project(exp(u_new), V, solver_type="mumps")
    relative_error = errornorm(n_num, n_exact, "l2") / norm(n_exact, "l2")
    with open(files.error_file, "a") as fh:
        fh.write(f"h_max = {mesh.hmax()}\t dt = {dt.time_step}\t relative_error = {relative_error}\n")
    return n_num, relative_error


if __name__ == "__main__":
    print(main()[1])
