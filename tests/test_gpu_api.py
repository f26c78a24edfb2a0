"""The reference's own time loop (tests/integrated_tests/streamer_discharge/fedm_streamer.py:312-348) written against the
reference-named API of fedm_b200 (functions.adaptive_solver / adaptive_timestep, Problem, PETScSNESSolver,
FunctionAssigner, files.error_file) on the GPU; results must equal the oracle's."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200 import dolfin_shim as df
from fedm_b200 import file_io
from fedm_b200.functions import Problem, adaptive_solver, adaptive_timestep
from fedm_b200.streamer import StreamerCase, U_W


def test_streamer_time_loop_through_reference_api(tmp_path):
    from oracle.streamer import StreamerProblem, streamer_mesh
    n = 24
    case = StreamerCase(n=n)                       # mesh, model, Dirichlet set, preconditioner (the "form compilation" step)
    case.set_initial_state()
    hot = case.hot
    file_io.files.output_folder_path = tmp_path
    # ---- objects the reference script creates (fedm_streamer.py:141-172, 284-305)
    ME, V = df.FunctionSpace(hot, 3), df.FunctionSpace(hot, 1)
    u_new, u_old, u_old1 = df.Function(ME), df.Function(ME), df.Function(ME)
    u_new.vector().t.copy_(case.x_new)
    u_old.vector().t.copy_(case.x_new)
    assigner = df.FunctionAssigner([V, V, V], ME)
    variable_list_new = [df.Function(V) for _ in range(3)]
    variable_list_old = [df.Function(V) for _ in range(3)]
    dt = df.Expression("time_step", time_step=5e-12, degree=0)
    dt_old = df.Expression("time_step", time_step=1e30, degree=0)
    problem = Problem(hot, dt, dt_old, u_old, u_old1, n_hist=2, bc_values=lambda t: [0.0, U_W])
    nonlinear_solver = df.PETScSNESSolver()
    nonlinear_solver.parameters["relative_tolerance"] = 1e-4
    nonlinear_solver.parameters["linear_solver"] = "mumps"
    nonlinear_solver.parameters["maximum_iterations"] = 20
    t, T_final, ttol, dt_min, dt_max = 0.0, 2.4e-11, 1e-3, 1e-15, 5e-12
    error, max_error = [0.0] * 2, [1] * 3
    # ---- time loop, verbatim structure of fedm_streamer.py:312-348
    while t < T_final * (1 + 1e-6):
        u_old1.assign(u_old)
        u_old.assign(u_new)
        assigner.assign(variable_list_old, u_old)
        t = adaptive_solver(nonlinear_solver, problem, t, dt, dt_old, u_new, u_old, variable_list_new, variable_list_old,
                            assigner, error, file_io.files.error_file, max_error, ttol, dt_min,
                            time_dependent_arguments=[], approximation="LFA")
        dt_old.time_step = dt.time_step
        dt.time_step = adaptive_timestep(dt.time_step, max_error, ttol, dt_min, dt_max)
        max_error[2] = max_error[1]
        max_error[1] = max_error[0]
    rows = [[float(v) for v in ln.split()] for ln in (tmp_path / "relative error.log").read_text().splitlines()]
    prob = StreamerProblem(streamer_mesh(n))
    x_ref, rows_ref, _ = prob.run(T_final=T_final)
    assert len(rows) == len(rows_ref) == 5
    for (e, a, b), (er, ar, br) in zip(rows, rows_ref):
        assert a == ar and b == br and e == pytest.approx(er, rel=1e-7)
    x = u_new.vector().get_local()
    for k in range(3):
        assert np.abs(x[k::3] - x_ref[k::3]).max() <= 1e-8 * np.abs(x_ref[k::3]).max()
    # per-field views were filled by the assigner
    assert np.array_equal(variable_list_new[1].vector().get_local(), x[1::3])
    assert df.norm(u_new.vector()) == pytest.approx(np.linalg.norm(x), rel=1e-13)
