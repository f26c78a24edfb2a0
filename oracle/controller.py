"""Step-size controllers, the BDF2 error estimator and the accept/reject loop
(oracle -- test infrastructure only).

Restates fedm/functions.py:915-927 (PID), :930-937 (PI.3.4), :940-951 (H211b), the estimator at
:1062-1064 and the retry logic at :1077-1127 of the reference.  The PID controller is pinned
bit-exactly by both golden ``relative error.log`` files (tests/test_oracle_controller.py).
"""
import numpy as np

DOLFIN_EPS = 3.0e-16


def adaptive_timestep(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    e0, e1, e2 = error[0], error[1], error[2]
    fac = (e1 / e0) ** 0.075 * (tol / e0) ** 0.175 * (e1 ** 2 / (e0 * e2)) ** 0.01
    dt = dt * fac
    return max(min(dt, dt_max), dt_min)


def adaptive_timestep_PI34(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    dt = dt * ((0.8 * tol / error[0]) ** (0.3 / 3) * (0.8 * error[1] / error[0]) ** (0.4 / 3))
    return max(min(dt, dt_max), dt_min)


def adaptive_timestep_H211b(dt, dt_old, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    dt = dt * ((0.8 * tol / error[0]) ** (1 / 12) * (0.8 * tol / error[1]) ** (1 / 12)
               * (dt / dt_old) ** (-1 / 4))
    return max(min(dt, dt_max), dt_min)


def error_estimate(v_new, v_old):
    """l2 ratio with DOLFIN_EPS added entry-wise, exactly as ``df.norm(a - b + EPS)/df.norm(b + EPS)``."""
    num = np.linalg.norm(v_new - v_old + DOLFIN_EPS)
    den = np.linalg.norm(v_old + DOLFIN_EPS)
    return float(num / den)


def format_error_row(err, dt_old, dt):
    """Byte-compatible row of ``relative error.log`` (fedm/functions.py:1068)."""
    return f"{err:<23}  {dt_old:<23}  {dt:<23}\n"


class SolveFailed(RuntimeError):
    """What PETScSNESSolver raises on non-convergence (error_on_nonconvergence=True)."""


def adaptive_attempts(solve_fn, reset_fn, t, dt, ttol, dt_min, max_error, log_rows=None, dt_old=None):
    """The accept/reject recursion of ``adaptive_solver`` written as a loop.

    solve_fn(t_new, dt) -> error estimate (raises SolveFailed on Newton failure);
    reset_fn() restores u_new <- u_old.  Returns (t_new, dt_used, n_attempts)."""
    attempts = 0
    while True:
        attempts += 1
        t_new = t + dt
        try:
            err = solve_fn(t_new, dt)
            if log_rows is not None:
                log_rows.append((err, dt_old, dt))
            max_error[0] = max([err])
            if err >= ttol:
                dt = dt * (0.5 * ttol / max_error[0])
                failed = True
            else:
                failed = False
        except SolveFailed:
            dt = dt * 0.5
            failed = True
        if not failed:
            return t_new, dt, attempts
        if dt < dt_min:
            raise SystemExit("Minimum time-step size reached, program is terminating.")
        reset_fn()
