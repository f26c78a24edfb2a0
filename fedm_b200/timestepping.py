"""Adaptive BDF2 time stepping on top of the device Newton solver.

Mirrors the control flow of ``fedm.functions.adaptive_solver`` / ``adaptive_timestep``
(fedm/functions.py:915-927, 958-1130): attempt, estimate, log a row of ``relative error.log``, accept or shrink
and retry, ``SystemExit`` below ``dt_min``.  The heavy work of every attempt is one call into libfedmx.
"""
import sys

from . import hotpath as _hp
from .lib import FedmxError


class ErrorGreaterThanTTOL(Exception):
    pass


def format_error_row(err, dt_old, dt):
    return f"{err:<23}  {dt_old:<23}  {dt:<23}\n"


def attempt_until_accepted(solve_attempt, reset, t, dt, dt_old, ttol, dt_min, max_error, error_log=None):
    """solve_attempt(t_new, dt) -> relative error (raises on solver failure); reset() restores u_new <- u_old.
    Returns (t_new, dt_used, attempts)."""
    attempts = 0
    while True:
        attempts += 1
        try:
            t_new = t + dt
            err = solve_attempt(t_new, dt)
            if error_log is not None:
                error_log(format_error_row(err, dt_old, dt))
            max_error[0] = max([err, 0.0])
            if err >= ttol:
                raise ErrorGreaterThanTTOL
            return t_new, dt, attempts
        except FedmxError:
            # a library failure (CUDA / NCCL / peer-transport error, bad argument) is not a solver failure: retrying with a
            # smaller step would only hide it -- and would desynchronise the ranks of a partitioned run.  Fatal.
            raise
        except Exception as exc:                      # same catch-all as the reference (fedm/functions.py:1080)
            if isinstance(exc, ErrorGreaterThanTTOL):
                dt = dt * (0.5 * ttol / max_error[0])
            else:
                print(f"[fedm_b200] attempt with dt = {dt:g} failed ({type(exc).__name__}: {exc}); halving the step",
                      file=sys.stderr, flush=True)
                dt = dt * 0.5
            if dt < dt_min:
                raise SystemExit("Minimum time-step size reached, program is terminating.")
            reset()


def adaptive_timestep(dt, error, tol=1e-4, dt_min=1e-13, dt_max=1e-9):
    return _hp.pid_step(dt, error, tol, dt_min, dt_max)
