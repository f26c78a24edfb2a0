"""CPU baseline for bench.py (oracle -- test infrastructure / reported baseline only, never the product path).

The literal reference (FEniCS 2019.1.0 + PETSc under mpirun) cannot run in this image (no dolfin/PETSc/MPI,
no network), so the reported CPU number is the oracle port timed on the box's host cores
on a BOUNDED sample of the bench workload: the same streamer BDF2 step on a smaller n x n grid, scaled to the
full mesh by the DOF ratio (assumes linear cost in DOF, which favours the CPU: Krylov/LU costs grow faster).
Preference order: the compiled C/OpenMP restatement (oracle/csrc, kind "port", all cores) when it has been
built, else the numpy/scipy oracle (kind "port", 1 core).
"""
import ctypes
import os
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def _c_port():
    so = HERE / "csrc" / "libstreamer_ref.so"
    if not so.exists():
        return None
    try:
        return ctypes.CDLL(str(so))
    except OSError:
        return None


def run(steps=1, warmup=0, full_n=1825, sample_n=0):
    full_dof = 3 * (full_n + 1) ** 2
    lib = _c_port()
    if lib is not None:
        from . import streamer_c
        return streamer_c.timed_steps(lib, steps, warmup, full_n, sample_n or 512)
    from .streamer import StreamerProblem, streamer_mesh
    n = sample_n or 96
    prob = StreamerProblem(streamer_mesh(n))
    total = max(1, steps) + warmup
    times = []

    def cb(step, t, x):
        times.append(time.perf_counter())

    t0 = time.perf_counter()
    prob.run(max_steps=total, callback=cb)           # includes set-up; per-step times from the callback
    marks = [t0] + times
    # the first interval contains the initial Poisson solve + pattern set-up: excluded when there is a later step
    per = np.diff(marks)
    used = per[max(warmup, 1 if len(per) > 1 else 0):] if len(per) > 1 else per
    sps_sample = 1.0 / float(np.mean(used))
    dof = 3 * (n + 1) ** 2
    return {"value": sps_sample * dof / full_dof, "cores": 1, "kind": "port",
            "sample": f"numpy/scipy oracle, {len(used)} BDF2 step(s) on a {n}x{n} grid ({dof} DOF, exact sparse LU = the "
                      f"reference examples' 'mumps' mode), scaled to {full_dof} DOF by the DOF ratio; "
                      f"host has {os.cpu_count()} cores, 1 used"}
