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


def _replica(args):
    """One independent copy of the sample workload (runs in its own process, numpy threads pinned to 1)."""
    n, total, warmup = args
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"
    from .streamer import StreamerProblem, streamer_mesh
    prob = StreamerProblem(streamer_mesh(n))
    times = []
    t0 = time.perf_counter()
    prob.run(max_steps=total, callback=lambda step, t, x: times.append(time.perf_counter()))
    per = np.diff([t0] + times)
    # the first interval contains the initial Poisson solve + pattern set-up: excluded when there is a later step
    used = per[max(warmup, 1 if len(per) > 1 else 0):] if len(per) > 1 else per
    return float(np.mean(used)), len(used)


def run(steps=1, warmup=0, full_n=1825, sample_n=0, procs=0):
    """CPU arm.  The oracle's sparse LU is single-threaded, so ALL host cores are used the way ``mpirun -np <cores>`` of
    independent jobs would: one replica of the sample workload per core, throughput summed (no communication cost is
    charged to the CPU side, which favours it)."""
    full_dof = 3 * (full_n + 1) ** 2
    lib = _c_port()
    if lib is not None:
        from . import streamer_c
        return streamer_c.timed_steps(lib, steps, warmup, full_n, sample_n or 512)
    import multiprocessing as mp
    n = sample_n or 96
    total = max(1, steps) + warmup
    cores = procs or max(1, min(os.cpu_count() or 1, 64))
    if cores > 1:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_replica, [(n, total, warmup)] * cores)
    else:
        res = [_replica((n, total, warmup))]
    sps_sample = sum(1.0 / r[0] for r in res)                      # aggregate steps/s over the replicas
    dof = 3 * (n + 1) ** 2
    return {"value": sps_sample * dof / full_dof, "cores": cores, "kind": "port",
            "sample": f"numpy/scipy oracle, {cores} independent replica(s) (one per host core), {res[0][1]} BDF2 step(s) each on a "
                      f"{n}x{n} grid ({dof} DOF, exact sparse LU = the reference examples' 'mumps' mode), aggregate "
                      f"throughput scaled to {full_dof} DOF by the DOF ratio; host has {os.cpu_count()} cores"}
