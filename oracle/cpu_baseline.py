"""CPU baseline for bench.py (oracle -- test infrastructure / reported baseline only, never the product path).

The literal reference (FEniCS 2019.1.0 + PETSc under mpirun) cannot run in this image (no dolfin/PETSc/MPI, no network;
SURVEY.md section 8c).  The reported CPU number is the compiled C++/OpenMP restatement of the same hot path
(oracle/csrc/streamer_ref.cpp via oracle/streamer_c.py): owner-computes assembly and GMRES(30) + block-Jacobi/ILU(0) -- the
reference test's KSP (fedm_streamer.py:32) with PETSc's parallel default PC -- on ALL host cores, one OpenMP thread per
"rank".  It is parity-checked against the numpy oracle (tests/test_oracle_cport.py).

Two entry points:
  full(...)    bench.py --impl reference: the bench configuration ITSELF (n = 1825, 10.0 M DOF), timed per accepted BDF2
               step; the start state is the oracle's (IC + sparse-LU initial Poisson solve, fedm_streamer.py:194-223)
  sample(...)  the `cpu_baseline` object of the GPU line: a bounded sample (~10-30 s): one BDF2 step each on two smaller grids
               of the same family, extrapolated to the bench mesh with the MEASURED cost exponent (not assumed linear)
"""
import math
import os
import time

import numpy as np


def _initial_state(mesh):
    from .streamer import StreamerProblem
    return StreamerProblem(mesh).initial_state()


def _time_steps(n, steps, warmup, threads, budget_s):
    from . import streamer_c
    from .streamer import streamer_mesh
    t0 = time.perf_counter()
    mesh = streamer_mesh(n)
    x0 = _initial_state(mesh)
    prob = streamer_c.StreamerC(mesh, threads)
    setup_s = time.perf_counter() - t0
    marks = []
    t_start = time.perf_counter()

    def on_step(k, t, x):
        marks.append(time.perf_counter())
        return not (k >= warmup + 1 and marks[-1] - t_start > budget_s)

    _, rows, stats = prob.run(x0, max_steps=warmup + steps, on_step=on_step)
    per = np.diff([t_start] + marks)
    timed = per[warmup:] if len(per) > warmup else per[-1:]
    out = {"s_per_step": float(np.mean(timed)), "timed_steps": int(len(timed)), "warmup_done": int(len(per) - len(timed)),
           "cores": prob.threads, "setup_s": setup_s, "dof": 3 * mesh.num_vertices,
           "newton_its_per_step": float(np.mean([s[0] for s in stats[-len(timed):]])),
           "krylov_its_per_step": float(np.mean([s[1] for s in stats[-len(timed):]])), "phase_seconds_total": prob.timers()}
    prob.close()
    return out


def full(steps=1, warmup=0, n=1825, threads=0, budget_s=200.0):
    """The bench configuration itself on all host cores.  Warm-up is capped at one step and the run stops at the wall-clock
    budget (one step at 10 M DOF takes minutes on a CPU): the number of steps actually timed is returned and printed."""
    r = _time_steps(n, steps, min(warmup, 1), threads, budget_s)
    r["value"] = 1.0 / r["s_per_step"]
    r["kind"] = "port"
    r["sample"] = (f"C++/OpenMP restatement (oracle/csrc/streamer_ref.cpp) on the bench mesh itself, n={n} ({r['dof']} DOF): "
                   f"{r['cores']} threads = {r['cores']} 'ranks' (owner-computes assembly, GMRES(30) rtol 1e-5 + block-Jacobi/ILU(0) = the "
                   f"reference test's KSP under mpirun); {r['timed_steps']} timed accepted BDF2 step(s) after {r['warmup_done']} warm-up "
                   f"step(s) ({steps} requested; wall-clock budget {budget_s:.0f} s), {r['krylov_its_per_step']:.0f} Krylov and "
                   f"{r['newton_its_per_step']:.1f} Newton iterations per step; set-up incl. the sparse-LU initial Poisson solve "
                   f"{r['setup_s']:.0f} s (untimed); host has {os.cpu_count()} cores")
    return r


def sample(full_n=1825, sizes=(256, 512), threads=0):
    """Bounded sample for the GPU line's `cpu_baseline`: one timed step (after one warm-up step) on each of two smaller grids;
    the cost exponent p in t ~ DOF^p is MEASURED from the two and used to place the number on the bench mesh."""
    rs = [_time_steps(n, 1, 1, threads, 1e9) for n in sizes]
    full_dof = 3 * (full_n + 1) ** 2
    p = math.log(rs[1]["s_per_step"] / rs[0]["s_per_step"]) / math.log(rs[1]["dof"] / rs[0]["dof"])
    t_full = rs[1]["s_per_step"] * (full_dof / rs[1]["dof"]) ** p
    return {"value": 1.0 / t_full, "cores": rs[1]["cores"], "kind": "port", "exponent": p,
            "measured": [{"n": n, "dof": r["dof"], "s_per_step": r["s_per_step"], "krylov_its_per_step": r["krylov_its_per_step"]}
                         for n, r in zip(sizes, rs)],
            "sample": f"C++/OpenMP restatement (oracle/csrc/streamer_ref.cpp), {rs[1]['cores']} threads, GMRES(30) rtol 1e-5 + "
                      f"block-Jacobi/ILU(0): one accepted BDF2 step (after one warm-up step) on the n={sizes[0]} and n={sizes[1]} grids "
                      f"({rs[0]['s_per_step']:.2f} s, {rs[1]['s_per_step']:.2f} s; {rs[0]['krylov_its_per_step']:.0f} / "
                      f"{rs[1]['krylov_its_per_step']:.0f} Krylov iterations), extrapolated to {full_dof} DOF with the measured exponent "
                      f"t ~ DOF^{p:.2f}; `bench.py --impl reference` times the full mesh itself; host has {os.cpu_count()} cores"}
