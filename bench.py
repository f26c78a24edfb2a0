#!/usr/bin/env python
"""Headline benchmark: BDF2 steps/s (+ Jacobian-assembly GDOF/s) of the streamer case on a synthetic
~10 M-DOF mesh (BASELINE.json `metric`), one process per GPU.

    python bench.py --gpus N --steps K --warmup W            # this implementation
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on the host cores

One JSON line on stdout (rank 0).  A "step" is one accepted BDF2 time step of
tests/integrated_tests/streamer_discharge/fedm_streamer.py:312-348 (state shift, Newton solve with
GMRES(30)+block-Jacobi at PETSc's rtol 1e-5 -- the reference *test's* solver semantics --, error estimate,
PID controller).  Inputs are larger than L2 (matrix 1.7 GB, vectors 80 MB each), no L2 flush needed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FULL_N = 1825            # n x n "right" grid -> 3 334 276 nodes, 10 002 828 DOF


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled every 10 ms from a thread of this process (the
    timed region of a multi-GPU run lasts a few hundred ms -- too short for a freshly spawned `nvidia-smi -lms` to report);
    falls back to `nvidia-smi` if the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.stop_flag, self.th, self.proc, self.lines = False, None, None, []

    def _poll(self):
        import pynvml as nv
        h = None
        try:                                               # the CUDA ordinal may differ from NVML's (CUDA_VISIBLE_DEVICES): use the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(get_reasons(h))
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.th = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.rd = threading.Thread(target=lambda: [self.lines.append(ln.strip()) for ln in self.proc.stdout], daemon=True)
            self.rd.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.th is not None:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "NVML, 10 ms period, timed region only"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 50"}


def algorithmic_bytes(nv, nc, n_eq=3, zero_entries=1):
    """SURVEY 8(d), for the storage format actually used: SELL-32 node blocks whose structurally zero entry
    (d R_e / d U_i of the streamer system: 1 of 9) is neither written by the assembly nor read by the SpMV."""
    ne = nv + nc - 1                                  # Euler (planar, one boundary component)
    nnzb = nv + 2 * ne
    nnz = n_eq * n_eq * nnzb
    stored = nnz - zero_entries * nnzb
    asm = 12 * nc + 16 * nv + 3 * 8 * n_eq * nv + 8 * stored + 8 * n_eq * nv
    # Krylov SpMV = fused  z = Dinv (A v)  + extraction of the multigrid right-hand side: matrix values + column indices +
    # slice table + v and z + the node-block inverses (n_eq^2 doubles per node) + one extracted component
    spmv = 8 * stored + 4 * nnzb + 4 * nv + 16 * n_eq * nv + 8 * n_eq * n_eq * nv + 8 * nv
    return asm, spmv, nnz


def workload_config(n, ny=None):
    """The workload both arms run -- identical dict in the GPU line and in the `--impl reference` line."""
    ny = n if ny is None else ny
    nv, nc = (n + 1) * (ny + 1), 2 * n * ny
    shape = f"n={n}" if ny == n else f"{n}x{ny} cells"
    return {"workload": f"synthetic streamer mesh {shape} (right grid), one accepted BDF2 step, Newton rtol 1e-4, GMRES(30) rtol 1e-5 "
                        "(reference test semantics)",
            "dof": 3 * nv, "cells": nc, "nnz": 9 * (nv + 2 * (nv + nc - 1)), "l2": "inputs larger than L2 (no flush)"}


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the compiled restatement of the reference's path (oracle/csrc) on all host cores, on the bench mesh itself."""
    from oracle import cpu_baseline
    res = cpu_baseline.full(steps=args.steps, warmup=args.warmup, n=args.n, budget_s=args.ref_budget_s)
    line = {"impl": "reference", "metric": "streamer BDF2 steps/s (synthetic 10.0M-DOF mesh)", "value": res["value"],
            "unit": "steps/s", "n_gpus": args.gpus, "steps": res["timed_steps"], "warmup": res["warmup_done"],
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 / res["value"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(args.n),
            "solver": {"linear": "GMRES(30), classical Gram-Schmidt, left PC, rtol 1e-5", "preconditioner": "block Jacobi (one block per "
                       "thread) + ILU(0) -- PETSc's default under mpirun", "newton_its_per_step": res["newton_its_per_step"],
                       "krylov_its_per_step": res["krylov_its_per_step"], "phase_seconds_total": res["phase_seconds_total"],
                       "setup_s": res["setup_s"]},
            "cpu_baseline": {"value": res["value"], "unit": "steps/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="fedm_b200")
    ap.add_argument("--n", type=int, default=FULL_N, help="grid size (n x n cells pairs); 1825 -> 10.0M DOF")
    ap.add_argument("--ref-budget-s", type=float, default=200.0, help="wall-clock budget of the CPU arm's timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity trajectory (1-GPU / exact-solve re-run)")
    ap.add_argument("--precond", default="amg", choices=["amg", "ilu0", "bjacobi"],
                    help="amg (default): two-stage AMG + node-block Jacobi; ilu0: the reference test's literal PC (PETSc default for "
                         "linear_solver='gmres': ILU(0), block Jacobi across GPUs) -- reports its Krylov iterations per step")
    ap.add_argument("--weak", action="store_true", help="weak scaling: n*sqrt(N) x n*sqrt(N) cells of the same domain (isotropic "
                                                          "refinement), i.e. ~10 M DOF PER GPU; the default is strong scaling of the fixed n x n mesh")
    ap.add_argument("--profile-detail", action="store_true", help="nested event timing of multigrid levels and collectives")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return 0

    import numpy as np
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from fedm_b200.streamer import StreamerCase

    t_setup = time.time()
    n_base = args.n
    if args.weak:                                  # same DOF per GPU: refine both directions by sqrt(N) (cells stay isotropic)
        args.n = int(round(args.n * world ** 0.5))
    ny = args.n
    case = StreamerCase(n=args.n, ny=ny, device=local_rank, world=world, rank=rank, precond=args.precond)
    case.set_initial_state(bench_initial_potential=True)
    setup_s = time.time() - t_setup
    hot = case.hot
    nv, nc = case.global_vertices, case.global_cells
    dof = 3 * nv

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        case.step()
    # ---- timed region: K accepted steps, device-timed, max over ranks
    hot.profile_enable(2 if args.profile_detail else 1)
    hot.profile_get()
    hot.launch_count(reset=True)
    n_stats0 = len(case.stats)
    clocks = ClockSampler(local_rank)
    clocks.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    attempts = 0
    for _ in range(args.steps):
        attempts += case.step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    launches = hot.launch_count()
    prof = hot.profile_get()
    hot.profile_enable(False)
    if world > 1:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    stats = case.stats[n_stats0:]
    newton_its = sum(s[0] for s in stats)
    krylov_its = sum(s[1] for s in stats)
    # state after warmup + K accepted steps, global vertex order (gathered over the ranks): input of the parity check below
    x_timed = case.gather_state().clone()
    steps_taken = args.warmup + args.steps

    # ---- end-to-end: same steps through the host-buffer API (H2D of the three state vectors, D2H of the result)
    hb = case.make_host_buffers()
    case.step_host(hb)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        case.step_host(hb)
    torch.cuda.synchronize()          # the trailing re-upload of the last step runs on a copy stream: inside the timed region
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        tms = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms.item())

    if rank != 0:
        if world > 1:
            dist.barrier()                        # rank 0 runs the 1-GPU parity trajectory meanwhile
            dist.destroy_process_group()
        return 0

    # ---- parity of what was timed (outside every timed region).  One GPU: the same code with exact linear solves
    # (ksp_rtol 1e-13 = the examples' 'mumps' semantics) for the same number of steps; N GPUs: the un-partitioned run.
    parity = None
    if not args.no_parity:
        kw = dict(ksp_rtol=1e-13, ksp_max_it=50000) if world == 1 else {}
        ref_case = StreamerCase(n=args.n, ny=ny, device=local_rank, **kw)    # always the default preconditioner
        ref_case.set_initial_state(bench_initial_potential=True)
        for _ in range(steps_taken):                                   # (run() would stop at its default T_final)
            ref_case.step()
        Xr = ref_case.x_new.view(-1, 3)
        rel = [float((x_timed[:, k] - Xr[:, k]).abs().max() / Xr[:, k].abs().max()) for k in range(3)]
        dts = [r[2] for r in case.rows[:steps_taken]]
        dts_ref = [r[2] for r in ref_case.rows[:steps_taken]]
        parity = {"against": "same code, ksp_rtol=1e-13 (exact linear solves), 1 GPU" if world == 1 else
                             "un-partitioned 1-GPU run, same solver semantics",
                  "steps": steps_taken, "max_rel_diff": max(rel), "per_component": rel,
                  "dt_sequence_max_rel_diff": max(abs(a / b - 1.0) for a, b in zip(dts, dts_ref)),
                  "newton_counts_equal": [s_[0] for s_ in case.stats[:steps_taken]] == [s_[0] for s_ in ref_case.stats[:steps_taken]],
                  "krylov_its_per_step_ref": sum(s_[1] for s_ in ref_case.stats) / max(1, steps_taken)}
        ref_case.hot.close()
        del ref_case

    peak, peak_src = peaks()
    asm_b, spmv_b, nnz = algorithmic_bytes(nv, nc)
    asm_b_8d, _, _ = algorithmic_bytes(nv, nc, zero_entries=0)          # full 9-entry blocks, 2.133 GB at n = 1825
    # DRAM traffic / pipe utilisation of the two dominant kernels come from the committed `ncu --set full` capture of this very
    # configuration (they cannot be measured inside a timed run); the file names its source report
    ncu = {}
    ncu_file = ROOT / "profiles" / "r2_ncu_summary.json"
    if ncu_file.exists() and world == 1 and args.n == FULL_N and args.precond == "amg":
        ncu = json.loads(ncu_file.read_text())
    n_sp, ms_sp = prof["spmv"]
    n_aj, ms_aj = prof["assemble_jacobian"]
    n_ar, ms_ar = prof["assemble_residual"]
    n_or, ms_or = prof["orthogonalisation"]
    n_pc, ms_pc = prof["preconditioner"]
    # per-rank algorithmic bytes (strong scaling: the mesh is split over the ranks)
    spmv_gbs = (spmv_b / world) * n_sp / (ms_sp * 1e-3) / 1e9 if ms_sp > 0 else None
    asm_gbs = (asm_b / world) * n_aj / (ms_aj * 1e-3) / 1e9 if ms_aj > 0 else None
    line = {
        "metric": "streamer BDF2 steps/s (synthetic 10.0M-DOF mesh)",
        "value": args.steps / (ms * 1e-3), "unit": "steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.weak else "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.n, ny),
        "solver": {"linear": "GMRES(30), classical Gram-Schmidt with one reduction per iteration, left PC, rtol 1e-5",
                   "preconditioner": {"amg": "two-stage: AMG V(1,1) on the potential block + node-block Jacobi",
                                      "ilu0": "ILU(0) per GPU block (multicolour ordering), block Jacobi across GPUs: the reference test's PC",
                                      "bjacobi": "node-block Jacobi"}[args.precond],
                   "matrix_format": "SELL-32 3x3 node blocks, FP64; the structurally zero block entry (1 of 9) is kept in the "
                                    "exported dolfin pattern but not streamed",
                   "parallelism": f"vertex partition x{world}" if world > 1 else "single GPU"},
        "gdof_steps_per_s": dof * args.steps / (ms * 1e-3) / 1e9,        # throughput in unknowns advanced per second (weak-scaling figure)
        "weak_scaling": {"n_per_gpu_base": n_base, "n": args.n, "dof_per_gpu": dof / world} if args.weak else None,
        "attempts": attempts, "newton_its_per_step": newton_its / args.steps, "krylov_its_per_step": krylov_its / args.steps,
        "jacobian_assembly_gdof_s": (dof * n_aj / (ms_aj * 1e-3) / 1e9) if ms_aj > 0 else None,
        "residual_assembly_gdof_s": (dof * n_ar / (ms_ar * 1e-3) / 1e9) if ms_ar > 0 else None,
        "kernel_ms_per_step": {"spmv": ms_sp / args.steps, "orthogonalisation": ms_or / args.steps,
                               "preconditioner": ms_pc / args.steps, "assemble_jacobian": ms_aj / args.steps, "assemble_residual": ms_ar / args.steps,
                               # nested break-down: multigrid levels inside "preconditioner", collectives inside all classes
                               "mg_fine": prof["mg_fine"][1] / args.steps, "mg_mid": prof["mg_mid"][1] / args.steps,
                               "mg_coarse": prof["mg_coarse"][1] / args.steps, "comm": prof["comm"][1] / args.steps},
        "roofline": {"kernel": "spmv_kernel<3,true,true> (SELL-32 block SpMV + node-block Jacobi + multigrid RHS extraction)", "bound": "hbm",
                     "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak if spmv_gbs else None,
                     "traffic": ncu.get("spmv", {}).get("dram_bytes_per_launch"),
                     "traffic_source": ncu.get("spmv", {}).get("source"),
                     "peak_source": peak_src, "launches": n_sp,
                     "algorithmic_bytes_per_launch": spmv_b / world},
        "roofline_assembly": {"kernel": "assemble_tiled_kernel<StreamerModel,true> (cell-batch staging in shared memory)", "bound": "hbm",
                              "achieved": asm_gbs, "peak": peak, "unit": "GB/s", "frac": asm_gbs / peak if asm_gbs else None,
                              # the same launch time against SURVEY 8(d)'s byte count (all nine block entries: 213.2 B/DOF)
                              "achieved_survey_8d": asm_gbs * asm_b_8d / asm_b if asm_gbs else None,
                              "frac_survey_8d": asm_gbs * asm_b_8d / asm_b / peak if asm_gbs else None,
                              "frac_of_8tb_spec_survey_8d": asm_gbs * asm_b_8d / asm_b / 8000.0 if asm_gbs else None,
                              "traffic": ncu.get("assemble_jacobian", {}).get("dram_bytes_per_launch"),
                              "fp64_pipe_active_frac": ncu.get("assemble_jacobian", {}).get("fp64_pipe_active_frac"),
                              "traffic_source": ncu.get("assemble_jacobian", {}).get("source"),
                              "launches": n_aj, "algorithmic_bytes_per_launch": asm_b / world,
                              "algorithmic_bytes_per_launch_survey_8d": asm_b_8d / world},
        "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": "steps/s",
                "h2d_bytes_per_step": int(2 * 8 * dof), "d2h_bytes_per_step": int(8 * dof + 8)},
        "gpu_launches": launches, "clocks": clk, "setup_s": setup_s,
        "parity" if world == 1 else "parity_vs_1gpu": parity,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import cpu_baseline
            res = cpu_baseline.sample(full_n=args.n)
            line["cpu_baseline"] = {"value": res["value"], "unit": "steps/s", "cores": res["cores"], "kind": res["kind"],
                                    "sample": res["sample"], "exponent": res["exponent"], "measured": res["measured"]}
        except Exception as exc:                                # the baseline must never sink the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
