"""How fast is the USER-SCRIPT path (dolfin shim + form compiler + libfedmx) on the bench mesh?  Runs examples/streamer_discharge.py
unchanged on an n x n grid and times every `nonlinear_solver.solve` call (one BDF2 attempt).  Run on a GPU box:
    python tools/time_script_path.py 1825 4"""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "compat"), str(ROOT / "examples")]
import torch

import dolfin

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
times = []
orig = dolfin.PETScSNESSolver.solve


def timed_solve(self, problem, x):
    torch.cuda.synchronize()
    t0 = time.time()
    out = orig(self, problem, x)
    torch.cuda.synchronize()
    times.append((time.time() - t0, out[0]))
    return out


dolfin.PETScSNESSolver.solve = timed_solve
import streamer_discharge as sd

t0 = time.time()
with tempfile.TemporaryDirectory() as tmp:
    sd.main(n=n, input_dir=ROOT / "tests" / "golden" / "streamer_discharge_input", output_dir=Path(tmp), linear_solver="gmres", max_steps=steps)
total = time.time() - t0
dof = 3 * (n + 1) ** 2
warm = [t for t, _ in times[1:]] or [times[0][0]]
print(f"script path, n={n} ({dof} DOF): total {total:.1f} s (set-up, JIT load, AMG hierarchy, {len(times)} solves); "
      f"solve calls {[round(t * 1e3, 1) for t, _ in times]} ms, Newton its {[k for _, k in times]}; "
      f"warm solve {1e3 * sum(warm) / len(warm):.1f} ms -> {len(warm) / sum(warm):.1f} solves/s", flush=True)
