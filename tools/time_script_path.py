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
asm = {}
orig = dolfin.PETScSNESSolver.solve


def timed_solve(self, problem, x):
    torch.cuda.synchronize()
    t0 = time.time()
    out = orig(self, problem, x)
    torch.cuda.synchronize()
    times.append((time.time() - t0, out[0]))
    if len(times) == steps:                                      # last solve: time the generated assembly kernels alone
        prob = next(iter(self._compiled.values()))[1]
        xt = prob._x_local if prob.local is not None else x.function.vector().store.t
        for name, fn in (("jacobian", lambda: prob.hot.jacobian(xt, 1.0, 1.0)), ("residual", lambda: prob.hot.residual(xt, 1.0, 1.0))):
            fn()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            torch.cuda.synchronize()
            ev[0].record()
            for _ in range(10):
                fn()
            ev[1].record()
            torch.cuda.synchronize()
            asm[name] = ev[0].elapsed_time(ev[1]) / 10
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
      f"warm solve {1e3 * sum(warm) / len(warm):.1f} ms -> {len(warm) / sum(warm):.1f} solves/s; generated assembly kernels "
      f"(incl. block-Jacobi update) {', '.join(f'{k} {v:.2f} ms' for k, v in asm.items())}", flush=True)
