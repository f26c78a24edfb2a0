"""Generates tests/golden/streamer_n256_T1e-10.npz: the CPU oracle's run of the parity configuration that SURVEY.md 8(d) /
BASELINE.md name for the streamer case (n = 256 'right' grid, 198 147 DOF, T_final = 1e-10 => 21 accepted steps, exact
linear solves = the reference examples' 'mumps' mode; time loop of
tests/integrated_tests/streamer_discharge/fedm_streamer.py:312-353).

The oracle needs ~10 minutes for this run, too slow for the test-suite, so its outputs are committed as a fixture:
  rows        the 21 rows of `relative error.log` (error, dt_old, dt)
  newton_its  Newton iterations per attempt
  idx, sample state after the last step at every 5th node (all three components), FP64
  norms       per component: [l2 norm, max |.|, sum] over ALL nodes (checks the un-sampled nodes too)
Also writes the reject-path fixtures (forced rejections with a tiny ttol and forced Newton failures with max_it = 1 at
n = 32; control flow of fedm/functions.py:1077-1127).

    python tools/make_golden_streamer256.py            # run once in the build container; pure oracle, no GPU
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.streamer import StreamerProblem, streamer_mesh  # noqa: E402

OUT = ROOT / "tests" / "golden"


def main(n=256, T_final=1e-10):
    t0 = time.time()
    prob = StreamerProblem(streamer_mesh(n))
    x, rows, its = prob.run(T_final=T_final, callback=lambda s, t, x: print(f"step {s} t={t:.4e} ({time.time() - t0:.0f} s)", flush=True))
    X = x.reshape(-1, 3)
    idx = np.arange(0, X.shape[0], 5)
    norms = np.array([[np.linalg.norm(X[:, k]), np.abs(X[:, k]).max(), X[:, k].sum()] for k in range(3)])
    np.savez_compressed(OUT / f"streamer_n{n}_T1e-10.npz", rows=np.array(rows, dtype=np.float64),
                        newton_its=np.array([h[0] for h in its]), idx=idx, sample=X[idx], norms=norms, n=n)
    print("wrote", OUT / f"streamer_n{n}_T1e-10.npz", len(rows), "rows", time.time() - t0, "s")


def reject_fixtures(n=32):
    """Forced rejections: ttol = 2e-5 is below the first step's estimate (7.4e-5 at dt = 5e-12 on this grid), so attempts are
    rejected and dt shrinks by 0.5*ttol/err (fedm/functions.py:1085-1096; two rejections in 4 steps); forced Newton
    failure: max_it = 1 makes the first six solves fail (dolfin raises, fedm halves dt, :1094-1096) until the step is small
    enough (5e-12/64) for one iteration to reach rtol 1e-4."""
    out = {}
    prob = StreamerProblem(streamer_mesh(n))
    x, rows, its = prob.run(ttol=2e-5, max_steps=4)
    out["ttol_rows"] = np.array(rows, dtype=np.float64)
    out["ttol_state"] = x
    out["ttol_newton"] = np.array([h[0] for h in its])
    prob = StreamerProblem(streamer_mesh(n))
    x, rows, its = prob.run(max_it=1, max_steps=3)
    out["maxit_rows"] = np.array(rows, dtype=np.float64)
    out["maxit_state"] = x
    out["maxit_newton"] = np.array([h[0] for h in its])
    np.savez_compressed(OUT / f"streamer_n{n}_reject.npz", **out)
    print("reject fixtures:", len(out["ttol_rows"]), "rows (ttol),", len(out["maxit_rows"]), "rows (max_it)")
    print(out["ttol_rows"])
    print(out["maxit_rows"], out["maxit_newton"])


if __name__ == "__main__":
    if "--reject" in sys.argv:
        reject_fixtures()
    else:
        main()
