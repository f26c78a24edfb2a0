"""Generates tests/golden/tof_160x320.npz: the CPU oracle's run of BASELINE configs[1] = examples/time_of_flight/fedm-tof.py
(160 x 320 'right' grid of [0,5e-4] x [0,1e-3], 51 681 DOF, dt = 1e-12, t0 = 2.5e-9 -> T_final = 3e-9: 500 BDF2 steps;
reference: examples/time_of_flight/fedm-tof.py:57-61, 75-76, 87).

The oracle needs ~10 s per step, so the outputs are committed as a fixture (checkpoints after 20 and after 500 steps):
  idx, u20 / u500     ln n_e at every 3rd node
  norms20 / norms500  [l2 norm, max |.|, sum] over ALL nodes
  rel_l2_500          the reference's figure of merit: relative L2 error of project(exp(u)) against the analytic solution
  newton_its          Newton iterations per step

    nice python tools/make_golden_tof160.py            # build container only; pure oracle, no GPU
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle.tof import TofProblem  # noqa: E402

OUT = ROOT / "tests" / "golden" / "tof_160x320.npz"


def norms(x):
    return np.array([np.linalg.norm(x), np.abs(x).max(), x.sum()])


def main():
    t0 = time.time()
    p = TofProblem(160, 320, 5e-4, 1e-3)
    x20, its20 = p.run(T_final=2.52e-9)
    print("20 steps", time.time() - t0, flush=True)
    idx = np.arange(0, x20.size, 3)
    np.savez_compressed(OUT, idx=idx, u20=x20[idx], norms20=norms(x20), newton_its=np.array(its20))
    p = TofProblem(160, 320, 5e-4, 1e-3)
    x, its = p.run(T_final=3e-9)
    _, _, rel = p.postprocess(x)
    np.savez_compressed(OUT, idx=idx, u20=x20[idx], norms20=norms(x20), u500=x[idx], norms500=norms(x),
                        rel_l2_500=rel, newton_its=np.array(its))
    print("500 steps", time.time() - t0, "rel L2", rel, flush=True)


if __name__ == "__main__":
    main()
