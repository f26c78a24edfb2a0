"""Short driver for `ncu --set full` captures: one warm BDF2 step of the bench workload without the set-up Poisson solve
(the initial potential is the Laplace field), so that the first launches ncu sees are the streamer kernels."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np

from fedm_b200.streamer import BOX, U_W, StreamerCase

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1825
case = StreamerCase(n=n)
ui, ue = case.initial_log_densities()
x0 = np.stack([ui, ue, U_W * case.mesh.coords[:, 1] / BOX], axis=1).ravel()
case.set_initial_state(x0=x0)
case.step()
print("step done:", case.stats[-1])
