"""Wall-clock of the parity configurations of BASELINE.json on the GPU (they are tests, not bench lines; this only records how
long the reference's own integrated-test cases take on the device path).  Run on a GPU box."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from fedm_b200.glow import GlowCase
from fedm_b200.streamer import StreamerCase
from fedm_b200.tof import TofCase


def timed(label, make, run):
    t0 = time.time()
    case = make()
    torch.cuda.synchronize()
    t1 = time.time()
    out = run(case)
    torch.cuda.synchronize()
    t2 = time.time()
    print(f"{label}: set-up {t1 - t0:.2f} s, run {t2 - t1:.2f} s {out}", flush=True)


golden = ROOT / "tests" / "golden"
timed("time_of_flight 40x40, 100 BDF2 steps (reference test)", lambda: TofCase(), lambda c: (c.run(), "")[1])
timed("glow_discharge 100x100 crossed, 101 005 DOF, 6 adaptive steps (reference test)", lambda: GlowCase(golden / "glow_discharge_input", ksp_rtol=1e-12),
      lambda c: (c.run(), f"{len(c.rows)} steps, Newton its {[s[0] for s in c.stats]}")[1])
def streamer():
    c = StreamerCase(n=256)
    c.set_initial_state()
    return c
timed("streamer 256x256, 198 147 DOF, T_final 1e-10 (reference test semantics)", streamer, lambda c: f"{c.run()} steps")
