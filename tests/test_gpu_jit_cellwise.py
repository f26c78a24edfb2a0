"""The two assembly paths of a generated plug-in -- cell-wise (every cell evaluated once into the scratch records, rows gather:
fedm_b200/csrc/row_kernel.cuh) and row-wise (FEDMX_JIT_CELLS=0) -- must give the same Newton trajectory on the reference's
streamer script (tests/integrated_tests/streamer_discharge/fedm_streamer.py:296-353); each runs in its own process because
the library reads the switch once."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = Path(__file__).resolve().parent.parent
SCRIPT = """
import sys
from pathlib import Path
root = Path(sys.argv[1])
sys.path[:0] = [str(root), str(root / "compat"), str(root / "examples")]
import numpy as np
import dolfin
import streamer_discharge as sd
u, mesh = sd.main(n=40, input_dir=root / "tests" / "golden" / "streamer_discharge_input", output_dir=Path(sys.argv[2]),
                  linear_solver="mumps", max_steps=3)
np.save(sys.argv[3], u.vector()[:])
"""


def test_cellwise_and_rowwise_generated_assembly_agree(tmp_path):
    out = {}
    for mode in ("1", "0"):
        env = dict(os.environ, FEDMX_JIT_CELLS=mode)
        npy = tmp_path / f"u{mode}.npy"
        r = subprocess.run([sys.executable, "-c", SCRIPT, str(ROOT), str(tmp_path / f"out{mode}"), str(npy)], env=env,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[mode] = np.load(npy)
    for k in range(3):
        a, b = out["1"][k::3], out["0"][k::3]
        assert np.abs(a - b).max() <= 1e-10 * np.abs(b).max(), k
