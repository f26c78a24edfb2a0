"""User-level scripts (examples/) written against the FEDM API run through the dolfin shim, the form compiler (UFL-lite ->
JIT CUDA row kernel) and libfedmx on the GPU; results are compared with the CPU oracle / the reference goldens."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

ROOT = Path(__file__).resolve().parent.parent
os.environ.setdefault("FEDM_B200_QUIET", "1")


@pytest.fixture()
def cuda_shim():
    sys.path[:0] = [str(ROOT / "compat"), str(ROOT / "examples")]
    import dolfin
    dolfin.set_backend(None)               # default backend = CUDA
    dolfin.parameters["form_compiler"]["quadrature_degree"] = None   # scripts set this global; each test = a fresh process
    yield dolfin
    dolfin.set_backend(None)
    dolfin.parameters["form_compiler"]["quadrature_degree"] = None


def test_streamer_script_on_gpu_matches_oracle(cuda_shim, tmp_path, golden_dir):
    import streamer_discharge as sd
    from oracle.streamer import StreamerProblem, streamer_mesh
    n = 24
    u, mesh = sd.main(n=n, input_dir=golden_dir / "streamer_discharge_input", output_dir=tmp_path, linear_solver="mumps", max_steps=4)
    prob = StreamerProblem(streamer_mesh(n))
    x_ref, rows_ref, _ = prob.run(max_steps=4)
    x = u.vector()[:]
    for k in range(3):
        assert np.abs(x[k::3] - x_ref[k::3]).max() <= 1e-8 * np.abs(x_ref[k::3]).max(), k
    rows = [[float(v) for v in ln.split()] for ln in (tmp_path / "relative error.log").read_text().splitlines()]
    for (e, a, b), (er, ar, br) in zip(rows, rows_ref):
        assert a == ar and b == br and e == pytest.approx(er, rel=1e-7)


def test_time_of_flight_script_on_gpu_matches_fenics_golden(cuda_shim, tmp_path, golden_dir):
    """JIT path end to end against FEniCS output: UFL degree estimation (8 -> 25-point collapsed Gauss-Jacobi rule),
    P2-interpolated source Expression, Newton, project(exp(u)), errornorm."""
    import time_of_flight as tof
    n_num, rel = tof.main(output_dir=tmp_path)
    gold = np.load(golden_dir / "tof_electrons_f3199.npy")
    vals = n_num.vector()[:]
    assert np.abs(vals - gold).max() <= 1e-9 * gold.max()
    ref = float((golden_dir / "time_of_flight_relative_error.log").read_text().split("=")[-1])
    assert abs(rel - ref) <= 1e-8 * ref
