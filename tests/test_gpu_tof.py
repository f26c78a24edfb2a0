"""GPU time-of-flight run against the reference's own golden output (FEniCS-pinned, no oracle in between):
tests/integrated_tests/time_of_flight/20220707_results/electrons000000.vtu (PointData f_3199), and against the
oracle for the log-density field.  The reference test's bar is 1e-5 (test_time_of_flight.py:45-56)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200.tof import TofCase


def test_tof_matches_fenics_golden(golden_dir):
    case = TofCase()
    x = case.run()
    assert set(case.newton_its) == {3}                       # as FEniCS/oracle: 3 Newton iterations per step
    n_num = case.project_density(x)
    gold = np.load(golden_dir / "tof_electrons_f3199.npy")
    assert np.argmax(n_num) == 1476
    assert np.abs(n_num - gold).max() <= 1e-9 * gold.max()
    big = np.abs(gold) > 1e-6 * gold.max()
    assert np.abs((n_num[big] - gold[big]) / gold[big]).max() < 1e-5      # the reference's own acceptance bound


def test_tof_short_run_matches_oracle():
    from oracle.tof import TofProblem
    p = TofProblem()
    x_ref, its = p.run(T_final=2.505e-9)
    case = TofCase()
    x = case.run(T_final=2.505e-9).cpu().numpy()
    assert case.newton_its == its
    assert np.abs(x - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
