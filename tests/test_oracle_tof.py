"""Pins the oracle's whole numerical path (FIAT rules, P1 assembly, log-BDF2, Newton, projection) on the
reference's time-of-flight goldens: tests/integrated_tests/time_of_flight/20220707_results/electrons000000.vtu
(PointData f_3199) and 'relative error.log'.  The reference test asserts 1e-5 (test_time_of_flight.py:45-56);
the restatement reproduces FEniCS to ~1e-13."""
import numpy as np

from oracle.tof import TofProblem


def test_tof_matches_golden_field_and_error(golden_dir):
    p = TofProblem()
    x, its = p.run()
    assert set(its) == {3}
    n_num, _, rel = p.postprocess(x)
    gold = np.load(golden_dir / "tof_electrons_f3199.npy")
    assert np.argmax(n_num) == np.argmax(gold) == 1476
    assert np.abs(n_num - gold).max() <= 5e-12 * gold.max()
    ref_rel = float((golden_dir / "time_of_flight_relative_error.log").read_text().split("=")[-1])
    assert abs(rel - ref_rel) <= 1e-11 * ref_rel
    # the reference's own acceptance test (far-field ringing excluded where gold ~ 0)
    big = np.abs(gold) > 1e-6 * gold.max()
    assert np.abs((n_num[big] - gold[big]) / gold[big]).max() < 1e-5
