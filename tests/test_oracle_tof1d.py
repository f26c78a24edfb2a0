"""BASELINE.json configs[0] (examples/time_of_flight_1D of the reference, P2 interval elements, CPU only): the oracle against
the case's exact solution.  The reference has no fixture for this example, so the pins are the analytic swarm (method of
exact solutions, as the example itself does, fedm-tof_1d.py:155-160) and the expected orders of accuracy."""
import numpy as np

from oracle.tof1d import Tof1D, ln_swarm


def test_residual_jacobian_consistency():
    prob = Tof1D(n_cells=50)
    rng = np.random.default_rng(0)
    u = ln_swarm(prob.x, 1e-10) + rng.normal(0, 1e-2, prob.ndof)
    uo, uo1 = ln_swarm(prob.x, 9e-11), ln_swarm(prob.x, 8e-11)
    f = np.exp(ln_swarm(prob.x, 1e-10))
    b, A = prob.residual_jacobian(u, uo, uo1, f, 1e-11, 1.3e-11)
    d = rng.normal(0, 1.0, prob.ndof)
    eps = 1e-6
    bp, _ = prob.residual_jacobian(u + eps * d, uo, uo1, f, 1e-11, 1.3e-11, jac=False)
    bm, _ = prob.residual_jacobian(u - eps * d, uo, uo1, f, 1e-11, 1.3e-11, jac=False)
    fd = (bp - bm) / (2 * eps)
    assert np.abs(A @ d - fd).max() <= 1e-6 * np.abs(fd).max()


def test_reference_configuration_tracks_the_exact_swarm():
    """The example's own set-up (4000 cells, dt = 1e-11); a shorter end time keeps the CPU suite fast.  At this step size the
    error is the BDF2 time error (2.9e-3 at t = 4e-10, measured; it scales as dt^2, next test)."""
    prob = Tof1D()
    u, t, its = prob.run(T_final=4e-10)
    assert abs(t - 4e-10) <= 1e-6 * 4e-10 and max(its) <= 6
    assert 1e-3 < prob.l2_relative_error(u, t) < 4e-3


def test_second_order_in_time():
    """Variable-step BDF2 coefficients with the two BDF1 start-up steps of the script: halving dt divides the error by 4
    (the P2 spatial error, ~1e-6, is far below)."""
    errs = []
    for dt in (1e-11, 5e-12, 2.5e-12):
        prob = Tof1D(n_cells=2000)
        u, t, _ = prob.run(T_final=2e-10, dt=dt)
        errs.append(prob.l2_relative_error(u, t))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(1.9 < r < 2.1 for r in rates), (errs, rates)
