"""Pins oracle.controller against the reference's golden relative-error logs (bit-exact dt columns).

Reference: fedm/functions.py:915-927 (adaptive_timestep), fedm_streamer.py:343-348 / fedm_gd.py:462-470
(the max_error shift); goldens: tests/integrated_tests/{streamer,glow}_discharge/20220707_results/relative error.log.
"""
import numpy as np
import pytest

from oracle.controller import (adaptive_timestep, adaptive_timestep_H211b, adaptive_timestep_PI34,
                               error_estimate, format_error_row)


def _rows(path):
    return [[float(t) for t in line.split()] for line in path.read_text().splitlines() if line.strip()]


@pytest.mark.parametrize("case,ttol,dt_min,dt_max", [
    ("streamer_discharge", 1e-3, 1e-15, 5e-12),
    ("glow_discharge", 2e-3, 1e-15, 1e-8),
])
def test_pid_controller_bit_exact(golden_dir, case, ttol, dt_min, dt_max):
    rows = _rows(golden_dir / f"{case}_relative_error.log")
    max_error = [1, 1, 1]
    for i, (err, dt_old, dt) in enumerate(rows):
        if i > 0:
            assert dt_old == rows[i - 1][2]
        max_error[0] = max([err, 0.0])
        dt_next = adaptive_timestep(dt, max_error, ttol, dt_min, dt_max)
        max_error[2] = max_error[1]
        max_error[1] = max_error[0]
        if i + 1 < len(rows):
            assert dt_next == rows[i + 1][2], (i, dt_next, rows[i + 1][2])


def test_error_row_format_matches_golden_bytes(golden_dir):
    text = (golden_dir / "glow_discharge_relative_error.log").read_text().splitlines(keepends=True)
    for line in text:
        e, a, b = (float(t) for t in line.split())
        assert format_error_row(e, a, b) == line


def test_other_controllers_and_estimator():
    assert adaptive_timestep_PI34(1e-12, [1e-4, 1e-4, 1e-4], 1e-3, 1e-15, 1.0) == pytest.approx(
        1e-12 * (0.8 * 10) ** 0.1 * 0.8 ** (0.4 / 3))
    assert adaptive_timestep_H211b(1e-12, 1e-12, [1e-4, 1e-4], 1e-3, 1e-15, 1.0) == pytest.approx(
        1e-12 * 8 ** (1 / 12) * 8 ** (1 / 12))
    a = np.array([1.0, 2.0, 3.0])
    b = np.array([1.0, 2.5, 2.0])
    assert error_estimate(a, b) == np.linalg.norm(a - b + 3e-16) / np.linalg.norm(b + 3e-16)
