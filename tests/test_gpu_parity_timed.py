"""GPU parity on the configurations that are actually timed / named by BASELINE.md ("What's weak" 1-4).

* streamer n = 256, T_final = 1e-10 (21 accepted steps) -- SURVEY 8(d)'s named parity configuration -- against the committed
  oracle fixture tests/golden/streamer_n256_T1e-10.npz (tools/make_golden_streamer256.py), in both solver semantics:
  tight linear solves (the examples' 'mumps' mode) and the bench semantics (two-stage AMG, GMRES(30) rtol 1e-5);
  reference loop: tests/integrated_tests/streamer_discharge/fedm_streamer.py:312-353.
* the reject paths of fedm/functions.py:1077-1127 driven on the device: error >= ttol -> dt *= 0.5 ttol/err, and a Newton
  failure (max_it = 1) -> dt *= 0.5, both against the oracle's rows (tests/golden/streamer_n32_reject.npz).
* bench semantics at scale (n = 512, 788 k DOF): the rtol-1e-5 trajectory stays within 1e-8 of the rtol-1e-13 one.
* time of flight at the example's size 160 x 320 (BASELINE configs[1], examples/time_of_flight/fedm-tof.py:87).
* partitioned runs on 2/4/8 GPUs (spawned under torch.distributed.run when the box has the devices) against the same
  n = 256 fixture.
Tolerances: states 1e-8 relative (north_star), dt sequence bit-equal in the exact-solve mode, Newton counts equal.
"""
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

from fedm_b200.streamer import StreamerCase

ROOT = Path(__file__).resolve().parent.parent


def check_against_fixture(case, gold, state_tol, err_tol, dt_exact):
    X = case.x_new.view(-1, 3).cpu().numpy()
    rows, rows_ref = np.array(case.rows), gold["rows"]
    assert rows.shape == rows_ref.shape, (rows.shape, rows_ref.shape)
    assert [s[0] for s in case.stats] == gold["newton_its"].tolist()          # Newton iterations per attempt
    if dt_exact:
        assert np.array_equal(rows[:, 1:], rows_ref[:, 1:])                   # accepted dt sequence, bit for bit
    else:
        assert np.abs(rows[:, 1:] / rows_ref[:, 1:] - 1.0).max() <= 1e-6
    assert np.abs(rows[:, 0] / rows_ref[:, 0] - 1.0).max() <= err_tol
    worst = 0.0
    for k in range(3):
        ref = gold["sample"][:, k]
        worst = max(worst, float(np.abs(X[gold["idx"], k] - ref).max() / np.abs(ref).max()))
        l2, mx, sm = gold["norms"][k]
        assert np.linalg.norm(X[:, k]) == pytest.approx(l2, rel=state_tol)
        assert np.abs(X[:, k]).max() == pytest.approx(mx, rel=state_tol)
    assert worst <= state_tol, worst
    return worst


@pytest.fixture(scope="module")
def gold256(golden_dir):
    return np.load(golden_dir / "streamer_n256_T1e-10.npz")


def test_streamer_n256_21_steps_exact_solves(gold256):
    case = StreamerCase(n=256, ksp_rtol=1e-13, ksp_max_it=50000)
    case.set_initial_state()
    steps = case.run(T_final=1e-10)
    assert steps == 21
    worst = check_against_fixture(case, gold256, 1e-8, 1e-7, dt_exact=True)
    print("n=256, 21 steps, exact solves: max rel state error", worst)


def test_streamer_n256_21_steps_bench_semantics(gold256):
    """GMRES(30) rtol 1e-5 + two-stage AMG, Pythagoras-norm Gram-Schmidt: the semantics bench.py times."""
    case = StreamerCase(n=256)
    case.set_initial_state()
    steps = case.run(T_final=1e-10)
    assert steps == 21
    worst = check_against_fixture(case, gold256, 1e-8, 1e-5, dt_exact=False)
    print("n=256, 21 steps, bench semantics: max rel state error", worst)


def test_bench_semantics_trajectory_at_scale():
    """n = 512 (788 k DOF), 10 steps: loose (bench) vs tight Krylov tolerance, same code otherwise."""
    a, b = StreamerCase(n=512), StreamerCase(n=512, ksp_rtol=1e-13, ksp_max_it=50000)
    for c in (a, b):
        c.set_initial_state()
        c.run(max_steps=10)
    assert [s[0] for s in a.stats] == [s[0] for s in b.stats]
    Xa, Xb = a.x_new.view(-1, 3), b.x_new.view(-1, 3)
    for k in range(3):
        rel = float((Xa[:, k] - Xb[:, k]).abs().max() / Xb[:, k].abs().max())
        assert rel <= 1e-8, (k, rel)
    ra, rb = np.array(a.rows), np.array(b.rows)
    assert np.abs(ra / rb - 1.0).max() <= 1e-5


def test_forced_rejection_error_above_ttol(golden_dir):
    """ttol = 2e-5 < the first estimate (7.4e-5): attempts are rejected and dt shrinks by 0.5 ttol/err
    (fedm/functions.py:1085-1093); every attempt writes a log row, exactly as the oracle's."""
    g = np.load(golden_dir / "streamer_n32_reject.npz")
    case = StreamerCase(n=32, ksp_rtol=1e-13, ksp_max_it=50000)
    case.ttol = 2e-5
    case.set_initial_state()
    attempts = [case.step() for _ in range(4)]
    rows = np.array(case.rows)
    assert attempts == [2, 2, 1, 1] and rows.shape == g["ttol_rows"].shape
    assert np.abs(rows / g["ttol_rows"] - 1.0).max() <= 1e-6
    assert rows[0, 2] == g["ttol_rows"][0, 2] and rows[1, 1] == 1e30          # first retry keeps dt_old
    x, ref = case.x_new.cpu().numpy(), g["ttol_state"]
    for k in range(3):
        assert np.abs(x[k::3] - ref[k::3]).max() <= 1e-8 * np.abs(ref[k::3]).max(), k


def test_forced_newton_failure_halves_dt(golden_dir):
    """maximum_iterations = 1: the solver raises (DIVERGED_MAX_IT) six times, dt is halved each time
    (fedm/functions.py:1094-1096), then one iteration suffices at dt = 5e-12/64."""
    g = np.load(golden_dir / "streamer_n32_reject.npz")
    case = StreamerCase(n=32, ksp_rtol=1e-13, ksp_max_it=50000, newton_max_it=1)
    case.set_initial_state()
    attempts = [case.step() for _ in range(3)]
    rows = np.array(case.rows)
    assert attempts[0] == 7 and rows.shape == g["maxit_rows"].shape
    assert rows[0, 2] == 5e-12 / 64 == g["maxit_rows"][0, 2]
    assert np.abs(rows / g["maxit_rows"] - 1.0).max() <= 1e-6
    x, ref = case.x_new.cpu().numpy(), g["maxit_state"]
    for k in range(3):
        assert np.abs(x[k::3] - ref[k::3]).max() <= 1e-8 * np.abs(ref[k::3]).max(), k


def test_library_errors_are_fatal_not_retried():
    """A libfedmx failure inside an attempt must surface, not be turned into dt halving."""
    from fedm_b200.lib import FedmxError
    from fedm_b200.timestepping import attempt_until_accepted

    def boom(t, dt):
        raise FedmxError("simulated CUDA failure")

    with pytest.raises(FedmxError):
        attempt_until_accepted(boom, lambda: None, 0.0, 1e-12, 1e30, 1e-3, 1e-15, [1, 1, 1])


def test_tof_example_size_160x320(golden_dir):
    """BASELINE configs[1]: 160 x 320 grid, dt = 1e-12 from t0 = 2.5e-9; 20 steps and the full 500 steps (T = 3e-9)
    against the oracle fixture (tools/make_golden_tof160.py), plus the reference's figure of merit."""
    from fedm_b200.tof import TofCase
    path = golden_dir / "tof_160x320.npz"
    if not path.exists():
        pytest.skip("fixture not generated")
    g = np.load(path)
    case = TofCase(nx=160, ny=320, width=5e-4, height=1e-3)
    x = case.run(T_final=2.52e-9).cpu().numpy()
    assert case.newton_its == g["newton_its"][:20].tolist()
    assert np.abs(x[g["idx"]] - g["u20"]).max() <= 1e-9 * np.abs(g["u20"]).max()
    assert np.linalg.norm(x) == pytest.approx(g["norms20"][0], rel=1e-9)
    if "u500" in g.files:
        case = TofCase(nx=160, ny=320, width=5e-4, height=1e-3)
        x = case.run(T_final=3e-9).cpu().numpy()
        assert len(case.newton_its) == 500 and case.newton_its == g["newton_its"].tolist()
        assert np.abs(x[g["idx"]] - g["u500"]).max() <= 1e-8 * np.abs(g["u500"]).max()
        assert np.linalg.norm(x) == pytest.approx(g["norms500"][0], rel=1e-8)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_run_matches_fixture(world):
    """One process per GPU (the reference's `mpirun -np N`, README.md:63-67): n = 256, 21 steps, strips of >= 32 node rows."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    for mode, precond in (("exact", "amg"), ("bench", "amg"), ("bench", "ilu0")):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29540 + world), str(ROOT / "tools" / "mgpu_check.py"), "--golden", "--mode", mode,
               "--precond", precond]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "PASS" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_user_script_runs_partitioned_under_torchrun(world):
    """`torchrun -n N examples/streamer_discharge.py` = the reference's `mpirun -np N python3 fedm-streamer.py`: the shim
    partitions every solve; result == 1-GPU run to 1e-8, identical on all ranks, MPI.sum/max global (tools/script_mgpu_check.py)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    for case in ("streamer", "glow"):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(29560 + world), str(ROOT / "tools" / "script_mgpu_check.py"), "--case", case]
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "PASS" in r.stdout
