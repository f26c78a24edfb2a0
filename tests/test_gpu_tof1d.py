"""BASELINE configs[0] on the device: 1-D P2 time of flight (examples/time_of_flight_1D/fedm-tof_1d.py) against the CPU oracle
(oracle/tof1d.py, same discretisation: states 1e-9, identical Newton counts) and against the case's exact solution."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200.tof1d import Tof1DCase
from oracle.tof1d import Tof1D


def test_entries_match_oracle():
    o = Tof1D(n_cells=50)
    case = Tof1DCase(n_cells=50)
    rng = np.random.default_rng(0)
    from oracle.tof1d import ln_swarm, source
    u = ln_swarm(o.x, 1e-9) + rng.normal(0, 1e-3, o.ndof)
    uo, uo1 = ln_swarm(o.x, 0.9e-9), ln_swarm(o.x, 0.8e-9)
    f = source(o.x, 1e-9)
    b_ref, A_ref = o.residual_jacobian(u, uo, uo1, f, 1e-11, 1.3e-11)
    dev = case.hot.device
    for t, a in ((case.u_new, u), (case.u_old, uo), (case.u_old1, uo1), (case.f, f)):
        t.copy_(torch.from_numpy(a).to(dev))
    case._assemble(1e-11, 1.3e-11)
    b = case.res.cpu().numpy()
    band = case.band.cpu().numpy().reshape(-1, 5)
    A = A_ref.toarray()
    scale = np.abs(A) @ np.abs(u)
    assert (np.abs(b - b_ref) / scale).max() <= 1e-12
    for i in range(o.ndof):
        for k in range(5):
            j = i - 2 + k
            ref = A[i, j] if 0 <= j < o.ndof else 0.0
            assert abs(band[i, k] - ref) <= 1e-12 * np.abs(A[i]).max(), (i, k)


def test_example_run_matches_oracle_and_exact_solution():
    """The example's configuration (4000 cells, 8001 dofs, dt = 1e-11) for 30 steps."""
    o = Tof1D()
    u_ref, t_ref, its_ref = o.run(T_final=3e-10)
    case = Tof1DCase()
    u = case.run(T_final=3e-10).cpu().numpy()
    assert case.newton_its == its_ref
    assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max()
    err = case.l2_relative_error()
    assert err == pytest.approx(o.l2_relative_error(u_ref, t_ref), rel=1e-6) and err < 1e-2
