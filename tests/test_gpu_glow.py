"""GPU glow-discharge run (GlowModel kernels, table interpolation, reduced-field projection, adaptive BDF2) against
the reference's own goldens -- tests/integrated_tests/glow_discharge/20220707_results/{electrons,Ar_plus,Ar_star}.h5
and 'relative error.log' -- and, entry by entry, against the CPU oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200.glow import GlowCase, NF


def test_glow_assembly_matches_oracle(golden_dir):
    """Residual and Jacobian of the first step on a small mesh at a perturbed state (all blocks populated)."""
    from oracle.glow import GlowProblem
    nx = 12
    case = GlowCase(golden_dir / "glow_discharge_input", nx=nx, ny=nx, precond="bjacobi")
    prob = GlowProblem(golden_dir / "glow_discharge_input", nx=nx, ny=nx)
    n = prob.mesh.num_vertices
    rng = np.random.default_rng(5)
    X = prob.initial_state().reshape(-1, 5).copy()
    X[:, :4] += rng.normal(0, 1e-2, size=(n, 4))
    X[:, 4] = -30.0 * (1 - prob.mesh.coords[:, 1] / 0.01) + rng.normal(0, 0.5, size=n)
    Xo = X.copy(); Xo[:, :4] += rng.normal(0, 1e-3, size=(n, 4))
    Xo1 = Xo.copy(); Xo1[:, :4] += rng.normal(0, 1e-3, size=(n, 4))
    eps = np.exp(Xo[:, 0] - Xo[:, 3])
    dt, dt_old, t = 2e-13, 3e-13, 5e-13
    # oracle side
    prob.Uold, prob.Uold1, prob.eps = Xo, Xo1, eps
    prob.dt, prob.dt_old, prob.t = dt, dt_old, t
    prob.update_coefficients(eps, prob.reduced_field(X[:, 4]))
    x = X.ravel()
    b_ref, A_ref = prob.residual(x), prob.jacobian(x)
    # device side: same sequence as GlowCase.step
    dev = case.device
    case.x_new.copy_(torch.from_numpy(x).to(dev))
    case.x_old.copy_(torch.from_numpy(Xo.ravel()).to(dev))
    case.x_old1.copy_(torch.from_numpy(Xo1.ravel()).to(dev))
    case.hot.fields[:, 13] = torch.from_numpy(eps).to(dev)
    case._update_coefficients()
    redE = case.redE.cpu().numpy()
    redE_ref = prob.reduced_field(X[:, 4])
    assert np.abs(redE - redE_ref).max() <= 1e-11 * np.abs(redE_ref).max()
    for name, col in (("mu2", 10), ("D2", 11), ("mu3", 8), ("D3", 6), ("mu3d", 9), ("D3d", 7), ("k0", 14), ("kd3", 24), ("k4", 18)):
        got = case.hot.fields[:, col].cpu().numpy()
        assert np.abs(got - prob.coef[name]).max() <= 1e-12 * np.abs(prob.coef[name]).max(), name
    case.hot.bc_values = np.array([prob.U_w * (1.0 - np.exp(-t / 1e-9)), 0.0])
    case.hot.bdf2_history(4, case.x_old, case.x_old1, dt, dt_old)
    b = case.hot.residual(case.x_new, dt, dt_old, t).cpu().numpy()
    scale = np.abs(A_ref) @ np.abs(x)
    assert (np.abs(b - b_ref) / scale).max() <= 1e-11
    case.hot.jacobian(case.x_new, dt, dt_old, t)
    rowptr, colidx, vals = case.hot.matrix_bsr()
    A = sp.bsr_matrix((vals, colidx, rowptr), shape=A_ref.shape).tocsr()
    A.sort_indices()
    assert np.array_equal(A.indptr, A_ref.indptr) and np.array_equal(A.indices, A_ref.indices)
    rowid = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    rowmax = np.zeros(A.shape[0])
    np.maximum.at(rowmax, rowid, np.abs(A_ref.data))
    assert (np.abs(A.data - A_ref.data) <= 1e-10 * rowmax[rowid]).all(), float((np.abs(A.data - A_ref.data) / rowmax[rowid]).max())


def test_glow_matches_fenics_golden(golden_dir):
    """The reference's integrated test on the device: 6 accepted steps to t = 1e-11, tight linear solves ('mumps')."""
    case = GlowCase(golden_dir / "glow_discharge_input", ksp_rtol=1e-12)
    out = case.run()
    gold = [[float(t) for t in ln.split()] for ln in (golden_dir / "glow_discharge_relative_error.log").read_text().splitlines()]
    assert len(case.rows) == 6 and [s[0] for s in case.stats] == [2] * 6
    for (e, dto, dt), (ge, gdto, gdt) in zip(case.rows, gold):
        assert abs(e - ge) <= 1e-7 * ge                       # the reference test uses np.allclose (1e-5)
        assert abs(dt - gdt) <= 1e-8 * gdt
    g = np.load(golden_dir / "glow_h5.npz")
    d2v = np.empty(20201, dtype=np.int64)
    d2v[g["cell_dofs"]] = g["topology"].ravel()
    O = out.reshape(-1, 5)
    for name, k in (("electrons", 3), ("Ar_plus", 2), ("Ar_star", 1)):
        ref = g[name + "_1"]
        assert np.abs((O[d2v, k] - ref) / ref).max() <= 1e-8, name
