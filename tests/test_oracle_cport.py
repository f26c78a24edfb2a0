"""Pins the compiled CPU restatement (oracle/csrc/streamer_ref.cpp -- bench.py's CPU arm) on the numpy oracle:
assembled residual / Jacobian entries <= 1e-12 of the row's term scale, ILU(0) == oracle/newton.py:ilu0, and whole BDF2
steps (Newton counts, estimator rows, states).  Reference semantics: fedm/functions.py:188-202, fedm_streamer.py:32,302-353."""
import numpy as np
import pytest

from oracle import streamer_c
from oracle.newton import ilu0, ilu0_apply
from oracle.streamer import StreamerProblem, streamer_mesh

if not streamer_c.SO.exists():
    pytest.skip("oracle/csrc/libstreamer_ref.so not built (python -c 'import __graft_entry__ as g; g.build()')", allow_module_level=True)


def perturbed(prob, seed=0):
    rng = np.random.default_rng(seed)
    X = prob.initial_state().reshape(-1, 3).copy()
    X[:, :2] += rng.normal(0, 1e-3, size=(X.shape[0], 2))
    X[:, 2] += rng.normal(0, 1e-3 * 18750.0, size=X.shape[0])
    Xo = X.copy()
    Xo[:, :2] += rng.normal(0, 1e-3, size=(X.shape[0], 2))
    Xo1 = Xo.copy()
    Xo1[:, :2] += rng.normal(0, 1e-3, size=(X.shape[0], 2))
    return X.ravel(), Xo, Xo1


@pytest.mark.parametrize("n,p,threads", [(20, 1.0, 1), (17, 2.0, 3)])
def test_entries_match_numpy_oracle(n, p, threads):
    mesh = streamer_mesh(n, p)
    prob, cp = StreamerProblem(mesh), streamer_c.StreamerC(mesh, threads)
    x, xo, xo1 = perturbed(prob)
    for q in (prob, cp):
        q.dt, q.dt_old, q.Uold, q.Uold1 = 4e-12, 5e-12, xo, xo1
    A_ref, b_ref = prob.jacobian(x), prob.residual(x)
    A, b = cp.jacobian(x), cp.residual(x)
    A.sort_indices()
    assert np.array_equal(A.indptr, A_ref.indptr) and np.array_equal(A.indices, A_ref.indices)      # dolfin's pattern
    scale = np.abs(A_ref) @ np.abs(x)
    assert (np.abs(b - b_ref) / scale).max() <= 1e-12
    rowid = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    rowmax = np.zeros(A.shape[0])
    np.maximum.at(rowmax, rowid, np.abs(A_ref.data))
    assert (np.abs(A.data - A_ref.data) <= 1e-12 * rowmax[rowid]).all()
    cp.close()


def test_ilu0_single_block_equals_oracle_ilu0():
    mesh = streamer_mesh(10)
    prob, cp = StreamerProblem(mesh), streamer_c.StreamerC(mesh, 1)
    x, xo, xo1 = perturbed(prob)
    for q in (prob, cp):
        q.dt, q.dt_old, q.Uold, q.Uold1 = 4e-12, 5e-12, xo, xo1
    A = cp.jacobian(x)
    LU, diag = ilu0(A)
    r = np.random.default_rng(1).normal(size=x.size)
    z_ref = ilu0_apply(LU, diag, r)
    z = cp.pc_apply(r)
    assert np.abs(z - z_ref).max() <= 1e-12 * np.abs(z_ref).max()
    cp.close()


@pytest.mark.parametrize("threads", [1, 4])
def test_steps_match_numpy_oracle(threads):
    """Tight Krylov tolerance: the block-Jacobi/ILU(0)-GMRES run must land on the exact-solve trajectory (same Newton counts,
    same dt rows, states to 1e-8); at PETSc's rtol 1e-5 the estimator stays within 1e-6."""
    n = 24
    mesh = streamer_mesh(n)
    prob = StreamerProblem(mesh)
    x_ref, rows_ref, its_ref = prob.run(max_steps=3)
    cp = streamer_c.StreamerC(mesh, threads)
    x, rows, stats = cp.run(prob.initial_state(), max_steps=3, ksp_rtol=1e-13)
    assert [s[0] for s in stats] == [h[0] for h in its_ref]
    for (e, a, b), (er, ar, br) in zip(rows, rows_ref):
        assert a == ar and b == br and e == pytest.approx(er, rel=1e-7)
    for k in range(3):
        assert np.abs(x[k::3] - x_ref[k::3]).max() <= 1e-8 * np.abs(x_ref[k::3]).max()
    x5, rows5, stats5 = cp.run(prob.initial_state(), max_steps=3)
    assert [s[0] for s in stats5] == [h[0] for h in its_ref]
    for (e, a, b), (er, ar, br) in zip(rows5, rows_ref):
        assert a == ar and b == br and e == pytest.approx(er, rel=1e-6)
    cp.close()
