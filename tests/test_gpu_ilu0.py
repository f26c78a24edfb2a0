"""ILU(0) preconditioner on the device (fedm_b200/csrc/fedmx_ilu.cu) -- PETSc's default PC of the reference test's KSP
(tests/integrated_tests/streamer_discharge/fedm_streamer.py:32,302-305).  The device computes the exact ILU(0) of the matrix in
its multicolour elimination order; checked against the oracle's point ILU(0) (oracle/newton.py:ilu0, PETSc's algorithm) applied
to the identically permuted matrix, and through whole BDF2 steps (converged results do not depend on the preconditioner)."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

from fedm_b200.streamer import StreamerCase
from oracle.newton import ilu0, ilu0_apply, lu_solve
from oracle.streamer import StreamerProblem, streamer_mesh


def test_ilu0_equals_point_ilu0_of_the_permuted_matrix():
    n = 10
    case = StreamerCase(n=n, precond="ilu0")
    case.set_initial_state()
    x = case.x_new.clone()
    hot = case.hot
    hot.bdf2_history(2, x, x, 5e-12, 1e30)
    hot.jacobian(x, 5e-12, 1e30)                                   # assembles and (precond = ilu0) factors
    rowptr, colidx, vals = hot.matrix_bsr()
    nb = rowptr.size - 1
    A = sp.bsr_matrix((vals, colidx, rowptr), shape=(3 * nb, 3 * nb)).tocsr()
    n_colors, n_blocks, order = hot.ilu0_info()
    assert 3 <= n_colors <= 8 and n_blocks == colidx.size and sorted(order.tolist()) == list(range(nb))
    perm = (3 * order[:, None] + np.arange(3)[None, :]).ravel()    # node-interleaved dofs in elimination order
    Ap = A[perm][:, perm].tocsr()
    LU, diag = ilu0(Ap)
    rng = np.random.default_rng(0)
    r = rng.normal(size=3 * nb)
    z_ref = np.empty_like(r)
    z_ref[perm] = ilu0_apply(LU, diag, r[perm])
    z = hot.ilu0_solve(torch.from_numpy(r).to(hot.device)).cpu().numpy()
    assert np.abs(z - z_ref).max() <= 1e-11 * np.abs(z_ref).max()
    # as a Krylov preconditioner: same solution as the sparse LU
    b = hot.residual(x, 5e-12, 1e30)
    dx, its, conv, _ = hot.gmres(b, restart=30, rtol=1e-13, max_it=5000)
    assert conv
    dx_ref = lu_solve(A, b.cpu().numpy())
    assert np.linalg.norm(dx.cpu().numpy() - dx_ref) <= 1e-8 * np.linalg.norm(dx_ref)


def test_steps_with_ilu0_match_oracle_and_report_reference_mode_iterations():
    """Reference-test mode (GMRES(30) rtol 1e-5, ILU(0)): same Newton counts and dt rows as the exact-solve oracle, estimator
    within 1e-6; with tight tolerances the states agree to 1e-8 (the preconditioner never changes converged results)."""
    n = 32
    prob = StreamerProblem(streamer_mesh(n))
    x_ref, rows_ref, its_ref = prob.run(max_steps=3)
    case = StreamerCase(n=n, precond="ilu0")
    case.set_initial_state()
    case.run(max_steps=3)
    assert [s[0] for s in case.stats] == [h[0] for h in its_ref]
    for (e, a, b), (er, ar, br) in zip(case.rows, rows_ref):
        assert a == ar and b == br and e == pytest.approx(er, rel=1e-6)
    loose_its = sum(s[1] for s in case.stats)
    tight = StreamerCase(n=n, precond="ilu0", ksp_rtol=1e-13, ksp_max_it=50000)
    tight.set_initial_state()
    tight.run(max_steps=3)
    x = tight.x_new.cpu().numpy()
    for k in range(3):
        assert np.abs(x[k::3] - x_ref[k::3]).max() <= 1e-8 * np.abs(x_ref[k::3]).max()
    amg = StreamerCase(n=n)
    amg.set_initial_state()
    amg.run(max_steps=3)
    print("Krylov iterations for 3 steps: ILU(0)", loose_its, "two-stage AMG", sum(s[1] for s in amg.stats))
    assert loose_its >= sum(s[1] for s in amg.stats)
