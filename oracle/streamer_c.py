"""ctypes driver of oracle/csrc/libstreamer_ref.so -- the compiled, all-core CPU restatement of the streamer hot path
(oracle -- test infrastructure and the reported CPU baseline only; the product never imports this).

The problem definition (mesh, boundaries, Dirichlet nodes, initial condition, time loop, controller) is the numpy oracle's
(oracle/streamer.py, restating tests/integrated_tests/streamer_discharge/fedm_streamer.py); only the per-attempt Newton
solve -- assembly + GMRES(30) with block-Jacobi/ILU(0), the reference test's KSP (fedm_streamer.py:32) -- runs in the C++
library, with P OpenMP threads standing in for `mpirun -np P` (README.md:63-67).
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from .controller import SolveFailed, adaptive_attempts, adaptive_timestep, error_estimate
from .streamer import BOUNDARIES, BOX, U_W, initial_log_densities, streamer_mesh

SO = Path(__file__).resolve().parent / "csrc" / "libstreamer_ref.so"


def load():
    lib = C.CDLL(str(SO))
    vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
    lib.sref_create.restype = vp
    lib.sref_create.argtypes = [i64, vp, i64, vp, ci]
    lib.sref_destroy.argtypes = [vp]
    lib.sref_threads.argtypes = [vp]
    lib.sref_set_bc.argtypes = [vp, i64, vp, vp]
    lib.sref_set_facets.argtypes = [vp, i64, vp]
    lib.sref_set_history.argtypes = [vp, vp, vp, dbl, dbl]
    lib.sref_residual.argtypes = [vp, vp, vp]
    lib.sref_jacobian.argtypes = [vp, vp, vp]
    lib.sref_nnz.restype = i64
    lib.sref_nnz.argtypes = [vp]
    lib.sref_get_csr.argtypes = [vp, vp, vp, vp]
    lib.sref_factor.argtypes = [vp]
    lib.sref_pc_apply.argtypes = [vp, vp, vp]
    lib.sref_gmres.argtypes = [vp, vp, vp, ci, dbl, ci, C.POINTER(ci)]
    lib.sref_newton.argtypes = [vp, vp, dbl, dbl, dbl, ci, ci, dbl, ci, C.POINTER(ci), vp]
    lib.sref_timers.argtypes = [vp, vp, ci]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class StreamerC:
    """Same interface as oracle.streamer.StreamerProblem where the tests need it (residual, jacobian, run)."""

    def __init__(self, mesh, threads=0):
        self.lib = load()
        self.mesh = mesh
        threads = threads or (os.cpu_count() or 1)
        coords = np.ascontiguousarray(mesh.coords, dtype=np.float64)
        cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
        self.h = C.c_void_p(self.lib.sref_create(coords.shape[0], _p(coords), cells.shape[0], _p(cells), threads))
        self.threads = self.lib.sref_threads(self.h)
        self.ndof = 3 * coords.shape[0]
        z = coords[:, 1]
        cath = np.nonzero(np.abs(z) < 3e-16)[0]
        anod = np.nonzero(np.abs(z - BOX) < 3e-16)[0]
        nodes = np.concatenate([cath, anod]).astype(np.int32)
        vals = np.concatenate([np.zeros(cath.size), np.full(anod.size, U_W)])
        self.lib.sref_set_bc(self.h, nodes.size, _p(nodes), _p(vals))
        fv, fcell, fopp = mesh.boundary_facets()
        mk = mesh.mark_boundaries(BOUNDARIES)
        sel = np.nonzero((mk == 1) | (mk == 2))[0]                     # electron 'Neumann' drift flux on both electrodes
        opp = cells[fcell[sel], fopp[sel]]
        f3 = np.ascontiguousarray(np.stack([fv[sel, 0], fv[sel, 1], opp], axis=1), dtype=np.int32)
        self.lib.sref_set_facets(self.h, f3.shape[0], _p(f3))
        self.dt, self.dt_old = 5e-12, 1e30
        self.Uold = self.Uold1 = None

    def close(self):
        if self.h:
            self.lib.sref_destroy(self.h)
            self.h = None

    def _push_history(self):
        uo = np.ascontiguousarray(self.Uold, dtype=np.float64).ravel()
        uo1 = np.ascontiguousarray(self.Uold1, dtype=np.float64).ravel()
        self.lib.sref_set_history(self.h, _p(uo), _p(uo1), self.dt, self.dt_old)

    def residual(self, x):
        self._push_history()
        b = np.empty(self.ndof)
        self.lib.sref_residual(self.h, _p(np.ascontiguousarray(x)), _p(b))
        return b

    def jacobian(self, x):
        import scipy.sparse as sp
        self._push_history()
        self.lib.sref_jacobian(self.h, _p(np.ascontiguousarray(x)), None)
        nnz = self.lib.sref_nnz(self.h)
        ptr, col, val = np.empty(self.ndof + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        self.lib.sref_get_csr(self.h, _p(ptr), _p(col), _p(val))
        return sp.csr_matrix((val, col, ptr), shape=(self.ndof, self.ndof))

    def pc_apply(self, r):
        """block-Jacobi/ILU(0) of the matrix assembled by the last jacobian() call."""
        self.lib.sref_factor(self.h)
        z = np.zeros(self.ndof)
        self.lib.sref_pc_apply(self.h, _p(np.ascontiguousarray(r)), _p(z))
        return z

    def newton(self, x, rtol=1e-4, atol=1e-10, stol=1e-16, max_it=20, restart=30, ksp_rtol=1e-5, ksp_max_it=10000):
        self._push_history()
        kits = C.c_int(0)
        hist = np.zeros(64)
        its = self.lib.sref_newton(self.h, _p(x), rtol, atol, stol, max_it, restart, ksp_rtol, ksp_max_it, C.byref(kits), _p(hist))
        if its < 0:
            raise SolveFailed("Newton / linear solve did not converge")
        return its, kits.value, hist[:its + 1].copy()

    def timers(self, reset=False):
        t = np.zeros(4)
        self.lib.sref_timers(self.h, _p(t), 1 if reset else 0)
        return dict(zip(("assemble_jacobian_s", "assemble_residual_s", "ilu0_factor_s", "krylov_s"), t.tolist()))

    def run(self, x0, T_final=1e-10, ttol=1e-3, dt_min=1e-15, dt_max=5e-12, max_steps=None, ksp_rtol=1e-5, on_step=None, **newton_kw):
        """The time loop of fedm_streamer.py:312-353 (same as oracle.streamer.StreamerProblem.run) around sref_newton."""
        x_new = np.ascontiguousarray(x0, dtype=np.float64).copy()
        x_old = x_new.copy()
        t = 0.0
        self.dt, self.dt_old = 5e-12, 1e30
        max_error = [1, 1, 1]
        rows, stats = [], []
        steps = 0
        while t < T_final * (1 + 1e-6):
            self.Uold1 = x_old.copy()
            x_old = x_new.copy()
            self.Uold = x_old

            def solve_fn(t_new, dt):
                self.dt = dt
                its, kits, _ = self.newton(x_new, ksp_rtol=ksp_rtol, **newton_kw)
                stats.append((its, kits))
                return error_estimate(x_new[1::3], x_old[1::3])

            def reset_fn():
                x_new[:] = x_old

            t, self.dt, _ = adaptive_attempts(solve_fn, reset_fn, t, self.dt, ttol, dt_min, max_error, log_rows=rows, dt_old=self.dt_old)
            self.dt_old = self.dt
            self.dt = adaptive_timestep(self.dt, max_error, ttol, dt_min, dt_max)
            max_error[2] = max_error[1]
            max_error[1] = max_error[0]
            steps += 1
            if on_step is not None and on_step(steps, t, x_new) is False:
                break
            if max_steps is not None and steps >= max_steps:
                break
        return x_new, rows, stats
