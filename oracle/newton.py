"""Newton and linear solves with PETSc SNES/KSP semantics (oracle -- test infrastructure only).

Stand-ins for ``PETScSNESSolver.solve`` (reference call site fedm/functions.py:1047): newtonls with
the basic (full-step) line search, convergence on ||F||_2 (atol 1e-10 | rtol*||F0|| | stol), and
either an exact sparse LU (the reference's 'mumps' mode; rows are equilibrated first because they
scale with exp(u)) or left-preconditioned restarted GMRES + ILU(0) (the streamer
test's 'gmres' mode, PETSc defaults: restart 30, rtol 1e-5, classical Gram-Schmidt).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .controller import SolveFailed


def lu_solve(A, b):
    A = A.tocsr()
    s = np.abs(A).max(axis=1).toarray().ravel()
    s[s == 0.0] = 1.0
    Dr = sp.diags(1.0 / s)
    lu = spla.splu((Dr @ A).tocsc(), permc_spec="MMD_AT_PLUS_A")     # structurally symmetric FE pattern: 6x faster than COLAMD
    x = lu.solve(b / s)
    # one step of iterative refinement on the scaled system
    r = b / s - (Dr @ A) @ x
    x += lu.solve(r)
    return x


def ilu0(A):
    """ILU(0) of a CSR matrix in its given ordering (PETSc's default PC in serial).  Returns LU CSR data."""
    A = A.tocsr().copy()
    A.sort_indices()
    indptr, indices, data = A.indptr, A.indices, A.data
    n = A.shape[0]
    diag = np.empty(n, dtype=np.int64)
    for i in range(n):
        lo, hi = indptr[i], indptr[i + 1]
        d = lo + np.searchsorted(indices[lo:hi], i)
        diag[i] = d
    for i in range(n):
        lo, hi = indptr[i], indptr[i + 1]
        cols = indices[lo:hi]
        for kk in range(lo, diag[i]):
            k = indices[kk]
            data[kk] /= data[diag[k]]
            lik = data[kk]
            # row k upper part
            klo, khi = diag[k] + 1, indptr[k + 1]
            kc = indices[klo:khi]
            pos = np.searchsorted(cols, kc)
            ok = (pos < cols.size)
            ok[ok] &= cols[pos[ok]] == kc[ok]
            data[lo + pos[ok]] -= lik * data[klo:khi][ok]
    return A, diag


def ilu0_apply(LU, diag, r):
    indptr, indices, data = LU.indptr, LU.indices, LU.data
    n = r.size
    y = r.copy()
    for i in range(n):
        lo = indptr[i]
        d = diag[i]
        if d > lo:
            y[i] -= data[lo:d] @ y[indices[lo:d]]
    for i in range(n - 1, -1, -1):
        d, hi = diag[i], indptr[i + 1]
        if hi > d + 1:
            y[i] -= data[d + 1:hi] @ y[indices[d + 1:hi]]
        y[i] /= data[d]
    return y


def gmres_left(A, b, M_apply, restart=30, rtol=1e-5, atol=1e-50, maxit=10000, x0=None):
    """Left-preconditioned GMRES(restart), classical Gram-Schmidt, convergence on the
    preconditioned residual norm -- PETSc KSPGMRES defaults.  Returns (x, iterations, converged)."""
    n = b.size
    x = np.zeros(n) if x0 is None else x0.copy()
    its = 0
    r = M_apply(b - A @ x)
    beta0 = np.linalg.norm(r)
    tol = max(rtol * np.linalg.norm(M_apply(b)), atol)
    if beta0 <= tol:
        return x, 0, True
    while its < maxit:
        beta = np.linalg.norm(r)
        V = np.zeros((restart + 1, n))
        H = np.zeros((restart + 1, restart))
        cs, sn = np.zeros(restart), np.zeros(restart)
        g = np.zeros(restart + 1)
        g[0] = beta
        V[0] = r / beta
        k_used = 0
        done = False
        for k in range(restart):
            w = M_apply(A @ V[k])
            h = V[:k + 1] @ w
            w = w - h @ V[:k + 1]
            H[:k + 1, k] = h
            H[k + 1, k] = np.linalg.norm(w)
            if H[k + 1, k] != 0.0:
                V[k + 1] = w / H[k + 1, k]
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k] = d
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            its += 1
            k_used = k + 1
            if abs(g[k + 1]) <= tol or its >= maxit:
                done = abs(g[k + 1]) <= tol
                break
        y = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        x = x + y @ V[:k_used]
        if done:
            return x, its, True
        r = M_apply(b - A @ x)
    return x, its, False


def newton(residual_fn, jacobian_fn, x, rtol=1e-9, atol=1e-10, stol=1e-16, max_it=50,
           linear="lu", history=None, krylov_rtol=1e-5):
    """SNES newtonls/basic.  residual_fn(x)->b (BCs applied), jacobian_fn(x)->CSR (BCs applied).
    x is updated in place; returns (iterations, converged).  Raises SolveFailed like dolfin does."""
    b = residual_fn(x)
    fnorm0 = fnorm = np.linalg.norm(b)
    if history is not None:
        history.append(fnorm)
    if not np.isfinite(fnorm):
        raise SolveFailed("residual is not finite")
    if fnorm < atol:
        return 0, True
    for it in range(1, max_it + 1):
        A = jacobian_fn(x)
        if linear == "lu":
            dx = lu_solve(A, b)
        else:
            LU, diag = ilu0(A)
            dx, kits, ok = gmres_left(A, b, lambda r: ilu0_apply(LU, diag, r), rtol=krylov_rtol)
            if not ok:
                raise SolveFailed("linear solve failed")
        x -= dx
        b = residual_fn(x)
        fnorm = np.linalg.norm(b)
        if history is not None:
            history.append(fnorm)
        if not np.isfinite(fnorm):
            raise SolveFailed("residual is not finite")
        if fnorm < atol or fnorm <= rtol * fnorm0:
            return it, True
        if np.linalg.norm(dx) < stol * np.linalg.norm(x):
            return it, True
    raise SolveFailed(f"Newton did not converge in {max_it} iterations (|F|={fnorm:g})")
