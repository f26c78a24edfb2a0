"""Smoothed-aggregation AMG hierarchy for the potential block (set-up on the host, once; V-cycles run in libfedmx).

The Phi-Phi block of the coupled Jacobian, 2 pi r grad(u).grad(v) (fedm/functions.py:401), depends on the mesh
only, so its multigrid hierarchy is built a single time per problem.  It is the first stage of the two-stage
(CPR-style) preconditioner used by the device GMRES: multigrid on the elliptic Phi block, node-block Jacobi on
the transport unknowns.  The reference leaves this to PETSc's PC (ILU(0)/bjacobi, or hypre_amg as its scripts
suggest for iterative solves, fedm_streamer.py:306-307); only iteration counts depend on it, never the results.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

try:
    import numba
except Exception:                                                   # pragma: no cover
    numba = None


def _aggregate_py(indptr, indices, n):
    """Standard greedy aggregation (Vanek et al.): returns aggregate id per node and the number of aggregates."""
    agg = np.full(n, -1, dtype=np.int64)
    na = 0
    for i in range(n):                                              # pass 1: untouched neighbourhoods become aggregates
        if agg[i] >= 0 or indptr[i + 1] == indptr[i]:               # isolated nodes (Dirichlet rows) stay out of the coarse space
            continue
        free = True
        for q in range(indptr[i], indptr[i + 1]):
            if agg[indices[q]] >= 0:
                free = False
                break
        if free:
            agg[i] = na
            for q in range(indptr[i], indptr[i + 1]):
                agg[indices[q]] = na
            na += 1
    tmp = agg.copy()
    for i in range(n):                                              # pass 2: join a neighbouring aggregate
        if agg[i] >= 0:
            continue
        for q in range(indptr[i], indptr[i + 1]):
            j = indices[q]
            if tmp[j] >= 0:
                agg[i] = tmp[j]
                break
    for i in range(n):                                              # pass 3: leftovers
        if agg[i] >= 0 or indptr[i + 1] == indptr[i]:
            continue
        agg[i] = na
        for q in range(indptr[i], indptr[i + 1]):
            j = indices[q]
            if agg[j] < 0:
                agg[j] = na
        na += 1
    return agg, na


_aggregate = numba.njit(cache=True)(_aggregate_py) if numba is not None else _aggregate_py


def _strength(A, theta):
    """Symmetric strength-of-connection graph: |a_ij| >= theta sqrt(a_ii a_jj); returns CSR without diagonal."""
    A = A.tocoo()
    d = np.abs(A.tocsr().diagonal())
    keep = (A.row != A.col) & (np.abs(A.data) >= theta * np.sqrt(d[A.row] * d[A.col])) & (A.data != 0.0)
    S = sp.csr_matrix((np.ones(int(keep.sum()), dtype=np.int8), (A.row[keep], A.col[keep])), shape=A.shape)
    S = ((S + S.T) > 0).astype(np.int8).tocsr()
    S.sort_indices()
    return S


def _rho_dinv_a(A, dinv, iters=15, seed=0):
    """Spectral radius estimate of D^-1 A by power iteration (A is SPD here)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0.5, 1.0, A.shape[0])
    lam = 1.0
    for _ in range(iters):
        y = dinv * (A @ x)
        lam = np.linalg.norm(y) / np.linalg.norm(x)
        x = y / np.linalg.norm(y)
    return 1.1 * lam


def symmetric_dirichlet(K, bc_nodes):
    """Zero rows AND columns of constrained nodes, unit diagonal (the preconditioner wants an SPD operator)."""
    K = K.tocsr().copy()
    n = K.shape[0]
    free = np.ones(n)
    free[bc_nodes] = 0.0
    F = sp.diags(free)
    K = (F @ K @ F + sp.diags(1.0 - free)).tocsr()
    K.eliminate_zeros()
    K.sort_indices()
    return K


def build_hierarchy(K, theta=0.0, max_levels=10, coarse_size=600):
    """Returns a list of level dicts {A, P, R, dinv, omega} (P/R absent on the coarsest) and the dense coarse inverse."""
    levels = []
    A = K.tocsr()
    A.sort_indices()
    while True:
        d = A.diagonal()
        dinv = 1.0 / d
        rho = _rho_dinv_a(A, dinv)
        lvl = {"A": A, "dinv": dinv, "omega": (4.0 / 3.0) / rho}
        levels.append(lvl)
        n = A.shape[0]
        if n <= coarse_size or len(levels) >= max_levels:
            break
        S = _strength(A, theta)
        agg, na = _aggregate(S.indptr.astype(np.int64), S.indices.astype(np.int64), n)
        if na >= n or na == 0:
            break
        member = np.nonzero(agg >= 0)[0]
        cnt = np.bincount(agg[member], minlength=na).astype(np.float64)
        T = sp.csr_matrix((1.0 / np.sqrt(cnt[agg[member]]), (member, agg[member])), shape=(n, na))
        P = (T - sp.diags(lvl["omega"] * dinv) @ (A @ T)).tocsr()
        P.sort_indices()
        R = P.T.tocsr()
        R.sort_indices()
        lvl["P"], lvl["R"] = P, R
        A = (R @ A @ P).tocsr()
        A.sort_indices()
    Ac = levels[-1]["A"].toarray()
    coarse_inv = np.linalg.inv(Ac)
    return levels, coarse_inv


def vcycle_reference(levels, coarse_inv, b, nu=1):
    """Plain numpy V-cycle (used by CPU tests to validate the hierarchy; the product cycle runs on the GPU)."""
    def cyc(l, rhs):
        lv = levels[l]
        if l == len(levels) - 1:
            return coarse_inv @ rhs
        x = lv["omega"] * lv["dinv"] * rhs
        for _ in range(nu - 1):
            x = x + lv["omega"] * lv["dinv"] * (rhs - lv["A"] @ x)
        r = rhs - lv["A"] @ x
        x = x + lv["P"] @ cyc(l + 1, lv["R"] @ r)
        for _ in range(nu):
            x = x + lv["omega"] * lv["dinv"] * (rhs - lv["A"] @ x)
        return x
    return cyc(0, b)
