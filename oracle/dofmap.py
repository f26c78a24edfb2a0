"""dolfin 2019.1.0 serial P1 dof numbering (oracle -- test infrastructure only).

dolfin renumbers the nodes of a FunctionSpace with SCOTCH's Gibbs-Poole-Stockmeyer ordering
(DofMapBuilder -> SCOTCH::compute_gps; un-vendored third party).  SURVEY.md 4.4-3 derived the
equivalent "chain-BFS" from the golden ``cell_dofs`` of the glow-discharge HDF5 files, where it
reproduces all 20201 entries; tests/test_oracle_dofmap.py pins this restatement on that fixture.
Mixed spaces interleave components: dof = N_eq*node + comp (not pinned by any fixture).
"""
from collections import deque

import numpy as np


def vertex_adjacency(cells, n):
    c = np.asarray(cells, dtype=np.int64)
    e = np.concatenate([c[:, [0, 1]], c[:, [1, 2]], c[:, [0, 2]]], axis=0)
    e = np.concatenate([e, e[:, ::-1]], axis=0)
    key = np.unique(e[:, 0] * n + e[:, 1])
    src, dst = key // n, key % n
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, src + 1, 1)
    ptr = np.cumsum(ptr)
    return ptr, dst                              # neighbour lists sorted ascending


def bfs_levels(ptr, adj, root):
    n = ptr.size - 1
    lev = np.full(n, -1, dtype=np.int64)
    lev[root] = 0
    q = deque([root])
    last = root
    while q:
        v = q.popleft()
        last = v
        for w in adj[ptr[v]:ptr[v + 1]]:
            if lev[w] < 0:
                lev[w] = lev[v] + 1
                q.append(w)
    return lev, last


def pseudo_peripheral_root(ptr, adj, start):
    """GPS pseudo-peripheral search: BFS from the current vertex; while the BFS depth strictly exceeds the
    previous depth, move to a farthest vertex (lowest index among ties -- no tie occurs in the fixture)."""
    v, prev = start, -1
    while True:
        lev, _ = bfs_levels(ptr, adj, v)
        depth = int(lev.max())
        if depth <= prev:
            return v, lev
        prev = depth
        v = int(np.nonzero(lev == depth)[0][0])


def gps_numbering(cells, n, start=0):
    """Returns d2v (dof -> vertex) of the chain-BFS numbering."""
    ptr, adj = vertex_adjacency(cells, n)
    root, lev = pseudo_peripheral_root(ptr, adj, start)
    numbered = np.zeros(n, dtype=bool)
    queued = np.zeros(n, dtype=bool)
    d2v = []
    q = deque([root])
    queued[root] = True
    while q:
        v = q.popleft()
        if numbered[v]:
            continue
        while v is not None:
            numbered[v] = True
            d2v.append(v)
            nxt = None
            for w in adj[ptr[v]:ptr[v + 1]]:
                if numbered[w]:
                    continue
                if nxt is None and lev[w] == lev[v]:
                    nxt = int(w)
                elif not queued[w]:
                    queued[w] = True
                    q.append(int(w))
            v = nxt
    return np.asarray(d2v, dtype=np.int64)
