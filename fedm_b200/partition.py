"""Vertex partition of a triangle mesh across GPUs, ghost layer and halo plan (host, set-up only).

Mirrors what dolfin's MPI mesh decomposition provides to the reference under ``mpirun -np N`` (README.md:63-67 of
the reference): every rank owns a contiguous chunk of vertices (strips along one axis), keeps all cells touching
an owned vertex (one ghost-cell layer), so residual/Jacobian assembly is owner-computes with no communication;
ghost values of a vector are refreshed by one halo exchange (``update_ghost_values()``, fedm/functions.py:625).
Local numbering: owned vertices first (ascending global id), ghosts after (ascending global id).
"""
import numpy as np


class LocalMesh:
    """One rank's view: local coords/cells/facets + maps to the global numbering + its halo plan."""

    def __init__(self, rank, world, coords, cells, n_owned, local_to_global, facet_vertices, facet_markers,
                 neighbors, send_ptr, send_nodes, recv_ptr, recv_nodes, n_global_vertices, n_global_cells):
        self.rank, self.world = rank, world
        self.coords, self.cells, self.n_owned = coords, cells, n_owned
        self.local_to_global = local_to_global
        self.facet_vertices, self.facet_markers = facet_vertices, facet_markers
        self.neighbors, self.send_ptr, self.send_nodes = neighbors, send_ptr, send_nodes
        self.recv_ptr, self.recv_nodes = recv_ptr, recv_nodes
        self.n_global_vertices, self.n_global_cells = n_global_vertices, n_global_cells

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    def global_to_local(self):
        g2l = np.full(self.n_global_vertices, -1, dtype=np.int64)
        g2l[self.local_to_global] = np.arange(self.local_to_global.size)
        return g2l


def strip_owner(coords, world, axis=1):
    """Owner rank per vertex: vertices sorted by (coords[axis], other axis) and cut into `world` equal chunks."""
    n = coords.shape[0]
    order = np.lexsort((coords[:, 1 - axis], coords[:, axis]))
    owner = np.empty(n, dtype=np.int32)
    bounds = (np.arange(world + 1) * n) // world
    for p in range(world):
        owner[order[bounds[p]:bounds[p + 1]]] = p
    return owner


def rcb_owner(coords, world):
    """Owner rank per vertex by recursive coordinate bisection: the vertex set is cut across its longer side into two parts
    whose sizes are proportional to the number of ranks on each side, recursively -- balanced (+-1 vertex) and compact
    sub-domains for graded or unstructured meshes (dolfin delegates this to SCOTCH/ParMETIS)."""
    coords = np.asarray(coords, dtype=np.float64)
    owner = np.empty(coords.shape[0], dtype=np.int32)

    def split(idx, r0, nr):
        if nr == 1:
            owner[idx] = r0
            return
        ext = coords[idx].max(axis=0) - coords[idx].min(axis=0)
        axis = int(ext[1] > ext[0])
        order = idx[np.lexsort((idx, coords[idx, 1 - axis], coords[idx, axis]))]
        left_r = nr // 2
        cut = (idx.size * left_r) // nr
        split(order[:cut], r0, left_r)
        split(order[cut:], r0 + left_r, nr - left_r)

    split(np.arange(coords.shape[0]), 0, int(world))
    return owner


def partition_owner(coords, world, method="auto"):
    """'strip': equal slabs along the longer axis; 'rcb': recursive coordinate bisection; 'auto': slabs up to 8 ranks on
    near-square domains (one neighbour on each side, the shortest halo lists for <= 8 ranks), RCB otherwise."""
    ext = np.ptp(np.asarray(coords), axis=0)
    if method == "auto":
        method = "strip" if world <= 8 and max(ext) <= 4.0 * max(min(ext), 1e-300) else "rcb"
    if method == "strip":
        return strip_owner(coords, world, axis=int(ext[1] >= ext[0]))
    if method == "rcb":
        return rcb_owner(coords, world)
    raise ValueError("partition method must be 'auto', 'strip' or 'rcb'")


def build_local_mesh(coords, cells, facet_vertices, facet_markers, owner, rank, world):
    n = coords.shape[0]
    cells = np.asarray(cells)
    owned = np.nonzero(owner == rank)[0]
    touch = (owner[cells] == rank).any(axis=1)
    lcells_g = cells[touch]
    verts = np.unique(lcells_g)
    ghosts = verts[owner[verts] != rank]
    l2g = np.concatenate([owned, ghosts]).astype(np.int64)
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[l2g] = np.arange(l2g.size)
    lcells = g2l[lcells_g].astype(np.int32)
    fv = np.asarray(facet_vertices).reshape(-1, 2)
    keep = (g2l[fv] >= 0).all(axis=1) if fv.size else np.zeros(0, dtype=bool)
    lfv = g2l[fv[keep]].astype(np.int32)
    lfm = np.asarray(facet_markers)[keep].astype(np.int32) if fv.size else np.zeros(0, dtype=np.int32)
    # ---- halo plan: what I receive (my ghosts, grouped by owner) and what I send (my owned nodes that are ghosts elsewhere)
    recv_by = {}
    for q in np.unique(owner[ghosts]):
        recv_by[int(q)] = ghosts[owner[ghosts] == q]                      # ascending global id
    # my owned nodes that are ghosts on rank q = my vertices of the cells that also have a vertex owned by q (one pass over
    # the cells of mixed ownership instead of one pass per rank)
    send_by = {}
    oc = owner[lcells_g]                                                   # cells touching my vertices are the only candidates
    mixed = (oc != oc[:, :1]).any(axis=1)
    if mixed.any():
        cm, om = lcells_g[mixed], oc[mixed]
        pairs = []
        for a in range(3):
            for b in range(3):
                if a == b:
                    continue
                sel = (om[:, a] == rank) & (om[:, b] != rank)
                if sel.any():
                    pairs.append(np.stack([om[sel, b].astype(np.int64), cm[sel, a].astype(np.int64)], axis=1))
        if pairs:
            pr = np.unique(np.concatenate(pairs), axis=0)                  # sorted by (rank, node): ascending global id per rank
            for q in np.unique(pr[:, 0]):
                send_by[int(q)] = pr[pr[:, 0] == q, 1]
    neighbors = sorted(set(recv_by) | set(send_by))
    send_ptr, recv_ptr, send_nodes, recv_nodes = [0], [0], [], []
    for q in neighbors:
        s = send_by.get(q, np.zeros(0, dtype=np.int64))
        r = recv_by.get(q, np.zeros(0, dtype=np.int64))
        send_nodes.append(g2l[s])
        recv_nodes.append(g2l[r])
        send_ptr.append(send_ptr[-1] + s.size)
        recv_ptr.append(recv_ptr[-1] + r.size)
    cat = lambda parts: (np.concatenate(parts) if parts else np.zeros(0)).astype(np.int32)
    return LocalMesh(rank, world, np.ascontiguousarray(coords[l2g]), lcells, owned.size, l2g, lfv, lfm,
                     np.asarray(neighbors, dtype=np.int32), np.asarray(send_ptr, dtype=np.int32), cat(send_nodes),
                     np.asarray(recv_ptr, dtype=np.int32), cat(recv_nodes), n, cells.shape[0])


def peer_staging_offsets(rank, plans):
    """For the NVLink peer-memory transport (fedmx_peer_halo_set): `plans[q] = (neighbor ranks of q, recv_ptr of q)` for every
    rank q; returns, per neighbour of `rank` in its own neighbour order, the node offset of `rank`'s segment inside THAT
    neighbour's receive staging buffer (= the neighbour's recv_ptr entry for `rank`)."""
    mine = plans[rank][0]
    return [int(plans[q][1][list(plans[q][0]).index(rank)]) for q in mine]
