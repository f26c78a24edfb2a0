"""Host-side multi-GPU logic (no GPU needed): vertex partition, ghost layer, halo plan; plus a world_size-2 gloo run
that exchanges halos exactly as the NCCL path does (same plan, torch.distributed P2P on CPU tensors)."""
import os

import numpy as np
import pytest

from fedm_b200.mesh import rectangle
from fedm_b200.partition import build_local_mesh, strip_owner

BOUNDS = [["line", 0.0, 0.0, 0.0, 1.0], ["line", 1.0, 1.0, 0.0, 1.0], ["line", 0.0, 1.0, 0.0, 0.0], ["line", 0.0, 1.0, 1.0, 1.0]]


def parts(world, diag="right", n=9):
    m = rectangle((0, 0), (1, 1), n, n + 2, diag)
    fv = m.exterior_facets()
    fm = m.mark_lines(BOUNDS)
    owner = strip_owner(m.coords, world)
    return m, [build_local_mesh(m.coords, m.cells, fv, fm, owner, r, world) for r in range(world)], owner


@pytest.mark.parametrize("world,diag", [(1, "right"), (2, "right"), (4, "crossed"), (8, "right")])
def test_partition_invariants(world, diag):
    m, locs, owner = parts(world, diag)
    seen = np.zeros(m.num_vertices, dtype=int)
    for lm in locs:
        seen[lm.local_to_global[:lm.n_owned]] += 1
        # every cell touching an owned vertex is present, and only those
        want = (owner[m.cells] == lm.rank).any(axis=1).sum()
        assert lm.cells.shape[0] == want
        assert (lm.cells.min(axis=1) < lm.n_owned).all()
        # ghosts are exactly the non-owned vertices of local cells
        assert set(np.unique(lm.cells)) == set(range(lm.num_vertices))
        assert np.array_equal(lm.coords, m.coords[lm.local_to_global])
        assert (lm.send_nodes < lm.n_owned).all() and (lm.recv_nodes >= lm.n_owned).all()
        assert lm.recv_ptr[-1] == lm.num_vertices - lm.n_owned          # every ghost is received exactly once
        assert np.unique(lm.recv_nodes).size == lm.recv_nodes.size
    assert (seen == 1).all()
    # send/recv lists of neighbouring ranks match in global ids and order
    for a in locs:
        for i, q in enumerate(a.neighbors):
            b = locs[q]
            j = list(b.neighbors).index(a.rank)
            sa = a.local_to_global[a.send_nodes[a.send_ptr[i]:a.send_ptr[i + 1]]]
            rb = b.local_to_global[b.recv_nodes[b.recv_ptr[j]:b.recv_ptr[j + 1]]]
            assert np.array_equal(sa, rb)


def _gloo_worker(rank, world, port, ok):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, locs, owner = parts(world)
    lm = locs[rank]
    n_comp = 3
    g = np.arange(m.num_vertices * n_comp, dtype=np.float64).reshape(-1, n_comp) * 0.5 + 1.0     # "global vector"
    vec = torch.zeros((lm.num_vertices, n_comp), dtype=torch.float64)
    vec[:lm.n_owned] = torch.from_numpy(g[lm.local_to_global[:lm.n_owned]])
    reqs, rbufs = [], []
    for i, q in enumerate(lm.neighbors):
        s = torch.from_numpy(lm.send_nodes[lm.send_ptr[i]:lm.send_ptr[i + 1]].astype(np.int64))
        r = torch.from_numpy(lm.recv_nodes[lm.recv_ptr[i]:lm.recv_ptr[i + 1]].astype(np.int64))
        sb = vec[s].contiguous()
        rb = torch.empty((r.numel(), n_comp), dtype=torch.float64)
        reqs.append(dist.isend(sb, int(q)))
        reqs.append(dist.irecv(rb, int(q)))
        rbufs.append((r, rb))
    for rq in reqs:
        rq.wait()
    for r, rb in rbufs:
        vec[r] = rb
    good = np.array_equal(vec.numpy(), g[lm.local_to_global])
    # the Krylov dot product: owned entries + allreduce == the global dot product
    t = torch.tensor([float((vec[:lm.n_owned] ** 2).sum())], dtype=torch.float64)
    dist.all_reduce(t)
    good = good and abs(t.item() - float((g ** 2).sum())) <= 1e-12 * float((g ** 2).sum())
    # the NCCL unique id travels as 128 bytes from rank 0
    ident = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(ident, 0)
    good = good and bool((ident == torch.arange(128, dtype=torch.uint8)).all())
    ok[rank] = 1 if good else 0
    dist.destroy_process_group()


def test_halo_plan_world2_gloo():
    import torch.multiprocessing as mp
    world = 2
    ok = mp.get_context("spawn").Array("i", [0] * world)
    port = 29500 + (os.getpid() % 2000)
    procs = [mp.get_context("spawn").Process(target=_gloo_worker, args=(r, world, port, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_staging_offsets_put_every_segment_where_the_receiver_unpacks_it(world):
    """Host logic of the NVLink peer-memory transport (fedmx_peer_halo_set): a numpy emulation of the push (sender writes its
    segment at the remote offset of the receiver's staging buffer) followed by the receiver's unpack must equal the owners'
    values at every ghost node."""
    from fedm_b200.partition import peer_staging_offsets
    m, locs, owner = parts(world, "right", n=11)
    plans = [([int(q) for q in lm.neighbors], [int(v) for v in lm.recv_ptr]) for lm in locs]
    gvals = np.random.default_rng(world).normal(size=m.num_vertices)
    staging = [np.full(lm.recv_ptr[-1], np.nan) for lm in locs]
    for lm in locs:                                                       # push
        offs = peer_staging_offsets(lm.rank, plans)
        local = gvals[lm.local_to_global]
        for i, q in enumerate(lm.neighbors):
            seg = local[lm.send_nodes[lm.send_ptr[i]:lm.send_ptr[i + 1]]]
            staging[q][offs[i]:offs[i] + seg.size] = seg
    for lm in locs:                                                       # unpack
        assert not np.isnan(staging[lm.rank]).any()
        vec = np.zeros(lm.num_vertices)
        vec[lm.recv_nodes] = staging[lm.rank]
        assert np.array_equal(vec[lm.n_owned:], gvals[lm.local_to_global[lm.n_owned:]])


def test_rcb_partition_is_balanced_and_compact():
    """Recursive coordinate bisection (general / graded meshes): +-1 balance for any rank count, sub-domains are
    coordinate boxes, and the ghost layer is much smaller than for slabs when there are many ranks."""
    from fedm_b200.mesh import rectangle
    from fedm_b200.partition import build_local_mesh, partition_owner, rcb_owner, strip_owner
    m = rectangle((0.0, 0.0), (1.0, 1.0), 48, 48, "right", map_x=lambda s: s ** 2)
    for world in (2, 3, 5, 8, 16):
        owner = rcb_owner(m.coords, world)
        counts = np.bincount(owner, minlength=world)
        assert counts.max() - counts.min() <= 1 and counts.sum() == m.num_vertices
    fv, fm = m.exterior_facets(), np.ones(len(m.exterior_facets()), dtype=np.int32)
    ghosts = {}
    for name, owner in (("rcb", rcb_owner(m.coords, 16)), ("strip", strip_owner(m.coords, 16))):
        ghosts[name] = sum(build_local_mesh(m.coords, m.cells, fv, fm, owner, r, 16).num_vertices -
                           build_local_mesh(m.coords, m.cells, fv, fm, owner, r, 16).n_owned for r in range(0, 16, 5))
    assert ghosts["rcb"] < 0.6 * ghosts["strip"]
    assert np.array_equal(partition_owner(m.coords, 4, "strip"), strip_owner(m.coords, 4, axis=1))
    assert np.array_equal(partition_owner(m.coords, 16, "auto"), rcb_owner(m.coords, 16))
