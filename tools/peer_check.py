"""Multi-GPU check of the NVLink peer-memory collectives (fedmx_peer.cu) against NCCL, plus per-operation latencies.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/peer_check.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from fedm_b200.streamer import StreamerCase

os.environ["FEDMX_PEER"] = "1"          # the peer-memory transport is opt-in (NCCL is the default): create the heap for this check
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
case = StreamerCase(n=n, device=local, world=world, rank=rank, precond="bjacobi")
hot, lm = case.hot, case.local
dev = hot.device
ok = True
gen = torch.Generator(device=dev).manual_seed(1234 + rank)


def check(name, cond):
    global ok
    flag = torch.tensor([1 if cond else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"  {name}: {'ok' if flag.item() else 'FAIL'}", flush=True)
    ok = ok and bool(flag.item())


# ---- halo exchange: peer == NCCL, repeated with changing data (races show up as stale ghosts)
for n_comp in (1, 3):
    good = True
    for it in range(100):
        v = torch.rand(lm.num_vertices * n_comp, dtype=torch.float64, device=dev, generator=gen)
        a, b = v.clone(), v.clone()
        hot.peer_enable(False); hot.halo_exchange(a, n_comp)
        hot.peer_enable(True); hot.halo_exchange(b, n_comp)
        good = good and torch.equal(a, b)
    check(f"halo exchange n_comp={n_comp} x100", good and hot.peer_error() == 0)

# ---- small allreduce: identical on all ranks and equal to the rank-ordered sum
good = True
for it in range(200):
    m = 1 + (it * 7) % 128
    v = torch.rand(m, dtype=torch.float64, device=dev, generator=gen)
    parts = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(parts, v)
    ref = torch.zeros_like(v)
    for p in parts:
        ref += p
    hot.allreduce_sum(v)
    good = good and torch.equal(v, ref)
check("small allreduce x200 (bit-exact rank-ordered sum)", good and hot.peer_error() == 0)

# ---- big allreduce
good = True
big = max(200, lm.n_global_vertices // 2)
for it in range(50):
    m = big - it
    v = torch.rand(m, dtype=torch.float64, device=dev, generator=gen)
    parts = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(parts, v)
    ref = torch.zeros_like(v)
    for p in parts:
        ref += p
    hot.allreduce_sum(v)
    good = good and torch.equal(v, ref)
check(f"big allreduce x50 (n~{big})", good and hot.peer_error() == 0)


# ---- latencies
def timed(fn, reps=500):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


vec = torch.rand(lm.num_vertices * 3, dtype=torch.float64, device=dev)
small = torch.rand(8, dtype=torch.float64, device=dev)
bigv = torch.rand(big, dtype=torch.float64, device=dev)
res = {}
for mode in (False, True):
    hot.peer_enable(mode)
    res[mode] = (timed(lambda: hot.halo_exchange(vec, 3)), timed(lambda: hot.allreduce_sum(small)), timed(lambda: hot.allreduce_sum(bigv), 200))
if rank == 0:
    print(f"  latency us (halo n_comp=3, allreduce 8, allreduce {big}): NCCL {res[False][0]:.1f} {res[False][1]:.1f} {res[False][2]:.1f} | "
          f"peer {res[True][0]:.1f} {res[True][1]:.1f} {res[True][2]:.1f}", flush=True)
    print("peer_check", "PASS" if ok else "FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
