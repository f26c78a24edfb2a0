"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; tests/test_gpu_parity_timed.py spawns it):
the partitioned streamer run must reproduce the CPU oracle (and hence the 1-GPU run) to 1e-8.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/mgpu_check.py [--golden] [--mode exact|bench] [--n 32 --steps 4]

--golden: n = 256, T_final = 1e-10 (21 steps) against tests/golden/streamer_n256_T1e-10.npz (strips of >= 32 node rows at 8
ranks, so the partition is not degenerate); otherwise a small case against the oracle run on rank 0.
--mode bench: GMRES(30) rtol 1e-5 (what bench.py times) instead of tight linear solves.
"""
import argparse
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import torch.distributed as dist

from fedm_b200.streamer import StreamerCase

ap = argparse.ArgumentParser()
ap.add_argument("--golden", action="store_true")
ap.add_argument("--mode", default="exact", choices=["exact", "bench"])
ap.add_argument("--precond", default="amg", choices=["amg", "ilu0", "bjacobi"])
ap.add_argument("--n", type=int, default=32)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 256 if args.golden else args.n
kw = dict(ksp_rtol=1e-13, ksp_max_it=50000) if args.mode == "exact" else {}
case = StreamerCase(n=n, device=local, world=world, rank=rank, precond=args.precond, **kw)
case.set_initial_state()
if args.golden:
    case.run(T_final=1e-10)
else:
    case.run(max_steps=args.steps)
full = case.gather_state()                                       # (n_global, 3) on every rank
lm = case.local
# ghost consistency: every rank's ghost entries equal the owners' values
mine = case.x_new.view(-1, 3)
ghost_err = (mine - full[torch.from_numpy(lm.local_to_global).to(mine.device)]).abs().max().item()
ok = ghost_err == 0.0
if rank == 0:
    X = full.cpu().numpy()
    rows = np.array(case.rows)
    if args.golden:
        g = np.load(ROOT / "tests" / "golden" / "streamer_n256_T1e-10.npz")
        errs = [float(np.abs(X[g["idx"], k] - g["sample"][:, k]).max() / np.abs(g["sample"][:, k]).max()) for k in range(3)]
        nrm = [abs(np.linalg.norm(X[:, k]) / g["norms"][k, 0] - 1.0) for k in range(3)]
        rows_ref, its_ref = g["rows"], g["newton_its"].tolist()
    else:
        from oracle.streamer import StreamerProblem, streamer_mesh
        prob = StreamerProblem(streamer_mesh(n))
        x_ref, rows_ref, its = prob.run(max_steps=args.steps)
        Xr = x_ref.reshape(-1, 3)
        errs = [float(np.abs(X[:, k] - Xr[:, k]).max() / np.abs(Xr[:, k]).max()) for k in range(3)]
        nrm = [0.0]
        rows_ref, its_ref = np.array(rows_ref), [h[0] for h in its]
    same_newton = [s[0] for s in case.stats] == its_ref
    if args.mode == "exact":
        rows_ok = rows.shape == rows_ref.shape and np.array_equal(rows[:, 1:], rows_ref[:, 1:]) and \
            np.abs(rows[:, 0] / rows_ref[:, 0] - 1).max() <= 1e-7
    else:
        rows_ok = rows.shape == rows_ref.shape and np.abs(rows / rows_ref - 1).max() <= 1e-5
    ok = ok and max(errs) <= 1e-8 and max(nrm) <= 1e-8 and same_newton and rows_ok
    print(f"mgpu_check world={world} n={n} mode={args.mode}: rel errors {errs}, ghost_err {ghost_err}, newton its equal "
          f"{same_newton}, log rows ok {rows_ok}, krylov its/step {np.mean([s[1] for s in case.stats]):.1f} -> "
          f"{'PASS' if ok else 'FAIL'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=mine.device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
