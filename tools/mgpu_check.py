"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
the partitioned streamer run must reproduce the CPU oracle (and hence the 1-GPU run) to 1e-8.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_check.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from fedm_b200.streamer import StreamerCase

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, steps = 32, 4
case = StreamerCase(n=n, device=local, world=world, rank=rank, ksp_rtol=1e-13, ksp_max_it=50000)
case.set_initial_state()
case.run(max_steps=steps)
lm = case.local
owned = case.x_new.view(-1, 3)[:lm.n_owned].contiguous()
full = torch.zeros((lm.n_global_vertices, 3), dtype=torch.float64, device=owned.device)
full[torch.from_numpy(lm.local_to_global[:lm.n_owned]).to(owned.device)] = owned
dist.all_reduce(full)
# ghost consistency: every rank's ghost entries equal the owners' values
mine = case.x_new.view(-1, 3)
ghost_err = (mine - full[torch.from_numpy(lm.local_to_global).to(owned.device)]).abs().max().item()
ok = ghost_err == 0.0
if rank == 0:
    from oracle.streamer import StreamerProblem, streamer_mesh
    prob = StreamerProblem(streamer_mesh(n))
    x_ref, rows_ref, its_ref = prob.run(max_steps=steps)
    x = full.cpu().numpy().ravel()
    errs = [float(np.abs(x[k::3] - x_ref[k::3]).max() / np.abs(x_ref[k::3]).max()) for k in range(3)]
    same_newton = [s[0] for s in case.stats] == [h[0] for h in its_ref]
    rows_ok = all(a == ar and b == br and abs(e - er) <= 1e-7 * er for (e, a, b), (er, ar, br) in zip(case.rows, rows_ref))
    ok = ok and max(errs) < 1e-8 and same_newton and rows_ok
    print(f"mgpu_check world={world}: rel errors {errs}, ghost_err {ghost_err}, newton its equal {same_newton}, "
          f"log rows ok {rows_ok}, krylov its {[s[1] for s in case.stats]} -> {'PASS' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
