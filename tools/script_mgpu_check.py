"""Drop-in SCRIPT path on several GPUs (run under torchrun; tests/test_gpu_parity_timed.py spawns it):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 \
        tools/script_mgpu_check.py [--n 128 --steps 4]

Every rank executes the UNCHANGED user script examples/streamer_discharge.py (FEDM API, dolfin shim) -- the reference's
`mpirun -np N python3 fedm-streamer.py` (README.md:63-67).  The shim keeps the vectors replicated and partitions every solve
(fedm_b200/dolfin/cuda_backend.py); the result must equal the single-GPU hand-kernel run of the same case to 1e-8 and be
identical on all ranks; `MPI.sum / max` must give global values (fedm/utils.py:22-35).
"""
import argparse
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "compat"), str(ROOT / "examples")]
import numpy as np
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="streamer", choices=["streamer", "glow"])
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])

import dolfin                                                     # the shim; creates the process group on first MPI use

if args.case == "glow":
    # the reference's glow-discharge test case (5 coupled equations, tables, adaptive BDF2, XDMF/HDF5 checkpoints) under N ranks:
    # rank 0 writes the files; they must equal the FEniCS goldens like the 1-GPU run does (tests/test_gpu_shim.py)
    import glow_discharge as gd
    from fedm_b200.hdf5_lite import read_h5
    gold = np.load(ROOT / "tests" / "golden" / "glow_h5.npz")
    tmp = tempfile.mkdtemp() if rank == 0 else None
    box = [tmp]
    dist_ = dolfin.MPI._dist()
    dist_.broadcast_object_list(box, src=0)
    out = Path(box[0])
    gd.main(input_dir=ROOT / "tests" / "golden" / "glow_discharge_input", output_dir=out)
    dolfin.MPI.barrier()
    ok = True
    if rank == 0:
        errs = {}
        for name in ("electrons", "Ar_plus", "Ar_star"):
            vecs = read_h5(out / "number density" / name / f"{name}.h5", name)
            ref = gold[name + "_1"]
            errs[name] = float(np.abs((vecs[1].ravel() - ref) / ref).max())
        rows = (out / "relative error.log").read_text().splitlines()
        ok = max(errs.values()) <= 1e-8 and len(rows) == 6
        print(f"script_mgpu_check world={world} case=glow: rel errors vs FEniCS golden HDF5 {errs}, log rows {len(rows)} -> "
              f"{'PASS' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=torch.device("cuda", local))
    dist_.all_reduce(flag, op=dist_.ReduceOp.MIN)
    dist_.barrier()
    dist_.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)

import streamer_discharge as sd

with tempfile.TemporaryDirectory() as tmp:
    u_new, mesh = sd.main(n=args.n, input_dir=ROOT / "tests" / "golden" / "streamer_discharge_input",
                          output_dir=Path(tmp) / f"rank{rank}", linear_solver="gmres", max_steps=args.steps)
x = u_new.vector().store.t
# replicated result: identical on all ranks
ref = x.clone()
dist.broadcast(ref, 0)
same = bool(torch.equal(ref, x))
n_cells = dolfin.MPI.sum(dolfin.MPI.comm_world, mesh.num_cells())
hmax = dolfin.MPI.max(dolfin.MPI.comm_world, mesh.hmax())
ok = same and n_cells == 2 * args.n * args.n and hmax == mesh.hmax()
prob = sd.main.last["solver"]._compiled[id(sd.main.last["problem"])][1]
partitioned = prob.local is not None and prob.local.world == world and prob.hot.n_owned < (args.n + 1) ** 2
if rank == 0:
    from fedm_b200.streamer import StreamerCase
    case = StreamerCase(n=args.n, device=local)
    case.set_initial_state()
    case.run(max_steps=args.steps)
    X, Xr = x.view(-1, 3), case.x_new.view(-1, 3)
    errs = [float((X[:, k] - Xr[:, k]).abs().max() / Xr[:, k].abs().max()) for k in range(3)]
    ok = ok and partitioned and max(errs) <= 1e-8
    print(f"script_mgpu_check world={world} n={args.n}: rel errors vs 1-GPU hand-kernel run {errs}, identical on all ranks {same}, "
          f"partitioned solve {partitioned} (owned rows on rank 0: {prob.hot.n_owned}), MPI.sum(num_cells) = {n_cells} -> "
          f"{'PASS' if ok else 'FAIL'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=x.device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
