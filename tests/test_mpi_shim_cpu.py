"""dolfin.MPI of the shim over torch.distributed, world size 2 on the CPU (gloo): the reductions FEDM's utilities rely on
(`mesh_statistics`: MPI.sum(num_cells), MPI.max(hmax), MPI.min(hmin), fedm/utils.py:22-35) and rank-0-only output."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

WORKER = '''import sys
sys.path.insert(0, ROOT_PATH)
from fedm_b200 import dolfin as df
rank, size = df.MPI.rank(df.MPI.comm_world), df.MPI.size(df.MPI.comm_world)
assert size == 2
assert df.MPI.sum(df.MPI.comm_world, rank + 1) == 3
assert df.MPI.max(df.MPI.comm_world, 10.0 * (rank + 1)) == 20.0
assert df.MPI.min(df.MPI.comm_world, 10.0 * (rank + 1)) == 10.0
mesh = df.RectangleMesh(df.Point(0.0, 0.0), df.Point(1.0, 2.0), 7, 9)
# every rank holds the whole mesh; the local cell counts are shares that add up to the global count
assert df.MPI.sum(df.MPI.comm_world, mesh.num_cells()) == 2 * 7 * 9
assert df.MPI.max(df.MPI.comm_world, mesh.hmax()) == mesh.hmax()
df.MPI.barrier()
sys.stdout.write("rank%dok " % rank)
'''


def test_mpi_shim_reductions_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.replace("ROOT_PATH", repr(str(ROOT))))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29691", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank0ok" in r.stdout and "rank1ok" in r.stdout
