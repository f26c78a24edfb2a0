"""Micro-benchmark of the assembly / SpMV / orthogonalisation kernels at bench size (run on a GPU box)."""
import sys, time, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from fedm_b200.streamer import StreamerCase

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1825
case = StreamerCase(n=n, precond="bjacobi")
ui, ue = case.initial_log_densities()
x0 = np.stack([ui, ue, 18750.0 * case.mesh.coords[:, 1] / 0.0125 + 50 * np.sin(case.mesh.coords[:, 0] * 2000)], axis=1).ravel()
x = torch.from_numpy(x0).to(case.device)
hot = case.hot
hot.bdf2_history(2, x, x, 5e-12, 1e30)
b = hot.new_vector()
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
dof = 3 * case.mesh.num_vertices
tj = timeit(lambda: hot.lib.fedmx_assemble_jacobian(hot.h, x.data_ptr(), 5e-12, 1e30, 0.0, None, 0, None))
tr = timeit(lambda: hot.residual(x, 5e-12, 1e30, out=b))
ts = timeit(lambda: hot.spmv(x, b))
print(f"n={n} dof={dof} MINB={os.environ.get('FEDMX_ASM_MINB','default')}: jacobian {tj:.3f} ms ({dof/tj/1e6:.2f} GDOF/s, {2.1328e9*(dof/10002828)/tj/1e6:.0f} GB/s)  residual {tr:.3f} ms  spmv {ts:.3f} ms")
# bit-level check of the kernel variant selected by FEDMX_ASM_MINB on a small mesh: print a checksum of the assembled matrix
small = StreamerCase(n=96, precond="bjacobi")
ui, ue = small.initial_log_densities()
xs = torch.from_numpy(np.stack([ui, ue, 18750.0 * small.mesh.coords[:, 1] / 0.0125 + 50 * np.sin(small.mesh.coords[:, 0] * 2000)], axis=1).ravel()).to(small.device)
small.hot.bdf2_history(2, xs, xs, 5e-12, 1e30)
small.hot.jacobian(xs, 5e-12, 1e30)
vals = small.hot.matrix_bsr()[2]
import hashlib
print("matrix sha1", hashlib.sha1(vals.tobytes()).hexdigest()[:16], "sum", float(vals.sum()))
