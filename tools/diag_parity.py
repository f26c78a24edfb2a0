"""Diagnostic: bench semantics (GMRES rtol 1e-5) vs exact linear solves at the bench size -- where do the trajectories differ?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from fedm_b200.streamer import StreamerCase

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1825
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 25
a = StreamerCase(n=n)
b = StreamerCase(n=n, ksp_rtol=1e-13, ksp_max_it=50000)
for c in (a, b):
    c.set_initial_state(bench_initial_potential=True)
for s in range(steps):
    a.step(); b.step()
    Xa, Xb = a.x_new.view(-1, 3), b.x_new.view(-1, 3)
    d = (Xa - Xb).abs()
    k = int(d[:, 1].argmax())
    print(f"step {s+1}: newton {a.stats[-1][0]}/{b.stats[-1][0]} krylov {a.stats[-1][1]}/{b.stats[-1][1]} fnorm0 {a.stats[-1][2]:.3e} fnorm {a.stats[-1][3]:.3e}/{b.stats[-1][3]:.3e} "
          f"max|d ln ne| {float(d[k,1]):.3e} at node {k} (r,z)=({a.mesh.coords[k,0]:.4e},{a.mesh.coords[k,1]:.4e}) ln ne = {float(Xa[k,1]):.4f}/{float(Xb[k,1]):.4f} "
          f"max|d ln ni| {float(d[:,0].max()):.2e} max|d phi| {float(d[:,2].max()):.2e} "
          f"density-weighted rel diff ne {float(((Xa[:,1].exp()-Xb[:,1].exp()).abs()).max()/Xb[:,1].exp().max()):.2e}", flush=True)
