"""Diagnostic: Newton/Krylov behaviour of the streamer step versus mesh size (run on a GPU box)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fedm_b200.streamer import StreamerCase

for n in [int(a) for a in sys.argv[2:]] or [128, 256, 512]:
    t0 = time.time()
    case = StreamerCase(n=n, ksp_max_it=3000, precond=sys.argv[1])
    t1 = time.time()
    case.set_initial_state(bench_initial_potential=True)
    t2 = time.time()
    print(f"n={n} setup {t1-t0:.1f}s poisson {t2-t1:.1f}s stats(its, kits, f0, f)={case.initial_poisson_stats}", flush=True)
    for s in range(3):
        k0 = len(case.stats)
        torch.cuda.synchronize(); ts = time.time()
        try:
            att = case.step()
        except SystemExit as e:
            print("  step failed:", e, case.stats[k0:][:6]); break
        torch.cuda.synchronize()
        print(f"  step {s}: {time.time()-ts:.3f}s attempts={att} dt={case.dt_old:g} err={case.rows[-1][0]:.3e} stats={case.stats[k0:]}", flush=True)
    case.hot.close()
