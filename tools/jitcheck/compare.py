"""Compare host<tag>.bin (g++) with gpurun_out/dev<tag>.bin (B200): usage  compare.py <neq>:<tag> ..."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
ok = True
for arg in sys.argv[1:]:
    n, tag = arg.split(":")
    n = int(n)
    per = 2 * n + 3 * n * n
    h = np.fromfile(ROOT / "tools" / "jitcheck" / f"host{tag}.bin").reshape(-1, per)
    d = np.fromfile(ROOT / "gpurun_out" / f"dev{tag}.bin").reshape(-1, per)
    rel = np.abs(h - d) / np.abs(h).max(axis=1, keepdims=True)
    bad = np.argwhere(rel > 1e-10)
    print(f"{tag} (n_eq={n}): max |host-dev| / row scale = {rel.max():.3e}, mismatches {len(bad)}")
    ok = ok and len(bad) == 0
sys.exit(0 if ok else 1)
