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
    h = np.fromfile(ROOT / "tools" / "jitcheck" / f"host{tag}.bin")
    d = np.fromfile(ROOT / "gpurun_out" / f"dev{tag}.bin")
    assert h.size == d.size, (h.size, d.size)
    if h.size > 256 * per:                      # cell-wise records follow the 256 row cases
        hc, dc = h[256 * per:], d[256 * per:]
        scale = np.abs(hc).max()
        print(f"{tag}: cell records max |host-dev| / max = {np.abs(hc - dc).max() / scale:.3e}")
        ok = ok and np.abs(hc - dc).max() <= 1e-10 * scale
    h, d = h[:256 * per].reshape(-1, per), d[:256 * per].reshape(-1, per)
    rel = np.abs(h - d) / np.abs(h).max(axis=1, keepdims=True)
    bad = np.argwhere(rel > 1e-10)
    print(f"{tag} (n_eq={n}): max |host-dev| / row scale = {rel.max():.3e}, mismatches {len(bad)}")
    ok = ok and len(bad) == 0
sys.exit(0 if ok else 1)
