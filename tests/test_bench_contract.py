"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads, and the algorithmic-byte
model of the roofline matches SURVEY.md section 8(d) for the bench mesh."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--sample-n", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("streamer BDF2 steps/s") and d["dtype"] == "f64" and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "24x24" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_algorithmic_bytes_model():
    sys.path.insert(0, str(ROOT))
    import bench
    nv, nc = 1826 * 1826, 2 * 1825 * 1825
    asm, spmv, nnz = bench.algorithmic_bytes(nv, nc)
    assert nnz == 209_927_934                                        # 9 * (N_v + 2 N_e)
    asm_all, spmv_all, _ = bench.algorithmic_bytes(nv, nc, zero_entries=0)
    assert abs(asm_all - 2.133e9) < 2e6                              # 2.133 GB for the full 9-entry blocks
    assert asm == asm_all - 8 * nnz // 9 and spmv == spmv_all - 8 * nnz // 9
