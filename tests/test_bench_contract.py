"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads, and the algorithmic-byte
model of the roofline matches SURVEY.md section 8(d) for the bench mesh."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    """`--impl reference` = the compiled CPU restatement on the SAME workload (here shrunk with --n so that it runs in seconds)."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--n", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"].startswith("streamer BDF2 steps/s") and d["dtype"] == "f64" and d["scaling"] == "strong"
    assert d["steps"] == 2 and d["warmup"] == 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "n=24" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, str(ROOT))
    import bench
    assert d["config"] == bench.workload_config(24)               # the very dict the GPU arm prints
    assert d["solver"]["krylov_its_per_step"] > 0 and d["solver"]["newton_its_per_step"] >= 1


def test_cpu_baseline_sample_reports_measured_exponent():
    sys.path.insert(0, str(ROOT))
    from oracle import cpu_baseline
    r = cpu_baseline.sample(full_n=64, sizes=(16, 32), threads=2)
    assert r["kind"] == "port" and r["value"] > 0 and 0.3 < r["exponent"] < 3.0 and len(r["measured"]) == 2
    assert "measured exponent" in r["sample"]


def test_algorithmic_bytes_model():
    sys.path.insert(0, str(ROOT))
    import bench
    nv, nc = 1826 * 1826, 2 * 1825 * 1825
    asm, spmv, nnz = bench.algorithmic_bytes(nv, nc)
    assert nnz == 209_927_934                                        # 9 * (N_v + 2 N_e)
    asm_all, spmv_all, _ = bench.algorithmic_bytes(nv, nc, zero_entries=0)
    assert abs(asm_all - 2.133e9) < 2e6                              # 2.133 GB for the full 9-entry blocks
    assert asm == asm_all - 8 * nnz // 9 and spmv == spmv_all - 8 * nnz // 9
