"""Turns an `ncu --set full` report into the committed evidence: a text summary per kernel under profiles/ and the JSON file
bench.py reads its `traffic` / `fp64_pipe_active_frac` fields from (run in the build container: `ncu -i` needs no GPU).

    python tools/ncu_summary.py gpurun_out/r2_prof_final.ncu-rep profiles/r2_ncu_dominant_kernels.txt profiles/r2_ncu_summary.json
"""
import csv
import io
import json
import subprocess
import sys

rep, out_txt, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


groups = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = "assemble_jacobian" if "assemble_tiled_kernel" in name and ", 1," in name.replace("(bool)", "").replace("(int)", "") else \
            "assemble_residual" if "assemble_tiled_kernel" in name else "spmv" if "spmv_kernel" in name else name.split("(")[0][-40:]
    groups.setdefault(short, []).append(r)
lines, summary = [f"source report: {rep} (ncu --set full --clock-control none; per-launch values are cold-cache, serialised)\n"], {}
for short, rs in groups.items():
    lines.append(f"== {short}: {rs[0][idx['Kernel Name']][:150]}   ({len(rs)} launch(es) captured)")
    for k in KEYS:
        if k in idx:
            vals = [r[idx[k]] for r in rs]
            lines.append(f"   {k:<85} {', '.join(vals)} {units[idx[k]]}")
    rd = [to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) for r in rs]
    wr = [to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in rs]
    summary[short] = {"dram_bytes_per_launch": sum(rd) / len(rd) + sum(wr) / len(wr),
                      "fp64_pipe_active_frac": sum(float(r[idx["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]]) for r in rs) / len(rs) / 100.0,
                      "duration_us_under_ncu": sum(float(r[idx["gpu__time_duration.sum"]]) for r in rs) / len(rs),
                      "source": out_txt}
    lines.append("")
open(out_txt, "w").write("\n".join(lines))
json.dump(summary, open(out_json, "w"), indent=1)
print(json.dumps(summary, indent=1))
