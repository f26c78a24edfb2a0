"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (profiles/*.txt).
usage: python tools/summarize_launches.py gpurun_out/launches.csv "header comment" > profiles/rN_launches.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ik, ig, iv = H.index("Kernel Name"), H.index("Grid Size"), H.index("Metric Value")
agg = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    if len(r) <= iv or not r[iv].replace(",", "").replace(".", "").isdigit():
        continue
    name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
    n += 1
tot = sum(a[1] for a in agg.values())
for c in sys.argv[2:]:
    print("#", c)
print(f"# total kernel time {tot / 1e6:.2f} ms over {n} launches (per-launch times are cold-cache/serialised under ncu: compare SHARES)")
print(f"{'total_us':>12} {'share':>6} {'launches':>8} {'avg_us':>9}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1] / 1e3:12.1f} {100 * a[1] / tot:5.1f}% {a[0]:8d} {a[1] / a[0] / 1e3:9.1f}  {k[:110]}")
