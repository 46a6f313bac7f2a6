"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share.  Usage: python tools/summarize_launches.py gpurun_out/launches.csv [top]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name][0] += 1
    tot[name][1] += v * scale
total = sum(v[1] for v in tot.values())
n = sum(v[0] for v in tot.values())
print("launches: %d   total device time: %.3f ms (cold-cache, serialised under ncu: compare SHARES)" % (n, total / 1e3))
for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%6.2f%%  %9.1f us  x%-4d %s" % (100 * us / total, us, c, name[:110]))
