"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/summarize_launches.py launches.csv [skip_first_n]
"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((int(r["ID"]), r["Kernel Name"], val * scale))
rows.sort()
rows = rows[skip:]
agg = defaultdict(lambda: [0.0, 0])
for _, name, us in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = re.sub(r"mdimg::\(anonymous namespace\)::", "", short)
    agg[short][0] += us
    agg[short][1] += 1
total = sum(v[0] for v in agg.values())
print(f"{len(rows)} launches, {total/1e3:.3f} ms total device time (serialised, cold cache)")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for name, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{name[:60]:60s} {c:8d} {us:12.1f} {us/c:10.2f} {100*us/total:6.2f}%")
