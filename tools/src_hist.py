"""Per-CUDA-source-line instruction counts / stall samples of one kernel from an ncu report.
    python tools/src_hist.py report.ncu-rep kernel_regex [top_n]
"""
import csv
import io
import re
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
rows = []
cur_file = None
func = None
seen_funcs = set()
i = 0
recs = {}
for ln in lines:
    if ln.startswith('"File Path"'):
        cur_file = next(csv.reader([ln]))[1]
        continue
    if ln.startswith('"Function Name"'):
        func = next(csv.reader([ln]))[1]
        continue
    if ln.startswith('"Line No"'):
        hdr = next(csv.reader([ln]))
        continue
    if func is None or not re.search(pat, func):
        continue
    r = next(csv.reader([ln]))
    if len(r) < 9 or not r[0]:
        continue
    try:
        n = float(r[7]); smp = float(r[6])
    except ValueError:
        continue
    key = (func[:40], cur_file.split("/")[-1], int(r[0]), r[1].strip()[:90])
    a = recs.setdefault(key, [0.0, 0.0])
    a[0] += n; a[1] += smp
# ncu lists each kernel instance; keep totals per (file,line)
tot = sum(v[0] for v in recs.values()) or 1
ts = sum(v[1] for v in recs.values()) or 1
print(f"total warp-instr {tot:.0f}, samples {ts:.0f}")
for (fn, f, line, src), (n, smp) in sorted(recs.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot:6.2f}% instr {100*smp/ts:6.2f}% smp  {f}:{line:<4d} {src}")
