"""Instruction mix of one kernel from an ncu report's source page (SASS).
    python tools/sass_hist.py report.ncu-rep [kernel-index]
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for bi, blk in enumerate(blocks[1:]):
    lines = blk.splitlines()
    name = lines[0][:90]
    rd = csv.DictReader(io.StringIO("\n".join(lines[1:])))
    agg = defaultdict(float)
    tot = 0.0
    samples = defaultdict(float)
    for r in rd:
        try:
            n = float(r["Instructions Executed"] or 0)
        except ValueError:
            continue
        op = r["Source"].strip().split()[0] if r["Source"].strip() else "?"
        if op.startswith("@"):
            op = r["Source"].strip().split()[1]
        op = op.split(".")[0]
        agg[op] += n
        tot += n
        try:
            samples[op] += float(r["# Samples"] or 0)
        except ValueError:
            pass
    print(f"\n== kernel {bi}: {name}\n   total warp-instructions {tot:.0f}")
    st = sum(samples.values()) or 1
    for op, n in sorted(agg.items(), key=lambda kv: -kv[1])[:28]:
        print(f"   {op:12s} {n:12.0f} {100*n/tot:6.2f}%   samples {100*samples[op]/st:5.1f}%")
