"""Resident throughput of the C2 stack pipeline against chunk size and worker count.
    python tools/chunk_sweep.py [slices]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stack  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ops = get_ops()
base = np.stack([synth.ct_slice(1000 + z, z / 128) for z in range(128)])
raw = np.tile(base, (-(-n // 128), 1, 1))[:n]
dev = torch.from_numpy(raw.view(np.int16)).to(ops.device)
plan = synth.plan_full()
for chunk, workers in [(512, 2), (512, 4), (256, 4), (256, 8), (128, 4), (128, 8), (342, 3), (1024, 1), (171, 6)]:
    last = None
    for _ in range(2):
        last = process_stack(dev, plan, chunk=chunk, ops=ops, workers=workers)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        last = process_stack(dev, plan, chunk=chunk, ops=ops, workers=workers)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"chunk {chunk:5d} workers {workers}: {ms:7.2f} ms per {n}-slice stack = {n * 512 * 512 / ms / 1e3:8.1f} Mpx/s", flush=True)
