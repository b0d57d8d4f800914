"""Small fixed workload for ncu: one warm-up + one measured pass of the stack pipeline.
    python tools/ncu_target.py [slices] [passes]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stack  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops = get_ops()
raw = np.stack([synth.ct_slice(1000 + z, z / n) for z in range(n)])
raw_dev = torch.from_numpy(raw.view(np.int16)).to(ops.device)
plan = synth.plan_full()
for _ in range(passes):
    res = process_stack(raw_dev, plan, chunk=n, ops=ops)
torch.cuda.synchronize()
print("launches", ops.lib.mdimg_launch_count(), "tv iters mean", res.tv_iterations.mean(), "failed", int(res.failed.sum()))
