"""Small fixed workload for ncu: mdimg_metrics (+NIQE) / mdimg_quality / mdimg_fullref on n CT slices.
    python tools/ncu_metrics.py [slices] [passes]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ops = get_ops()
raw = np.stack([synth.ct_slice(1000 + z, z / n) for z in range(min(n, 64))])
raw = np.tile(raw, (-(-n // raw.shape[0]), 1, 1))[:n]
x = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device))
y = torch.clamp(x * 0.9 + 0.02, 0, 1).contiguous()
for _ in range(passes):
    rows = ops.metrics(x, with_niqe=True)
    q = ops.quality(y, niqe=True)
    fr = ops.fullref(x, y)
torch.cuda.synchronize()
print("launches", ops.lib.mdimg_launch_count(), float(rows[0, 0]), float(q[0, 0]), float(fr[0, 0]))
