"""ncu target: compute_metrics (+ NIQE) and compute_validation extras on a 512-slice CT chunk.
    python tools/ncu_metrics.py [slices]
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ops = get_ops()
raw = np.stack([synth.ct_slice(1000 + z, z / 64) for z in range(64)])
raw = np.concatenate([raw] * max(1, n // 64), 0)[:n]
x = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device))
y = torch.empty_like(x)
ops.gamma(x, y, 0.9)
for _ in range(2):
    rows = ops.metrics(x, with_niqe=True)
    fr = ops.fullref(x, y)
torch.cuda.synchronize()
print("ok", rows.shape, float(fr[0, 0]))
