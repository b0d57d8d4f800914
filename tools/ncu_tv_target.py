"""TV-Chambolle alone on a stack, for an ncu capture of k_tvp: python tools/ncu_tv_target.py [slices]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ops = get_ops()
raw = np.stack([synth.ct_slice(1000 + z, z / 64) for z in range(64)])
raw = np.concatenate([raw] * max(1, n // 64), 0)[:n]
x = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device))
y = torch.empty_like(x)
it = ops.tv_chambolle(x, y, 0.05, eps=0.0, max_iter=12)
torch.cuda.synchronize()
print("tv iterations", it.float().mean().item())
