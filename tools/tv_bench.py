"""Per-body cost of the TV-Chambolle kernels: (41 bodies - 1 body) / 40 on a 512-slice chunk.
    MDIMG_TV_K=2|4 python tools/tv_bench.py [slices] [size]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
size = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ops = get_ops()
raw = np.stack([synth.ct_slice(1000 + z, z / 64, size=size) for z in range(64)])
raw = np.concatenate([raw] * (n // 64), 0) if n >= 64 else raw[:n]
x = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device))
y = torch.empty_like(x)


def tv_ms(iters, eps=0.0):
    ops.tv_chambolle(x, y, 0.05, eps=eps, max_iter=iters)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    it = ops.tv_chambolle(x, y, 0.05, eps=eps, max_iter=iters)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b), it


t41, _ = tv_ms(41)
t1, _ = tv_ms(1)
per = (t41 - t1) / 40
px = x.numel()
tfull, it = tv_ms(200, eps=2e-4)
print(f"K={os.environ.get('MDIMG_TV_K', 'default')} {n}x{size}x{size}: {per:.4f} ms/body  "
      f"{px / per / 1e6:.1f} Gpx/s/body  algorithmic {20 * px / per / 1e6:.0f} GB/s; "
      f"full run (eps 2e-4) {tfull:.2f} ms, mean iters {it.float().mean().item():.1f}")
