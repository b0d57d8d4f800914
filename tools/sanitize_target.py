"""Small workload for compute-sanitizer: the round-2 kernels on awkward shapes."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stack  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

ops = get_ops()
rng = np.random.default_rng(0)
for shape in [(3, 94, 141), (2, 9, 11), (2, 64, 200), (1, 130, 257), (2, 72, 1000), (1, 600, 200), (3, 128, 128)]:
    x = torch.from_numpy(rng.random(shape, dtype=np.float32)).to(ops.device)
    rows = ops.metrics(x, with_niqe=True)
    q = ops.quality(x, niqe=True)
    sel = torch.tensor([shape[0] - 1], dtype=torch.int32, device=ops.device)
    rows2 = ops.metrics(x, with_niqe=True, sel=sel)
    out = torch.empty_like(x)
    ops.wavelet_denoise(x, out, mode="soft")
    ops.wavelet_denoise(x, out, mode="hard")
    ops.light_denoise(x, out, 0.3)
    if shape[1] >= 7 and shape[2] >= 7:
        ops.fullref(x, out)
    for ks in (8, 16, 32, 7):          # interior (vectorised) and border regions of the CLAHE histogram kernel
        ops.clahe(x, out, 0.02, ks)
    raw = torch.from_numpy(rng.integers(0, 4096, shape, dtype=np.uint16).view(np.int16)).to(ops.device)
    ops.normalize(raw)
    torch.cuda.synchronize()
    print("ok", shape, float(rows[0, 0]), float(q[0, 1]))
raw = np.stack([synth.ct_slice(1000 + z, z / 3, size=128) for z in range(3)])
res = process_stack(torch.from_numpy(raw.view(np.int16)).to(ops.device), synth.plan_full(), chunk=2, ops=ops)
torch.cuda.synchronize()
print("pipeline ok", res.labels[0])
