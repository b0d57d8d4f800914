"""mdimg_metrics on 512 slices with different MDIMG_METRICS_SUB settings (one child process each)."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from mdimg_b200 import synth
    from mdimg_b200.stack import get_ops
    ops = get_ops()
    n = 512
    raw = np.stack([synth.ct_slice(1000 + z, z / 64) for z in range(64)])
    raw = np.tile(raw, (8, 1, 1))
    x = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device))
    for _ in range(3):
        rows = ops.metrics(x, with_niqe=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        rows = ops.metrics(x, with_niqe=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"SUB={os.environ.get('MDIMG_METRICS_SUB')}: {e0.elapsed_time(e1) / 5:.3f} ms per call, checksum {float(rows.nan_to_num().sum()):.9e}")
else:
    for sub in ("0", "16", "24", "32", "48", "64", "128"):
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MDIMG_METRICS_SUB=sub), check=True)
