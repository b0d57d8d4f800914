"""A/B of the metrics kernels: strip kernel (default) against the previous tile kernels
(MDIMG_METRICS_TILES=1, read once per process, hence two child processes).

    python tools/ab_metrics.py            # runs both children, compares rows, prints timings
Writes gpurun_out/ab_metrics.txt.
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(out_path: str) -> None:
    import torch
    from mdimg_b200 import synth
    from mdimg_b200.stack import get_ops
    ops = get_ops()
    dev = ops.device
    rng = np.random.default_rng(7)
    cases = {}
    ims = {
        "ct512": np.stack([synth.ct_slice(1000 + z, z / 8).astype(np.float32) / 4095.0 for z in range(8)]),
        "unit256": np.stack([synth.unit_image(4000 + k, 256) for k in range(3)]),
        "odd94x141": np.stack([synth.unit_image(4001, 256)[:94, :141]]),
        "tiny9x11": rng.random((2, 9, 11), dtype=np.float32),
        "cr600": np.stack([synth.radiograph(2000, 600).astype(np.float32) / 4095.0]),
        "wide40x700": rng.random((1, 40, 700), dtype=np.float32),
        "flat": np.full((1, 64, 64), 0.25, np.float32),
        "zeros": np.zeros((1, 64, 64), np.float32),
        "negative": (rng.random((1, 100, 130), dtype=np.float32) - 0.3),
        "tinyvals": (rng.random((1, 64, 200), dtype=np.float32) * 1e-25).astype(np.float32),
    }
    for name, a in ims.items():
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        cases[f"{name}|metrics"] = ops.metrics(t, with_niqe=True).cpu().numpy()
        cases[f"{name}|quality"] = ops.quality(t, niqe=True).cpu().numpy()
    # timing: 512 slices 512x512
    n = 512
    base = np.stack([synth.ct_slice(1000 + z, z / 64).astype(np.float32) / 4095.0 for z in range(64)])
    t = torch.from_numpy(np.tile(base, (n // 64, 1, 1))).to(dev)
    for fn, label in ((lambda: ops.metrics(t, with_niqe=True), "metrics"), (lambda: ops.quality(t, niqe=True), "quality")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        cases[f"time|{label}"] = np.array([e0.elapsed_time(e1) / 5])
    np.savez(out_path, **cases)


def main() -> None:
    if len(sys.argv) > 2 and sys.argv[1] == "child":
        child(sys.argv[2])
        return
    out_dir = ROOT / "gpurun_out"
    out_dir.mkdir(exist_ok=True)
    res = {}
    for mode in ("0", "1"):
        path = out_dir / f"ab_metrics_{mode}.npz"
        env = dict(os.environ, MDIMG_METRICS_TILES=mode)
        subprocess.run([sys.executable, __file__, "child", str(path)], check=True, env=env)
        res[mode] = np.load(path)
    lines = []
    names = ["sigma", "lap_var", "std", "pct_low", "pct_high", "entropy", "edge_density", "grad_mean", "grad_std",
             "snr", "cnr", "lap_energy", "hist_spread", "local_contrast", "grad_strength", "grad_entropy", "mean",
             "edge_ratio", "niqe", "var_of_var", "gmax", "p05", "p95", "t90"]
    worst = 0.0
    for key in res["0"].files:
        a, b = res["0"][key], res["1"][key]
        if key.startswith("time|"):
            lines.append(f"{key:24s} strip {a[0]:8.3f} ms   tiles {b[0]:8.3f} ms   ({b[0] / a[0]:.2f}x)")
            continue
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-12)
        rel = np.where(np.isnan(a) & np.isnan(b), 0.0, rel)
        rel = np.where(a == b, 0.0, rel)
        col = np.nanmax(rel, axis=0)
        bad = [(names[i] if key.endswith("metrics") else ("edge_ratio", "niqe")[i], float(col[i])) for i in range(len(col)) if not col[i] <= 0]
        worst = max(worst, float(np.nanmax(col))) if np.isfinite(np.nanmax(col)) else float("inf")
        lines.append(f"{key:24s} max rel diff {float(np.nanmax(col)):.3e}   nonzero: {bad}")
    lines.append(f"worst relative difference between the two kernel generations: {worst:.3e}")
    text = "\n".join(lines)
    print(text)
    (out_dir / "ab_metrics.txt").write_text(text + "\n")


if __name__ == "__main__":
    main()
