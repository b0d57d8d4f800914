"""Per-operator device-time breakdown of the stack pipeline (CUDA events around every C-ABI call).
    python tools/profile_ops.py [slices] [chunk ...]
"""
from __future__ import annotations

import json
import os
import sys
import time
from collections import defaultdict
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
# per-operator timing needs the step calls issued from Python: mdimg_enhance would be one opaque call
os.environ["MDIMG_NATIVE_ENGINE"] = "0"
sys.path.insert(0, str(ROOT))

from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stack  # noqa: E402
from mdimg_b200.stack import StackOps, get_ops  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    chunks = [int(c) for c in sys.argv[2:]] or [10, 32, 128]
    ops = get_ops()
    raw = np.stack([synth.ct_slice(1000 + z, z / n) for z in range(n)])
    raw_dev = torch.from_numpy(raw.view(np.int16)).to(ops.device)
    plan = synth.plan_full()
    report = {}
    orig_call = StackOps._call
    for chunk in chunks:
        process_stack(raw_dev, plan, chunk=chunk, ops=ops)     # warm-up
        torch.cuda.synchronize()
        events = []

        def timed_call(self, fn, *args, _events=events):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            orig_call(self, fn, *args)
            b.record()
            _events.append((fn.__name__ if hasattr(fn, "__name__") else str(fn), a, b))

        StackOps._call = timed_call
        l0 = ops.lib.mdimg_launch_count()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = process_stack(raw_dev, plan, chunk=chunk, ops=ops)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        StackOps._call = orig_call
        agg = defaultdict(float)
        cnt = defaultdict(int)
        for name, a, b in events:
            agg[name] += a.elapsed_time(b)
            cnt[name] += 1
        total = e0.elapsed_time(e1)
        launches = ops.lib.mdimg_launch_count() - l0
        print(f"\n== chunk {chunk}: {n} slices, device {total:.1f} ms, wall {wall:.1f} ms, "
              f"{n*512*512/total/1e3:.1f} Mpx/s, {launches} launches, sum(ops) {sum(agg.values()):.1f} ms")
        it_mean = float(res.tv_iterations.mean())
        # algorithmic bytes per pixel and call of every operator
        alg = {"mdimg_normalize_u16": 8, "mdimg_metrics": 4, "mdimg_fullref": 8, "mdimg_wavelet_denoise": 12,
               "mdimg_light_denoise": 16, "mdimg_clahe": 16, "mdimg_clahe_gamma": 16, "mdimg_gamma": 8, "mdimg_unsharp": 8,
               "mdimg_bilateral": 8, "mdimg_tv_chambolle": 20 * it_mean + 4, "mdimg_clip01": 8,
               "mdimg_estimate_sigma": 4, "mdimg_quality": 4}
        peak = 6552.6
        try:
            peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass
        px = n * 512 * 512
        for name, ms in sorted(agg.items(), key=lambda kv: -kv[1]):
            b = alg.get(name)
            roof = ""
            if b:
                n_chunks = max(1, -(-n // chunk))
                gbs = b * px * (cnt[name] / n_chunks) / (ms / 1e3) / 1e9     # every call covers one chunk
                roof = f"  alg {b:6.1f} B/px  {gbs:7.0f} GB/s  {100 * gbs / peak:5.1f}% of HBM peak"
            print(f"   {name:28s} {ms:9.2f} ms  {100*ms/total:5.1f}%  calls {cnt[name]:5d}  "
                  f"{ms/ n * 1e3:8.1f} us/slice{roof}")
        report[chunk] = {"total_ms": total, "wall_ms": wall, "launches": int(launches),
                         "ops": {k: v for k, v in agg.items()},
                         "tv_iters_mean": float(res.tv_iterations.mean())}
    out = ROOT / "gpurun_out" / "profile_ops.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
