"""The other BASELINE.json configurations (parity-tested in tests/test_gpu_large.py), timed on one GPU:
  C3  64 synthetic 3000x3000 uint16 radiographs, P_cr (CLAHE + unsharp heavy) + metrics + validation
  C5  metrics-only sweep (compute_metrics + compute_validation against a gamma-0.9 copy), 256^2 .. 4096^2,
      batches of ~256 Mpx per point
Prints one JSON object; CUDA-event timing after a warm-up pass, inputs resident in HBM.
    python tools/bench_configs.py [c3_images] [c5_mpx]
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.batch import process_stack  # noqa: E402
from mdimg_b200.engine import Engine  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402

PEAK = 6552.6
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mpx5 = float(sys.argv[2]) if len(sys.argv) > 2 else 256.0
    ops = get_ops()
    out = {"hbm_peak_gbs": PEAK}

    # ---- C1: one 512x512 slice through the drop-in functions (host numpy in / out), latency ----
    import time

    from mdimg_b200.pipeline import dicom_io as gio
    from mdimg_b200.pipeline import enhancement as genh
    from mdimg_b200.pipeline import metrics as gmet
    raw1 = synth.ct_slice(1000, 0.0)
    planf = synth.plan_full()

    def c1_chain():
        x = gio.normalize_image(raw1)
        m = gmet.compute_metrics(x)
        issues = gmet.detect_issues(m)
        enh, labels = genh.apply_enhancements(x, issues)
        val = gmet.compute_validation(x, enh)
        return gmet.compute_objective_score(val)

    def wall(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    x1 = gio.normalize_image(raw1)
    out["C1"] = {
        "deterministic_chain_ms": wall(c1_chain),          # normalize -> metrics -> issues -> enhance -> validate -> score
        "compute_metrics_ms": wall(lambda: gmet.compute_metrics(x1)),
        "apply_enhancements_from_params_P_full_ms": wall(lambda: genh.apply_enhancements_from_params(x1, planf)),
        "compute_validation_ms": wall(lambda: gmet.compute_validation(x1, x1)),
        "note": "wall-clock per call on one 512x512 slice, numpy in / numpy out (H2D + kernels + D2H)"}

    # ---- C3 ----
    base = [synth.radiograph(2000 + i) for i in range(min(n3, 8))]
    raw = np.stack([base[i % len(base)] for i in range(n3)])
    dev = torch.from_numpy(raw.view(np.int16)).to(ops.device)
    plan = synth.plan_cr()
    ms = timed(lambda: process_stack(dev, plan, chunk=16, ops=ops), reps=2)
    px = float(n3) * 3000 * 3000
    # P_cr = 72 B/px algorithmic
    out["C3"] = {"images": n3, "size": 3000, "ms": ms, "mpx_per_s": px / ms / 1e3, "images_per_s": n3 / ms * 1e3,
                 "algorithmic_GBps": 72 * px / ms / 1e6, "frac_of_hbm_peak": 72 * px / ms / 1e6 / PEAK,
                 "note": f"{len(base)} distinct images repeated to {n3}; chunk 16 images"}
    # per-operator device time of one C3 pass (CUDA events around every C-ABI call)
    from collections import defaultdict
    from mdimg_b200.stack import StackOps
    events = []
    orig_call = StackOps._call

    def timed_call(self, fn, *args):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig_call(self, fn, *args)
        b.record()
        events.append((fn.__name__, a, b))

    StackOps._call = timed_call
    os.environ["MDIMG_NATIVE_ENGINE"] = "0"        # step calls issued from Python, so that each one is timed
    res3 = process_stack(dev, plan, chunk=16, ops=ops, workers=1)
    torch.cuda.synchronize()
    os.environ.pop("MDIMG_NATIVE_ENGINE", None)
    StackOps._call = orig_call
    agg, cnt = defaultdict(float), defaultdict(int)
    for name, a, b in events:
        agg[name] += a.elapsed_time(b)
        cnt[name] += 1
    out["C3"]["ops_ms"] = {k: round(v, 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])}
    out["C3"]["ops_calls"] = dict(cnt)
    out["C3"]["guards"] = {"halo": int(res3.packed[:, 50].sum()), "noise": int(res3.packed[:, 51].sum()),
                           "over": int(res3.packed[:, 52].sum())}
    del dev
    torch.cuda.empty_cache()

    # ---- C5 ----
    eng = Engine(ops)
    out["C5"] = []
    for k, size in enumerate((256, 512, 1024, 2048, 4096)):
        n = max(1, int(round(mpx5 * 1e6 / (size * size))))
        distinct = min(n, 8)
        imgs = np.stack([synth.unit_image(4000 + k * 16 + i, size) for i in range(distinct)])
        x = torch.from_numpy(np.ascontiguousarray(imgs[np.arange(n) % distinct])).to(ops.device)
        y = torch.empty_like(x)
        ops.gamma(x, y, 0.9)

        def metrics_only():
            return ops.metrics(x)

        def validation():
            rb, ra, fr = eng.validation_rows(x, y)
            return rb.cpu()

        ms_m = timed(metrics_only)
        ms_v = timed(validation)
        px = float(n) * size * size
        out["C5"].append({"size": size, "images": n, "compute_metrics_ms": ms_m,
                          "compute_metrics_mpx_per_s": px / ms_m / 1e3,
                          "compute_metrics_GBps_alg4": 4 * px / ms_m / 1e6,
                          "compute_metrics_frac": 4 * px / ms_m / 1e6 / PEAK,
                          "compute_validation_ms": ms_v, "compute_validation_mpx_per_s": px / ms_v / 1e3,
                          "compute_validation_GBps_alg16": 16 * px / ms_v / 1e6,
                          "compute_validation_frac": 16 * px / ms_v / 1e6 / PEAK})
        del x, y
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "bench_configs.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
