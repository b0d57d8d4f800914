"""Developer parity report: every operator of the CUDA path vs the CPU oracle, with the actual
error numbers (pytest only says pass/fail).  Run on a GPU box:  python tests/gpu_check.py [out.json]
"""

from __future__ import annotations

import json
import sys
import time
import traceback
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mdimg_b200 import synth  # noqa: E402
from mdimg_b200.engine import METRIC_KEYS, Engine  # noqa: E402
from mdimg_b200.stack import get_ops  # noqa: E402
from oracle import exposure as oex  # noqa: E402
from oracle import filters as oflt  # noqa: E402
from oracle import ref_enhancement as oenh  # noqa: E402
from oracle import ref_metrics as omet  # noqa: E402
from oracle import restoration as ores  # noqa: E402
from oracle.fullref import peak_signal_noise_ratio, structural_similarity  # noqa: E402

REPORT = {}


def images():
    ims = {
        "clean64": synth.fixture_clean(),
        "noisy64": synth.fixture_noisy(),
        "lowc64": synth.fixture_low_contrast(),
        "ct512": omet.normalize_image(synth.ct_slice(1000)),
        "unit256": synth.unit_image(4000, 256),
    }
    rng = np.random.default_rng(5)
    odd = synth.unit_image(4001, 256)[:94, :141].copy()
    ims["odd94x141"] = np.ascontiguousarray(odd + rng.normal(0, 0.01, odd.shape).astype(np.float32)).clip(0, 1)
    cr = omet.normalize_image(synth.radiograph(2000, 600))
    ims["cr600"] = cr
    return ims


def dev(ops, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(ops.device)[None].contiguous()


def cmp_img(name, got, ref):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    finite = np.isfinite(ref)
    rec = {
        "max_abs": float(np.nanmax(d)) if d.size else 0.0,
        "n_diff": int((got != ref).sum()),
        "n": int(ref.size),
        "nan_mismatch": int((np.isnan(got) != np.isnan(ref)).sum()),
    }
    REPORT[name] = rec
    print(f"{name:48s} max_abs={rec['max_abs']:.3e} n_diff={rec['n_diff']}/{rec['n']}", flush=True)
    return rec


def section(fn):
    try:
        fn()
    except Exception:  # noqa: BLE001
        REPORT[fn.__name__ + "_error"] = traceback.format_exc()
        print("ERROR in", fn.__name__)
        traceback.print_exc()


def main():
    ops = get_ops()
    eng = Engine(ops)
    ims = images()

    def t_normalize():
        raw = synth.ct_slice(1000)
        got = ops.normalize(torch.from_numpy(raw.view(np.int16)).to(ops.device)[None])[0].cpu().numpy()
        cmp_img("normalize_u16/ct512", got, omet.normalize_image(raw))
        rawf = (synth.unit_image(1, 200) * 37 - 5).astype(np.float32)
        got = ops.normalize(dev(ops, rawf))[0].cpu().numpy()
        cmp_img("normalize_f32/unit200", got, omet.normalize_image(rawf))
        const = np.full((33, 47), 3.0, np.float32)
        cmp_img("normalize_f32/const", ops.normalize(dev(ops, const))[0].cpu().numpy(), omet.normalize_image(const))

    def t_metrics():
        for k, im in ims.items():
            rows = ops.metrics(dev(ops, im), with_niqe=True)[0].cpu().numpy()
            ref = omet.compute_metrics(im)
            worst = 0.0
            for i, key in enumerate(METRIC_KEYS):
                r, g = ref[key], rows[i]
                rel = abs(g - r) / max(abs(r), 1e-12)
                REPORT[f"metrics/{k}/{key}"] = {"ref": r, "got": float(g), "rel": rel}
                worst = max(worst, rel if abs(r) > 1e-9 else abs(g - r))
                if rel > 1e-5 and abs(g - r) > 1e-9:
                    print(f"   metrics/{k}/{key}: ref={r!r} got={g!r} rel={rel:.2e}")
            er, nq = omet.compute_edge_ratio(im), omet.compute_niqe_approximation(im)
            REPORT[f"metrics/{k}/edge_ratio"] = {"ref": er, "got": float(rows[17])}
            REPORT[f"metrics/{k}/niqe"] = {"ref": nq, "got": float(rows[18])}
            print(f"metrics/{k:12s} worst_rel={worst:.2e} edge_ratio d={abs(rows[17]-er):.2e} niqe d={abs(rows[18]-nq):.2e}", flush=True)
            q = ops.quality(dev(ops, im), niqe=True)[0].cpu().numpy()
            REPORT[f"quality/{k}"] = {"edge_d": abs(q[0] - er), "niqe_d": abs(q[1] - nq)}
            s = float(ops.estimate_sigma(dev(ops, im))[0].item())
            sr = float(ores.estimate_sigma(im))
            REPORT[f"sigma/{k}"] = {"ref": sr, "got": s}
            print(f"   sigma ref={sr!r} got={s!r}  quality d=({abs(q[0]-er):.2e},{abs(q[1]-nq):.2e})")

    def t_fullref():
        for k, im in ims.items():
            other = np.clip(im ** np.float32(0.9) + np.float32(0.01), 0, 1).astype(np.float32)
            fr = ops.fullref(dev(ops, im), dev(ops, other))[0].cpu().numpy()
            s = float(structural_similarity(im, other, data_range=1.0))
            p = float(peak_signal_noise_ratio(im, other, data_range=1.0))
            REPORT[f"fullref/{k}"] = {"ssim_ref": s, "ssim": float(fr[0]), "psnr_ref": p, "psnr": float(fr[1])}
            print(f"fullref/{k:12s} ssim d={abs(fr[0]-s):.2e} psnr d={abs(fr[1]-p):.2e}")
        same = ops.fullref(dev(ops, ims["clean64"]), dev(ops, ims["clean64"]))[0].cpu().numpy()
        print("fullref identical:", same)
        REPORT["fullref/identical"] = {"ssim": float(same[0]), "psnr": float(same[1])}

    def t_wavelet():
        for k, im in ims.items():
            for mode in ("soft", "hard"):
                x = dev(ops, im)
                out = torch.empty_like(x)
                ops.wavelet_denoise(x, out, mode=mode)
                cmp_img(f"wavelet_{mode}/{k}", out[0].cpu().numpy(), ores.denoise_wavelet(im, mode=mode))
            x = dev(ops, im)
            out = torch.empty_like(x)
            sk = ops.light_denoise(x, out, 0.3)
            ref = oenh.light_denoise(im, 0.3)
            cmp_img(f"light_denoise/{k}", out[0].cpu().numpy(), ref)
            REPORT[f"light_denoise/{k}/skipped"] = int(sk[0].item())

    def t_clahe():
        for k, im in ims.items():
            for clip, ks in ((0.015, 16), (0.03, 32), (0.08, 48), (0.002, 4), (0.02, 7)):
                x = dev(ops, im)
                out = torch.empty_like(x)
                st = ops.clahe(x, out, clip, ks)
                ref = oex.equalize_adapthist(im, kernel_size=ks, clip_limit=clip)
                cmp_img(f"clahe_c{clip}_k{ks}/{k}", out[0].cpu().numpy(), ref)

    def t_pointwise():
        for k, im in ims.items():
            x = dev(ops, im)
            out = torch.empty_like(x)
            ops.gamma(x, out, 0.95)
            cmp_img(f"gamma0.95/{k}", out[0].cpu().numpy(), oex.adjust_gamma(im, 0.95))
            ops.gamma(x, out, 1.3, assume_nonneg=True)
            cmp_img(f"gamma1.3/{k}", out[0].cpu().numpy(), oex.adjust_gamma(im, 1.3))
        neg = ops.gamma(dev(ops, ims["clean64"] - 0.5), torch.empty_like(dev(ops, ims["clean64"])), 0.9)
        REPORT["gamma/negflag"] = int(neg[0].item())

    def t_unsharp():
        for k, im in ims.items():
            for r, a in ((0.8, 0.5), (2.0, 1.5), (3.0, 2.5), (0.2, 0.03)):
                x = dev(ops, im)
                out = torch.empty_like(x)
                ops.unsharp(x, out, r, a)
                cmp_img(f"unsharp_r{r}_a{a}/{k}", out[0].cpu().numpy(), oflt.unsharp_mask(im, r, a))
        im = ims["noisy64"] - np.float32(0.3)
        x = dev(ops, im)
        out = torch.empty_like(x)
        ops.unsharp(x, out, 0.8, 0.5)
        cmp_img("unsharp_negative/noisy64", out[0].cpu().numpy(), oflt.unsharp_mask(im, 0.8, 0.5))

    def t_bilateral():
        for k, im in ims.items():
            for d_ in (5, 9, 3, 4):
                x = dev(ops, im)
                out = torch.empty_like(x)
                ops.bilateral(x, out, d_, 0.05, 0.05)
                cmp_img(f"bilateral_d{d_}/{k}", out[0].cpu().numpy(), oenh.bilateral_filter(im, d_, 0.05, 0.05))

    def t_tv():
        for k, im in ims.items():
            for w in (0.05, 0.15, 0.01):
                x = dev(ops, im)
                out = torch.empty_like(x)
                it = ops.tv_chambolle(x, out, w)
                ref, rit = ores.denoise_tv_chambolle(im, w, return_iters=True)
                rec = cmp_img(f"tv_w{w}/{k}", out[0].cpu().numpy(), ref)
                rec["iters"] = int(it[0].item())
                rec["iters_ref"] = int(rit)
                print(f"      iters got={rec['iters']} ref={rec['iters_ref']}")

    def t_pipeline():
        plan = synth.plan_full()
        for k in ("noisy64", "ct512", "unit256", "odd94x141"):
            im = ims[k]
            res = eng.enhance_from_params(dev(ops, im), plan)
            ref, labels = oenh.apply_enhancements_from_params(im, plan)
            rec = cmp_img(f"P_full/{k}", res.image[0].cpu().numpy(), ref)
            d = np.abs(res.image[0].cpu().numpy().astype(np.float64) - ref)
            rec["frac_gt_1lsb"] = float((d > 1.0 / 65535).mean())
            rec["labels_equal"] = labels == res.labels[0]
            rec["tv_iters"] = None if res.tv_iterations is None else int(res.tv_iterations[0])
            print("      labels equal:", rec["labels_equal"], "frac>1LSB:", rec["frac_gt_1lsb"], res.labels[0])
        plan2 = synth.plan_cr()
        for k in ("cr600", "ct512"):
            im = ims[k]
            res = eng.enhance_from_params(dev(ops, im), plan2)
            ref, labels = oenh.apply_enhancements_from_params(im, plan2)
            rec = cmp_img(f"P_cr/{k}", res.image[0].cpu().numpy(), ref)
            rec["labels_equal"] = labels == res.labels[0]
            print("      labels equal:", rec["labels_equal"], res.labels[0], labels)
        for issues in (["noise"], ["blur"], ["low_contrast", "clipping_low"], ["noise", "blur", "clipping_high"]):
            for k in ("noisy64", "ct512"):
                im = ims[k]
                res = eng.enhance_from_issues(dev(ops, im), issues)
                ref, labels = oenh.apply_enhancements(im, issues)
                rec = cmp_img(f"issues_{'+'.join(issues)}/{k}", res.image[0].cpu().numpy(), ref)
                rec["labels_equal"] = labels == res.labels[0]
                print("      labels equal:", rec["labels_equal"])

    def t_batch_consistency():
        # a stack of different slices must give the same rows as slice-by-slice calls
        stack = np.stack([omet.normalize_image(synth.ct_slice(1000 + z, z / 8)) for z in range(8)])
        xs = torch.from_numpy(stack).to(ops.device)
        rows = ops.metrics(xs, with_niqe=True).cpu().numpy()
        worst = 0.0
        for z in range(8):
            r1 = ops.metrics(xs[z:z + 1].contiguous(), with_niqe=True)[0].cpu().numpy()
            worst = max(worst, float(np.nanmax(np.abs(rows[z] - r1) / np.maximum(np.abs(r1), 1e-12))))
        REPORT["batch_consistency/metrics_worst_rel"] = worst
        print("batch consistency metrics worst rel", worst)
        sel = torch.tensor([1, 5, 6], dtype=torch.int32, device=ops.device)
        out = xs.clone()
        ops.bilateral(xs, out, 5, 0.05, 0.05, sel=sel)
        changed = [bool((out[z] != xs[z]).any().item()) for z in range(8)]
        REPORT["batch_consistency/sel_changed"] = changed
        print("sel changed:", changed)

    t0 = time.time()
    for fn in (t_normalize, t_metrics, t_fullref, t_wavelet, t_clahe, t_pointwise, t_unsharp, t_bilateral,
               t_tv, t_pipeline, t_batch_consistency):
        section(fn)
        torch.cuda.synchronize()
    REPORT["_seconds"] = time.time() - t0
    REPORT["_launches"] = int(ops.lib.mdimg_launch_count())
    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "gpurun_out" / "gpu_check.json"
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(REPORT, indent=1, default=str))
    print("wrote", out)


if __name__ == "__main__":
    main()
