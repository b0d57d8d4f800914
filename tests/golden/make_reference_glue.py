"""Golden vectors from the REFERENCE'S OWN SOURCE for everything above the scikit-image leaves.

scikit-image / PyWavelets are not installed (no network), so `pipeline/metrics.py` and
`pipeline/enhancement.py` cannot be imported as they are.  This script installs a stand-in
`skimage` package whose leaf functions (`filters.laplace/sobel_h/sobel_v/unsharp_mask`,
`exposure.equalize_adapthist/adjust_gamma`, `restoration.estimate_sigma/denoise_wavelet/
denoise_tv_chambolle`, `metrics.structural_similarity/peak_signal_noise_ratio`) forward to the
oracle's restatements of those leaves, and then imports and RUNS the reference's real modules from
/root/reference: its control flow, numpy glue, percentile / histogram / entropy code, safeguards,
clamping, label strings, validation and scoring arithmetic are the reference's own bytes.

What the vectors pin: `oracle/ref_metrics.py` and `oracle/ref_enhancement.py` (the restated
control flow) against the reference itself.  What they do not pin: the leaves, which are the same
oracle code on both sides (parity of those stays unpinned).

Run in the build container only (needs /root/reference):
    python tests/golden/make_reference_glue.py
"""

from __future__ import annotations

import json
import sys
import types
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from mdimg_b200 import synth  # noqa: E402
from oracle import exposure as oex  # noqa: E402
from oracle import filters as oflt  # noqa: E402
from oracle import fullref as ofr  # noqa: E402
from oracle import restoration as ores  # noqa: E402


def install_skimage_stand_in() -> None:
    """No-op when the real scikit-image is importable: the vectors then pin the leaves as well."""
    try:
        import skimage  # noqa: F401
        import skimage.restoration  # noqa: F401  (needs PyWavelets)
        print("using the installed scikit-image", skimage.__version__, "- the vectors pin the leaves too")
        return
    except Exception:  # noqa: BLE001
        pass
    sk = types.ModuleType("skimage")
    filters = types.ModuleType("skimage.filters")
    filters.laplace = lambda image: oflt.laplace(image)
    filters.sobel_h = lambda image: oflt.sobel_h(image)
    filters.sobel_v = lambda image: oflt.sobel_v(image)
    filters.unsharp_mask = lambda image, radius=1.0, amount=1.0: oflt.unsharp_mask(image, radius, amount)
    exposure = types.ModuleType("skimage.exposure")
    exposure.equalize_adapthist = (
        lambda image, kernel_size=None, clip_limit=0.01, nbins=256:
        oex.equalize_adapthist(image, kernel_size=kernel_size, clip_limit=clip_limit))
    exposure.adjust_gamma = lambda image, gamma=1, gain=1: oex.adjust_gamma(image, gamma, gain)
    restoration = types.ModuleType("skimage.restoration")

    def estimate_sigma(image, average_sigmas=False, *, channel_axis=None):
        assert channel_axis is None
        return ores.estimate_sigma(image)

    def denoise_wavelet(image, sigma=None, wavelet="db1", mode="soft", wavelet_levels=None,
                        convert2ycbcr=False, method="BayesShrink", rescale_sigma=True, *, channel_axis=None):
        assert channel_axis is None and wavelet == "db1" and method == "BayesShrink" and rescale_sigma
        return ores.denoise_wavelet(image, sigma=sigma, mode=mode)

    def denoise_tv_chambolle(image, weight=0.1, eps=2.0e-4, max_num_iter=200, *, channel_axis=None):
        assert channel_axis is None
        return ores.denoise_tv_chambolle(image, weight, eps, max_num_iter)

    restoration.estimate_sigma = estimate_sigma
    restoration.denoise_wavelet = denoise_wavelet
    restoration.denoise_tv_chambolle = denoise_tv_chambolle
    metrics = types.ModuleType("skimage.metrics")
    metrics.structural_similarity = lambda a, b, data_range=None: ofr.structural_similarity(a, b, data_range=data_range)
    metrics.peak_signal_noise_ratio = lambda a, b, data_range=None: ofr.peak_signal_noise_ratio(a, b, data_range=data_range)
    sk.filters, sk.exposure, sk.restoration, sk.metrics = filters, exposure, restoration, metrics
    for m in (sk, filters, exposure, restoration, metrics):
        sys.modules[m.__name__] = m


def plans(schemas):
    """Candidate plans: P_full, P_cr, a hard-threshold plan with out-of-bounds parameters (clamping),
    a plan whose op order differs from the fixed step order, and one that only sharpens hard (halo)."""
    P, E = schemas.EnhancementParams, schemas.EnhancementPlan

    def mk(ops, **kw):
        return E(recommended_ops=ops, params=P(**kw), rationale="golden")

    return {
        "p_full": mk(["denoise", "clahe", "gamma", "unsharp", "post_denoise", "bilateral", "tv_denoise"],
                     clahe_clip_limit=0.015, clahe_tile_size=16, gamma=0.95, unsharp_radius=0.8,
                     unsharp_amount=0.5, denoise_mode="soft", post_denoise_strength=0.3, bilateral_d=5,
                     bilateral_sigma_color=0.05, bilateral_sigma_space=0.05, tv_denoise_weight=0.05),
        "p_cr": mk(["clahe", "unsharp"], clahe_clip_limit=0.03, clahe_tile_size=32, unsharp_radius=2.0,
                   unsharp_amount=1.5),
        "p_clamped_hard": mk(["denoise", "gamma", "unsharp"], denoise_mode="hard", gamma=1.5,
                             unsharp_radius=3.0, unsharp_amount=2.5, clahe_clip_limit=0.08,
                             clahe_tile_size=48, post_denoise_strength=0.0),
        "p_reordered": mk(["unsharp", "gamma", "clahe", "tv_denoise"], clahe_clip_limit=0.01,
                          clahe_tile_size=8, gamma=1.05, unsharp_radius=1.0, unsharp_amount=2.0,
                          tv_denoise_weight=0.02),
        "p_sharpen_only": mk(["unsharp"], unsharp_radius=1.5, unsharp_amount=2.5),
    }


def plan_to_json(plan) -> dict:
    return {"recommended_ops": list(plan.recommended_ops), "params": plan.params.model_dump()}


def main() -> None:
    assert REFERENCE.exists(), "needs the reference checkout"
    install_skimage_stand_in()
    sys.path.insert(0, str(REFERENCE))
    warnings.filterwarnings("ignore")
    import pipeline.enhancement as renh  # the reference's own modules
    import pipeline.metrics as rmet
    import pipeline.schemas as rsch

    ims = {"clean64": synth.fixture_clean(), "noisy64": synth.fixture_noisy(), "lowc64": synth.fixture_low_contrast()}
    ct = synth.ct_slice(1000, 0.25, size=96)
    x = ct.astype(np.float32)
    ims["ct96"] = (x - x.min()) / (x.max() - x.min())      # pipeline/dicom_io.py:84-91 (pydicom is absent too)
    out = {"metrics": {}, "issues": {}, "niqe": {}, "edge_ratio": {}, "from_issues": {}, "plans": {},
           "validation": {}, "score": {}}
    arrays = {}
    pl = plans(rsch)
    out["plan_defs"] = {k: plan_to_json(v) for k, v in pl.items()}
    for name, im in ims.items():
        m = rmet.compute_metrics(im)
        out["metrics"][name] = m
        out["issues"][name] = rmet.detect_issues(m)
        out["niqe"][name] = rmet.compute_niqe_approximation(im)
        out["edge_ratio"][name] = rmet.compute_edge_ratio(im)
        for issues in (out["issues"][name], ["noise", "blur"], ["low_contrast", "clipping_low"],
                       ["clipping_high", "blur"], []):
            key = f"{name}|{','.join(issues)}"
            enh, labels = renh.apply_enhancements(im, list(issues))
            arrays[f"issues|{key}"] = enh
            out["from_issues"][key] = labels
        for pname, plan in pl.items():
            key = f"{name}|{pname}"
            try:
                enh, labels = renh.apply_enhancements_from_params(im, plan)
            except ValueError as exc:      # data-dependent errors are part of the observable behaviour
                out["plans"][key] = {"error": f"ValueError: {exc}"}
                continue
            arrays[f"plan|{key}"] = enh
            out["plans"][key] = labels
            val = rmet.compute_validation(im, enh)
            score, breakdown = rmet.compute_objective_score(val)
            out["validation"][key] = val
            out["score"][key] = {"score": score, "breakdown": breakdown}
    (HERE / "reference_glue.json").write_text(json.dumps(out, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o)))
    np.savez_compressed(HERE / "reference_glue.npz", **arrays)
    print(f"wrote {len(arrays)} arrays, {len(out['validation'])} validations")


if __name__ == "__main__":
    main()
