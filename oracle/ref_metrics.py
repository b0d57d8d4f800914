"""CPU restatement of the reference's ``pipeline/metrics.py`` and ``normalize_image``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``; PARITY UNPINNED for the
skimage/pywt leaves).  Every function cites the reference lines it follows; the control
flow and the order of floating-point operations are the reference's, the code is not.
"""

from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
from scipy.ndimage import uniform_filter

from . import filters as flt
from .fullref import peak_signal_noise_ratio, structural_similarity
from .restoration import estimate_sigma

# pipeline/metrics.py:25-34
THRESHOLDS = {
    "noise_sigma": 0.08,
    "blur_lap_var": 0.001,
    "low_contrast_std": 0.12,
    "clip_pct": 0.01,
    "ssim": 0.70,
    "psnr": 22.0,
    "quality_improvement": 0.10,
}

#: key order of the dict returned by ``compute_metrics`` (pipeline/metrics.py:90-109)
METRIC_KEYS = (
    "sigma", "lap_var", "std", "pct_low", "pct_high", "entropy", "edge_density",
    "gradient_mag_mean", "gradient_mag_std", "snr_proxy", "cnr_proxy", "laplacian_energy",
    "histogram_spread", "local_contrast_std", "gradient_strength", "gradient_entropy",
)


def normalize_image(image: np.ndarray) -> np.ndarray:
    """pipeline/dicom_io.py:84-91."""
    img = image.astype(np.float32)
    lo, hi = float(np.min(img)), float(np.max(img))
    if hi - lo < 1e-8:
        return np.zeros_like(img, dtype=np.float32)
    return (img - lo) / (hi - lo)


def _entropy_from_counts(counts: np.ndarray, empty_is_zero: bool = False) -> float:
    counts = counts[counts > 0]
    if empty_is_zero and counts.size == 0:
        return 0.0
    p = counts / counts.sum()
    return float(-np.sum(p * np.log2(p)))


def shannon_entropy(image: np.ndarray, bins: int = 256) -> float:
    """pipeline/metrics.py:112-117."""
    counts, _ = np.histogram(image.ravel(), bins=bins, range=(0.0, 1.0))
    return _entropy_from_counts(counts)


def local_contrast_std(image: np.ndarray, patch: int = 7) -> float:
    """pipeline/metrics.py:120-129."""
    m = uniform_filter(image, size=patch)
    q = uniform_filter(image**2, size=patch)
    return float(np.std(np.sqrt(np.maximum(q - m**2, 0))))


def gradient_strength(grad: np.ndarray) -> float:
    """pipeline/metrics.py:132-138."""
    t = float(np.percentile(grad, 90))
    strong = grad[grad >= t]
    return float(np.mean(strong)) if strong.size else 0.0


def gradient_entropy(grad: np.ndarray, bins: int = 128) -> float:
    """pipeline/metrics.py:141-151."""
    counts, _ = np.histogram(grad.ravel(), bins=bins, range=(0.0, float(grad.max()) + 1e-8))
    return _entropy_from_counts(counts, empty_is_zero=True)


def edge_density(image: np.ndarray, frac: float = 0.1) -> float:
    """pipeline/metrics.py:154-158."""
    g = flt.gradient_magnitude(image)
    t = frac * g.max() if g.max() > 0 else 0
    return float(np.mean(g > t))


def compute_metrics(image: np.ndarray) -> Dict[str, float]:
    """pipeline/metrics.py:42-109 — the 16 no-reference metrics."""
    out: Dict[str, float] = {}
    sigma = float(estimate_sigma(image))
    lap = flt.laplace(image)
    grad = flt.gradient_magnitude(image)
    p05, p95 = float(np.percentile(image, 5)), float(np.percentile(image, 95))
    q25, q75 = float(np.percentile(image, 25)), float(np.percentile(image, 75))
    out["sigma"] = sigma
    out["lap_var"] = float(np.var(lap))
    out["std"] = float(np.std(image))
    out["pct_low"] = float(np.mean(image <= 0.01))
    out["pct_high"] = float(np.mean(image >= 0.99))
    out["entropy"] = shannon_entropy(image)
    out["edge_density"] = edge_density(image)
    out["gradient_mag_mean"] = float(np.mean(grad))
    out["gradient_mag_std"] = float(np.std(grad))
    out["snr_proxy"] = float(np.mean(image) / max(sigma, 1e-8))
    out["cnr_proxy"] = float((p95 - p05) / max(sigma, 1e-8))
    out["laplacian_energy"] = float(np.mean(lap**2))
    out["histogram_spread"] = q75 - q25
    out["local_contrast_std"] = local_contrast_std(image)
    out["gradient_strength"] = gradient_strength(grad)
    out["gradient_entropy"] = gradient_entropy(grad)
    return {k: out[k] for k in METRIC_KEYS}


def detect_issues(metrics: Dict[str, float]) -> List[str]:
    """pipeline/metrics.py:166-179."""
    tests = (
        ("noise", metrics["sigma"] > THRESHOLDS["noise_sigma"]),
        ("blur", metrics["lap_var"] < THRESHOLDS["blur_lap_var"]),
        ("low_contrast", metrics["std"] < THRESHOLDS["low_contrast_std"]),
        ("clipping_low", metrics["pct_low"] > THRESHOLDS["clip_pct"]),
        ("clipping_high", metrics["pct_high"] > THRESHOLDS["clip_pct"]),
    )
    return [name for name, hit in tests if hit]


def compute_edge_ratio(image: np.ndarray) -> float:
    """pipeline/metrics.py:213-217."""
    lap_abs = np.abs(flt.laplace(image))
    grad = flt.gradient_magnitude(image)
    return float(np.mean(lap_abs) / (np.mean(grad) + 1e-8))


def compute_niqe_approximation(image: np.ndarray) -> float:
    """pipeline/metrics.py:187-210."""
    m = uniform_filter(image, size=16)
    q = uniform_filter(image**2, size=16)
    local_var = np.maximum(q - m**2, 0)
    var_of_var = float(np.std(local_var) / (np.mean(local_var) + 1e-8))
    ratio = compute_edge_ratio(image)
    return float(var_of_var + max(0, ratio - 1.0) * 10)


_PAIRED = (  # (validation key stem, metric key) -> *_before / *_after / *_change
    ("entropy", "entropy"), ("snr", "snr_proxy"), ("cnr", "cnr_proxy"),
)
_PAIRED_EXTRA = (
    ("local_contrast", "local_contrast_std"), ("gradient_strength", "gradient_strength"),
    ("gradient_entropy", "gradient_entropy"),
)


def compute_validation(original: np.ndarray, enhanced: np.ndarray) -> Dict[str, object]:
    """pipeline/metrics.py:225-329."""
    mb = compute_metrics(original)
    ma = compute_metrics(enhanced)
    ssim = float(structural_similarity(original, enhanced, data_range=1.0))
    psnr = float(peak_signal_noise_ratio(original, enhanced, data_range=1.0))
    niqe_b = compute_niqe_approximation(original)
    niqe_a = compute_niqe_approximation(enhanced)
    niqe_ok = niqe_a <= niqe_b

    eps = 1e-8
    contrast_gain = (ma["std"] - mb["std"]) / max(mb["std"], eps)
    sharpness_gain = (ma["lap_var"] - mb["lap_var"]) / max(mb["lap_var"], eps)
    noise_reduction = (mb["sigma"] - ma["sigma"]) / max(mb["sigma"], eps)
    edge_ratio_after = compute_edge_ratio(enhanced)
    qi = float(0.35 * contrast_gain + 0.35 * sharpness_gain + 0.30 * noise_reduction)

    ok_ssim = ssim >= THRESHOLDS["ssim"]
    ok_psnr = psnr >= THRESHOLDS["psnr"]
    ok_gain = qi >= THRESHOLDS["quality_improvement"]
    passes = (ok_ssim and ok_psnr) or (ok_ssim and ok_gain) or (ok_psnr and ok_gain and niqe_ok)

    res: Dict[str, object] = {
        "ssim": ssim, "psnr": psnr, "quality_improvement": qi,
        "meets_ssim": ok_ssim, "meets_psnr": ok_psnr, "meets_improvement": ok_gain,
        "passes": passes,
        "niqe_before": niqe_b, "niqe_after": niqe_a, "niqe_improved": niqe_ok,
        "contrast_gain": contrast_gain, "sharpness_gain": sharpness_gain,
        "noise_change": -noise_reduction,
    }
    for stem, key in _PAIRED:
        res[f"{stem}_before"] = mb[key]
        res[f"{stem}_after"] = ma[key]
        res[f"{stem}_change"] = ma[key] - mb[key]
    res["edge_density_change"] = ma["edge_density"] - mb["edge_density"]
    res["histogram_spread_change"] = ma["histogram_spread"] - mb["histogram_spread"]
    res["laplacian_energy_before"] = mb["laplacian_energy"]
    res["laplacian_energy_after"] = ma["laplacian_energy"]
    res["edge_ratio"] = edge_ratio_after
    for stem, key in _PAIRED_EXTRA:
        res[f"{stem}_before"] = mb[key]
        res[f"{stem}_after"] = ma[key]
        res[f"{stem}_change"] = ma[key] - mb[key]
    res["metrics_before"] = mb
    res["metrics_after"] = ma
    return res


def compute_objective_score(validation: dict) -> Tuple[float, dict]:
    """pipeline/metrics.py:337-408."""
    def f(key: str) -> float:
        return float(validation.get(key, 0))

    def capped(x: float, cap: float) -> float:
        return max(0.0, min(x, cap))

    passes = bool(validation.get("passes", False))
    parts = {
        "contrast_gain": f("contrast_gain"),
        "sharpness_gain": f("sharpness_gain"),
        "noise_penalty": max(0.0, f("noise_change")),
        "niqe_degradation": max(0.0, f("niqe_after") - f("niqe_before")),
        "halo_penalty": max(0.0, f("edge_ratio") - 1.0) * 5.0,
        "entropy_penalty": max(0.0, abs(f("entropy_change")) - 0.5) * 2.0,
        "snr_reward": capped(f("snr_change") * 0.1, 0.5),
        "hs_reward": capped(f("histogram_spread_change") * 0.5, 0.3),
        "local_contrast_reward": capped(f("local_contrast_change") * 0.3, 0.3),
        "gradient_strength_reward": capped(f("gradient_strength_change") * 0.2, 0.2),
        "gradient_entropy_penalty": max(0.0, abs(f("gradient_entropy_change")) - 0.3) * 1.5,
    }
    score = (
        0.35 * parts["contrast_gain"]
        + 0.35 * parts["sharpness_gain"]
        - 0.30 * parts["noise_penalty"]
        - 5.0 * parts["niqe_degradation"]
        - 10.0 * (0 if passes else 1)
        - parts["halo_penalty"]
        - parts["entropy_penalty"]
        + parts["snr_reward"]
        + parts["hs_reward"]
        + parts["local_contrast_reward"]
        + parts["gradient_strength_reward"]
        - parts["gradient_entropy_penalty"]
    )
    breakdown = {k: round(v, 4) for k, v in parts.items()}
    breakdown["passes"] = passes
    return round(float(score), 4), breakdown


def ingest_frames(raw: np.ndarray, slope=None, intercept=None, monochrome1: bool = False) -> np.ndarray:
    """Pixel path of ``load_dicom`` (pipeline/dicom_io.py:44-49) followed by ``normalize_image`` on
    every frame: pydicom's ``apply_modality_lut`` rescale (float64 multiply, float64 add — pydicom is
    an un-vendored dependency, restated from pydicom/pixel_data_handlers/util.py; PARITY UNPINNED),
    ``.astype(float32)``, ``image.max() - image`` over the whole array for MONOCHROME1."""
    arr = raw
    if slope is not None and intercept is not None:
        arr = arr.astype(np.float64) * float(slope)
        arr += float(intercept)
    img = arr.astype(np.float32)
    if monochrome1:
        img = img.max() - img
    if img.ndim == 2:
        return normalize_image(img)
    return np.stack([normalize_image(f) for f in img])
