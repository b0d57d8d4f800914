"""CPU restatement of the reference's ``pipeline/enhancement.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``; PARITY UNPINNED for the
skimage/pywt leaves).  The step order, gating, clamping, safeguards and op-label strings
are the reference's (cited per function); the implementation is table-driven rather than
the reference's inline chain.
"""

from __future__ import annotations

import logging
from typing import Callable, Dict, List, Tuple

import numpy as np

from . import ref_metrics as rm
from .exposure import adjust_gamma, equalize_adapthist
from .filters import unsharp_mask
from .restoration import denoise_tv_chambolle, denoise_wavelet, estimate_sigma

logger = logging.getLogger(__name__)

# pipeline/enhancement.py:32-42
ENHANCEMENT_PARAMS = {
    "clahe_clip_limit": 0.015,
    "clahe_tile_size": 16,
    "gamma_brighten": 0.95,
    "gamma_darken": 1.05,
    "unsharp_radius": 0.8,
    "unsharp_amount": 0.5,
    "denoise_sigma": None,
    "denoise_wavelet_mode": "soft",
    "post_denoise_strength": 0.3,
}

# pipeline/schemas.py:16-28
PARAM_BOUNDS = {
    "clahe_clip_limit": (0.002, 0.08),
    "clahe_tile_size": (4, 48),
    "gamma": (0.6, 1.5),
    "unsharp_radius": (0.2, 3.0),
    "unsharp_amount": (0.03, 2.5),
    "post_denoise_strength": (0.0, 0.8),
    "bilateral_d": (0, 13),
    "bilateral_sigma_color": (0.005, 0.20),
    "bilateral_sigma_space": (0.005, 0.20),
    "tv_denoise_weight": (0.0, 0.15),
}


def halo_detected(enhanced: np.ndarray, max_edge_ratio: float = 1.5) -> bool:
    """pipeline/enhancement.py:50-52."""
    return rm.compute_edge_ratio(enhanced) > max_edge_ratio


def noise_amplified(original: np.ndarray, enhanced: np.ndarray, max_ratio: float = 1.3) -> bool:
    """pipeline/enhancement.py:55-63."""
    s0 = float(estimate_sigma(original))
    s1 = float(estimate_sigma(enhanced))
    if s0 < 1e-8:
        return False
    return s1 > s0 * max_ratio


def over_processed(original: np.ndarray, enhanced: np.ndarray, max_drop: float = 0.5) -> bool:
    """pipeline/enhancement.py:66-72."""
    return (rm.compute_niqe_approximation(enhanced)
            - rm.compute_niqe_approximation(original)) > max_drop


def light_denoise(image: np.ndarray, strength: float = 0.3) -> np.ndarray:
    """pipeline/enhancement.py:80-94."""
    s = float(estimate_sigma(image))
    if s < 0.001:
        return image
    den = denoise_wavelet(image, mode="soft", sigma=s * 0.5)
    return ((1 - strength) * image + strength * den).astype(np.float32)


def bilateral_filter(image: np.ndarray, d: int = 5, sigma_color: float = 0.05,
                     sigma_space: float = 0.05) -> np.ndarray:
    """pipeline/enhancement.py:102-143 (the only leaf routine that lives in the reference)."""
    if d <= 0:
        return image
    d = min(d, 9)
    if d % 2 == 0:
        d += 1
    r = d // 2
    h, w = image.shape
    padded = np.pad(image, r, mode="reflect")
    num = np.zeros_like(image)
    den = np.zeros_like(image)
    yy, xx = np.mgrid[-r : r + 1, -r : r + 1]
    w_space = np.exp(-(xx**2 + yy**2) / (2 * sigma_space**2 * d**2))
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            nb = padded[r + dy : r + dy + h, r + dx : r + dx + w]
            diff = image - nb
            wgt = w_space[dy + r, dx + r] * np.exp(-(diff**2) / (2 * sigma_color**2))
            num += wgt * nb
            den += wgt
    return (num / (den + 1e-10)).astype(np.float32)


def apply_enhancements(image: np.ndarray, issues: List[str]) -> Tuple[np.ndarray, List[str]]:
    """pipeline/enhancement.py:151-227."""
    P = ENHANCEMENT_PARAMS
    cur = image.copy()
    labels: List[str] = []
    has = set(issues).__contains__

    if has("noise"):
        cur = denoise_wavelet(cur, mode=P["denoise_wavelet_mode"])
        labels.append("Wavelet denoise (pre)")
    if has("low_contrast") or has("clipping_low") or has("clipping_high"):
        k = P["clahe_tile_size"]
        cur = equalize_adapthist(cur, clip_limit=P["clahe_clip_limit"], kernel_size=k)
        labels.append(f"CLAHE (clip={P['clahe_clip_limit']}, tile={k})")
    if has("clipping_low") and not has("clipping_high"):
        cur = adjust_gamma(cur, gamma=P["gamma_brighten"])
        labels.append(f"Gamma brighten ({P['gamma_brighten']})")
    elif has("clipping_high") and not has("clipping_low"):
        cur = adjust_gamma(cur, gamma=P["gamma_darken"])
        labels.append(f"Gamma darken ({P['gamma_darken']})")
    if has("blur"):
        cur = unsharp_mask(cur, radius=P["unsharp_radius"], amount=P["unsharp_amount"])
        labels.append(f"Unsharp mask (r={P['unsharp_radius']}, a={P['unsharp_amount']})")
    if has("blur") and P["post_denoise_strength"] > 0:
        cur = light_denoise(cur, strength=P["post_denoise_strength"])
        labels.append(f"Light denoise (post, s={P['post_denoise_strength']})")
    cur = np.clip(cur, 0.0, 1.0)

    if noise_amplified(image, cur):
        logger.warning("Noise amplification detected — applying corrective denoise.")
        cur = np.clip(light_denoise(cur, strength=0.4), 0.0, 1.0)
        labels.append("Auto-corrective denoise (noise guard)")
    return cur.astype(np.float32), labels


def clamp_plan_params(p) -> Dict[str, object]:
    """pipeline/enhancement.py:249-263 — clamp the ten numeric parameters to PARAM_BOUNDS."""
    def c(name: str):
        lo, hi = PARAM_BOUNDS[name]
        return max(lo, min(hi, getattr(p, name)))

    return {
        "clip_limit": c("clahe_clip_limit"),
        "tile_size": int(c("clahe_tile_size")),
        "gamma": c("gamma"),
        "u_radius": c("unsharp_radius"),
        "u_amount": c("unsharp_amount"),
        "dn_mode": p.denoise_mode if p.denoise_mode in ("soft", "hard") else "soft",
        "post_str": c("post_denoise_strength"),
        "bilateral_d": int(c("bilateral_d")),
        "bilateral_sc": c("bilateral_sigma_color"),
        "bilateral_ss": c("bilateral_sigma_space"),
        "tv_weight": c("tv_denoise_weight"),
    }


_STEP_ORDER = ("denoise", "clahe", "gamma", "unsharp", "post_denoise", "bilateral", "tv_denoise")


def _step_table(q: Dict[str, object], u_amount: float):
    """name -> (enabled, transform, label) for the seven steps (enhancement.py:268-312)."""
    steps: Dict[str, Tuple[bool, Callable[[np.ndarray], np.ndarray], str]] = {
        "denoise": (True, lambda x: denoise_wavelet(x, mode=q["dn_mode"]),
                    f"Wavelet denoise (pre, mode={q['dn_mode']})"),
        "clahe": (True, lambda x: equalize_adapthist(x, clip_limit=q["clip_limit"],
                                                     kernel_size=q["tile_size"]),
                  f"CLAHE (clip={q['clip_limit']:.4f}, tile={q['tile_size']})"),
        "gamma": (abs(q["gamma"] - 1.0) > 1e-4, lambda x: adjust_gamma(x, gamma=q["gamma"]),
                  f"Gamma {'brighten' if q['gamma'] < 1.0 else 'darken'} ({q['gamma']:.3f})"),
        "unsharp": (True, lambda x: unsharp_mask(x, radius=q["u_radius"], amount=u_amount),
                    f"Unsharp mask (r={q['u_radius']:.2f}, a={u_amount:.2f})"),
        "post_denoise": (q["post_str"] > 0, lambda x: light_denoise(x, strength=q["post_str"]),
                         f"Light denoise (post, s={q['post_str']:.2f})"),
        "bilateral": (q["bilateral_d"] > 0,
                      lambda x: bilateral_filter(x, d=q["bilateral_d"],
                                                 sigma_color=q["bilateral_sc"],
                                                 sigma_space=q["bilateral_ss"]),
                      f"Bilateral (d={q['bilateral_d']}, sc={q['bilateral_sc']:.3f}, "
                      f"ss={q['bilateral_ss']:.3f})"),
        "tv_denoise": (q["tv_weight"] > 0,
                       lambda x: denoise_tv_chambolle(x, weight=q["tv_weight"]),
                       f"TV denoise (w={q['tv_weight']:.4f})"),
    }
    return steps


def apply_enhancements_from_params(image: np.ndarray, plan) -> Tuple[np.ndarray, List[str]]:
    """pipeline/enhancement.py:235-369."""
    q = clamp_plan_params(plan.params)
    ops = [op.lower().strip() for op in plan.recommended_ops]
    cur = image.copy()
    labels: List[str] = []

    steps = _step_table(q, q["u_amount"])
    for name in _STEP_ORDER:  # fixed order, gated by membership (enhancement.py:268-312)
        enabled, fn, label = steps[name]
        if name in ops and enabled:
            cur = fn(cur)
            labels.append(label)
    cur = np.clip(cur, 0.0, 1.0)

    if "unsharp" in ops and halo_detected(cur):  # enhancement.py:319-353
        logger.warning("Halo detected (edge_ratio > 1.5) — re-applying with halved unsharp_amount.")
        reduced = q["u_amount"] * 0.5
        redo = _step_table(q, reduced)
        cur = image.copy()
        for op in ops:  # the plan's own order, duplicates included
            if op in redo and redo[op][0]:
                cur = redo[op][1](cur)
        cur = np.clip(cur, 0.0, 1.0)
        labels.append(f"[safeguard] Unsharp reduced to {reduced:.2f}")

    if noise_amplified(image, cur):  # enhancement.py:356-360
        logger.warning("Noise amplification detected — applying corrective denoise.")
        cur = np.clip(light_denoise(cur, strength=0.4), 0.0, 1.0)
        labels.append("Auto-corrective denoise (noise guard)")

    if over_processed(image, cur, 0.5):  # enhancement.py:363-367
        logger.warning("Over-processing detected (NIQE degraded >0.5). Blending back.")
        cur = np.clip(0.6 * cur + 0.4 * image, 0.0, 1.0)
        labels.append("Blend-back 40% original (over-processing guard)")
    return cur.astype(np.float32), labels
