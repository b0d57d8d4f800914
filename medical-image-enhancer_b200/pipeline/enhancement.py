"""Drop-in for the reference's ``pipeline/enhancement.py``: same names, signatures, op labels.

``apply_enhancements`` (issue-gated defaults) and ``apply_enhancements_from_params`` (plan-driven,
seven steps + three safeguards) run on the GPU through libmdimg_b200.so.  Inputs are 2-D float32
arrays in [0, 1], never mutated; the result is a new float32 array clipped to [0, 1] plus the list
of applied-operation labels.
"""

from __future__ import annotations

from typing import List, Tuple, TYPE_CHECKING

import numpy as np
import torch

from ..engine import ENHANCEMENT_PARAMS, Engine  # noqa: F401
from ..stack import get_ops
from .metrics import compute_edge_ratio, compute_niqe_approximation, _to_stack

if TYPE_CHECKING:
    from .schemas import EnhancementPlan


def _estimate_sigma(image: np.ndarray) -> float:
    ops = get_ops()
    return float(ops.estimate_sigma(_to_stack(image, ops))[0].item())


def _check_halo(enhanced: np.ndarray, max_edge_ratio: float = 1.5) -> bool:
    """True if halo artefacts are detected (pipeline/enhancement.py:50-52)."""
    return compute_edge_ratio(enhanced) > max_edge_ratio


def _check_noise_amplification(original: np.ndarray, enhanced: np.ndarray, max_ratio: float = 1.3) -> bool:
    """pipeline/enhancement.py:55-63."""
    before, after = _estimate_sigma(original), _estimate_sigma(enhanced)
    if before < 1e-8:
        return False
    return after > before * max_ratio


def _check_over_processing(original: np.ndarray, enhanced: np.ndarray, max_niqe_degradation: float = 0.5) -> bool:
    """pipeline/enhancement.py:66-72."""
    return (compute_niqe_approximation(enhanced) - compute_niqe_approximation(original)) > max_niqe_degradation


def _light_denoise(image: np.ndarray, strength: float = 0.3) -> np.ndarray:
    """pipeline/enhancement.py:80-94 — returns the input object itself when sigma < 0.001."""
    ops = get_ops()
    src = _to_stack(image, ops)
    dst = torch.empty_like(src)
    skipped = ops.light_denoise(src, dst, strength)
    if bool(skipped[0].item()):
        return image
    return dst[0].cpu().numpy()


def _bilateral_filter(image: np.ndarray, d: int = 5, sigma_color: float = 0.05,
                      sigma_space: float = 0.05) -> np.ndarray:
    """pipeline/enhancement.py:102-143."""
    if d <= 0:
        return image
    ops = get_ops()
    src = _to_stack(image, ops)
    dst = torch.empty_like(src)
    ops.bilateral(src, dst, d, sigma_color, sigma_space)
    return dst[0].cpu().numpy()


def apply_enhancements(image: np.ndarray, issues: List[str]) -> Tuple[np.ndarray, List[str]]:
    """Issue-gated conservative enhancement (pipeline/enhancement.py:151-227)."""
    ops = get_ops()
    res = Engine(ops).enhance_issues(_to_stack(image, ops), list(issues))
    return res.image[0].cpu().numpy(), res.labels[0]


def apply_enhancements_from_params(image: np.ndarray, plan: "EnhancementPlan") -> Tuple[np.ndarray, List[str]]:
    """Plan-driven enhancement with PARAM_BOUNDS clamping and safeguards (pipeline/enhancement.py:235-369)."""
    ops = get_ops()
    res = Engine(ops).enhance_plan(_to_stack(image, ops), plan)
    return res.image[0].cpu().numpy(), res.labels[0]
