"""Drop-in for the reference's ``pipeline/metrics.py``: same names, signatures, return types.

Every image-sized computation runs on the GPU (``mdimg_metrics`` / ``mdimg_quality`` /
``mdimg_fullref`` in libmdimg_b200.so); only the scalar bookkeeping the reference also does in
Python stays on the host.  Inputs are 2-D float32 arrays normalised to [0, 1] and are never
mutated; outputs are Python floats / bools.
"""

from __future__ import annotations

from typing import Dict

import numpy as np
import torch

from ..engine import (METRIC_KEYS, MC_EDGE_RATIO, MC_NIQE, THRESHOLDS, Engine, metrics_dict,  # noqa: F401
                      objective_score, validation_dict)
from ..stack import get_ops


def _to_stack(image: np.ndarray, ops) -> torch.Tensor:
    arr = np.asarray(image)
    if arr.ndim != 2:
        raise ValueError(f"expected a 2-D image, got shape {arr.shape}")
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    return torch.from_numpy(arr).to(ops.device)[None].contiguous()


def compute_metrics(image: np.ndarray) -> Dict[str, float]:
    """The 16 no-reference quality metrics (reference: pipeline/metrics.py:42-109)."""
    ops = get_ops()
    rows = ops.metrics(_to_stack(image, ops))
    return metrics_dict(rows[0].cpu().numpy())


def detect_issues(metrics: Dict[str, float]) -> list[str]:
    """Threshold tests of pipeline/metrics.py:166-179 (scalar, host)."""
    checks = (
        ("noise", metrics["sigma"] > THRESHOLDS["noise_sigma"]),
        ("blur", metrics["lap_var"] < THRESHOLDS["blur_lap_var"]),
        ("low_contrast", metrics["std"] < THRESHOLDS["low_contrast_std"]),
        ("clipping_low", metrics["pct_low"] > THRESHOLDS["clip_pct"]),
        ("clipping_high", metrics["pct_high"] > THRESHOLDS["clip_pct"]),
    )
    return [name for name, hit in checks if hit]


def compute_niqe_approximation(image: np.ndarray) -> float:
    """No-reference naturalness score, lower is better (pipeline/metrics.py:187-210)."""
    ops = get_ops()
    return float(ops.quality(_to_stack(image, ops), niqe=True)[0, 1].item())


def compute_edge_ratio(image: np.ndarray) -> float:
    """mean|laplace| / (mean|sobel| + 1e-8) (pipeline/metrics.py:213-217)."""
    ops = get_ops()
    return float(ops.quality(_to_stack(image, ops), niqe=False)[0, 0].item())


def compute_validation(original: np.ndarray, enhanced: np.ndarray) -> Dict[str, object]:
    """Full- and no-reference comparison of original vs enhanced (pipeline/metrics.py:225-329)."""
    ops = get_ops()
    a = _to_stack(original, ops)
    b = _to_stack(enhanced, ops)
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    eng = Engine(ops)
    rb, ra, fr = eng.validation_rows(a, b)
    rb, ra, fr = rb[0].cpu().numpy(), ra[0].cpu().numpy(), fr[0].cpu().numpy()
    with np.errstate(divide="ignore"):
        return validation_dict(metrics_dict(rb), metrics_dict(ra), float(fr[0]), float(fr[1]),
                               float(rb[MC_NIQE]), float(ra[MC_NIQE]), float(ra[MC_EDGE_RATIO]))


def compute_objective_score(validation: dict) -> tuple[float, dict]:
    """Scalar objective used by the tuning loop (pipeline/metrics.py:337-408)."""
    return objective_score(validation)
