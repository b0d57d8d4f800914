"""Seeded synthetic inputs for tests and benchmarks.

Generated on the host with numpy's ``default_rng`` so the CPU oracle and the GPU path see the same
bytes.  Nothing here is part of the product path.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy import ndimage as ndi


def ct_slice(seed: int = 1000, z: float = 0.0, size: int = 512) -> np.ndarray:
    """C1/C2: uint16 CT-like slice: air outside a centred ellipse, soft tissue inside, three
    discs, Gaussian blur sigma 1.5, N(0, 25) noise, 12-bit clip.  `z` in [0, 1) varies the
    anatomy smoothly along a stack."""
    rng = np.random.default_rng(seed)
    s = size / 512.0
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    cy = cx = (size - 1) / 2.0
    a = (200.0 - 40.0 * abs(2 * z - 1)) * s
    b = (160.0 - 30.0 * abs(2 * z - 1)) * s
    img = np.zeros((size, size), np.float32)
    inside = ((xx - cx) / a) ** 2 + ((yy - cy) / b) ** 2 <= 1.0
    img[inside] = 1024.0
    ang = 2 * np.pi * z
    for k, val in enumerate((1324.0, 824.0, 1624.0)):
        th = ang + k * 2 * np.pi / 3
        dx, dy = 90.0 * s * np.cos(th), 70.0 * s * np.sin(th)
        disc = (xx - cx - dx) ** 2 + (yy - cy - dy) ** 2 <= (30.0 * s) ** 2
        img[disc] = val
    img = ndi.gaussian_filter(img, 1.5)
    img = img + rng.normal(0.0, 25.0, img.shape).astype(np.float32)
    return np.clip(img, 0, 4095).astype(np.uint16)


def ct_stack(n: int, seed0: int = 1000, size: int = 512) -> np.ndarray:
    """C2/C4: [n, size, size] uint16, seeds seed0 + z."""
    out = np.empty((n, size, size), np.uint16)
    for z in range(n):
        out[z] = ct_slice(seed0 + z, z / max(n, 1), size)
    return out


def radiograph(seed: int = 2000, size: int = 3000) -> np.ndarray:
    """C3: uint16 CR/DX-like image: sum of 6 random 2-D cosines (range 300-3500) times a soft
    collimator mask, N(0, 40) noise, 12-bit clip."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32)
    yy /= size
    xx /= size
    field = np.zeros((size, size), np.float32)
    for _ in range(6):
        fx, fy = rng.uniform(0.5, 3.0, 2)
        ph = rng.uniform(0, 2 * np.pi)
        amp = rng.uniform(0.5, 1.0)
        field += amp * np.cos(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
    field = (field - field.min()) / (field.max() - field.min())
    field = 300.0 + 3200.0 * field
    edge = 0.04
    mx = np.clip(np.minimum(xx, 1 - xx) / edge, 0, 1)
    my = np.clip(np.minimum(yy, 1 - yy) / edge, 0, 1)
    img = field * (mx * my) + rng.normal(0.0, 40.0, field.shape).astype(np.float32)
    return np.clip(img, 0, 4095).astype(np.uint16)


def unit_image(seed: int, size: int) -> np.ndarray:
    """C5: float32 [0,1] image: uniform noise blended 50/50 with a smooth cosine field."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:size, 0:size].astype(np.float32) / np.float32(size)
    field = np.zeros((size, size), np.float32)
    for _ in range(6):
        fx, fy = rng.uniform(0.5, 3.0, 2)
        ph = rng.uniform(0, 2 * np.pi)
        field += np.cos(2 * np.pi * (fx * xx + fy * yy) + ph).astype(np.float32)
    field = (field - field.min()) / (field.max() - field.min())
    noise = rng.random((size, size), dtype=np.float32)
    return (0.5 * noise + 0.5 * field).astype(np.float32)


# The three 64x64 fixtures of the reference's tests/conftest.py:9-32 (same seeds and formulas).
def fixture_clean() -> np.ndarray:
    rng = np.random.default_rng(42)
    g = np.linspace(0.1, 0.9, 64 * 64).reshape(64, 64).astype(np.float32)
    g += rng.normal(0, 0.005, g.shape).astype(np.float32)
    return np.clip(g, 0.0, 1.0)


def fixture_noisy() -> np.ndarray:
    rng = np.random.default_rng(99)
    base = np.full((64, 64), 0.5, dtype=np.float32)
    return np.clip(base + rng.normal(0, 0.15, base.shape).astype(np.float32), 0.0, 1.0)


def fixture_low_contrast() -> np.ndarray:
    return np.full((64, 64), 0.5, dtype=np.float32) + np.float32(0.01) * \
        np.random.default_rng(7).standard_normal((64, 64)).astype(np.float32)


def plan_full():
    """P_full of SURVEY.md §8d: all seven steps."""
    from .pipeline.schemas import EnhancementParams, EnhancementPlan
    return EnhancementPlan(
        recommended_ops=["denoise", "clahe", "gamma", "unsharp", "post_denoise", "bilateral", "tv_denoise"],
        params=EnhancementParams(clahe_clip_limit=0.015, clahe_tile_size=16, gamma=0.95, unsharp_radius=0.8,
                                 unsharp_amount=0.5, denoise_mode="soft", post_denoise_strength=0.3,
                                 bilateral_d=5, bilateral_sigma_color=0.05, bilateral_sigma_space=0.05,
                                 tv_denoise_weight=0.05),
    )


def plan_cr():
    """P_cr of SURVEY.md §8d: CLAHE + unsharp heavy."""
    from .pipeline.schemas import EnhancementParams, EnhancementPlan
    return EnhancementPlan(
        recommended_ops=["clahe", "unsharp"],
        params=EnhancementParams(clahe_clip_limit=0.03, clahe_tile_size=32, unsharp_radius=2.0,
                                 unsharp_amount=1.5),
    )


def shape_of(arr: np.ndarray) -> Tuple[int, int]:
    return int(arr.shape[-2]), int(arr.shape[-1])
