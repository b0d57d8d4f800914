"""Restatement of ``skimage.metrics.structural_similarity`` / ``peak_signal_noise_ratio``.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED for the wrapper logic (scikit-image not
installed); the box filters are scipy's own ``uniform_filter``.  Follows
``skimage/metrics/_structural_similarity.py`` and ``skimage/metrics/simple_metrics.py``
with the arguments the reference passes (``pipeline/metrics.py:232-233``:
``data_range=1.0``, everything else default: win_size 7, uniform window, sample
covariance, K1 0.01, K2 0.03).
"""

from __future__ import annotations

import numpy as np
from scipy.ndimage import uniform_filter


def structural_similarity(im1: np.ndarray, im2: np.ndarray, data_range: float = 1.0,
                          win_size: int = 7, full: bool = False):
    if im1.shape != im2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(im1.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    ft = np.float32 if (im1.dtype == np.float32 and im2.dtype == np.float32) else np.float64
    im1 = im1.astype(ft, copy=False)
    im2 = im2.astype(ft, copy=False)
    ndim = im1.ndim
    npx = win_size**ndim
    cov_norm = npx / (npx - 1)

    ux = uniform_filter(im1, size=win_size)
    uy = uniform_filter(im2, size=win_size)
    uxx = uniform_filter(im1 * im1, size=win_size)
    uyy = uniform_filter(im2 * im2, size=win_size)
    uxy = uniform_filter(im1 * im2, size=win_size)
    vx = cov_norm * (uxx - ux * ux)
    vy = cov_norm * (uyy - uy * uy)
    vxy = cov_norm * (uxy - ux * uy)

    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    a1, a2, b1, b2 = (2 * ux * uy + c1, 2 * vxy + c2, ux**2 + uy**2 + c1, vx + vy + c2)
    s = (a1 * a2) / (b1 * b2)
    pad = (win_size - 1) // 2
    mssim = s[pad:-pad, pad:-pad].mean(dtype=np.float64)
    if full:
        return mssim, s
    return mssim


def peak_signal_noise_ratio(image_true: np.ndarray, image_test: np.ndarray,
                            data_range: float = 1.0):
    ft = np.float32 if (image_true.dtype == np.float32 and image_test.dtype == np.float32) \
        else np.float64
    a = np.asarray(image_true, dtype=ft)
    b = np.asarray(image_test, dtype=ft)
    err = np.mean((a - b) ** 2, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return 10 * np.log10((data_range**2) / err)
