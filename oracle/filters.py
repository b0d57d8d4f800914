"""Restatement of ``skimage.filters`` routines used by the reference hot path.

TEST INFRASTRUCTURE ONLY.  The arithmetic below is scipy.ndimage's own (called
directly); only the kernels and the wrapper logic are restated from
``skimage/filters/edges.py`` (``laplace``, ``sobel_h``, ``sobel_v``),
``skimage/filters/_gaussian.py`` and ``skimage/filters/_unsharp_mask.py`` — for those
wrappers the status is PARITY UNPINNED (scikit-image not installed).

Reference call sites: ``pipeline/metrics.py:48,62,156,203-204,215-216``;
``pipeline/enhancement.py:202,290,338``.
"""

from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi

_LAPLACE = np.array([[0, -1, 0], [-1, 4, -1], [0, -1, 0]], dtype=np.float64)
# sobel_h: smooth [1,2,1]/4 along axis 1, derivative [1,0,-1] along axis 0.
_SOBEL_H = np.array([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=np.float64) / 4.0
_SOBEL_V = _SOBEL_H.T.copy()


def _as_float(image: np.ndarray) -> np.ndarray:
    image = np.asarray(image)
    if image.dtype in (np.float32, np.float64):
        return image
    if image.dtype == np.float16:
        return image.astype(np.float32)
    return image.astype(np.float64)


def laplace(image: np.ndarray) -> np.ndarray:
    """``filters.laplace(image)`` (ksize=3, no mask): ndi.convolve with the 5-point kernel."""
    image = _as_float(image)
    return ndi.convolve(image, _LAPLACE.astype(image.dtype), mode="reflect")


def sobel_h(image: np.ndarray) -> np.ndarray:
    image = _as_float(image)
    return ndi.convolve(image, _SOBEL_H.astype(image.dtype), mode="reflect")


def sobel_v(image: np.ndarray) -> np.ndarray:
    image = _as_float(image)
    return ndi.convolve(image, _SOBEL_V.astype(image.dtype), mode="reflect")


def gradient_magnitude(image: np.ndarray) -> np.ndarray:
    """``np.sqrt(sobel_h(image)**2 + sobel_v(image)**2)`` as written at ``metrics.py:62``."""
    return np.sqrt(sobel_h(image) ** 2 + sobel_v(image) ** 2)


def gaussian(image: np.ndarray, sigma: float) -> np.ndarray:
    """``filters.gaussian(image, sigma=sigma, mode='reflect')`` (truncate=4.0)."""
    image = _as_float(image)
    return ndi.gaussian_filter(image, sigma, mode="reflect", truncate=4.0)


def gaussian_weights(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """scipy's ``_gaussian_kernel1d(sigma, 0, radius)`` with radius = int(truncate*sigma + 0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x**2)
    return phi / phi.sum()


def unsharp_mask(image: np.ndarray, radius: float = 1.0, amount: float = 1.0) -> np.ndarray:
    """``filters.unsharp_mask(image, radius, amount)``, single channel, preserve_range=False."""
    fimg = _as_float(image)
    vrange = (-1.0, 1.0) if np.any(fimg < 0) else (0.0, 1.0)
    blurred = gaussian(fimg, radius)
    result = fimg + (fimg - blurred) * amount
    return np.clip(result, vrange[0], vrange[1], out=result)
