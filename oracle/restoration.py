"""Restatement of ``skimage.restoration`` routines used by the reference hot path.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (scikit-image>=0.21 is an un-vendored
dependency, ``requirements.txt:4``; not installed here).  Follows
``skimage/restoration/_denoise.py``: ``estimate_sigma`` / ``_sigma_est_dwt``,
``denoise_wavelet`` / ``_wavelet_threshold`` / ``_bayes_thresh`` and
``denoise_tv_chambolle`` / ``_denoise_tv_chambolle_nd``.

Reference call sites: ``pipeline/metrics.py:47``; ``pipeline/enhancement.py:59-60,82,86,
169,270,311,328,349``.
"""

from __future__ import annotations

import numpy as np
from scipy import stats

from . import wavelets as wv

#: ``scipy.stats.norm.ppf(0.75)`` — a numpy float64 scalar, as in skimage.
_GAUSS_Q75 = stats.norm.ppf(0.75)


def _sigma_est_dwt(detail_coeffs: np.ndarray):
    """MAD estimate: median(|d|, d != 0) / norm.ppf(.75).  float32 median / float64 -> float64."""
    nz = detail_coeffs[np.nonzero(detail_coeffs)]
    return np.median(np.abs(nz)) / _GAUSS_Q75


def estimate_sigma(image: np.ndarray):
    """``estimate_sigma(image, channel_axis=None, average_sigmas=True)`` for a 2-D image."""
    return _sigma_est_dwt(wv.dwtn_db2_dd(image))


def _bayes_thresh(details: np.ndarray, var):
    dvar = np.mean(details * details)
    eps = np.finfo(details.dtype).eps
    return var / np.sqrt(max(dvar - var, eps))


def denoise_wavelet(image: np.ndarray, sigma=None, mode: str = "soft") -> np.ndarray:
    """``denoise_wavelet(image, channel_axis=None, rescale_sigma=True, mode=mode[, sigma=sigma])``
    with the defaults wavelet='db1', method='BayesShrink', wavelet_levels=None.  Float input
    is neither rescaled nor clipped."""
    image = np.asarray(image)
    if image.dtype == np.float16:
        image = image.astype(np.float32)
    h, w = image.shape
    levels = max(wv.haar_max_level(image.shape) - 3, 1)
    coeffs = wv.haar_wavedec2(image, levels)
    dcoeffs = coeffs[1:]
    if sigma is None:
        sigma = _sigma_est_dwt(dcoeffs[-1]["dd"])
    var = sigma**2
    shrink = wv.threshold_soft if mode == "soft" else wv.threshold_hard
    denoised = [coeffs[0]]
    for lvl in dcoeffs:
        denoised.append({k: shrink(lvl[k], _bayes_thresh(lvl[k], var)) for k in lvl})
    out = wv.haar_waverec2(denoised)[:h, :w]
    return out.astype(image.dtype)


def denoise_tv_chambolle(image: np.ndarray, weight: float = 0.1, eps: float = 2.0e-4,
                         max_num_iter: int = 200, return_iters: bool = False):
    """``denoise_tv_chambolle(image, weight=weight, channel_axis=None)`` on a 2-D float32 image."""
    image = np.asarray(image)
    if image.dtype.kind != "f":
        raise TypeError("oracle handles float images only")
    if image.dtype == np.float16:
        image = image.astype(np.float32)
    ndim = image.ndim
    p = np.zeros((ndim,) + image.shape, dtype=image.dtype)
    g = np.zeros_like(p)
    d = np.zeros_like(image)
    out = image
    i = 0
    e_init = e_prev = None
    while i < max_num_iter:
        if i > 0:
            d = -p.sum(0)
            d[1:, :] += p[0, :-1, :]
            d[:, 1:] += p[1, :, :-1]
            out = image + d
        else:
            out = image
        energy = (d**2).sum()
        g[0, :-1, :] = np.diff(out, axis=0)
        g[1, :, :-1] = np.diff(out, axis=1)
        norm = np.sqrt((g**2).sum(axis=0))[np.newaxis, ...]
        energy += weight * norm.sum()
        tau = 1.0 / (2.0 * ndim)
        norm *= tau / weight
        norm += 1.0
        p -= tau * g
        p /= norm
        energy /= float(image.size)
        if i == 0:
            e_init = energy
            e_prev = energy
        else:
            if np.abs(e_prev - energy) < eps * e_init:
                break
            e_prev = energy
        i += 1
    if return_iters:
        # number of loop bodies executed
        return out, min(i + 1, max_num_iter)
    return out
