"""Restatement of ``skimage.exposure`` routines used by the reference hot path.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (scikit-image>=0.21 un-vendored, not installed).
Follows ``skimage/exposure/_adapthist.py`` (``equalize_adapthist``, ``_clahe``,
``clip_histogram``, ``map_histogram``), ``skimage/exposure/exposure.py``
(``rescale_intensity``, ``adjust_gamma``) and ``skimage/util/dtype.py`` (``img_as_uint``).

Reference call sites: ``pipeline/enhancement.py:183,194,197,277,284,332,336``.
"""

from __future__ import annotations

import math

import numpy as np

NR_OF_GRAY = 2**14


def img_as_uint(image: np.ndarray) -> np.ndarray:
    """float -> uint16: range check, multiply by 65535 in float32, rint, clip, cast."""
    image = np.asarray(image)
    if image.dtype.kind != "f":
        raise TypeError("oracle handles float images only")
    if np.min(image) < -1.0 or np.max(image) > 1.0:
        raise ValueError("Images of type float must be between -1 and 1.")
    comp = image.dtype if image.dtype.itemsize >= 2 else np.float32
    out = np.multiply(image, 65535, dtype=comp)
    np.rint(out, out=out)
    np.clip(out, 0, 65535, out=out)
    return out.astype(np.uint16)


def _rescale_u16_to_gray(image_u16: np.ndarray) -> np.ndarray:
    """``np.round(rescale_intensity(image_u16, out_range=(0, NR_OF_GRAY-1))).astype(uint16)``:
    float64 arithmetic, divide first then scale, round half to even."""
    imin, imax = float(image_u16.min()), float(image_u16.max())
    omin, omax = 0.0, float(NR_OF_GRAY - 1)
    img = np.clip(image_u16, imin, imax)
    if imin != imax:
        img = (img - imin) / (imax - imin)
        res = (img * (omax - omin) + omin).astype(np.float64)
    else:
        res = np.clip(img, omin, omax).astype(np.float64)
    return np.round(res).astype(np.uint16)


def _rescale_float_to_unit(image: np.ndarray) -> np.ndarray:
    """``rescale_intensity(image)`` for a non-negative float32 image: in_range 'image',
    out_range 'dtype' -> (0, 1); python-float scalars act as float32 (NEP 50)."""
    imin, imax = float(np.min(image)), float(np.max(image))
    omin, omax = (0.0, 1.0) if imin >= 0 else (-1.0, 1.0)
    img = np.clip(image, imin, imax)
    if imin != imax:
        img = (img - imin) / (imax - imin)
        return (img * (omax - omin) + omin).astype(image.dtype)
    return np.clip(img, omin, omax).astype(image.dtype)


def clip_histogram(hist: np.ndarray, clip_limit: int) -> np.ndarray:
    """Clip one contextual-region histogram and redistribute the excess (in place)."""
    excess_mask = hist > clip_limit
    excess = hist[excess_mask]
    n_excess = excess.sum() - excess.size * clip_limit
    hist[excess_mask] = clip_limit

    bin_incr = n_excess // hist.size
    upper = clip_limit - bin_incr

    low_mask = hist < upper
    n_excess -= hist[low_mask].size * bin_incr
    hist[low_mask] += bin_incr

    mid_mask = np.logical_and(hist >= upper, hist < clip_limit)
    mid = hist[mid_mask]
    n_excess += mid.sum() - mid.size * clip_limit
    hist[mid_mask] = clip_limit

    while n_excess > 0:
        prev_n_excess = n_excess
        for index in range(hist.size):
            under_mask = hist < clip_limit
            step_size = max(1, np.count_nonzero(under_mask) // n_excess)
            under_mask = under_mask[index::step_size]
            hist[index::step_size][under_mask] += 1
            n_excess -= np.count_nonzero(under_mask)
            if n_excess <= 0:
                break
        if prev_n_excess == n_excess:
            break
    return hist


def map_histogram(hist: np.ndarray, min_val: int, max_val: int, n_pixels: int) -> np.ndarray:
    out = np.cumsum(hist, axis=-1).astype(float)
    out *= (max_val - min_val) / n_pixels
    out += min_val
    np.clip(out, a_min=None, a_max=max_val, out=out)
    return out.astype(int)


def _clahe(image: np.ndarray, kernel_size, clip_limit: float, nbins: int,
           return_internals: bool = False):
    ndim = image.ndim
    dtype = image.dtype
    pad_start = [k // 2 for k in kernel_size]
    pad_end = [(k - s % k) % k + int(np.ceil(k / 2.0)) for k, s in zip(kernel_size, image.shape)]
    image = np.pad(image, [[a, b] for a, b in zip(pad_start, pad_end)], mode="reflect")

    bin_size = 1 + NR_OF_GRAY // nbins
    lut = np.arange(NR_OF_GRAY, dtype=np.min_scalar_type(NR_OF_GRAY))
    lut //= bin_size
    image = lut[image]

    ns_hist = [int(s / k) - 1 for s, k in zip(image.shape, kernel_size)]
    hist_blocks_shape = np.array([ns_hist, kernel_size]).T.flatten()
    hist_axis_order = np.array([np.arange(0, ndim * 2, 2), np.arange(1, ndim * 2, 2)]).flatten()
    hist_slices = [slice(k // 2, k // 2 + n * k) for k, n in zip(kernel_size, ns_hist)]
    hist_blocks = image[tuple(hist_slices)].reshape(hist_blocks_shape)
    hist_blocks = np.transpose(hist_blocks, axes=hist_axis_order)
    hist_block_assembled_shape = hist_blocks.shape
    hist_blocks = hist_blocks.reshape((math.prod(ns_hist), -1))

    kernel_elements = math.prod(kernel_size)
    if clip_limit > 0.0:
        clim = int(np.clip(clip_limit * kernel_elements, 1, None))
    else:
        clim = np.iinfo(hist_blocks.dtype).max

    hist = np.apply_along_axis(np.bincount, -1, hist_blocks, minlength=nbins)
    raw_hist = hist.copy() if return_internals else None
    hist = np.apply_along_axis(clip_histogram, -1, hist, clip_limit=clim)
    hist = map_histogram(hist, 0, NR_OF_GRAY - 1, kernel_elements)
    hist = hist.reshape(hist_block_assembled_shape[:ndim] + (-1,))

    map_array = np.pad(hist, [[1, 1] for _ in range(ndim)] + [[0, 0]], mode="edge")

    ns_proc = [int(s / k) for s, k in zip(image.shape, kernel_size)]
    blocks_shape = np.array([ns_proc, kernel_size]).T.flatten()
    blocks_axis_order = np.array([np.arange(0, ndim * 2, 2), np.arange(1, ndim * 2, 2)]).flatten()
    blocks = image.reshape(blocks_shape)
    blocks = np.transpose(blocks, axes=blocks_axis_order)
    blocks_flattened_shape = blocks.shape
    blocks = np.reshape(blocks, (math.prod(ns_proc), math.prod(blocks.shape[ndim:])))

    coeffs = np.meshgrid(*tuple([np.arange(k) / k for k in kernel_size[::-1]]), indexing="ij")
    coeffs = [np.transpose(c).flatten() for c in coeffs]
    inv_coeffs = [1 - c for c in coeffs]

    result = np.zeros(blocks.shape, dtype=np.float32)
    for edge in np.ndindex(*([2] * ndim)):
        edge_maps = map_array[tuple([slice(e, e + n) for e, n in zip(edge, ns_proc)])]
        edge_maps = edge_maps.reshape((math.prod(ns_proc), -1))
        edge_mapped = np.take_along_axis(edge_maps, blocks, axis=-1)
        edge_coeffs = np.prod([[inv_coeffs, coeffs][e][d] for d, e in enumerate(edge[::-1])], 0)
        result += (edge_mapped * edge_coeffs).astype(result.dtype)

    result = result.astype(dtype)
    result = result.reshape(blocks_flattened_shape)
    rebuild_order = np.array([np.arange(0, ndim), np.arange(ndim, ndim * 2)]).T.flatten()
    result = np.transpose(result, axes=rebuild_order)
    result = result.reshape(image.shape)
    unpad = tuple([slice(a, s - b) for a, b, s in zip(pad_start, pad_end, image.shape)])
    if return_internals:
        return result[unpad], {"raw_hist": raw_hist, "maps": hist, "clim": clim,
                               "ns_hist": ns_hist, "binned": image}
    return result[unpad]


def equalize_adapthist(image: np.ndarray, kernel_size=None, clip_limit: float = 0.01,
                       nbins: int = 256, return_internals: bool = False):
    """CLAHE as in skimage >= 0.19 (16-bit -> 14-bit working image, 256 bins)."""
    image = np.asarray(image)
    float_dtype = np.float32 if image.dtype in (np.float16, np.float32) else np.float64
    img = img_as_uint(image)
    img = _rescale_u16_to_gray(img)
    if kernel_size is None:
        kernel_size = tuple([max(s // 8, 1) for s in img.shape])
    elif np.isscalar(kernel_size):
        kernel_size = (kernel_size,) * img.ndim
    elif len(kernel_size) != img.ndim:
        raise ValueError(f"Incorrect value of `kernel_size`: {kernel_size}")
    kernel_size = [int(k) for k in kernel_size]
    res = _clahe(img, kernel_size, clip_limit, nbins, return_internals=return_internals)
    internals = None
    if return_internals:
        res, internals = res
        internals["quantised"] = img
        internals["stage_u16"] = res.copy()
    res = res.astype(float_dtype, copy=False)
    out = _rescale_float_to_unit(res)
    if return_internals:
        return out, internals
    return out


def adjust_gamma(image: np.ndarray, gamma: float = 1, gain: float = 1) -> np.ndarray:
    """``adjust_gamma`` for float input: ValueError on negatives, ((x/1.0)**gamma)*1.0*gain."""
    if gamma < 0:
        raise ValueError("Gamma should be a non-negative real number.")
    image = np.asarray(image)
    if np.any(image < 0):
        raise ValueError(
            "Image Correction methods work correctly only on images with non-negative values. "
            "Use skimage.exposure.rescale_intensity."
        )
    scale = 1.0
    return (((image / scale) ** gamma) * scale * gain).astype(image.dtype)
