"""Restatement of the PyWavelets routines the reference reaches through scikit-image.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: PyWavelets
(pinned only as ``PyWavelets>=1.4`` in the reference's ``requirements.txt:5``) is not
installed here, so this follows its published algorithm:

* ``pywt/_extensions/c/convolution.template.c`` ``downsampling_convolution``: output
  ``o`` of a decimating pass is ``sum_j filter[j] * x~[2*o + 1 - j]`` accumulated in the
  data's own precision (float32 data -> float32 filters and float32 accumulation, ``j``
  ascending), ``x~`` the half-sample symmetric extension (mode ``'symmetric'``), and the
  output has ``(N + F - 1) // 2`` samples.
* ``pywt/_multilevel.py`` ``wavedecn`` / ``waverecn`` and ``pywt/_multidim.py``
  ``dwtn`` / ``idwtn``: axis 0 is decomposed first, then axis 1; reconstruction runs the
  axes in reverse; mixed float32/float64 operands are promoted to float64.
* ``pywt/_thresholding.py`` ``soft`` / ``hard``.

Call sites in the reference: ``estimate_sigma`` (``pipeline/metrics.py:47``,
``pipeline/enhancement.py:59-60,82``) uses ``dwtn(x, 'db2')['dd']``; ``denoise_wavelet``
(``pipeline/enhancement.py:86,169,270,328``) uses ``wavedecn/waverecn`` with ``'db1'``.
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

# Decomposition filters as PyWavelets stores them (double); float32 passes use the
# float32 rounding of the same numbers.
_SQRT1_2 = 0.7071067811865476
DB1_DEC_LO = (_SQRT1_2, _SQRT1_2)
DB1_DEC_HI = (-_SQRT1_2, _SQRT1_2)
DB1_REC_LO = (_SQRT1_2, _SQRT1_2)
DB1_REC_HI = (_SQRT1_2, -_SQRT1_2)
DB2_DEC_LO = (-0.12940952255126037, 0.2241438680420134, 0.8365163037378079, 0.48296291314453416)
DB2_DEC_HI = (-0.48296291314453416, 0.8365163037378079, -0.2241438680420134, -0.12940952255126037)


def _work_dtype(x: np.ndarray) -> np.dtype:
    # pywt keeps float32 as float32 and turns everything else into float64.
    return np.dtype(np.float32) if x.dtype == np.float32 else np.dtype(np.float64)


def dwt_axis(x: np.ndarray, filt: Tuple[float, ...], axis: int) -> np.ndarray:
    """One decimating analysis pass along ``axis`` (mode 'symmetric')."""
    dt = _work_dtype(x)
    x = np.moveaxis(np.asarray(x, dtype=dt), axis, -1)
    n = x.shape[-1]
    flen = len(filt)
    n_out = (n + flen - 1) // 2
    i_max = 2 * (n_out - 1) + 1
    left = flen - 2
    right = max(i_max - (n - 1), 0)
    pad = [(0, 0)] * (x.ndim - 1) + [(left, right)]
    ext = np.pad(x, pad, mode="symmetric")
    f = [dt.type(c) for c in filt]

    def tap(j: int) -> np.ndarray:
        start = 1 - j + left
        return ext[..., start : start + 2 * n_out : 2]

    acc = f[0] * tap(0)
    for j in range(1, flen):
        acc = acc + f[j] * tap(j)
    if flen == 4 and n % 2 == 1 and n >= 3:
        # Right-overhang output i = N + 2: pywt walks the mirrored part first
        # (filter[2]*x[N-1], filter[1]*x[N-2], filter[0]*x[N-3]) and then filter[3]*x[N-1].
        last = f[2] * x[..., n - 1]
        last = last + f[1] * x[..., n - 2]
        last = last + f[0] * x[..., n - 3]
        last = last + f[3] * x[..., n - 1]
        acc[..., n_out - 1] = last
    return np.moveaxis(acc, -1, axis)


def dwtn_db2_dd(x: np.ndarray) -> np.ndarray:
    """``pywt.dwtn(x, 'db2')['dd']``: high-pass along axis 0, then along axis 1."""
    return dwt_axis(dwt_axis(x, DB2_DEC_HI, 0), DB2_DEC_HI, 1)


def haar_dwt2(x: np.ndarray) -> Dict[str, np.ndarray]:
    """``pywt.dwtn(x, 'db1')``; key letters are (axis 0, axis 1)."""
    lo0 = dwt_axis(x, DB1_DEC_LO, 0)
    hi0 = dwt_axis(x, DB1_DEC_HI, 0)
    return {
        "aa": dwt_axis(lo0, DB1_DEC_LO, 1),
        "ad": dwt_axis(lo0, DB1_DEC_HI, 1),
        "da": dwt_axis(hi0, DB1_DEC_LO, 1),
        "dd": dwt_axis(hi0, DB1_DEC_HI, 1),
    }


def _haar_idwt_axis(a: np.ndarray, d: np.ndarray, axis: int) -> np.ndarray:
    """``pywt.idwt_axis`` for db1: out = upconv(a, rec_lo) accumulated, then += upconv(d, rec_hi)."""
    if a.dtype != d.dtype:
        a = a.astype(np.float64)
        d = d.astype(np.float64)
    dt = _work_dtype(a)
    a = np.moveaxis(np.asarray(a, dtype=dt), axis, -1)
    d = np.moveaxis(np.asarray(d, dtype=dt), axis, -1)
    n = a.shape[-1]
    out = np.empty(a.shape[:-1] + (2 * n,), dtype=dt)
    lo = [dt.type(c) for c in DB1_REC_LO]
    hi = [dt.type(c) for c in DB1_REC_HI]
    out[..., 0::2] = lo[0] * a + hi[0] * d
    out[..., 1::2] = lo[1] * a + hi[1] * d
    return np.moveaxis(out, -1, axis)


def haar_idwt2(c: Dict[str, np.ndarray]) -> np.ndarray:
    """``pywt.idwtn`` for db1: axis 1 first ('aa'+'ad' -> 'a', 'da'+'dd' -> 'd'), then axis 0."""
    a = _haar_idwt_axis(c["aa"], c["ad"], 1)
    d = _haar_idwt_axis(c["da"], c["dd"], 1)
    return _haar_idwt_axis(a, d, 0)


def haar_max_level(shape: Tuple[int, ...]) -> int:
    """``pywt.dwtn_max_level(shape, 'db1')`` = floor(log2(min dim)) for a 2-tap filter."""
    n = min(shape)
    if n < 1:
        return 0
    return int(math.floor(math.log2(n)))


def haar_wavedec2(x: np.ndarray, level: int) -> List:
    """``pywt.wavedecn(x, 'db1', level=level)`` -> [cA_L, {ad,da,dd}_L, ..., {..}_1]."""
    a = np.asarray(x)
    details = []
    for _ in range(level):
        c = haar_dwt2(a)
        a = c.pop("aa")
        details.append(c)
    details.reverse()
    return [a] + details


def haar_waverec2(coeffs: List) -> np.ndarray:
    """``pywt.waverecn``: coarsest to finest, trimming the running approximation to the
    stored detail shape when it is one sample longer (odd lengths)."""
    a = coeffs[0]
    for idx, d in enumerate(coeffs[1:]):
        if idx > 0:
            shp = d["dd"].shape
            a = a[: shp[0], : shp[1]]
        full = dict(d)
        full["aa"] = a
        a = haar_idwt2(full)
    return a


def threshold_soft(data: np.ndarray, value) -> np.ndarray:
    """``pywt.threshold(data, value, 'soft')``: data * clip(1 - value/|data|, 0, None)."""
    data = np.asarray(data)
    magnitude = np.absolute(data)
    with np.errstate(divide="ignore", invalid="ignore"):
        shrink = 1 - value / magnitude
        shrink.clip(min=0, max=None, out=shrink)
        shrink = data * shrink
    return shrink


def threshold_hard(data: np.ndarray, value) -> np.ndarray:
    """``pywt.threshold(data, value, 'hard')``: where(|data| < value, 0, data)."""
    data = np.asarray(data)
    return np.where(np.less(np.absolute(data), value), 0, data)
