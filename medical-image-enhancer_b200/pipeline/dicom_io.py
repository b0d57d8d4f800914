"""``normalize_image`` — drop-in for ``pipeline/dicom_io.py:84-91`` (the windowing/normalisation
step that produces the hot path's input) — and ``save_visuals`` (``dicom_io.py:99-126``) as a GPU
mosaic + PNG.  DICOM file parsing and report text stay in the reference."""

from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from .. import png
from ..stack import get_ops


def save_visuals(original: np.ndarray, enhanced: np.ndarray, out_dir: str, base_name: str) -> Dict[str, str]:
    """Side-by-side before/after PNG, same signature and return value as the reference
    (``{"before_after": path}``, file ``<base_name>_before_after.png``).  The reference renders a
    matplotlib figure (titles, resampling to a 1500x750 canvas); this writes the two panels at native
    resolution with matplotlib's gray colormap quantisation (each panel autoscaled to its own
    min / max, 256 levels), computed on the GPU."""
    a = np.ascontiguousarray(original, dtype=np.float32)
    b = np.ascontiguousarray(enhanced, dtype=np.float32)
    if a.ndim != 2 or a.shape != b.shape:
        raise ValueError("save_visuals expects two 2-D images of the same shape")
    ops = get_ops()
    mosaic = ops.mosaic(torch.from_numpy(a[None]).to(ops.device), torch.from_numpy(b[None]).to(ops.device))
    os.makedirs(out_dir, exist_ok=True)
    figure_path = os.path.join(out_dir, f"{base_name}_before_after.png")
    png.write_gray8(figure_path, mosaic[0].cpu().numpy())
    return {"before_after": figure_path}


def normalize_image(image: np.ndarray) -> np.ndarray:
    """Normalise pixel values to [0, 1] (float32); a constant image becomes all zeros."""
    arr = np.asarray(image)
    if arr.ndim != 2:
        raise ValueError(f"normalize_image expects a 2-D image, got shape {arr.shape}")
    ops = get_ops()
    if arr.dtype == np.uint16:
        dev = torch.from_numpy(np.ascontiguousarray(arr).view(np.int16)).to(ops.device)
    else:
        dev = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(ops.device)
    out = ops.normalize(dev[None])
    return out[0].cpu().numpy()


def ingest_stack(raw: np.ndarray, slope=None, intercept=None, monochrome1: bool = False) -> np.ndarray:
    """Pixel path of the reference's ``load_dicom`` (pipeline/dicom_io.py:44-49) fused with
    ``normalize_image`` for every frame of a 16-bit stack: modality rescale
    ``float32(float64(raw) * slope + intercept)``, MONOCHROME1 inversion against the maximum of the
    whole pixel array, per-frame normalisation to [0, 1].  Unlike the reference, which keeps only
    the middle frame of a multi-frame object (dicom_io.py:72-73), every frame is returned."""
    arr = np.asarray(raw)
    if arr.dtype not in (np.uint16, np.int16):
        raise ValueError("ingest_stack expects int16 or uint16 samples")
    if arr.ndim == 2:
        arr = arr[None]
    ops = get_ops()
    dev = torch.from_numpy(np.ascontiguousarray(arr).view(np.int16)).to(ops.device)
    if arr.dtype == np.uint16:
        dev = dev.view(torch.uint16)
    return ops.ingest(dev, slope, intercept, monochrome1).cpu().numpy()
