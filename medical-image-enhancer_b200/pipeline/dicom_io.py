"""``normalize_image`` — drop-in for ``pipeline/dicom_io.py:84-91`` (the windowing/normalisation
step that produces the hot path's input).  File I/O, plotting and report text stay in the
reference."""

from __future__ import annotations

import numpy as np
import torch

from ..stack import get_ops


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
