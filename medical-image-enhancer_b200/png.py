"""Minimal 8-bit grayscale PNG writer (zlib from the standard library; matplotlib / PIL are not
needed on the product path).  Used by the report mosaics that replace the reference's matplotlib
figure (pipeline/dicom_io.py:99-126)."""

from __future__ import annotations

import struct
import zlib

import numpy as np


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def encode_gray8(image: np.ndarray, level: int = 6) -> bytes:
    """PNG bytes of a 2-D uint8 array (colour type 0, bit depth 8, filter 0 on every row)."""
    arr = np.ascontiguousarray(image)
    if arr.ndim != 2 or arr.dtype != np.uint8:
        raise ValueError("encode_gray8 expects a 2-D uint8 array")
    h, w = arr.shape
    raw = np.empty((h, w + 1), np.uint8)
    raw[:, 0] = 0                      # filter type "None"
    raw[:, 1:] = arr
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)
    return (b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level))
            + _chunk(b"IEND", b""))


def write_gray8(path: str, image: np.ndarray, level: int = 6) -> None:
    with open(path, "wb") as f:
        f.write(encode_gray8(image, level))
