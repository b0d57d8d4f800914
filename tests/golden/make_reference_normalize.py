"""Golden vectors for normalize_image from the REFERENCE'S OWN function body (pipeline/dicom_io.py:84-91).

pipeline/dicom_io.py imports pydicom and matplotlib (absent here), but normalize_image itself is pure
numpy: its source is cut out of the reference file with `ast` and executed as it is.
    python tests/golden/make_reference_normalize.py
"""

from __future__ import annotations

import ast
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
REFERENCE = Path("/root/reference/pipeline/dicom_io.py")

from mdimg_b200 import synth  # noqa: E402


def reference_function(name: str):
    src = REFERENCE.read_text()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), str(REFERENCE), "exec"), ns)   # noqa: S102
    return ns[name]


def cases() -> dict:
    rng = np.random.default_rng(123)
    ct = synth.ct_slice(1000, 0.25, size=96)
    return {
        "ct96_u16": ct,
        "small_int": np.array([[10, 20], [30, 40]], dtype=np.int64),          # tests/test_detection.py:107-112
        "constant_u16": np.full((32, 48), 1234, np.uint16),                     # tests/test_detection.py:114-117
        "int16_negative": (rng.integers(-1024, 3071, (64, 80))).astype(np.int16),
        "float32_wide": (rng.standard_normal((48, 64)) * 1e3).astype(np.float32),
        "float64_tiny_range": (1.0 + rng.random((40, 40)) * 1e-9),
        "u16_full_range": rng.integers(0, 65536, (64, 64)).astype(np.uint16),
        "odd_size_u16": rng.integers(0, 4096, (37, 53)).astype(np.uint16),
    }


def main() -> None:
    assert REFERENCE.exists(), "needs the reference checkout"
    fn = reference_function("normalize_image")
    out = {}
    for name, arr in cases().items():
        out[f"in|{name}"] = arr
        res = fn(arr)
        assert res.dtype == np.float32 and res.shape == arr.shape
        out[f"out|{name}"] = res
    np.savez_compressed(HERE / "reference_normalize.npz", **out)
    print(len(out) // 2, "cases")


if __name__ == "__main__":
    main()
