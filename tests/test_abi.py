"""The C-ABI library builds, loads without a GPU and exports every symbol include/mdimg_b200.h
declares; the product path fails loudly when no CUDA device is present (no CPU fallback)."""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mdimg_b200.h"


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from mdimg_b200 import _lib
    return _lib.load_library()


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mdimg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ("mdimg_init", "mdimg_last_error", "mdimg_workspace_bytes", "mdimg_normalize_u16",
                 "mdimg_metrics", "mdimg_estimate_sigma", "mdimg_quality", "mdimg_fullref",
                 "mdimg_wavelet_denoise", "mdimg_clahe", "mdimg_gamma", "mdimg_unsharp",
                 "mdimg_light_denoise", "mdimg_bilateral", "mdimg_tv_chambolle"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in mdimg_b200.h but not exported"


def test_binding_table_matches_header(lib):
    from mdimg_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_workspace_queries_need_no_gpu(lib):
    from mdimg_b200 import _lib
    for op, param in ((_lib.OP_METRICS, 0), (_lib.OP_WAVELET, 0), (_lib.OP_CLAHE, 16), (_lib.OP_TV, 200),
                      (_lib.OP_LIGHT_DENOISE, 0), (_lib.OP_SIGMA, 0)):
        small = lib.mdimg_workspace_bytes(op, 1, 64, 64, param)
        big = lib.mdimg_workspace_bytes(op, 8, 512, 512, param)
        assert 0 < small < big
    assert lib.mdimg_workspace_bytes(_lib.OP_BILATERAL, 4, 64, 64, 0) == 0


def test_invalid_arguments_are_reported(lib):
    rc = lib.mdimg_clip01(None, None, 1, 0, 5, None, 0, None)
    assert rc == 1
    assert b"invalid stack shape" in lib.mdimg_last_error()


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mdimg_b200 import _lib
    rc = lib.mdimg_init(0)
    assert rc == _lib.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.mdimg_last_error()
    from mdimg_b200.pipeline.metrics import compute_metrics
    import numpy as np
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        compute_metrics(np.zeros((64, 64), np.float32))


def test_product_path_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: the package, the
    import shim, the tools and the examples must not."""
    for folder in ("medical-image-enhancer_b200", "mdimg_b200", "tools", "examples"):
        for py in (ROOT / folder).rglob("*.py"):
            src = py.read_text()
            assert "import oracle" not in src and "from oracle" not in src, py
    for c_src in list((ROOT / "medical-image-enhancer_b200" / "csrc").glob("*.cu*")) + list((ROOT / "examples").glob("*.c")):
        assert "oracle/" not in c_src.read_text(), c_src          # no include of, or path into, oracle/


def test_plan_clamp_never_truncates_the_operation_list(lib):
    """mdimg_plan_clamp is host arithmetic (no GPU): PARAM_BOUNDS clamping, and a list longer than
    ops[] is an error instead of a silently shortened plan (the halo safeguard replays every entry)."""
    import ctypes as C
    from mdimg_b200 import _lib
    q = _lib.EnhancePlan()
    q.n_ops = _lib.MAX_PLAN_OPS
    q.gamma, q.clahe_clip_limit, q.unsharp_amount, q.tv_denoise_weight = 9.0, 0.0, -1.0, 1.0
    assert lib.mdimg_plan_clamp(C.byref(q)) == 0
    assert (q.n_ops, q.gamma, q.clahe_clip_limit, q.unsharp_amount, q.tv_denoise_weight) == (
        _lib.MAX_PLAN_OPS, 1.5, 0.002, 0.03, 0.15)
    q.n_ops = _lib.MAX_PLAN_OPS + 1
    assert lib.mdimg_plan_clamp(C.byref(q)) == 1
    assert b"capacity" in lib.mdimg_last_error()
    header = HEADER.read_text()
    assert f"#define MDIMG_MAX_PLAN_OPS {_lib.MAX_PLAN_OPS}" in header
