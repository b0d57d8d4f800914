"""Shared fixtures.  GPU tests are marked ``gpu`` (run with ``-m gpu`` on a B200); everything else
runs on the CPU-only box."""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


LSB16 = 1.0 / 65535


def assert_within_lsb(got, ref, what="", lsb=LSB16):
    """north_star bar for pixels: max-abs <= 1 LSB of a 16-bit export, every pixel.  On failure the
    worst pixel and the number of offenders are reported."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.size == 0:
        return
    err = np.abs(got - ref)
    worst = np.unravel_index(int(np.argmax(err)), err.shape)
    assert err[worst] <= lsb, (
        f"{what}: max |diff| = {err[worst]:.3e} = {err[worst] / LSB16:.2f} LSB at {worst} "
        f"(got {got[worst]!r}, ref {ref[worst]!r}); {int((err > lsb).sum())} of {err.size} pixels over the bar")


@pytest.fixture(scope="session")
def synth():
    from mdimg_b200 import synth as s
    return s


# The reference's own fixtures (tests/conftest.py:9-32), same seeds and formulas.
@pytest.fixture
def synthetic_image_clean(synth) -> np.ndarray:
    return synth.fixture_clean()


@pytest.fixture
def synthetic_image_noisy(synth) -> np.ndarray:
    return synth.fixture_noisy()


@pytest.fixture
def synthetic_image_low_contrast(synth) -> np.ndarray:
    return synth.fixture_low_contrast()


@pytest.fixture(scope="session")
def images(synth):
    """Named float32 [0,1] test images: the three 64x64 reference fixtures, a 512x512 CT slice, a
    256x256 noisy field, an odd-sized crop and a 600x600 radiograph."""
    from oracle import ref_metrics as omet
    rng = np.random.default_rng(5)
    odd = synth.unit_image(4001, 256)[:94, :141].copy()
    return {
        "clean64": synth.fixture_clean(),
        "noisy64": synth.fixture_noisy(),
        "lowc64": synth.fixture_low_contrast(),
        "ct512": omet.normalize_image(synth.ct_slice(1000)),
        "unit256": synth.unit_image(4000, 256),
        "odd94x141": np.ascontiguousarray(odd + rng.normal(0, 0.01, odd.shape).astype(np.float32)).clip(0, 1),
        "cr600": omet.normalize_image(synth.radiograph(2000, 600)),
    }


@pytest.fixture(scope="session")
def ops():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mdimg_b200.stack import get_ops
    return get_ops()


@pytest.fixture(scope="session")
def dev(ops):
    import torch

    def _dev(a):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(ops.device)[None].contiguous()
    return _dev
