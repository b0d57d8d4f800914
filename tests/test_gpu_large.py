"""BASELINE.json's other configurations as parity cases, at their real sizes where the oracle
finishes in seconds, and through size-independent properties where it does not:
  C3  3000x3000 radiograph, CLAHE + unsharp heavy (P_cr)
  C5  metrics-only sweep 256^2 ... 4096^2 (validation-agent path)
  C2/C4  512x512 stacks: chunking / worker invariance, per-slice independence, idempotence."""

from __future__ import annotations

import numpy as np
import pytest
import torch

from mdimg_b200.engine import METRIC_KEYS
from conftest import assert_within_lsb
from oracle import ref_enhancement as oenh
from oracle import ref_metrics as omet

pytestmark = pytest.mark.gpu
LSB16 = 1.0 / 65535


@pytest.fixture(scope="module")
def radiograph(synth):
    return omet.normalize_image(synth.radiograph(2000, 3000))


def test_c3_radiograph_p_cr_matches_oracle(ops, synth, radiograph):
    from mdimg_b200.pipeline.enhancement import apply_enhancements_from_params
    got, labels = apply_enhancements_from_params(radiograph, synth.plan_cr())
    ref, ref_labels = oenh.apply_enhancements_from_params(radiograph, synth.plan_cr())
    assert labels == ref_labels
    assert got.shape == (3000, 3000) and got.dtype == np.float32
    assert np.abs(got - ref).max() <= LSB16


def test_c3_radiograph_metrics_and_validation(ops, dev, radiograph):
    other = np.clip(radiograph ** np.float32(0.9), 0, 1).astype(np.float32)
    row = ops.metrics(dev(radiograph), with_niqe=True)[0].cpu().numpy()
    ref = omet.compute_metrics(radiograph)
    for i, key in enumerate(METRIC_KEYS):
        assert row[i] == pytest.approx(ref[key], rel=1e-5, abs=1e-9), key
    from mdimg_b200.pipeline.metrics import compute_validation
    v = compute_validation(radiograph, other)
    rv = omet.compute_validation(radiograph, other)
    for key in ("ssim", "psnr", "niqe_before", "niqe_after", "edge_ratio", "quality_improvement"):
        assert v[key] == pytest.approx(rv[key], rel=1e-5, abs=1e-8), key
    assert v["passes"] == rv["passes"]


@pytest.mark.parametrize("size", [256, 1024, 2048, 4096])
def test_c5_metrics_sweep(ops, dev, synth, size):
    im = synth.unit_image(4000 + size, size)
    row = ops.metrics(dev(im), with_niqe=True)[0].cpu().numpy()
    ref = omet.compute_metrics(im)
    for i, key in enumerate(METRIC_KEYS):
        assert row[i] == pytest.approx(ref[key], rel=1e-5, abs=1e-9), (size, key)
    assert row[17] == pytest.approx(omet.compute_edge_ratio(im), rel=1e-5)
    assert row[18] == pytest.approx(omet.compute_niqe_approximation(im), rel=1e-5)
    copy = np.clip(im ** np.float32(0.9), 0, 1).astype(np.float32)
    fr = ops.fullref(dev(im), dev(copy))[0].cpu().numpy()
    from oracle.fullref import peak_signal_noise_ratio, structural_similarity
    assert fr[0] == pytest.approx(float(structural_similarity(im, copy, data_range=1.0)), rel=1e-9)
    assert fr[1] == pytest.approx(float(peak_signal_noise_ratio(im, copy, data_range=1.0)), rel=1e-9)


@pytest.fixture(scope="module")
def ct_stack(synth, ops):
    raw = np.stack([synth.ct_slice(3000 + z, z / 96) for z in range(96)])
    return raw, torch.from_numpy(raw.view(np.int16)).to(ops.device)


def test_c2_chunking_and_worker_invariance(ops, synth, ct_stack):
    """The result of a stack must not depend on how it is cut into chunks or on how many streams
    drive the chunks: pixels bit-identical, rows equal up to the order of float64 atomics."""
    from mdimg_b200.batch import process_stack
    raw, dev_raw = ct_stack
    plan = synth.plan_full()
    a = process_stack(dev_raw, plan, chunk=96, workers=1, ops=ops)
    b = process_stack(dev_raw, plan, chunk=32, workers=2, ops=ops)
    c = process_stack(dev_raw, plan, chunk=17, workers=3, ops=ops)
    for other in (b, c):
        assert torch.equal(a.enhanced, other.enhanced)
        np.testing.assert_allclose(a.packed, other.packed, rtol=1e-9, atol=1e-12)
        assert a.labels == other.labels
    assert not a.failed.any()
    assert (a.tv_iterations >= 2).all() and (a.tv_iterations <= 200).all()
    e = a.enhanced
    assert float(e.min()) >= 0.0 and float(e.max()) <= 1.0


def test_c2_slices_are_independent_and_match_the_oracle(ops, synth, ct_stack):
    from mdimg_b200.batch import process_stack
    raw, dev_raw = ct_stack
    plan = synth.plan_full()
    res = process_stack(dev_raw, plan, chunk=96, ops=ops)
    perm = torch.arange(95, -1, -1, device=ops.device)
    rev = process_stack(dev_raw[perm].contiguous(), plan, chunk=96, ops=ops)
    assert torch.equal(res.enhanced, rev.enhanced[perm])
    for z in (0, 47, 95):
        x = omet.normalize_image(raw[z])
        ref, ref_labels = oenh.apply_enhancements_from_params(x, plan)
        assert res.labels[z] == ref_labels
        assert_within_lsb(res.enhanced[z].cpu().numpy(), ref, f"C2 slice {z}")
        got = res.metrics_before(z)
        want = omet.compute_metrics(x)
        for k in METRIC_KEYS:
            assert got[k] == pytest.approx(want[k], rel=1e-5, abs=1e-9), k


def test_idempotence_and_fixed_points(ops, ct_stack):
    _, dev_raw = ct_stack
    x = ops.normalize(dev_raw[:8].contiguous())
    assert torch.equal(ops.normalize(x), x)                      # already spans [0, 1] exactly
    y = torch.empty_like(x)
    ops.clip01(x, y)
    assert torch.equal(y, x)
    ops.clahe(x, y, 0.015, 16)
    assert float(y.min()) == 0.0 and float(y.max()) == 1.0
    fr = ops.fullref(x, x).cpu().numpy()
    assert np.allclose(fr[:, 0], 1.0) and np.isinf(fr[:, 1]).all()
    z = torch.empty_like(x)
    it = ops.tv_chambolle(x, z, 1e-6)
    assert float((z - x).abs().max()) < 1e-4 and int(it.max()) <= 200


@pytest.mark.parametrize("shape,cap", [((3000, 3000), 5), ((1000, 1502), 8), ((513, 770), 11)])
def test_tv_packed_kernels_at_radiograph_sizes(ops, synth, shape, cap):
    """Strip grids far from the 512x512 case (row strips that do not divide the height, 50 column
    strips, an odd height): the field after `cap` bodies equals the oracle's bit for bit."""
    from oracle import restoration as ores
    h, w = shape
    base = omet.normalize_image(synth.radiograph(2000, 3000))
    im = np.ascontiguousarray(base[:h, :w])
    x = torch.from_numpy(im[None]).to(ops.device)
    out = torch.empty_like(x)
    iters = int(ops.tv_chambolle(x, out, 0.1, eps=0.0, max_iter=cap)[0].item())
    ref, ref_iters = ores.denoise_tv_chambolle(im, 0.1, eps=0.0, max_num_iter=cap, return_iters=True)
    assert iters == ref_iters == cap
    np.testing.assert_array_equal(out[0].cpu().numpy(), ref)
