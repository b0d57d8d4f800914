"""The reference's own hot-path tests (tests/test_metrics.py, tests/test_pipeline.py,
tests/test_detection.py of Hiresh444/medical-image-enhancer), restated against the drop-in
modules, plus end-to-end parity of both enhancement entry points with the oracle."""

from __future__ import annotations

import numpy as np
import pytest

from conftest import assert_within_lsb
from oracle import ref_enhancement as oenh
from oracle import ref_metrics as omet

pytestmark = pytest.mark.gpu

LSB16 = 1.0 / 65535
EXPECTED_KEYS = {
    "sigma", "lap_var", "std", "pct_low", "pct_high", "entropy", "edge_density", "gradient_mag_mean",
    "gradient_mag_std", "snr_proxy", "cnr_proxy", "laplacian_energy", "histogram_spread",
    "local_contrast_std", "gradient_strength", "gradient_entropy",
}


@pytest.fixture(scope="module")
def api(ops):
    from mdimg_b200.pipeline import dicom_io, enhancement, metrics, schemas
    return type("Api", (), {"metrics": metrics, "enhancement": enhancement, "schemas": schemas,
                            "dicom_io": dicom_io})


# ---- tests/test_metrics.py -------------------------------------------------------------------
def test_returns_all_16_keys(api, synthetic_image_clean):
    m = api.metrics.compute_metrics(synthetic_image_clean)
    assert set(m) == EXPECTED_KEYS and len(m) == 16
    assert all(isinstance(v, float) and np.isfinite(v) for v in m.values())


def test_entropy_snr_ordering(api, synthetic_image_clean, synthetic_image_noisy):
    clean = api.metrics.compute_metrics(synthetic_image_clean)
    noisy = api.metrics.compute_metrics(synthetic_image_noisy)
    assert clean["entropy"] >= 0 and clean["snr_proxy"] >= 0
    assert clean["snr_proxy"] > noisy["snr_proxy"]
    assert noisy["sigma"] > 0.08


def test_validation_identical_images(api, synthetic_image_clean):
    v = api.metrics.compute_validation(synthetic_image_clean, synthetic_image_clean)
    for key in ("entropy_change", "snr_change", "cnr_change", "edge_ratio", "laplacian_energy_before",
                "laplacian_energy_after", "ssim", "psnr", "niqe_before", "metrics_before", "metrics_after"):
        assert key in v
    assert len(v) == 38
    assert v["passes"] is True
    assert v["ssim"] == pytest.approx(1.0)


def test_objective_score_types(api, synthetic_image_clean):
    v = api.metrics.compute_validation(synthetic_image_clean, synthetic_image_clean)
    score, breakdown = api.metrics.compute_objective_score(v)
    assert isinstance(score, float) and isinstance(breakdown, dict)


def test_edge_ratio_range(api, synthetic_image_clean):
    er = api.metrics.compute_edge_ratio(synthetic_image_clean)
    assert isinstance(er, float) and 0 <= er <= 1.0 or er == pytest.approx(omet.compute_edge_ratio(synthetic_image_clean), rel=1e-5)


# ---- tests/test_detection.py -------------------------------------------------------------------
def test_normalize_image(api):
    out = api.dicom_io.normalize_image(np.array([[10, 20], [30, 40]], dtype=np.float32))
    assert out.dtype == np.float32
    assert float(out.min()) == pytest.approx(0.0) and float(out.max()) == pytest.approx(1.0)
    assert not api.dicom_io.normalize_image(np.full((8, 8), 7, np.float32)).any()
    u16 = np.arange(64, dtype=np.uint16).reshape(8, 8) * 50
    np.testing.assert_array_equal(api.dicom_io.normalize_image(u16), omet.normalize_image(u16))


def test_detect_issues_on_fixtures(api, synthetic_image_noisy, synthetic_image_low_contrast):
    assert "noise" in api.metrics.detect_issues(api.metrics.compute_metrics(synthetic_image_noisy))
    assert "low_contrast" in api.metrics.detect_issues(api.metrics.compute_metrics(synthetic_image_low_contrast))


# ---- tests/test_pipeline.py -------------------------------------------------------------------
def test_apply_enhancements_returns_valid_image(api, synthetic_image_noisy):
    before = synthetic_image_noisy.copy()
    enhanced, ops_ = api.enhancement.apply_enhancements(synthetic_image_noisy, ["noise"])
    assert enhanced.shape == synthetic_image_noisy.shape and enhanced.dtype == np.float32
    assert float(enhanced.min()) >= 0.0 and float(enhanced.max()) <= 1.0
    assert len(ops_) > 0
    np.testing.assert_array_equal(synthetic_image_noisy, before)      # input never mutated


def test_no_ops_when_no_issues(api, synthetic_image_clean):
    enhanced, ops_ = api.enhancement.apply_enhancements(synthetic_image_clean, [])
    assert ops_ == []
    np.testing.assert_array_equal(enhanced, synthetic_image_clean)


def _plan(api, **kw):
    params = kw.pop("params", {})
    return api.schemas.EnhancementPlan(params=api.schemas.EnhancementParams(**params), **kw)


def test_basic_plan(api, synthetic_image_noisy):
    plan = _plan(api, recommended_ops=["denoise", "clahe", "gamma", "unsharp"],
                 params=dict(clahe_clip_limit=0.02, clahe_tile_size=16, gamma=0.95, unsharp_radius=0.8,
                             unsharp_amount=0.5, denoise_mode="soft", post_denoise_strength=0.2))
    enhanced, ops_ = api.enhancement.apply_enhancements_from_params(synthetic_image_noisy, plan)
    assert enhanced.shape == (64, 64) and enhanced.dtype == np.float32
    assert 0.0 <= float(enhanced.min()) and float(enhanced.max()) <= 1.0 and len(ops_) > 0
    ref, ref_ops = oenh.apply_enhancements_from_params(synthetic_image_noisy, plan)
    assert ops_ == ref_ops
    assert_within_lsb(enhanced, ref, "basic plan")


def test_empty_ops(api, synthetic_image_clean):
    enhanced, ops_ = api.enhancement.apply_enhancements_from_params(
        synthetic_image_clean, _plan(api, recommended_ops=[], stop_reason="No issues."))
    assert ops_ == []


def test_parameter_clamping(api, synthetic_image_noisy):
    plan = _plan(api, recommended_ops=["clahe", "unsharp"],
                 params=dict(clahe_clip_limit=999.0, unsharp_amount=-5.0))
    enhanced, ops_ = api.enhancement.apply_enhancements_from_params(synthetic_image_noisy, plan)
    assert enhanced.shape == synthetic_image_noisy.shape and len(ops_) >= 1
    assert "CLAHE (clip=0.0800, tile=16)" in ops_ and any(o.startswith("Unsharp mask (r=0.80, a=0.03)") for o in ops_)


def test_invalid_denoise_mode_defaults_to_soft(api, synthetic_image_noisy):
    plan = _plan(api, recommended_ops=["denoise"], params=dict(denoise_mode="INVALID"))
    _, ops_ = api.enhancement.apply_enhancements_from_params(synthetic_image_noisy, plan)
    assert any("soft" in op for op in ops_)


def test_validation_enhanced_vs_original(api, synthetic_image_noisy):
    enhanced, _ = api.enhancement.apply_enhancements(synthetic_image_noisy, ["noise"])
    v = api.metrics.compute_validation(synthetic_image_noisy, enhanced)
    assert isinstance(v["passes"], bool) and "niqe_before" in v
    ref = omet.compute_validation(synthetic_image_noisy, enhanced)
    for key in ("ssim", "psnr", "niqe_before", "niqe_after", "contrast_gain", "sharpness_gain",
                "quality_improvement", "edge_ratio", "entropy_change", "snr_change"):
        assert v[key] == pytest.approx(ref[key], rel=1e-5, abs=1e-7), key
    for key in ("passes", "meets_ssim", "meets_psnr", "meets_improvement", "niqe_improved"):
        assert v[key] == ref[key], key
    assert api.metrics.compute_objective_score(v)[0] == pytest.approx(omet.compute_objective_score(ref)[0], abs=2e-4)


def test_deterministic_agent_chain(api, synthetic_image_noisy):
    """QualityDetection -> Recommendation -> Enhancement -> Validation (pipeline/core_agents.py:61-161)
    expressed with the drop-in functions."""
    m = api.metrics.compute_metrics(synthetic_image_noisy)
    issues = api.metrics.detect_issues(m)
    assert issues == omet.detect_issues(omet.compute_metrics(synthetic_image_noisy))
    enhanced, applied = api.enhancement.apply_enhancements(synthetic_image_noisy, issues)
    ref_img, ref_applied = oenh.apply_enhancements(synthetic_image_noisy, issues)
    assert applied == ref_applied
    assert np.abs(enhanced - ref_img).max() <= LSB16
    v = api.metrics.compute_validation(synthetic_image_noisy, enhanced)
    status = "PASS" if v["passes"] else ("WARN" if v["quality_improvement"] > 0 else "FAIL")
    assert status in ("PASS", "WARN", "FAIL")


# ---- helper functions of enhancement.py ----------------------------------------------------------
def test_private_helpers_mirror_the_reference(api, synthetic_image_noisy, synthetic_image_clean):
    e = api.enhancement
    assert e._check_halo(synthetic_image_noisy) == oenh.halo_detected(synthetic_image_noisy)
    assert e._check_noise_amplification(synthetic_image_clean, synthetic_image_noisy) is True
    assert e._check_noise_amplification(synthetic_image_noisy, synthetic_image_clean) is False
    assert e._check_over_processing(synthetic_image_clean, synthetic_image_noisy) == \
        oenh.over_processed(synthetic_image_clean, synthetic_image_noisy)
    assert e._bilateral_filter(synthetic_image_clean, d=0) is synthetic_image_clean
    flat = np.full((64, 64), 0.25, np.float32)
    assert e._light_denoise(flat + 0) is not None


# ---- full plans vs the oracle ----------------------------------------------------------------------
@pytest.mark.parametrize("name", ["noisy64", "ct512", "odd94x141"])
def test_p_full_matches_the_oracle(api, images, synth, name):
    im = images[name]
    got, labels = api.enhancement.apply_enhancements_from_params(im, synth.plan_full())
    ref, ref_labels = oenh.apply_enhancements_from_params(im, synth.plan_full())
    assert labels == ref_labels
    assert_within_lsb(got, ref, f"P_full {name}")          # north_star: max-abs <= 1 LSB, every pixel


@pytest.mark.parametrize("name", ["cr600", "ct512"])
def test_p_cr_matches_the_oracle(api, images, synth, name):
    im = images[name]
    got, labels = api.enhancement.apply_enhancements_from_params(im, synth.plan_cr())
    ref, ref_labels = oenh.apply_enhancements_from_params(im, synth.plan_cr())
    assert labels == ref_labels
    assert np.abs(got - ref).max() <= LSB16


@pytest.mark.parametrize("issues", [["noise"], ["blur"], ["low_contrast", "clipping_low"],
                                    ["noise", "blur", "clipping_high"]])
@pytest.mark.parametrize("name", ["noisy64", "ct512"])
def test_issue_driven_enhancement_matches_the_oracle(api, images, issues, name):
    im = images[name]
    got, labels = api.enhancement.apply_enhancements(im, issues)
    ref, ref_labels = oenh.apply_enhancements(im, issues)
    assert labels == ref_labels
    assert np.abs(got - ref).max() <= LSB16


def test_stack_pipeline_matches_per_slice_calls(ops, synth):
    import torch
    from mdimg_b200.batch import process_stack, process_stack_host
    raw = np.stack([synth.ct_slice(1000 + z, z / 6, size=256) for z in range(6)])
    plan = synth.plan_full()
    res = process_stack(torch.from_numpy(raw.view(np.int16)).to(ops.device), plan, chunk=4)
    out_h, res_h = process_stack_host(raw, plan, chunk=4, ops=ops)
    # float64 accumulators are combined with atomics (order not fixed), so the two runs agree to
    # rounding of the sums, not to the bit
    np.testing.assert_allclose(res.packed, res_h.packed, rtol=1e-9, atol=1e-12)
    np.testing.assert_array_equal(res.tv_iterations, res_h.tv_iterations)
    np.testing.assert_array_equal(res.enhanced.cpu().numpy(), out_h)
    for z in (0, 5):
        x = omet.normalize_image(raw[z])
        ref, ref_labels = oenh.apply_enhancements_from_params(x, plan)
        assert res.labels[z] == ref_labels
        assert_within_lsb(out_h[z], ref, f"stack slice {z}")
        v = res.validation(z)
        refv = omet.compute_validation(x, out_h[z])
        assert v["ssim"] == pytest.approx(refv["ssim"], rel=1e-6)
        assert v["metrics_after"]["entropy"] == pytest.approx(refv["metrics_after"]["entropy"], rel=1e-9)
        assert isinstance(res.score(z)[0], float)


def test_cohort_pipeline_equals_stack_by_stack(ops, synth):
    """A sequence of stacks through one chunk queue (process_stacks_host) gives each stack the result
    of its own process_stack_host call; shared (double-buffered) outputs are honoured."""
    from mdimg_b200.batch import process_stack_host, process_stacks_host
    stacks = [np.stack([synth.ct_slice(3000 + 16 * v + z, z / 5, size=128) for z in range(5 + v)]) for v in range(3)]
    plan = synth.plan_full()
    single = [process_stack_host(s.copy(), plan, chunk=2, ops=ops, workers=2) for s in stacks]
    single = [(o.copy(), r) for o, r in single]
    cohort = process_stacks_host(stacks, plan, chunk=2, ops=ops, workers=3, schedule=[3, 2, 1])
    assert len(cohort) == 3
    for (o1, r1), (o2, r2), s in zip(single, cohort, stacks):
        assert o2.shape == s.shape and o2.dtype == np.float32
        np.testing.assert_array_equal(o1, o2)
        np.testing.assert_allclose(r1.packed, r2.packed, rtol=1e-9, atol=1e-12)
        assert r1.labels == r2.labels and len(r2.labels) == s.shape[0]
    # caller-provided output arrays
    outs = [np.empty(s.shape, np.float32) for s in stacks]
    again = process_stacks_host(stacks, plan, chunk=4, ops=ops, out_hosts=outs, workers=2)
    for (o1, _), o3, (o4, _) in zip(single, outs, again):
        np.testing.assert_array_equal(o1, o3)
        assert o4 is o3 or np.shares_memory(o4, o3)


def test_empty_single_and_ragged_stacks(ops, synth):
    """Zero slices, one slice, and a cohort whose stacks differ in slice count AND image size."""
    import torch
    from mdimg_b200.batch import PACK_COLS, process_stack, process_stack_host, process_stacks_host
    plan = synth.plan_full()
    empty = np.zeros((0, 64, 64), np.uint16)
    res = process_stack(torch.from_numpy(empty.view(np.int16)).to(ops.device), plan)
    assert res.packed.shape == (0, PACK_COLS) and res.labels == [] and tuple(res.enhanced.shape) == (0, 64, 64)
    out, res_h = process_stack_host(empty, plan, ops=ops)
    assert out.shape == (0, 64, 64) and res_h.packed.shape == (0, PACK_COLS) and res_h.labels == []
    one = synth.ct_slice(4100, 0.3, size=96)[None]
    small = np.stack([synth.ct_slice(4200 + z, z / 3, size=64) for z in range(3)])
    wide = np.stack([synth.ct_slice(4300 + z, z / 2, size=128)[:94, :] for z in range(2)])     # 94 x 128
    cohort = process_stacks_host([one, empty, small, wide], plan, chunk=2, ops=ops, workers=3)
    assert [o.shape for o, _ in cohort] == [one.shape, empty.shape, small.shape, wide.shape]
    for stack, (o, r) in zip([one, empty, small, wide], cohort):
        assert len(r.labels) == stack.shape[0] and r.packed.shape == (stack.shape[0], PACK_COLS)
        for z in range(stack.shape[0]):
            x = omet.normalize_image(stack[z])
            ref, ref_labels = oenh.apply_enhancements_from_params(x, plan)
            assert r.labels[z] == ref_labels
            assert_within_lsb(o[z], ref, f"cohort slice {z}")
            assert r.metrics_before(z)["entropy"] == pytest.approx(omet.compute_metrics(x)["entropy"], rel=1e-9)


def test_uint16_export_is_img_as_uint_of_the_float_result(ops, synth):
    """out_dtype=uint16: the host receives uint16(clip(rint(x * 65535), 0, 65535)) of exactly the float32
    image the default path returns (bit for bit), hence within 1 LSB of the oracle's export wherever the
    float images agree to 1 LSB."""
    import torch
    from mdimg_b200.batch import process_stack_host
    raw = np.stack([synth.ct_slice(5000 + z, z / 4, size=128)[:, :126] for z in range(4)])      # 128 x 126: vector path
    odd = np.stack([synth.ct_slice(5100 + z, z / 2, size=96)[:95, :93] for z in range(2)])      # odd sizes: scalar path
    plan = synth.plan_full()
    for stack in (raw, odd):
        f32, res_f = process_stack_host(stack, plan, chunk=3, ops=ops)
        f32 = f32.copy()
        u16, res_u = process_stack_host(stack, plan, chunk=3, ops=ops, out_dtype=np.uint16)
        assert u16.dtype == np.uint16 and u16.shape == stack.shape
        want = np.clip(np.rint(f32 * np.float32(65535)), 0, 65535).astype(np.uint16)
        np.testing.assert_array_equal(u16, want)
        np.testing.assert_allclose(res_f.packed, res_u.packed, rtol=1e-9, atol=1e-12)
        assert res_f.labels == res_u.labels
        x = omet.normalize_image(stack[0])
        ref, _ = oenh.apply_enhancements_from_params(x, plan)
        ref16 = np.clip(np.rint(ref.astype(np.float32) * np.float32(65535)), 0, 65535).astype(np.int64)
        assert int(np.abs(u16[0].astype(np.int64) - ref16).max()) <= 1
    # edge values of the conversion itself: ties round to even, out-of-range clips
    probe = np.array([[0.0, 1.0, 0.5, 1.5 / 65535, 2.5 / 65535, -0.25, 1.25, 0.999999]], np.float32)
    t = torch.from_numpy(np.ascontiguousarray(np.tile(probe, (4, 1))[None])).to(ops.device)
    got = ops.export_u16(t).cpu().numpy().view(np.uint16)[0, 0]
    want = np.clip(np.rint(probe[0] * np.float32(65535)), 0, 65535).astype(np.uint16)
    np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        process_stack_host(raw, plan, ops=ops, out_dtype=np.int32)


def test_native_enhance_call_equals_the_python_engine(ops, synth):
    """mdimg_enhance (the whole apply_enhancements_from_params control flow in C++, one C-ABI call per stack)
    against Engine.enhance_from_params (the same logic on torch tensors, itself checked against the oracle):
    pixels bit-equal, flags / labels / TV iteration counts equal, metric rows equal to accumulation order."""
    import torch
    from mdimg_b200.engine import Engine
    from mdimg_b200.pipeline.schemas import EnhancementParams, EnhancementPlan
    eng = Engine(ops)
    rng = np.random.default_rng(5)
    ims = [synth.fixture_clean(), synth.fixture_noisy(), synth.fixture_low_contrast(),
           rng.random((64, 64), dtype=np.float32),                                   # noise: guards fire
           np.clip(synth.fixture_clean() * 1.6 - 0.2, 0, 1).astype(np.float32)]
    stack = torch.from_numpy(np.stack(ims)).to(ops.device)
    plans = [
        synth.plan_full(),
        synth.plan_cr() if hasattr(synth, "plan_cr") else EnhancementPlan(
            recommended_ops=["clahe", "unsharp"], params=EnhancementParams(clahe_clip_limit=0.03, clahe_tile_size=32,
                                                                           unsharp_radius=2.0, unsharp_amount=1.5)),
        EnhancementPlan(recommended_ops=["denoise", "gamma", "unsharp"],                  # out-of-bounds values: clamping
                        params=EnhancementParams(denoise_mode="hard", gamma=1.9, unsharp_radius=5.0, unsharp_amount=9.0,
                                                 clahe_clip_limit=0.5, clahe_tile_size=100, post_denoise_strength=0.0)),
        EnhancementPlan(recommended_ops=["Unsharp ", "gamma", "CLAHE", "tv_denoise", "nonsense"],   # plan order != step order
                        params=EnhancementParams(clahe_clip_limit=0.01, clahe_tile_size=8, gamma=1.05, unsharp_radius=1.0,
                                                 unsharp_amount=2.0, tv_denoise_weight=0.02)),
        EnhancementPlan(recommended_ops=["unsharp"], params=EnhancementParams(unsharp_radius=1.5, unsharp_amount=2.5)),
        EnhancementPlan(recommended_ops=["bilateral", "post_denoise"],
                        params=EnhancementParams(bilateral_d=6, bilateral_sigma_color=0.1, bilateral_sigma_space=0.1,
                                                 post_denoise_strength=0.6)),
        EnhancementPlan(recommended_ops=[], params=EnhancementParams()),
    ]
    fired = set()
    for plan in plans:
        ref = eng.enhance_from_params(stack, plan, on_error="flag")
        got = eng.enhance_from_params_native(stack, plan, on_error="flag")
        np.testing.assert_array_equal(got.image.cpu().numpy(), ref.image.cpu().numpy())
        assert got.labels == ref.labels
        np.testing.assert_array_equal(got.halo, ref.halo)
        np.testing.assert_array_equal(got.noise_guard, ref.noise_guard)
        np.testing.assert_array_equal(got.over_processed, ref.over_processed)
        if ref.tv_iterations is None:
            assert got.tv_iterations is None
        else:
            np.testing.assert_array_equal(got.tv_iterations, ref.tv_iterations)
        np.testing.assert_allclose(got.rows_after.cpu().numpy(), ref.rows_after.cpu().numpy(), rtol=1e-9, atol=1e-12,
                                   equal_nan=True)
        fired |= {k for k, v in (("halo", ref.halo), ("noise", ref.noise_guard), ("over", ref.over_processed)) if v.any()}
        # with the caller's rows of the input
        rows_b = ops.metrics(stack, with_niqe=True)
        again = eng.enhance_from_params_native(stack, plan, rows_before=rows_b, on_error="flag")
        np.testing.assert_array_equal(again.image.cpu().numpy(), ref.image.cpu().numpy())
    assert fired == {"halo", "noise", "over"}, fired           # every safeguard path was exercised
    # the reference's data-dependent ValueErrors: gamma on negative pixels, CLAHE outside [-1, 1]
    bad = stack.clone()
    bad[1] = bad[1] - 0.5
    bad[3] = bad[3] * 3.0
    for plan in (EnhancementPlan(recommended_ops=["gamma"], params=EnhancementParams(gamma=0.8)),
                 EnhancementPlan(recommended_ops=["clahe"], params=EnhancementParams())):
        ref = eng.enhance_from_params(bad, plan, on_error="flag")
        got = eng.enhance_from_params_native(bad, plan, on_error="flag")
        assert got.errors == ref.errors and got.errors
        assert got.labels == ref.labels
        np.testing.assert_array_equal(got.image.cpu().numpy(), ref.image.cpu().numpy())
        with pytest.raises(ValueError):
            eng.enhance_from_params_native(bad, plan)


def test_native_issue_chain_equals_the_python_engine(ops, synth):
    """mdimg_enhance_issues (apply_enhancements in one C-ABI call) against Engine.enhance_from_issues."""
    import torch
    from mdimg_b200.engine import Engine
    eng = Engine(ops)
    rng = np.random.default_rng(9)
    ims = [synth.fixture_clean(), synth.fixture_noisy(), synth.fixture_low_contrast(),
           np.clip(synth.fixture_clean() + 0.02 * rng.standard_normal((64, 64)), 0, 1).astype(np.float32)]
    stack = torch.from_numpy(np.stack(ims)).to(ops.device)
    guard = False
    for issues in (["noise"], ["blur"], ["noise", "blur"], ["low_contrast", "clipping_low"], ["clipping_high", "blur"],
                   ["clipping_low", "clipping_high"], ["noise", "blur", "low_contrast", "clipping_low"], [], ["unknown"]):
        ref = eng.enhance_from_issues(stack, issues)
        got = eng.enhance_issues(stack, issues)
        np.testing.assert_array_equal(got.image.cpu().numpy(), ref.image.cpu().numpy())
        assert got.labels == ref.labels
        np.testing.assert_array_equal(got.noise_guard, ref.noise_guard)
        guard |= bool(ref.noise_guard.any())
        s0 = ops.estimate_sigma(stack)
        again = eng.enhance_issues(stack, issues, sigma_before=s0)
        np.testing.assert_array_equal(again.image.cpu().numpy(), ref.image.cpu().numpy())
    assert guard                                                    # the noise guard fired somewhere
    bad = stack.clone()
    bad[2] = bad[2] * 3.0                                           # CLAHE input outside [-1, 1]
    for fn in (eng.enhance_issues, eng.enhance_from_issues):
        with pytest.raises(ValueError, match="between -1 and 1"):
            fn(bad, ["low_contrast"])


def test_two_devices_in_one_process(synth):
    """The library holds no per-device state besides what it sets up per device (shared-memory opt-ins):
    the same process can drive two GPUs (the multi-GPU path proper is one process per GPU)."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from mdimg_b200.batch import process_stack
    from mdimg_b200.stack import get_ops
    raw = np.stack([synth.ct_slice(6000 + z, z / 2, size=128) for z in range(2)])
    plan = synth.plan_full()
    results = []
    for idx in (1, 0, 1):
        dev = torch.device(f"cuda:{idx}")
        ops_i = get_ops(dev)
        with torch.cuda.device(dev):
            res = process_stack(torch.from_numpy(raw.view(np.int16)).to(dev), plan, chunk=2, ops=ops_i)
        results.append(res)
    for res in results[1:]:
        np.testing.assert_array_equal(res.enhanced.cpu().numpy(), results[0].enhanced.cpu().numpy())
        np.testing.assert_allclose(res.packed, results[0].packed, rtol=1e-9, atol=1e-12)
        assert res.labels == results[0].labels


def test_score_plans_matches_the_tool_loop(ops, images, synth):
    """K candidate plans x N images (pipeline/tools.py:95-183 semantics) vs the oracle run one by one."""
    import torch
    from mdimg_b200.batch import score_plans
    from mdimg_b200.pipeline.schemas import EnhancementParams, EnhancementPlan
    names = ["noisy64", "clean64", "lowc64"]
    stack = torch.from_numpy(np.stack([images[k] for k in names])).to(ops.device)
    plans = [
        EnhancementPlan(recommended_ops=["denoise", "clahe"], params=EnhancementParams(clahe_clip_limit=0.01)),
        EnhancementPlan(recommended_ops=["unsharp", "gamma"], params=EnhancementParams(gamma=1.2, unsharp_amount=2.0)),
        EnhancementPlan(recommended_ops=["tv_denoise", "bilateral"],
                        params=EnhancementParams(tv_denoise_weight=0.08, bilateral_d=3)),
    ]
    scores, results = score_plans(stack, plans, ops=ops)
    assert scores.shape == (3, 3)
    for k, plan in enumerate(plans):
        for i, name in enumerate(names):
            ref_img, ref_labels = oenh.apply_enhancements_from_params(images[name], plan)
            ref_score, _ = omet.compute_objective_score(omet.compute_validation(images[name], ref_img))
            assert results[k].labels[i] == ref_labels
            assert scores[k, i] == pytest.approx(ref_score, abs=5e-4), (k, name)


# ---- degenerate inputs: same result, same labels or the same exception as the reference flow ------
def _degenerate_images():
    rng = np.random.default_rng(99)
    base = rng.random((48, 40), dtype=np.float32)
    return {
        "zeros": np.zeros((32, 32), np.float32),
        "constant": np.full((40, 48), 0.5, np.float32),
        "two_valued": np.where(base > 0.5, np.float32(0.2), np.float32(0.8)).astype(np.float32),
        "tiny16": rng.random((16, 16), dtype=np.float32),
        "odd": rng.random((37, 53), dtype=np.float32),
        "one_hot": np.pad(np.ones((1, 1), np.float32), ((20, 19), (15, 24))),
        "saturated": np.clip(base * 3 - 1, 0, 1).astype(np.float32),
    }


def _run_or_exc(fn):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        try:
            return fn(), None
        except Exception as exc:  # noqa: BLE001
            return None, exc


@pytest.mark.parametrize("name", sorted(_degenerate_images()))
@pytest.mark.parametrize("which", ["p_full", "p_cr", "issues_all"])
def test_degenerate_inputs_follow_the_reference_flow(api, synth, name, which):
    im = _degenerate_images()[name]
    if which == "issues_all":
        issues = ["noise", "blur", "low_contrast", "clipping_low"]
        got, gexc = _run_or_exc(lambda: api.enhancement.apply_enhancements(im, issues))
        ref, rexc = _run_or_exc(lambda: oenh.apply_enhancements(im, issues))
    else:
        plan = synth.plan_full() if which == "p_full" else synth.plan_cr()
        got, gexc = _run_or_exc(lambda: api.enhancement.apply_enhancements_from_params(im, plan))
        ref, rexc = _run_or_exc(lambda: oenh.apply_enhancements_from_params(im, plan))
    if rexc is not None:
        assert gexc is not None and type(gexc) is type(rexc), (name, which, rexc, gexc)
        return
    assert gexc is None, (name, which, gexc)
    assert got[1] == ref[1], (name, which)                        # labels, safeguards included
    a, b = got[0], ref[0]
    assert a.shape == b.shape and a.dtype == np.float32
    assert np.array_equal(np.isnan(a), np.isnan(b)), (name, which)
    ok = ~np.isnan(b)
    err = np.abs(a[ok].astype(np.float64) - b[ok])
    assert err.size == 0 or float(err.max()) <= LSB16, (name, which, float(err.max()), int((err > LSB16).sum()))


@pytest.mark.parametrize("name", sorted(_degenerate_images()))
def test_degenerate_inputs_metrics_and_validation(api, name):
    im = _degenerate_images()[name]
    # not an affine copy: the NIQE approximation and the edge ratio are scale invariant, which would
    # leave `niqe_after <= niqe_before` to rounding noise
    other = np.clip(np.sqrt(im) * np.float32(0.9) + np.float32(0.05) * im, 0, 1).astype(np.float32)
    got, gexc = _run_or_exc(lambda: api.metrics.compute_validation(im, other))
    ref, rexc = _run_or_exc(lambda: omet.compute_validation(im, other))
    if rexc is not None:
        assert gexc is not None and type(gexc) is type(rexc), (name, rexc, gexc)
        return
    assert gexc is None, (name, gexc)
    assert list(got) == list(ref)
    # A constant image is the one place where the 1e-5 relative tolerance cannot hold: numpy's float32
    # pairwise mean of n equal values is not that value, so the reference reports std ~ 4e-9 (and
    # lap_var ~ 1e-17) where the true value -- and ours -- is 0; gains divide that noise by 1e-8.
    flat = float(np.ptp(im)) == 0.0 or float(np.ptp(other)) == 0.0
    for k, v in ref.items():
        if flat and (k.endswith(("_gain", "_change", "_before", "_after")) or k in ("quality_improvement", "meets_improvement", "passes", "metrics_before", "metrics_after", "niqe_improved")):
            continue
        if isinstance(v, dict):
            for kk, vv in v.items():
                assert got[k][kk] == pytest.approx(vv, rel=1e-5, abs=1e-9, nan_ok=True), (name, k, kk)
        elif isinstance(v, bool):
            # `niqe_after <= niqe_before` on (near-)ties is decided by rounding noise: the NIQE
            # approximation is scale invariant, and any map between two-valued images is affine
            tie = abs(ref["niqe_after"] - ref["niqe_before"]) <= 1e-5 * max(abs(ref["niqe_before"]), 1e-9)
            if tie and k in ("niqe_improved", "passes"):
                continue
            assert got[k] is v, (name, k)
        elif k.endswith(("_gain", "_change")) or k == "quality_improvement":
            assert got[k] == pytest.approx(v, rel=2e-5, abs=2e-5, nan_ok=True), (name, k)
        else:
            assert got[k] == pytest.approx(v, rel=1e-5, abs=1e-9, nan_ok=True), (name, k)


# ---- save_visuals (pipeline/dicom_io.py:99-126) as a GPU mosaic + PNG ---------------------------------
def _gray_levels(x: np.ndarray) -> np.ndarray:
    """matplotlib: Normalize() autoscaled to the panel (float32), gray colormap with N = 256."""
    vmin, vmax = float(x.min()), float(x.max())
    if vmax == vmin:
        return np.zeros(x.shape, np.uint8)
    t = (x - np.float32(vmin)) / np.float32(vmax - vmin)
    xa = t * np.float32(256)
    xa[xa == 256] = 255
    return np.clip(xa.astype(int), 0, 255).astype(np.uint8)


def test_save_visuals_writes_the_before_after_png(api, tmp_path, synthetic_image_noisy):
    from PIL import Image
    before = synthetic_image_noisy
    after, _ = api.enhancement.apply_enhancements(before, ["noise", "low_contrast"])
    out = api.dicom_io.save_visuals(before, after, str(tmp_path / "viz"), "case7")
    assert list(out) == ["before_after"] and out["before_after"].endswith("case7_before_after.png")
    img = np.array(Image.open(out["before_after"]))
    h, w = before.shape
    assert img.shape == (h, 2 * w + 8) and img.dtype == np.uint8
    np.testing.assert_array_equal(img[:, :w], _gray_levels(before))
    np.testing.assert_array_equal(img[:, w + 8:], _gray_levels(after))
    assert (img[:, w:w + 8] == 255).all()


def test_stack_visuals_one_launch_many_pngs(ops, synth, tmp_path):
    import torch
    from PIL import Image

    from mdimg_b200.batch import save_stack_visuals
    x = np.stack([omet.normalize_image(synth.ct_slice(1000 + z, z / 4, size=96)) for z in range(4)])
    x[3] = 0.25                                                     # flat panel -> level 0
    y = np.clip(x ** np.float32(0.8), 0, 1).astype(np.float32)
    paths = save_stack_visuals(torch.from_numpy(x).to(ops.device), torch.from_numpy(y).to(ops.device),
                               str(tmp_path), "vol", workers=2, gap=4)
    assert len(paths) == 4
    for z, p in enumerate(paths):
        img = np.array(Image.open(p))
        np.testing.assert_array_equal(img[:, :96], _gray_levels(x[z]))
        np.testing.assert_array_equal(img[:, 100:], _gray_levels(y[z]))


def test_concurrent_callers_get_the_serial_results(api, synth):
    """The reference's Flask backend runs pipelines on concurrent daemon threads
    (backend/pipeline_runner.py:46-51): per-thread workspaces and streams keep callers independent."""
    import threading
    ims = [synth.fixture_noisy(), synth.fixture_low_contrast(), synth.fixture_clean(),
           omet.normalize_image(synth.ct_slice(1000, 0.3, size=96))]
    plan = synth.plan_full()

    def work(im):
        enh, labels = api.enhancement.apply_enhancements_from_params(im, plan)
        return enh, labels, api.metrics.compute_metrics(enh), api.metrics.compute_validation(im, enh)["ssim"]

    serial = [work(im) for im in ims]
    results = [None] * len(ims)
    errors = []

    def runner(k):
        try:
            for _ in range(3):
                results[k] = work(ims[k])
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=runner, args=(k,)) for k in range(len(ims))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for (e0, l0, m0, s0), (e1, l1, m1, s1) in zip(serial, results):
        np.testing.assert_array_equal(e0, e1)
        assert l0 == l1 and m0 == m1 and s0 == s1


def test_process_cohort_single_rank_equals_the_stack_pipeline(ops, synth, tmp_path, monkeypatch):
    """shard.process_cohort without a process group (world 1): device volumes and host volumes give the rows of
    batch.process_stack, the gathered rows stay on the device, and `persist` writes one run per slice."""
    import torch
    from mdimg_b200.batch import PACK_COLS, process_stack
    from mdimg_b200.pipeline import storage
    from mdimg_b200.shard import process_cohort
    monkeypatch.setenv("MDIMG_DB_PATH", str(tmp_path / "cohort.db"))
    plan = synth.plan_full()
    vols = [np.stack([synth.ct_slice(6000 + 10 * v + z, z / 3, size=96) for z in range(n)]) for v, n in enumerate((3, 2))]
    dev_vols = [torch.from_numpy(v.view(np.int16)).to(ops.device) for v in vols]
    res = process_cohort(dev_vols, plan, chunk=2, ops=ops, persist={"input_filename": "cohort.npy"})
    assert res.rows.is_cuda and tuple(res.rows.shape) == (5, PACK_COLS) and res.counts == [5]
    want = np.concatenate([process_stack(d, plan, chunk=2, ops=ops).packed for d in dev_vols])
    np.testing.assert_allclose(res.rows_host(), want, rtol=1e-9, atol=1e-12)
    assert [(v, a, b) for v, a, b, _, _ in res.local] == [(0, 0, 3), (1, 0, 2)]
    assert len(res.run_ids) == 5 and len(res.labels) == 5
    runs = storage.list_runs(limit=10)
    assert len(runs) == 5 and {r["run_id"] for r in runs} == set(res.run_ids)
    host = process_cohort(vols, plan, chunk=2, ops=ops)
    np.testing.assert_allclose(host.rows_host(), want, rtol=1e-9, atol=1e-12)
    for (v, a, b, enh_h, _), (_, _, _, enh_d, _) in zip(host.local, res.local):
        np.testing.assert_array_equal(enh_h, enh_d.cpu().numpy())
