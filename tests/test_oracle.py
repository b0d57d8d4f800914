"""The CPU oracle against what can be pinned without scikit-image / PyWavelets: mathematical
identities, hand-computed answers, cv2 / scipy cross-checks and committed golden vectors
(tests/golden/, produced by tests/golden/make_golden.py)."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import exposure as oex
from oracle import filters as oflt
from oracle import ref_enhancement as oenh
from oracle import ref_metrics as omet
from oracle import restoration as ores
from oracle import wavelets as wv
from oracle.fullref import peak_signal_noise_ratio, structural_similarity

GOLDEN = Path(__file__).resolve().parent / "golden"


# ---------------------------------------------------------------- wavelets
@pytest.mark.parametrize("shape", [(64, 64), (47, 94), (33, 17), (128, 96)])
def test_haar_perfect_reconstruction(shape):
    x = np.random.default_rng(0).random(shape).astype(np.float64)
    lv = max(wv.haar_max_level(shape) - 3, 1)
    rec = wv.haar_waverec2(wv.haar_wavedec2(x, lv))[: shape[0], : shape[1]]
    assert np.abs(rec - x).max() < 1e-12


def test_haar_is_orthogonal_on_even_sizes():
    x = np.random.default_rng(1).random((64, 32))
    c = wv.haar_dwt2(x)
    energy = sum(float((v**2).sum()) for v in c.values())
    assert energy == pytest.approx(float((x**2).sum()), rel=1e-12)


def test_haar_hand_computed_2x2():
    x = np.array([[1.0, 2.0], [3.0, 5.0]])
    c = wv.haar_dwt2(x)
    assert c["aa"][0, 0] == pytest.approx(11.0 / 2)
    assert c["ad"][0, 0] == pytest.approx((1 - 2 + 3 - 5) / 2)     # detail along axis 1
    assert c["da"][0, 0] == pytest.approx((1 + 2 - 3 - 5) / 2)     # detail along axis 0
    assert c["dd"][0, 0] == pytest.approx((1 - 2 - 3 + 5) / 2)


def test_haar_float32_stays_float32_and_levels():
    x = np.random.default_rng(2).random((512, 512)).astype(np.float32)
    coeffs = wv.haar_wavedec2(x, max(wv.haar_max_level(x.shape) - 3, 1))
    assert len(coeffs) == 7 and coeffs[0].shape == (8, 8)
    assert all(v.dtype == np.float32 for lvl in coeffs[1:] for v in lvl.values())
    assert [wv.haar_max_level((n, n)) - 3 for n in (64, 256, 512, 1024, 3000, 4096)] == [3, 5, 6, 7, 8, 9]


def test_db2_dd_matches_direct_convolution():
    x = np.random.default_rng(3).random((40, 37))
    n0, n1 = x.shape
    hi = np.array(wv.DB2_DEC_HI)

    def ref_axis(a):   # full convolution of the symmetric extension, odd samples
        ext = np.pad(a, (3, 3), mode="symmetric")
        full = np.convolve(ext, hi)            # full[k] = sum_j hi[j] ext[k-j]
        n_out = (len(a) + 3) // 2
        return np.array([full[(2 * o + 1) + 3] for o in range(n_out)])

    t = np.apply_along_axis(ref_axis, 0, x)
    dd = np.apply_along_axis(ref_axis, 1, t)
    got = wv.dwtn_db2_dd(x)
    assert got.shape == ((n0 + 3) // 2, (n1 + 3) // 2)
    assert np.abs(got - dd).max() < 1e-12


def test_db2_annihilates_linear_ramps_inside():
    yy, xx = np.mgrid[0:64, 0:64].astype(np.float64)
    dd = wv.dwtn_db2_dd(0.3 * xx + 0.1 * yy + 2.0)
    assert np.abs(dd[2:-2, 2:-2]).max() < 1e-12


def test_estimate_sigma_recovers_gaussian_noise():
    rng = np.random.default_rng(4)
    x = (0.5 + rng.normal(0, 0.1, (256, 256))).astype(np.float32)
    assert float(ores.estimate_sigma(x)) == pytest.approx(0.1, rel=0.05)


def test_thresholds():
    d = np.array([-3.0, -0.5, 0.0, 0.5, 2.0], np.float32)
    np.testing.assert_allclose(wv.threshold_soft(d, 1.0), [-2.0, 0, 0, 0, 1.0], atol=1e-7)
    np.testing.assert_array_equal(wv.threshold_hard(d, 1.0), [-3.0, 0, 0, 0, 2.0])


def test_denoise_wavelet_reduces_noise_and_keeps_dtype(synthetic_image_noisy):
    out = ores.denoise_wavelet(synthetic_image_noisy)
    assert out.dtype == np.float32 and out.shape == (64, 64)
    assert out.std() < synthetic_image_noisy.std()


# ---------------------------------------------------------------- scipy-backed stencils vs cv2
def test_sobel_laplace_match_cv2_reflect(images):
    cv2 = pytest.importorskip("cv2")
    x = images["unit256"]
    sh = oflt.sobel_h(x)
    sv = oflt.sobel_v(x)
    lap = oflt.laplace(x)
    c_h = cv2.Sobel(x, cv2.CV_32F, 0, 1, ksize=3, borderType=cv2.BORDER_REFLECT) / 4.0
    c_v = cv2.Sobel(x, cv2.CV_32F, 1, 0, ksize=3, borderType=cv2.BORDER_REFLECT) / 4.0
    c_l = cv2.Laplacian(x, cv2.CV_32F, ksize=1, borderType=cv2.BORDER_REFLECT)
    assert np.abs(np.abs(sh) - np.abs(c_h)).max() < 1e-6
    assert np.abs(np.abs(sv) - np.abs(c_v)).max() < 1e-6
    assert np.abs(np.abs(lap) - np.abs(c_l)).max() < 2e-6


def test_gaussian_weights_match_scipy():
    for sigma in (0.2, 0.8, 2.0, 3.0):
        w = oflt.gaussian_weights(sigma)
        r = len(w) // 2
        imp = np.zeros(4 * r + 1)
        imp[2 * r] = 1.0
        ref = ndi.gaussian_filter1d(imp, sigma, truncate=4.0)
        np.testing.assert_allclose(ref[r: 3 * r + 1], w, rtol=0, atol=1e-16)


def test_unsharp_clip_range():
    x = np.random.default_rng(5).random((32, 32)).astype(np.float32)
    out = oflt.unsharp_mask(x, 2.0, 2.5)
    assert out.min() >= 0.0 and out.max() <= 1.0
    out = oflt.unsharp_mask(x - np.float32(0.5), 2.0, 2.5)
    assert out.min() >= -1.0 and out.min() < 0


# ---------------------------------------------------------------- CLAHE
def test_clip_histogram_respects_the_limit_and_never_removes_mass_below_it():
    """skimage's redistribute loop may hand out MORE than the clipped excess (its strided sweep
    stops only after n_excess went <= 0), so pixel mass is not conserved; what does hold: no bin
    ends above the limit, and bins are only ever raised towards it."""
    rng = np.random.default_rng(6)
    for _ in range(50):
        h = rng.multinomial(256, rng.dirichlet(np.full(256, 0.05))).astype(np.int64)
        clim = int(rng.integers(1, 40))
        out = oex.clip_histogram(h.copy(), clim)
        assert out.max() <= clim
        assert (out >= np.minimum(h, clim)).all()
    h = np.full(256, 1, np.int64)
    np.testing.assert_array_equal(oex.clip_histogram(h.copy(), 3), h)     # nothing to clip


def test_clahe_structure(images):
    x = images["ct512"]
    out, info = oex.equalize_adapthist(x, kernel_size=16, clip_limit=0.015, return_internals=True)
    assert out.dtype == np.float32 and out.shape == x.shape
    assert out.min() == 0.0 and out.max() == 1.0
    assert info["ns_hist"] == [32, 32] and info["clim"] == 3
    assert (info["raw_hist"].sum(axis=-1) == 256).all()
    assert info["quantised"].max() == 16383 and info["quantised"].min() == 0
    assert info["maps"].max() <= 16383
    stage = info["stage_u16"]
    np.testing.assert_array_equal(out, ((stage.astype(np.float32) - np.float32(stage.min()))
                                        / np.float32(float(stage.max()) - float(stage.min()))))


def test_clahe_odd_geometry_and_errors(images):
    x = images["odd94x141"]
    for k in (4, 7, 16, 48):
        out = oex.equalize_adapthist(x, kernel_size=k, clip_limit=0.02)
        assert out.shape == x.shape and 0.0 <= out.min() and out.max() <= 1.0
    with pytest.raises(ValueError):
        oex.equalize_adapthist(x * 3, kernel_size=16)
    with pytest.raises(ValueError):
        oex.adjust_gamma(x - 1, 0.9)


# ---------------------------------------------------------------- full-reference + TV
def test_ssim_psnr_properties(images):
    x = images["unit256"]
    y = np.clip(x + np.float32(0.05), 0, 1)
    assert structural_similarity(x, x) == pytest.approx(1.0)
    assert structural_similarity(x, y) == pytest.approx(structural_similarity(y, x), rel=1e-6)
    assert np.isinf(peak_signal_noise_ratio(x, x))
    assert peak_signal_noise_ratio(x, y) == pytest.approx(10 * np.log10(1 / np.mean((x - y).astype(np.float64) ** 2)))


def test_tv_chambolle_properties(images):
    x = images["noisy64"]
    out, it = ores.denoise_tv_chambolle(x, 0.1, return_iters=True)

    def tv(a):
        return np.abs(np.diff(a, axis=0)).sum() + np.abs(np.diff(a, axis=1)).sum()
    assert 1 <= it <= 200 and out.dtype == np.float32
    assert tv(out) < tv(x)
    assert out.mean() == pytest.approx(x.mean(), abs=1e-4)


# ---------------------------------------------------------------- golden vectors
def _golden():
    return json.loads((GOLDEN / "oracle_fixtures.json").read_text())


@pytest.mark.parametrize("name", ["clean64", "noisy64", "lowc64"])
def test_oracle_metrics_match_golden(images, name):
    g = _golden()["metrics"][name]
    m = omet.compute_metrics(images[name])
    for k, v in g.items():
        assert m[k] == pytest.approx(v, rel=1e-6, abs=1e-9), k


@pytest.mark.parametrize("name", ["clean64", "noisy64", "lowc64"])
def test_oracle_enhancement_matches_golden(images, synth, name):
    g = _golden()["p_full"][name]
    out, labels = oenh.apply_enhancements_from_params(images[name], synth.plan_full())
    assert labels == g["labels"]
    assert float(out.astype(np.float64).sum()) == pytest.approx(g["sum"], rel=1e-5)
    ref = np.load(GOLDEN / f"p_full_{name}.npy")
    assert np.abs(out - ref).max() <= 1.0 / 65535


def test_numpy_scipy_known_answers():
    """The numpy/scipy conventions the CUDA kernels hard-code."""
    g = _golden()["conventions"]
    x = np.arange(10, dtype=np.float32)
    assert ndi.uniform_filter(x, size=7).tolist() == g["uniform7"]
    assert ndi.uniform_filter(x, size=16).tolist() == g["uniform16"]
    assert np.pad(np.arange(4), 2, mode="reflect").tolist() == [2, 1, 0, 1, 2, 3, 2, 1]
    assert np.pad(np.arange(4), 2, mode="symmetric").tolist() == [1, 0, 0, 1, 2, 3, 3, 2]
    assert float(np.percentile(np.arange(11, dtype=np.float32), 25)) == 2.5


# ----------------------------------------------------------------------------------------------
# The restated control flow against the reference's OWN SOURCE (tests/golden/make_reference_glue.py:
# the reference's pipeline/metrics.py and pipeline/enhancement.py executed with the skimage leaves
# replaced by the oracle's leaf restatements).  Everything above the leaves must agree exactly.
# ----------------------------------------------------------------------------------------------
def _glue():
    return (json.loads((GOLDEN / "reference_glue.json").read_text()), np.load(GOLDEN / "reference_glue.npz"))


def _glue_images(synth):
    ims = {"clean64": synth.fixture_clean(), "noisy64": synth.fixture_noisy(), "lowc64": synth.fixture_low_contrast()}
    ims["ct96"] = omet.normalize_image(synth.ct_slice(1000, 0.25, size=96))
    return ims


def _plan_from_json(defn):
    from mdimg_b200.pipeline.schemas import EnhancementParams, EnhancementPlan
    return EnhancementPlan(recommended_ops=defn["recommended_ops"], params=EnhancementParams(**defn["params"]))


def _same(a, b):
    if isinstance(a, float) and isinstance(b, float):
        return a == b or (np.isnan(a) and np.isnan(b))
    return a == b


GLUE_IMAGES = ["clean64", "noisy64", "lowc64", "ct96"]


@pytest.mark.parametrize("name", GLUE_IMAGES)
def test_restated_metrics_equal_reference_source(synth, name):
    g, _ = _glue()
    im = _glue_images(synth)[name]
    m = omet.compute_metrics(im)
    assert list(m) == list(g["metrics"][name])                   # same 16 keys, same order
    for k, v in g["metrics"][name].items():
        assert m[k] == v, (k, m[k], v)                           # bit for bit
    assert omet.detect_issues(m) == g["issues"][name]
    assert omet.compute_niqe_approximation(im) == g["niqe"][name]
    assert omet.compute_edge_ratio(im) == g["edge_ratio"][name]


@pytest.mark.parametrize("name", GLUE_IMAGES)
def test_restated_issue_driven_enhancement_equals_reference_source(synth, name):
    g, arrs = _glue()
    im = _glue_images(synth)[name]
    keys = [k for k in g["from_issues"] if k.startswith(name + "|")]
    assert len(keys) >= 4
    for key in keys:
        issues = [s for s in key.split("|", 1)[1].split(",") if s]
        out, labels = oenh.apply_enhancements(im, issues)
        assert labels == g["from_issues"][key], key
        np.testing.assert_array_equal(out, arrs[f"issues|{key}"], err_msg=key)


@pytest.mark.parametrize("name", GLUE_IMAGES)
def test_restated_plan_enhancement_validation_and_score_equal_reference_source(synth, name):
    g, arrs = _glue()
    im = _glue_images(synth)[name]
    for pname, defn in g["plan_defs"].items():
        key = f"{name}|{pname}"
        plan = _plan_from_json(defn)
        want = g["plans"][key]
        if isinstance(want, dict):                                # the reference raised
            with pytest.raises(ValueError) as ei:
                oenh.apply_enhancements_from_params(im, plan)
            assert f"ValueError: {ei.value}" == want["error"]
            continue
        out, labels = oenh.apply_enhancements_from_params(im, plan)
        assert labels == want, key
        np.testing.assert_array_equal(out, arrs[f"plan|{key}"], err_msg=key)
        val = omet.compute_validation(im, out)
        gv = g["validation"][key]
        assert list(val) == list(gv), key                          # 38 entries, same order
        for k, v in gv.items():
            if isinstance(v, dict):
                assert val[k] == v, (key, k)
            else:
                assert _same(val[k], v), (key, k, val[k], v)
        score, breakdown = omet.compute_objective_score(val)
        assert score == g["score"][key]["score"] and breakdown == g["score"][key]["breakdown"], key


def test_normalize_image_is_the_references_own_output():
    """oracle normalize_image against vectors produced by the reference's own function body
    (tests/golden/make_reference_normalize.py): bit for bit, every dtype."""
    from pathlib import Path
    z = np.load(Path(__file__).resolve().parent / "golden" / "reference_normalize.npz")
    names = [k[3:] for k in z.files if k.startswith("in|")]
    assert len(names) == 8
    for name in names:
        got = omet.normalize_image(z[f"in|{name}"])
        want = z[f"out|{name}"]
        assert got.dtype == np.float32 and got.shape == want.shape
        np.testing.assert_array_equal(got, want, err_msg=name)
