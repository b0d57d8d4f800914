"""The CUDA drop-in against vectors produced by the REFERENCE'S OWN SOURCE
(tests/golden/make_reference_glue.py: pipeline/metrics.py and pipeline/enhancement.py of the
reference executed in the build container, scikit-image leaves supplied by the oracle).
The fixtures travel with the repository; nothing here reads /root/reference."""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
LSB16 = 1.0 / 65535
REL = 1.0e-5          # north_star tolerance on float metrics and scores
IMAGES = ["clean64", "noisy64", "lowc64", "ct96"]


@pytest.fixture(scope="module")
def glue():
    return json.loads((GOLDEN / "reference_glue.json").read_text()), np.load(GOLDEN / "reference_glue.npz")


@pytest.fixture(scope="module")
def api(ops):
    from mdimg_b200.pipeline import dicom_io, enhancement, metrics, schemas
    return type("Api", (), {"metrics": metrics, "enhancement": enhancement, "schemas": schemas, "dicom_io": dicom_io})


@pytest.fixture(scope="module")
def images(synth, api):
    ims = {"clean64": synth.fixture_clean(), "noisy64": synth.fixture_noisy(), "lowc64": synth.fixture_low_contrast()}
    ims["ct96"] = api.dicom_io.normalize_image(synth.ct_slice(1000, 0.25, size=96))
    return ims


def close(a, b, rel=REL):
    return abs(a - b) <= rel * max(abs(b), 1e-6) or (np.isnan(a) and np.isnan(b)) or a == b


def close_derived(key, a, b):
    """Gains / changes are differences of two nearly equal metrics divided by one of them: the
    1e-5 relative tolerance applies to the operands, i.e. it is an ABSOLUTE 2e-5 on ratios of O(1)
    quantities (sharpness_gain divides by lap_var ~ 1e-3..1e-2, hence relative there)."""
    if key.endswith(("_gain", "_change")) or key in ("quality_improvement",):
        return abs(a - b) <= 2e-5 * max(1.0, abs(b))
    return close(a, b)


@pytest.mark.parametrize("name", IMAGES)
def test_metrics_issues_niqe_edge_ratio(api, glue, images, name):
    g, _ = glue
    im = images[name]
    m = api.metrics.compute_metrics(im)
    assert list(m) == list(g["metrics"][name])
    for k, v in g["metrics"][name].items():
        assert close(m[k], v), (k, m[k], v)
    assert api.metrics.detect_issues(m) == g["issues"][name]
    assert close(api.metrics.compute_niqe_approximation(im), g["niqe"][name])
    assert close(api.metrics.compute_edge_ratio(im), g["edge_ratio"][name])


@pytest.mark.parametrize("name", IMAGES)
def test_issue_driven_enhancement(api, glue, images, name):
    g, arrs = glue
    im = images[name]
    for key in [k for k in g["from_issues"] if k.startswith(name + "|")]:
        issues = [s for s in key.split("|", 1)[1].split(",") if s]
        before = im.copy()
        out, labels = api.enhancement.apply_enhancements(im, issues)
        np.testing.assert_array_equal(im, before)                  # input never mutated
        assert labels == g["from_issues"][key], key
        assert out.dtype == np.float32 and out.shape == im.shape
        assert np.abs(out - arrs[f"issues|{key}"]).max() <= LSB16, key


@pytest.mark.parametrize("name", IMAGES)
def test_plans_validation_and_scores(api, glue, images, name):
    g, arrs = glue
    im = images[name]
    for pname, defn in g["plan_defs"].items():
        key = f"{name}|{pname}"
        plan = api.schemas.EnhancementPlan(recommended_ops=defn["recommended_ops"],
                                           params=api.schemas.EnhancementParams(**defn["params"]))
        want = g["plans"][key]
        if isinstance(want, dict):                                 # the reference raised ValueError here
            with pytest.raises(ValueError) as ei:
                api.enhancement.apply_enhancements_from_params(im, plan)
            assert f"ValueError: {ei.value}" == want["error"]
            continue
        out, labels = api.enhancement.apply_enhancements_from_params(im, plan)
        assert labels == want, key
        ref = arrs[f"plan|{key}"]
        err = np.abs(out.astype(np.float64) - ref)
        # north_star: <= 1 LSB of a 16-bit export on every pixel
        assert float(err.max()) <= LSB16, (key, float(err.max()), int((err > LSB16).sum()))
        val = api.metrics.compute_validation(im, ref)              # same pair of images as the reference
        gv = g["validation"][key]
        assert list(val) == list(gv), key
        for k, v in gv.items():
            if isinstance(v, dict):
                for kk, vv in v.items():
                    assert close(val[k][kk], vv), (key, k, kk, val[k][kk], vv)
            elif isinstance(v, bool):
                assert val[k] is v, (key, k)
            else:
                assert close_derived(k, val[k], v), (key, k, val[k], v)
        score, breakdown = api.metrics.compute_objective_score(val)
        assert abs(score - g["score"][key]["score"]) <= 2e-4, key    # scores are rounded to 4 decimals
        assert breakdown["passes"] == g["score"][key]["breakdown"]["passes"]


def test_normalize_image_against_the_references_own_output(api):
    """The drop-in normalize_image against the reference's own function body run on 8 inputs of different
    dtypes (tests/golden/make_reference_normalize.py): bit for bit."""
    z = np.load(GOLDEN / "reference_normalize.npz")
    for name in [k[3:] for k in z.files if k.startswith("in|")]:
        got = api.dicom_io.normalize_image(z[f"in|{name}"])
        want = z[f"out|{name}"]
        assert got.dtype == np.float32 and got.shape == want.shape
        np.testing.assert_array_equal(got, want, err_msg=name)
