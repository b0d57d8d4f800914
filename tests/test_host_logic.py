"""Host-side logic of the product path (no GPU): percentile plans, filter taps, parameter
clamping, labels, validation / scoring arithmetic and the multi-rank sharding + gather (gloo,
world_size 2)."""

from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from scipy import ndimage as ndi

from mdimg_b200 import engine
from mdimg_b200.shard import gather_labels, gather_rows, slice_range
from mdimg_b200.stack import bilateral_spatial, gaussian_taps, percentile_plan
from oracle import ref_enhancement as oenh
from oracle import ref_metrics as omet


def _lerp32(a, b, t):
    a, b, t = np.float32(a), np.float32(b), np.float32(t)
    d = b - a
    r = a + d * t
    if t >= 0.5:
        r = b - d * (np.float32(1) - t)
    return r


@pytest.mark.parametrize("n", [1, 2, 7, 64 * 64, 512 * 512, 3000 * 3000, 94 * 141, 4096 * 4096])
def test_percentile_plan_reproduces_numpy_bit_for_bit(n):
    x = np.random.default_rng(n).random(n).astype(np.float32)
    xs = np.sort(x)
    lo, hi, g = percentile_plan(n)
    for k, q in enumerate((5, 25, 75, 95, 90)):
        assert _lerp32(xs[lo[k]], xs[hi[k]], g[k]) == np.percentile(x, q)


def test_gaussian_taps_are_scipys():
    for sigma in (0.2, 0.8, 1.7, 3.0):
        taps = gaussian_taps(sigma)
        r = len(taps) - 1
        assert r == int(4 * sigma + 0.5)
        imp = np.zeros(4 * r + 3)
        imp[2 * r + 1] = 1.0
        ref = ndi.gaussian_filter1d(imp, sigma, truncate=4.0)
        np.testing.assert_array_equal(ref[2 * r + 1: 3 * r + 2], taps)


def test_bilateral_spatial_matches_reference_formula():
    for d, want in ((5, 5), (4, 5), (13, 9), (9, 9), (1, 1), (2, 3)):
        deff, w = bilateral_spatial(d, 0.05)
        assert deff == want and w.shape == (want, want) and w.dtype == np.float64
        r = want // 2
        assert w[r, r] == 1.0
        assert w[0, 0] == np.exp(-(2 * r * r) / (2 * 0.05**2 * want**2))


def test_param_bounds_contract():
    from mdimg_b200.pipeline.schemas import PARAM_BOUNDS, EnhancementParams
    assert set(PARAM_BOUNDS) == {
        "clahe_clip_limit", "clahe_tile_size", "gamma", "unsharp_radius", "unsharp_amount",
        "post_denoise_strength", "bilateral_d", "bilateral_sigma_color", "bilateral_sigma_space",
        "tv_denoise_weight"}
    assert all(lo < hi for lo, hi in PARAM_BOUNDS.values())
    assert PARAM_BOUNDS == oenh.PARAM_BOUNDS
    p = EnhancementParams()
    assert (p.clahe_clip_limit, p.clahe_tile_size, p.gamma, p.unsharp_radius, p.unsharp_amount,
            p.denoise_mode, p.post_denoise_strength, p.bilateral_d, p.tv_denoise_weight) == \
        (0.015, 16, 1.0, 0.8, 0.5, "soft", 0.3, 0, 0.0)


def test_clamping_and_labels_follow_the_reference():
    wild = SimpleNamespace(clahe_clip_limit=999.0, clahe_tile_size=1000, gamma=0.1, unsharp_radius=9.0,
                           unsharp_amount=-5.0, denoise_mode="INVALID", post_denoise_strength=2.0,
                           bilateral_d=40, bilateral_sigma_color=5.0, bilateral_sigma_space=0.0,
                           tv_denoise_weight=3.0)
    q = engine.ClampedParams.from_params(wild)
    ref = oenh.clamp_plan_params(wild)
    assert (q.clip_limit, q.tile_size, q.gamma, q.u_radius, q.u_amount, q.dn_mode, q.post_str,
            q.bilateral_d, q.bilateral_sc, q.bilateral_ss, q.tv_weight) == \
        (ref["clip_limit"], ref["tile_size"], ref["gamma"], ref["u_radius"], ref["u_amount"],
         ref["dn_mode"], ref["post_str"], ref["bilateral_d"], ref["bilateral_sc"], ref["bilateral_ss"],
         ref["tv_weight"])
    assert q.dn_mode == "soft" and q.clip_limit == 0.08 and q.u_amount == 0.03 and q.bilateral_d == 13
    table = oenh._step_table(ref, ref["u_amount"])
    for name in engine._STEP_ORDER:
        assert engine.Engine._label(name, q) == table[name][2]
        assert engine.Engine._enabled(name, q) == table[name][0]


def test_validation_and_score_arithmetic_match_the_oracle(synthetic_image_noisy):
    x = synthetic_image_noisy
    y, _ = oenh.apply_enhancements(x, ["noise"])
    ref = omet.compute_validation(x, y)
    got = engine.validation_dict(ref["metrics_before"], ref["metrics_after"], ref["ssim"], ref["psnr"],
                                 ref["niqe_before"], ref["niqe_after"], ref["edge_ratio"])
    assert list(got) == list(ref)
    assert len(got) == 38 and got["passes"] is ref["passes"]
    for k in ref:
        assert got[k] == ref[k], k
    assert engine.objective_score(got) == omet.compute_objective_score(ref)
    assert engine.METRIC_KEYS == omet.METRIC_KEYS and engine.THRESHOLDS == omet.THRESHOLDS


def test_detect_issues_thresholds():
    from mdimg_b200.pipeline.metrics import detect_issues
    ok = {"sigma": 0.03, "lap_var": 0.01, "std": 0.25, "pct_low": 0.005, "pct_high": 0.005}
    bad = {"sigma": 0.15, "lap_var": 0.0005, "std": 0.05, "pct_low": 0.05, "pct_high": 0.05}
    assert detect_issues(ok) == []
    assert detect_issues(bad) == ["noise", "blur", "low_contrast", "clipping_low", "clipping_high"]
    assert detect_issues(bad) == omet.detect_issues(bad)


def test_slice_range_partitions():
    for n in (0, 1, 7, 64, 1024, 8192):
        for world in (1, 2, 3, 8):
            spans = [slice_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert slice_range(8192, 3, 8) == (3072, 4096)


def _gloo_worker(rank, world, n_total, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = slice_range(n_total, rank, world)
    rows = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.ones((1, 5), dtype=torch.float64) + rank / 10
    out = gather_rows(rows, n_total)
    labels = gather_labels([[f"slice{i}", f"rank{rank}"] for i in range(a, b)])
    q.put((rank, (out.numpy(), labels)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_rows_gloo_world2(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, n_total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    a0, b0 = slice_range(n_total, 0, 2)
    for r in range(2):
        out, labels = results[r]
        assert labels == [[f"slice{i}", f"rank{0 if i < b0 else 1}"] for i in range(n_total)]
        assert out.shape == (n_total, 5)
        want = np.arange(n_total, dtype=np.float64)[:, None] * np.ones((1, 5))
        want[b0:] += 0.1
        np.testing.assert_allclose(out, want)


def test_png_writer_roundtrips_through_pil():
    """The report mosaics are written with a 30-line PNG encoder (zlib); PIL must read them back."""
    import io

    from PIL import Image

    from mdimg_b200 import png
    rng = np.random.default_rng(3)
    for shape in [(1, 1), (37, 53), (64, 136)]:
        a = rng.integers(0, 256, shape, dtype=np.uint8)
        got = np.array(Image.open(io.BytesIO(png.encode_gray8(a))))
        assert got.dtype == np.uint8 and np.array_equal(got, a)
    with pytest.raises(ValueError):
        png.encode_gray8(np.zeros((4, 4), np.float32))


def test_tapered_schedule_covers_the_stack():
    from mdimg_b200.batch import tapered_schedule
    for n in (1, 10, 64, 1000, 1024, 2048):
        for w in (1, 2, 4, 8):
            s = tapered_schedule(n, w)
            assert sum(s) == n and all(c > 0 for c in s)
            assert s == sorted(s, reverse=True) or s[-1] >= s[-2] or len(s) < 3   # non-increasing but for the remainder
    assert tapered_schedule(1024, 4)[:4] == [160, 160, 160, 160]


# ---- host-side entry points of the library (no device work: run on the CPU-only box) -----------
def _c_plan(**kw):
    from mdimg_b200 import _lib
    q = _lib.EnhancePlan()
    vals = dict(clahe_clip_limit=0.015, clahe_tile_size=16, gamma=1.0, unsharp_radius=0.8, unsharp_amount=0.5,
                denoise_hard=0, post_denoise_strength=0.3, bilateral_d=0, bilateral_sigma_color=0.05,
                bilateral_sigma_space=0.05, tv_denoise_weight=0.0)
    vals.update(kw)
    for k, v in vals.items():
        setattr(q, k, v)
    return q


def test_plan_clamp_is_param_bounds():
    """mdimg_plan_clamp == the reference's max(lo, min(hi, v)) over PARAM_BOUNDS (pipeline/schemas.py:16-28)."""
    import ctypes as C

    from mdimg_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(11)
    names = ["clahe_clip_limit", "clahe_tile_size", "gamma", "unsharp_radius", "unsharp_amount",
             "post_denoise_strength", "bilateral_d", "bilateral_sigma_color", "bilateral_sigma_space", "tv_denoise_weight"]
    for _ in range(200):
        raw = {}
        for nm in names:
            lo, hi = engine.PARAM_BOUNDS[nm]
            v = rng.uniform(lo - (hi - lo), hi + (hi - lo))
            raw[nm] = int(round(v)) if nm in ("clahe_tile_size", "bilateral_d") else float(v)
        q = _c_plan(**raw)
        assert lib.mdimg_plan_clamp(C.byref(q)) == 0
        want = engine.ClampedParams.from_params(SimpleNamespace(denoise_mode="soft", **raw))
        assert q.clahe_clip_limit == want.clip_limit and q.clahe_tile_size == want.tile_size and q.gamma == want.gamma
        assert q.unsharp_radius == want.u_radius and q.unsharp_amount == want.u_amount
        assert q.post_denoise_strength == want.post_str and q.bilateral_d == want.bilateral_d
        assert q.bilateral_sigma_color == want.bilateral_sc and q.bilateral_sigma_space == want.bilateral_ss
        assert q.tv_denoise_weight == want.tv_weight
        before = bytes(q)
        lib.mdimg_plan_clamp(C.byref(q))
        assert bytes(q) == before                               # idempotent


def test_default_tables_follow_numpy_and_scipy():
    """mdimg_enhance_tables_default: the percentile plan equals numpy's float32 plan exactly; the Gaussian
    taps and bilateral weights equal numpy's to the last bit or the one before (C exp vs numpy's exp)."""
    import ctypes as C

    from mdimg_b200 import _lib
    lib = _lib.load_library()
    for (h, w) in [(64, 64), (512, 512), (94, 141), (3000, 3000), (4096, 4096), (7, 9)]:
        for sigma, d, ss in [(0.8, 5, 0.05), (0.2, 0, 0.1), (3.0, 13, 0.2), (2.0, 8, 0.005), (1.37, 1, 0.07)]:
            q = _c_plan(unsharp_radius=sigma, bilateral_d=d, bilateral_sigma_space=ss)
            t = _lib.EnhanceTables()
            assert lib.mdimg_enhance_tables_default(C.byref(q), h, w, C.byref(t)) == 0
            lo, hi, g = percentile_plan(h * w)
            assert list(t.pct_lo) == [int(v) for v in lo] and list(t.pct_hi) == [int(v) for v in hi]
            assert [np.float32(v) for v in t.pct_gamma] == [np.float32(v) for v in g]
            taps = gaussian_taps(sigma)
            assert t.gauss_radius == len(taps) - 1
            np.testing.assert_allclose(np.array(t.gauss_taps[:len(taps)]), taps, rtol=2.3e-16, atol=0)
            if d > 0:
                deff, spatial = bilateral_spatial(d, ss)
                assert t.bilateral_d_eff == deff
                np.testing.assert_allclose(np.array(t.bilateral_spatial[:deff * deff]), spatial.ravel(), rtol=2.3e-16, atol=0)
            else:
                assert t.bilateral_d_eff == 0


def test_c_scalar_logic_equals_the_python_arithmetic():
    """mdimg_detect_issues / mdimg_validation_scalars_of / mdimg_objective_score (host functions of the
    library) against the statement-by-statement python restatement of pipeline/metrics.py:166-179,237-408
    (itself pinned against the reference's own bytes in tests/test_oracle.py): equal to the last bit."""
    import ctypes as C

    from mdimg_b200 import _lib
    from mdimg_b200.batch import ROW_COLS
    lib = _lib.load_library()
    rng = np.random.default_rng(21)
    dp = C.POINTER(C.c_double)
    n_pass = 0
    for trial in range(400):
        row = np.zeros(2 * ROW_COLS + 2)
        row[:2 * ROW_COLS] = rng.random(2 * ROW_COLS) * rng.choice([1e-3, 0.1, 1.0, 30.0], 2 * ROW_COLS)
        row[2 * ROW_COLS] = rng.uniform(0.4, 1.0)
        row[2 * ROW_COLS + 1] = rng.uniform(10.0, 50.0)
        if trial % 7 == 0:
            row[rng.integers(0, 2 * ROW_COLS)] = np.nan            # NaN metrics (all-zero image) must propagate alike
        if trial % 11 == 0:
            row[0] = 0.0                                           # sigma_before = 0: the eps floor
        if trial % 13 == 0:
            row[2 * ROW_COLS + 1] = np.inf                         # identical images
        mb = engine.metrics_dict(row[:ROW_COLS])
        ma = engine.metrics_dict(row[ROW_COLS:2 * ROW_COLS])
        want = engine.validation_dict(mb, ma, float(row[2 * ROW_COLS]), float(row[2 * ROW_COLS + 1]),
                                      float(row[engine.MC_NIQE]), float(row[ROW_COLS + engine.MC_NIQE]),
                                      float(row[ROW_COLS + engine.MC_EDGE_RATIO]))
        v = _lib.ValidationScalars()
        assert lib.mdimg_validation_scalars_of(row.ctypes.data_as(dp), C.byref(v)) == 0
        for key in ("ssim", "psnr", "quality_improvement", "niqe_before", "niqe_after", "contrast_gain", "sharpness_gain",
                    "noise_change", "entropy_change", "snr_change", "cnr_change", "edge_density_change",
                    "histogram_spread_change", "edge_ratio", "local_contrast_change", "gradient_strength_change",
                    "gradient_entropy_change"):
            a, b = getattr(v, key), want[key]
            assert a == b or (np.isnan(a) and np.isnan(b)), (key, a, b)
        for key in ("meets_ssim", "meets_psnr", "meets_improvement", "passes", "niqe_improved"):
            assert bool(getattr(v, key)) == bool(want[key]), key
        n_pass += bool(want["passes"])
        score = C.c_double()
        parts = (C.c_double * 11)()
        assert lib.mdimg_objective_score(C.byref(v), C.byref(score), parts) == 0
        ref_score, ref_parts = engine.objective_score(want)
        if not np.isnan(score.value):
            assert round(float(score.value), 4) == ref_score
        names = [k for k in ref_parts if k != "passes"]
        assert len(names) == 11
        for k, p in zip(names, parts):
            assert round(float(p), 4) == ref_parts[k] or (np.isnan(p) and np.isnan(ref_parts[k])), k
        mask = lib.mdimg_detect_issues(row.ctypes.data_as(dp))
        from mdimg_b200.pipeline.metrics import detect_issues
        got = [nm for nm, bit in _lib.ISSUE_BITS.items() if mask & bit]
        assert got == detect_issues(mb)
    assert 0 < n_pass < 400


def test_constants_are_the_references_own():
    """PARAM_BOUNDS, THRESHOLDS, ENHANCEMENT_PARAMS, the EnhancementParams defaults and the order of the 16
    metric keys against values read from the reference's modules (tests/golden/make_reference_constants.py)."""
    import json
    from pathlib import Path

    from mdimg_b200.pipeline.schemas import EnhancementParams
    ref = json.loads((Path(__file__).resolve().parent / "golden" / "reference_constants.json").read_text())
    assert {k: list(v) for k, v in engine.PARAM_BOUNDS.items()} == ref["PARAM_BOUNDS"]
    assert list(engine.PARAM_BOUNDS) == list(ref["PARAM_BOUNDS"])
    assert engine.THRESHOLDS == ref["THRESHOLDS"]
    assert engine.ENHANCEMENT_PARAMS == ref["ENHANCEMENT_PARAMS"]
    assert list(engine.METRIC_KEYS) == ref["metric_keys"]
    mine = EnhancementParams().model_dump()
    for k, v in ref["EnhancementParams_defaults"].items():
        assert mine[k] == v, k
    assert set(mine) <= set(ref["EnhancementParams_defaults"])


def test_chunk_spans_cover_every_slice_once():
    from mdimg_b200.batch import _spans_of, default_chunk
    for n in (0, 1, 5, 64, 1000, 1024):
        for chunk, schedule in ((7, None), (512, None), (1, None), (3, [4, 2, 1]), (9, [160, 124, 33]), (2, [1000000])):
            spans = _spans_of(n, chunk, schedule)
            assert [a for a, _ in spans] == [0] * (n > 0) + [b for _, b in spans[:-1]]
            assert (spans[-1][1] if spans else 0) == n and all(b > a for a, b in spans)
            if schedule:
                want = [schedule[k % len(schedule)] for k in range(len(spans))]
                assert [b - a for a, b in spans[:-1]] == want[:len(spans) - 1]
    # ~128 Mpx per chunk, at least one image, at most 4096 slices
    assert default_chunk(512, 512) == 512 and default_chunk(3000, 3000) == 14
    assert default_chunk(100000, 100000) == 1 and default_chunk(8, 8) == 4096
