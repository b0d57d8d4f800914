"""Batch persistence of the gathered result rows (reference
pipeline/storage.py:120-170 ``save_run`` and pipeline/runner.py:395-444 ``_persist_run``).

CPU only: SQLite + host arithmetic.  The first tests restate the reference's own
tests/test_storage.py for the run table; the interoperability tests execute the reference's
storage module itself (stdlib + numpy, importable here) when the checkout is present; the
``ValidationAgent`` decisions are pinned by vectors generated from the reference's own code
(tests/golden/make_validation_status.py)."""

from __future__ import annotations

import importlib.util
import json
import os
import sqlite3
from pathlib import Path

import numpy as np
import pytest

from mdimg_b200 import engine
from mdimg_b200.batch import PACK_COLS, ROW_COLS
from mdimg_b200.pipeline import storage
from mdimg_b200.shard import gather_labels
from oracle import ref_metrics as omet

GOLDEN = Path(__file__).resolve().parent / "golden"
REFERENCE_STORAGE = Path("/root/reference/pipeline/storage.py")


@pytest.fixture(autouse=True)
def _fresh_db(tmp_path, monkeypatch):
    monkeypatch.setenv("MDIMG_DB_PATH", str(tmp_path / "runs" / "mdimg.db"))
    storage.init_db()
    yield


def _record(rid, name="test.dcm", **kw):
    rec = dict(run_id=rid, input_filename=name, metadata_summary={"Modality": "CT"}, issues=["noise"],
               metrics_before={"sigma": 0.12}, metrics_after={"sigma": 0.05}, plan_json='{"ops": ["denoise"]}',
               validation={"ssim": 0.92, "psnr": 30.0, "passes": True}, applied_ops=["denoise"],
               explainability={"detected_issues": "Noise"}, report_path="outputs/test_report.md",
               before_after_path="outputs/test_before_after.png", agent_logs=[{"phase": "detection"}],
               status="PASS", genai_model="gpt-5-mini")
    rec.update(kw)
    return rec


def test_generate_run_id():
    ids = {storage.generate_run_id() for _ in range(200)}
    assert len(ids) == 200 and all(isinstance(i, str) and len(i) > 10 for i in ids)


def test_save_and_get_run():
    rid = storage.generate_run_id()
    storage.save_run(**_record(rid))
    row = storage.get_run(rid)
    assert row is not None
    assert row["input_filename"] == "test.dcm" and row["status"] == "PASS"
    assert row["issues"] == ["noise"] and row["metrics_after"] == {"sigma": 0.05}
    assert row["validation"]["passes"] is True and row["genai_llm_calls"] == 0
    assert storage.get_run("missing") is None


def test_list_runs_most_recent_first_and_paged():
    for i in range(3):
        storage.save_run(**_record(storage.generate_run_id(), name=f"file{i}.dcm", explainability=""))
    runs = storage.list_runs()
    assert len(runs) == 3
    assert [r["timestamp"] for r in runs] == sorted((r["timestamp"] for r in runs), reverse=True)
    assert len(storage.list_runs(limit=2)) == 2 and len(storage.list_runs(limit=5, offset=2)) == 1


def test_pending_then_status_then_replace():
    rid = storage.generate_run_id()
    storage.insert_pending_run(rid, "a.dcm")
    assert storage.get_run(rid)["status"] == "pending"
    storage.insert_pending_run(rid, "other.dcm")             # INSERT OR IGNORE
    assert storage.get_run(rid)["input_filename"] == "a.dcm"
    storage.update_run_status(rid, "running")
    assert storage.get_run(rid)["status"] == "running"
    storage.save_run(**_record(rid, name="a.dcm", status="WARN"))   # INSERT OR REPLACE
    assert storage.get_run(rid)["status"] == "WARN" and len(storage.list_runs()) == 1


def test_numpy_values_are_serialised():
    rid = storage.generate_run_id()
    storage.save_run(**_record(rid, metrics_before={"sigma": np.float32(0.25), "n": np.int64(3)},
                               validation={"passes": np.bool_(True), "v": np.arange(3), "nested": {"x": (np.float64(1.5),)}}))
    row = storage.get_run(rid)
    assert row["metrics_before"] == {"sigma": 0.25, "n": 3.0}
    assert row["validation"] == {"passes": True, "v": [0, 1, 2], "nested": {"x": [1.5]}}


def test_bulk_save_is_one_transaction():
    good = [_record(f"bulk{i:04d}", name=f"s{i}.dcm") for i in range(500)]
    assert storage.save_runs(good) == 500
    assert len(storage.list_runs(limit=1000)) == 500
    assert storage.save_runs([]) == 0
    # a record that cannot be written (NOT NULL input_filename) rolls the whole batch back
    bad = [_record(f"more{i:04d}") for i in range(10)] + [_record("broken", name=None)]
    with pytest.raises(sqlite3.IntegrityError):
        storage.save_runs(bad)
    assert len(storage.list_runs(limit=1000)) == 500 and storage.get_run("more0000") is None


def test_plain_text_explainability_stays_text():
    storage.save_run(**_record("txt", explainability="free text"))
    assert storage.get_run("txt")["explainability"] == "free text"


# ---- interoperability with the reference's own storage module ---------------------------------
def _reference_storage():
    if not REFERENCE_STORAGE.exists():
        pytest.skip("reference checkout not present (build container only)")
    spec = importlib.util.spec_from_file_location("_ref_storage", REFERENCE_STORAGE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rows_written_here_are_read_by_the_reference_and_back():
    ref = _reference_storage()
    ours = [_record(f"ours{i}", name=f"o{i}.dcm", metrics_after={"sigma": np.float32(0.5), "std": 0.1 * i})
            for i in range(4)]
    storage.save_runs(ours)
    for rec in ours:
        assert ref.get_run(rec["run_id"]) == storage.get_run(rec["run_id"])
    assert ref.list_runs() == storage.list_runs()
    theirs = _record("theirs", metrics_before={"sigma": np.float64(0.3)}, explainability="plain")
    ref.save_run(**theirs)
    assert storage.get_run("theirs") == ref.get_run("theirs")
    # same table definition: the reference's init_db on our file (and ours on theirs) changes nothing
    ref.init_db()
    storage.init_db()
    with sqlite3.connect(os.environ["MDIMG_DB_PATH"]) as conn:
        cols = [r[1] for r in conn.execute("PRAGMA table_info(runs)")]
    assert tuple(cols) == storage.RUN_COLUMNS


def test_encoded_row_equals_the_references_encoding():
    ref = _reference_storage()
    rec = _record("enc", metrics_before={"sigma": np.float32(0.12), "k": np.int32(2)},
                  validation={"passes": np.bool_(False), "ssim": np.float64(0.8)})
    storage.save_runs([rec])
    with sqlite3.connect(os.environ["MDIMG_DB_PATH"]) as conn:
        mine = conn.execute("SELECT * FROM runs WHERE run_id='enc'").fetchone()
    ref.save_run(**rec)
    with sqlite3.connect(os.environ["MDIMG_DB_PATH"]) as conn:
        his = conn.execute("SELECT * FROM runs WHERE run_id='enc'").fetchone()
    assert mine[0] == his[0] and mine[2:] == his[2:]          # every column but the timestamp


# ---- ValidationAgent decisions ----------------------------------------------------------------
def test_validation_status_matches_the_references_agent():
    cases = json.loads((GOLDEN / "validation_status.json").read_text())
    assert len(cases) == 96
    seen = set()
    for c in cases:
        got = storage.validation_status(c["validation"], c["issues"])
        assert got == c["result"], (c, got)
        assert list(got) == list(c["result"])                 # same field order as ValidationResult
        seen.add(got["status"])
    assert seen == {"PASS", "WARN", "FAIL"}


# ---- records from packed result rows ----------------------------------------------------------
def _packed_rows(n, seed=0):
    rng = np.random.default_rng(seed)
    packed = np.zeros((n, PACK_COLS))
    packed[:, :2 * ROW_COLS] = rng.random((n, 2 * ROW_COLS))
    packed[:, 2 * ROW_COLS] = rng.uniform(0.3, 1.0, n)         # ssim
    packed[:, 2 * ROW_COLS + 1] = rng.uniform(15.0, 45.0, n)   # psnr
    packed[0, 0] = 0.01                                        # sigma below / above the noise threshold
    packed[1, 0] = 0.5
    labels = [[f"op{j}" for j in range(i % 3)] for i in range(n)]
    return packed, labels


def test_stack_records_follow_the_reference_logic():
    packed, labels = _packed_rows(12)
    recs = storage.stack_records(packed, labels, "vol.dcm", plan_json='{"p": 1}', metadata_summary={"Modality": "CT"},
                                 first_slice=100)
    assert len(recs) == 12 and len({r["run_id"] for r in recs}) == 12
    for i, r in enumerate(recs):
        mb = engine.metrics_dict(packed[i, :ROW_COLS])
        ma = engine.metrics_dict(packed[i, ROW_COLS:2 * ROW_COLS])
        assert r["metrics_before"] == mb and r["metrics_after"] == ma
        assert r["issues"] == omet.detect_issues(mb)
        assert r["applied_ops"] == labels[i] and r["input_filename"] == "vol.dcm"
        assert r["metadata_summary"] == {"Modality": "CT", "slice_index": 100 + i}
        val = engine.validation_dict(mb, ma, float(packed[i, 2 * ROW_COLS]), float(packed[i, 2 * ROW_COLS + 1]),
                                     float(packed[i, engine.MC_NIQE]), float(packed[i, ROW_COLS + engine.MC_NIQE]),
                                     float(packed[i, ROW_COLS + engine.MC_EDGE_RATIO]))
        assert r["validation"] == storage.validation_status(val, r["issues"])
        assert r["status"] == r["validation"]["status"] in ("PASS", "WARN", "FAIL")
    assert "noise" not in recs[0]["issues"] and "noise" in recs[1]["issues"]
    with pytest.raises(ValueError):
        storage.stack_records(packed, labels[:-1], "vol.dcm")
    with pytest.raises(ValueError):
        storage.stack_records(packed, labels, "vol.dcm", run_ids=["a"])


def test_save_stack_round_trip():
    packed, labels = _packed_rows(64, seed=3)
    ids = storage.save_stack(packed, labels, "stack.dcm", run_ids=[f"s{i:03d}" for i in range(64)])
    assert ids == [f"s{i:03d}" for i in range(64)]
    rows = {r["run_id"]: r for r in storage.list_runs(limit=100)}
    assert len(rows) == 64
    for i in (0, 17, 63):
        r = rows[f"s{i:03d}"]
        assert r["metrics_after"] == engine.metrics_dict(packed[i, ROW_COLS:2 * ROW_COLS])   # float64 survives JSON
        assert r["applied_ops"] == labels[i] and r["metadata_summary"]["slice_index"] == i
        assert r["status"] == r["validation"]["status"]


def test_gather_labels_without_a_process_group_is_the_identity():
    labels = [["a"], [], ["b", "c"]]
    assert gather_labels(labels) == labels


# ---- two ranks: gather rows + labels, rank 0 persists the whole stack in one transaction --------
def _persist_worker(rank, world, n_total, port, db_path):
    import torch
    import torch.distributed as dist

    from mdimg_b200.shard import gather_labels, gather_rows, slice_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["MDIMG_DB_PATH"] = db_path
    dist.init_process_group("gloo", rank=rank, world_size=world)
    packed, labels = _packed_rows(n_total, seed=17)             # every rank builds the same table, keeps its span
    a, b = slice_range(n_total, rank, world)
    rows = gather_rows(torch.from_numpy(packed[a:b].copy()), n_total)
    labs = gather_labels(labels[a:b])
    if rank == 0:
        storage.save_stack(rows.numpy(), labs, "cohort.dcm", run_ids=[f"g{i:03d}" for i in range(n_total)],
                           plan_json='{"ops": []}')
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gather_and_rank0_persists(tmp_path):
    import torch.multiprocessing as mp
    n_total = 9                                                    # unequal spans: 5 + 4
    db = str(tmp_path / "gathered" / "mdimg.db")
    ctx = mp.get_context("spawn")
    port = 29700 + (os.getpid() % 1500)
    procs = [ctx.Process(target=_persist_worker, args=(r, 2, n_total, port, db)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    os.environ["MDIMG_DB_PATH"] = db
    packed, labels = _packed_rows(n_total, seed=17)
    rows = {r["run_id"]: r for r in storage.list_runs(limit=100)}
    assert sorted(rows) == [f"g{i:03d}" for i in range(n_total)]
    for i in range(n_total):
        r = rows[f"g{i:03d}"]
        assert r["metadata_summary"]["slice_index"] == i and r["applied_ops"] == labels[i]
        assert r["metrics_before"] == engine.metrics_dict(packed[i, :ROW_COLS])
        assert r["plan_json"] == '{"ops": []}' and r["status"] in ("PASS", "WARN", "FAIL")
