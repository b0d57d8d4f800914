"""Run persistence for batches.

The reference stores one row per processed image in the SQLite table ``runs``
(``pipeline/storage.py:38-58`` is the schema, ``storage.py:120-170`` ``save_run``), opening a
connection, writing one row and committing for every image (``pipeline/runner.py:395-444``
``_persist_run``).  A stack of 1024 slices gathered from 8 GPUs would pay 1024 connections and
1024 commits; here the gathered result rows are written with ONE connection, ONE transaction and
one ``executemany`` (``save_runs``).  The table layout, column encodings (JSON text) and the
``MDIMG_DB_PATH`` override are the reference's, so its readers (``get_run`` / ``list_runs``, the
Flask API) see batch rows exactly like single-image rows.  ``save_run`` keeps the reference's
signature and is the one-record case of ``save_runs``.

Chat messages and agent traces (the other two users of that database) are outside the image path
and are not mirrored.
"""

from __future__ import annotations

import json
import os
import sqlite3
import uuid
from contextlib import contextmanager
from datetime import datetime, timezone
from typing import Any, Dict, Iterable, Iterator, List, Mapping, Optional, Sequence

import numpy as np

#: column order of ``runs`` (reference pipeline/storage.py:38-58); the first is the primary key
RUN_COLUMNS = (
    "run_id", "timestamp", "input_filename", "metadata_summary", "issues", "metrics_before",
    "metrics_after", "plan_json", "validation", "applied_ops", "explainability", "report_path",
    "before_after_path", "agent_logs", "status", "genai_model", "genai_llm_calls",
)
_JSON_COLUMNS = ("metadata_summary", "issues", "metrics_before", "metrics_after", "validation",
                 "applied_ops", "agent_logs", "explainability")

_DDL = (
    """CREATE TABLE IF NOT EXISTS runs (
        run_id TEXT PRIMARY KEY,
        timestamp TEXT NOT NULL,
        input_filename TEXT NOT NULL,
        metadata_summary TEXT DEFAULT '{}',
        issues TEXT DEFAULT '[]',
        metrics_before TEXT DEFAULT '{}',
        metrics_after TEXT DEFAULT '{}',
        plan_json TEXT DEFAULT '',
        validation TEXT DEFAULT '{}',
        applied_ops TEXT DEFAULT '[]',
        explainability TEXT DEFAULT '{}',
        report_path TEXT DEFAULT '',
        before_after_path TEXT DEFAULT '',
        agent_logs TEXT DEFAULT '[]',
        status TEXT DEFAULT 'completed',
        genai_model TEXT DEFAULT '',
        genai_llm_calls INTEGER DEFAULT 0)""",
    "CREATE INDEX IF NOT EXISTS idx_runs_ts ON runs(timestamp)",
)


def _db_path() -> str:
    default = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                           "data", "mdimg.db")
    return os.environ.get("MDIMG_DB_PATH", default)


@contextmanager
def _session() -> Iterator[sqlite3.Connection]:
    """One connection = one transaction: committed on success, rolled back on error."""
    path = _db_path()
    folder = os.path.dirname(path)
    if folder:
        os.makedirs(folder, exist_ok=True)
    conn = sqlite3.connect(path)
    conn.row_factory = sqlite3.Row
    try:
        conn.execute("PRAGMA journal_mode=WAL")
        conn.execute("PRAGMA foreign_keys=ON")
        yield conn
        conn.commit()
    except BaseException:
        conn.rollback()
        raise
    finally:
        conn.close()


def init_db() -> None:
    """Create the ``runs`` table (and its timestamp index) when missing."""
    with _session() as conn:
        for stmt in _DDL:
            conn.execute(stmt)


def generate_run_id() -> str:
    return uuid.uuid4().hex[:12]


def _plain(obj: Any) -> Any:
    """numpy scalars / arrays / bools -> JSON-serialisable python values (recursively).  Like the
    reference, numpy integers become floats."""
    if isinstance(obj, Mapping):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, np.bool_):
        return bool(obj)
    if isinstance(obj, (np.floating, np.integer)):
        return float(obj)
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    return obj


def _encode(rec: Mapping[str, Any], stamp: str) -> tuple:
    """One record (keyword names of the reference's ``save_run``) -> one table row."""
    expl = rec.get("explainability", {})
    return (
        rec["run_id"],
        stamp,
        rec["input_filename"],
        json.dumps(rec.get("metadata_summary", {}), default=str),
        json.dumps(rec.get("issues", [])),
        json.dumps(_plain(rec.get("metrics_before", {}))),
        json.dumps(_plain(rec.get("metrics_after", {}))),
        rec.get("plan_json", ""),
        json.dumps(_plain(rec.get("validation", {}))),
        json.dumps(rec.get("applied_ops", [])),
        json.dumps(expl, default=str) if isinstance(expl, dict) else str(expl),
        rec.get("report_path", ""),
        rec.get("before_after_path", ""),
        json.dumps(rec.get("agent_logs", []), default=str),
        rec.get("status", "completed"),
        rec.get("genai_model", ""),
        int(rec.get("genai_llm_calls", 0)),
    )


_UPSERT = (f"INSERT OR REPLACE INTO runs ({', '.join(RUN_COLUMNS)}) "
           f"VALUES ({', '.join('?' * len(RUN_COLUMNS))})")


def save_runs(records: Iterable[Mapping[str, Any]]) -> int:
    """Bulk form of ``save_run``: every record in one transaction (all or nothing).  Returns the
    number of rows written.  The table is created on first use."""
    stamp = datetime.now(timezone.utc).isoformat()
    rows = [_encode(r, stamp) for r in records]
    if not rows:
        return 0
    with _session() as conn:
        for stmt in _DDL:
            conn.execute(stmt)
        conn.executemany(_UPSERT, rows)
    return len(rows)


def save_run(run_id: str, input_filename: str, metadata_summary: dict, issues: List[str],
             metrics_before: dict, metrics_after: dict, plan_json: str, validation: dict,
             applied_ops: List[str], explainability, report_path: str, before_after_path: str,
             agent_logs: List[dict], status: str = "completed", genai_model: str = "",
             genai_llm_calls: int = 0) -> None:
    """Reference signature (pipeline/storage.py:120-137): insert or replace one completed run."""
    save_runs([dict(run_id=run_id, input_filename=input_filename, metadata_summary=metadata_summary,
                    issues=issues, metrics_before=metrics_before, metrics_after=metrics_after,
                    plan_json=plan_json, validation=validation, applied_ops=applied_ops,
                    explainability=explainability, report_path=report_path,
                    before_after_path=before_after_path, agent_logs=agent_logs, status=status,
                    genai_model=genai_model, genai_llm_calls=genai_llm_calls)])


def insert_pending_run(run_id: str, input_filename: str) -> None:
    with _session() as conn:
        conn.execute("INSERT OR IGNORE INTO runs (run_id, timestamp, input_filename, status) VALUES (?, ?, ?, ?)",
                     (run_id, datetime.now(timezone.utc).isoformat(), input_filename, "pending"))


def update_run_status(run_id: str, status: str) -> None:
    with _session() as conn:
        conn.execute("UPDATE runs SET status = ? WHERE run_id = ?", (status, run_id))


def _decode(row: sqlite3.Row) -> Dict[str, Any]:
    rec = dict(row)
    for col in _JSON_COLUMNS:
        text = rec.get(col)
        if isinstance(text, str):
            try:
                rec[col] = json.loads(text)
            except ValueError:
                pass          # plain-text explainability stays text
    return rec


def get_run(run_id: str) -> Optional[Dict[str, Any]]:
    with _session() as conn:
        row = conn.execute("SELECT * FROM runs WHERE run_id = ?", (run_id,)).fetchone()
    return None if row is None else _decode(row)


def list_runs(limit: int = 100, offset: int = 0) -> List[Dict[str, Any]]:
    """Most recent first."""
    with _session() as conn:
        rows = conn.execute("SELECT * FROM runs ORDER BY timestamp DESC LIMIT ? OFFSET ?",
                            (limit, offset)).fetchall()
    return [_decode(r) for r in rows]


# ------------------------------------------------------------------------------------------
# Records from gathered result rows
# ------------------------------------------------------------------------------------------
def validation_status(validation: Mapping[str, Any], issues: Sequence[str]) -> Dict[str, Any]:
    """The scalar decisions ``ValidationAgent.run`` adds on top of ``compute_validation``
    (pipeline/core_agents.py:105-161): pass rule without detected issues, PASS / WARN / FAIL and
    the notes.  Returns the fields of the reference's ``ValidationResult``, which is what
    ``_persist_run`` stores in the ``validation`` column."""
    passes = bool(validation["passes"])
    meets_improvement = bool(validation["meets_improvement"])
    notes: List[str] = []
    if not issues:
        notes.append("No issues detected; enhancement not required.")
        passes = bool(validation["meets_ssim"] and validation["meets_psnr"])
        meets_improvement = True
    status = "PASS" if passes else "FAIL"
    if not passes and validation["quality_improvement"] > 0:
        status = "WARN"
        notes.append("Some improvement observed, but thresholds not fully met.")
    notes.append("Naturalness preserved (NIQE-approx stable or improved)." if validation.get("niqe_improved")
                 else "Warning: Naturalness may be degraded (possible over-processing).")
    noise_change = validation.get("noise_change", 0)
    if noise_change > 0.5:
        notes.append(f"Note: Noise increased by {noise_change * 100:.1f}% (sharpening side-effect).")
    keep = ("ssim", "psnr", "quality_improvement", "meets_ssim", "meets_psnr")
    out: Dict[str, Any] = {k: validation[k] for k in keep}
    out.update(meets_improvement=meets_improvement, passes=passes, status=status, notes=notes,
               niqe_before=validation.get("niqe_before", 0.0), niqe_after=validation.get("niqe_after", 0.0),
               niqe_improved=validation.get("niqe_improved", True),
               contrast_gain=validation.get("contrast_gain", 0.0),
               sharpness_gain=validation.get("sharpness_gain", 0.0), noise_change=noise_change)
    return out


def stack_records(packed: np.ndarray, labels: Sequence[Sequence[str]], input_filename: str,
                  plan_json: str = "", metadata_summary: Optional[dict] = None,
                  run_ids: Optional[Sequence[str]] = None, first_slice: int = 0) -> List[Dict[str, Any]]:
    """Per-slice ``save_run`` records from the packed result rows of a stack (``batch.StackResult``
    or the rows gathered from all ranks by ``shard.gather_rows``): metrics before / after, detected
    issues, the ``ValidationResult`` fields and status, the applied-operation labels.  Pure host
    arithmetic on the [N, PACK_COLS] float64 matrix; no image is touched."""
    from ..batch import StackResult          # local import: batch imports torch / the CUDA library
    from .metrics import detect_issues

    res = StackResult(enhanced=None, packed=np.asarray(packed, dtype=np.float64), labels=[list(l) for l in labels])
    n = res.packed.shape[0]
    if len(res.labels) != n:
        raise ValueError(f"{n} result rows but {len(res.labels)} label lists")
    if run_ids is not None and len(run_ids) != n:
        raise ValueError(f"{n} result rows but {len(run_ids)} run ids")
    records = []
    for i in range(n):
        mb, ma = res.metrics_before(i), res.metrics_after(i)
        issues = detect_issues(mb)
        with np.errstate(divide="ignore", invalid="ignore"):
            val = validation_status(res.validation(i), issues)
        meta = dict(metadata_summary or {})
        meta["slice_index"] = first_slice + i
        records.append(dict(
            run_id=run_ids[i] if run_ids is not None else generate_run_id(),
            input_filename=input_filename, metadata_summary=meta, issues=issues,
            metrics_before=mb, metrics_after=ma, plan_json=plan_json, validation=val,
            applied_ops=res.labels[i], explainability={}, report_path="", before_after_path="",
            agent_logs=[], status=val["status"]))
    return records


def save_stack(packed: np.ndarray, labels: Sequence[Sequence[str]], input_filename: str, **kwargs) -> List[str]:
    """``stack_records`` + ``save_runs``: one transaction for the whole stack.  Returns the run ids."""
    records = stack_records(packed, labels, input_filename, **kwargs)
    save_runs(records)
    return [r["run_id"] for r in records]
