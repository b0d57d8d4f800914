"""Helper process for the agent-chain tests: replays the reference's deterministic agent chain
(pipeline/core_agents.py:61-166) over the CUDA drop-in and prints one JSON document.

Where the reference checkout exists (the build container) its OWN agent classes are imported
unmodified after ``mdimg_b200.install_as_pipeline()``; on the GPU box (no /root/reference) the three
calls the agents make are issued directly and ``pipeline.storage.validation_status`` (pinned by 96
vectors against ``ValidationAgent.run``) supplies the decisions.

    python tests/agents_chain_runner.py identity        # CPU: which functions do the reference's agents hold?
    python tests/agents_chain_runner.py chain OUT.npz   # GPU: run the chain on the four golden inputs
"""

from __future__ import annotations

import dataclasses
import json
import logging
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE / "golden"))


def _plain(o):
    return o.item() if hasattr(o, "item") else str(o)


def identity() -> dict:
    """Import order A: install first, then import the agents.  Order B (a second interpreter run with
    'identity-late'): agents imported over stand-ins first, install afterwards -> rebinding."""
    import mdimg_b200
    from make_reference_agents import install_io_stubs
    from mdimg_b200.pipeline import dicom_io, enhancement, metrics
    install_io_stubs()
    late = len(sys.argv) > 2 and sys.argv[2] == "late"
    out = {"late": late}
    if late:
        # the reference's modules get imported with stand-in leaves first, as a host that imports
        # its pipeline before enabling the GPU path would have them
        from make_reference_glue import install_skimage_stand_in
        install_skimage_stand_in()
        sys.path.insert(0, str(REFERENCE))
        import pipeline.core_agents as agents
        assert agents.compute_metrics.__module__ == "pipeline.metrics"
        assert agents.compute_metrics is not metrics.compute_metrics
        out["installed"] = mdimg_b200.install_as_pipeline()
    else:
        out["installed"] = mdimg_b200.install_as_pipeline(reference_root=REFERENCE)
        import pipeline.core_agents as agents
    import pipeline
    import pipeline.dicom_io as ref_dio
    out["agents_file"] = agents.__file__
    out["same"] = {
        "compute_metrics": agents.compute_metrics is metrics.compute_metrics,
        "detect_issues": agents.detect_issues is metrics.detect_issues,
        "compute_validation": agents.compute_validation is metrics.compute_validation,
        "apply_enhancements": agents.apply_enhancements is enhancement.apply_enhancements,
        "normalize_image": ref_dio.normalize_image is dicom_io.normalize_image,
        "pkg.metrics": pipeline.metrics is metrics,
        "pkg.enhancement": pipeline.enhancement is enhancement,
        "report_builder_is_reference": ref_dio.build_markdown_report.__module__ == "pipeline.dicom_io"
        and Path(ref_dio.__file__).resolve().is_relative_to(REFERENCE),
    }
    from mdimg_b200 import engine
    out["logger"] = engine.logger.name
    # the tool layer imports more names (tools.py:21,106); it needs nothing else that is absent here
    try:
        import pipeline.tools as tools
        out["same"]["tools.compute_objective_score"] = tools.compute_objective_score is metrics.compute_objective_score
        out["same"]["tools.compute_validation"] = tools.compute_validation is metrics.compute_validation
    except Exception as exc:  # noqa: BLE001
        out["tools_import_error"] = f"{type(exc).__name__}: {exc}"
    return out


def chain(out_npz: str) -> dict:
    import mdimg_b200
    from make_reference_agents import chain_inputs, install_io_stubs, run_chain
    records = []

    class Keep(logging.Handler):
        def emit(self, record):
            records.append((record.name, record.getMessage()))

    for name in ("pipeline.enhancement", "mdimg_b200.enhancement"):
        lg = logging.getLogger(name)
        lg.setLevel(logging.WARNING)            # the root logger of this helper only prints errors
        lg.addHandler(Keep())
    use_reference = REFERENCE.exists()
    if use_reference:
        install_io_stubs()
        mdimg_b200.install_as_pipeline(reference_root=REFERENCE)
        import pipeline.core_agents as agents
    else:
        from mdimg_b200 import engine
        engine.set_logger_name("pipeline.enhancement")
    from mdimg_b200.pipeline import enhancement, metrics, storage
    out, arrays = {"used_reference_agents": use_reference}, {}
    for name, im in chain_inputs().items():
        if use_reference:
            det, rec, enh, val, report = run_chain(agents, im)
            entry = {"detection": {"metrics": det.metrics, "issues": det.issues},
                     "enhancement": {"applied_ops": enh.applied_ops, "metrics": enh.metrics},
                     "validation": dataclasses.asdict(val), "report": report}
            image = enh.image
        else:
            m = metrics.compute_metrics(im)                                   # QualityDetectionAgent.run
            issues = metrics.detect_issues(m)
            image, applied = enhancement.apply_enhancements(im, list(issues))  # EnhancementAgent.run
            m_after = metrics.compute_metrics(image)
            v = metrics.compute_validation(im, image)                        # ValidationAgent.run
            entry = {"detection": {"metrics": m, "issues": issues},
                     "enhancement": {"applied_ops": applied, "metrics": m_after},
                     "validation": storage.validation_status(v, issues), "report": None}
        arrays[name] = image
        out[name] = entry
    out["log_records"] = records
    np.savez_compressed(out_npz, **arrays)
    return out


if __name__ == "__main__":
    mode = sys.argv[1]
    logging.basicConfig(level=logging.ERROR)
    result = identity() if mode == "identity" else chain(sys.argv[2])
    print("RESULT_JSON " + json.dumps(result, default=_plain))
