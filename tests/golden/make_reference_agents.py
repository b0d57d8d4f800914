"""Golden transcript of the reference's deterministic agent chain, produced by the REFERENCE'S OWN
agent classes (`pipeline/core_agents.py:61-166`, imported unmodified) over the reference's own
`pipeline/metrics.py` / `pipeline/enhancement.py` / `pipeline/dicom_io.py`:

    QualityDetectionAgent -> RecommendationAgent -> EnhancementAgent -> ValidationAgent -> ReportAgent

exactly as `tests/test_pipeline.py:97-137` of the reference drives it.  scikit-image / PyWavelets /
pydicom / matplotlib are absent here, so the scikit-image leaves forward to the oracle's
restatements (same stand-in as make_reference_glue.py) and pydicom / matplotlib are empty stubs
(the chain never calls them).  The GPU suite replays the chain over the CUDA drop-in — through the
reference's agent classes themselves where /root/reference exists, through the three calls the
agents make otherwise — and compares with this transcript.

Run in the build container only (needs /root/reference):
    python tests/golden/make_reference_agents.py
"""

from __future__ import annotations

import dataclasses
import json
import sys
import types
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

from make_reference_glue import install_skimage_stand_in  # noqa: E402
from mdimg_b200 import synth  # noqa: E402


def install_io_stubs() -> None:
    """pydicom / matplotlib are imported at module level by pipeline/dicom_io.py; the agent chain
    never calls into them."""
    for name in ("pydicom", "pydicom.errors", "pydicom.pixel_data_handlers", "pydicom.pixel_data_handlers.util",
                 "matplotlib", "matplotlib.pyplot"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:  # noqa: BLE001
            pass
        m = types.ModuleType(name)
        if name.count(".") == 0 or name.endswith("pixel_data_handlers"):
            m.__path__ = []
        sys.modules[name] = m
    sys.modules["pydicom.errors"].__dict__.setdefault("InvalidDicomError", type("InvalidDicomError", (Exception,), {}))
    sys.modules["pydicom.pixel_data_handlers.util"].__dict__.setdefault("apply_modality_lut", lambda a, ds: a)
    sys.modules["matplotlib"].__dict__.setdefault("use", lambda *a, **k: None)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def chain_inputs():
    ims = {"clean64": synth.fixture_clean(), "noisy64": synth.fixture_noisy(), "lowc64": synth.fixture_low_contrast()}
    ct = synth.ct_slice(1000, 0.25, size=96).astype(np.float32)
    ims["ct96"] = (ct - ct.min()) / (ct.max() - ct.min())
    return ims


def run_chain(agents, image):
    """The reference's own driver, tests/test_pipeline.py:97-137."""
    detection = agents.QualityDetectionAgent().run(image)
    recommendations = agents.RecommendationAgent().run(detection)
    enhancement = agents.EnhancementAgent().run(image, recommendations)
    validation = agents.ValidationAgent().run(image, enhancement.image, detection)
    context = {
        "input_path": "test.dcm", "metadata": {"Modality": "CT"}, "issues": detection.issues,
        "recommendations": recommendations.recommendations, "applied_ops": enhancement.applied_ops,
        "metrics_before": detection.metrics, "metrics_after": enhancement.metrics,
        "validation": validation, "visuals": {}, "notes": validation.notes,
    }
    report = agents.ReportAgent().run(context)
    return detection, recommendations, enhancement, validation, report


def main() -> None:
    assert REFERENCE.exists(), "needs the reference checkout"
    install_skimage_stand_in()
    install_io_stubs()
    sys.path.insert(0, str(REFERENCE))
    warnings.filterwarnings("ignore")
    import pipeline.core_agents as agents       # the reference's own module, unmodified

    out, arrays = {}, {}
    for name, im in chain_inputs().items():
        det, rec, enh, val, report = run_chain(agents, im)
        arrays[name] = enh.image
        out[name] = {
            "detection": {"metrics": det.metrics, "issues": det.issues},
            "recommendations": {"recommendations": rec.recommendations, "mapping": rec.mapping},
            "enhancement": {"applied_ops": enh.applied_ops, "metrics": enh.metrics},
            "validation": dataclasses.asdict(val),
            "report": report,
        }
    (HERE / "reference_agents.json").write_text(
        json.dumps(out, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o)))
    np.savez_compressed(HERE / "reference_agents.npz", **arrays)
    print("wrote", list(out))


if __name__ == "__main__":
    main()
