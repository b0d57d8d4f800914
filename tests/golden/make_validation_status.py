"""Golden vectors for the scalar decisions of the reference's ``ValidationAgent.run``
(pipeline/core_agents.py:105-161), which is what ``_persist_run`` stores in the ``validation``
and ``status`` columns (pipeline/runner.py:395-444).

``pipeline/core_agents.py`` imports the image modules (scikit-image, pydicom, matplotlib: absent
here), so they are replaced by empty stand-ins whose ``compute_validation`` returns the case's
dict; the agent's own code then runs unmodified.  Run in the build container only:
    python tests/golden/make_validation_status.py
"""

from __future__ import annotations

import itertools
import json
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
REFERENCE = Path("/root/reference")


def main() -> None:
    assert REFERENCE.exists(), "needs the reference checkout"
    current = {}
    pkg = types.ModuleType("pipeline")
    pkg.__path__ = [str(REFERENCE / "pipeline")]
    met = types.ModuleType("pipeline.metrics")
    met.compute_metrics = met.detect_issues = None
    met.compute_validation = lambda original, enhanced: dict(current["v"])
    enh = types.ModuleType("pipeline.enhancement")
    enh.apply_enhancements = None
    dio = types.ModuleType("pipeline.dicom_io")
    dio.build_markdown_report = None
    for m in (pkg, met, enh, dio):
        sys.modules[m.__name__] = m
    import importlib
    agents = importlib.import_module("pipeline.core_agents")

    cases = []
    flags = list(itertools.product([False, True], repeat=5))
    for k, (ok_ssim, ok_psnr, ok_gain, niqe_ok, has_issues) in enumerate(flags):
        for qi, noise in ((-0.2, 0.1), (0.03, 0.75), (0.4, 0.5000001)):
            passes = (ok_ssim and ok_psnr) or (ok_ssim and ok_gain) or (ok_psnr and ok_gain and niqe_ok)
            v = {"ssim": 0.5 + 0.01 * k, "psnr": 20.0 + k, "quality_improvement": qi,
                 "meets_ssim": ok_ssim, "meets_psnr": ok_psnr, "meets_improvement": ok_gain,
                 "passes": passes, "niqe_before": 0.8, "niqe_after": 0.7 if niqe_ok else 0.9,
                 "niqe_improved": niqe_ok, "contrast_gain": 0.1 * k, "sharpness_gain": -0.05 * k,
                 "noise_change": noise}
            issues = ["noise", "blur"] if has_issues else []
            current["v"] = v
            res = agents.ValidationAgent().run(None, None, agents.DetectionResult(metrics={}, issues=issues))
            cases.append({"validation": v, "issues": issues, "result": dict(vars(res))})
    (HERE / "validation_status.json").write_text(json.dumps(cases, indent=0))
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
