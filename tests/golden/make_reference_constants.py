"""The reference's own constants for the hot path, read from its modules in the build container:
PARAM_BOUNDS and the EnhancementParams defaults (pipeline/schemas.py, imports as it is), THRESHOLDS
(pipeline/metrics.py:25-34) and ENHANCEMENT_PARAMS (pipeline/enhancement.py:32-42) -- the latter two
modules import scikit-image, which is supplied by the stand-in of make_reference_glue.py (constants do
not depend on it).
    python tests/golden/make_reference_constants.py
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
sys.path.insert(0, str(HERE.parent.parent))

from make_reference_glue import REFERENCE, install_skimage_stand_in  # noqa: E402


def main() -> None:
    assert REFERENCE.exists(), "needs the reference checkout"
    install_skimage_stand_in()
    sys.path.insert(0, str(REFERENCE))
    import pipeline.enhancement as renh
    import pipeline.metrics as rmet
    import pipeline.schemas as rsch
    import numpy as np
    keys = list(rmet.compute_metrics(np.linspace(0, 1, 64 * 64, dtype=np.float32).reshape(64, 64)))
    out = {
        "PARAM_BOUNDS": {k: list(v) for k, v in rsch.PARAM_BOUNDS.items()},
        "THRESHOLDS": dict(rmet.THRESHOLDS),
        "ENHANCEMENT_PARAMS": dict(renh.ENHANCEMENT_PARAMS),
        "EnhancementParams_defaults": rsch.EnhancementParams().model_dump(),
        "metric_keys": keys,
    }
    (HERE / "reference_constants.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out)[:300])


if __name__ == "__main__":
    main()
