"""Generates the golden vectors under tests/golden/.

The reference itself cannot be imported in the build container (scikit-image, PyWavelets and
pydicom are not installed and there is no network), so these vectors come from (a) numpy/scipy —
which ARE the reference's own arithmetic for box filters, percentiles and padding — and (b) the
oracle restatement, frozen here so that any later change to the oracle is detected.  They pin the
oracle against itself over time; they do not pin it against scikit-image (parity unpinned).

    python tests/golden/make_golden.py
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
from scipy import ndimage as ndi

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from mdimg_b200 import synth  # noqa: E402
from oracle import ref_enhancement as oenh  # noqa: E402
from oracle import ref_metrics as omet  # noqa: E402


def main() -> None:
    ims = {"clean64": synth.fixture_clean(), "noisy64": synth.fixture_noisy(),
           "lowc64": synth.fixture_low_contrast()}
    out = {"metrics": {}, "p_full": {}, "conventions": {}}
    plan = synth.plan_full()
    for name, im in ims.items():
        out["metrics"][name] = omet.compute_metrics(im)
        enh, labels = oenh.apply_enhancements_from_params(im, plan)
        out["p_full"][name] = {"labels": labels, "sum": float(enh.astype(np.float64).sum())}
        np.save(HERE / f"p_full_{name}.npy", enh)
    x = np.arange(10, dtype=np.float32)
    out["conventions"]["uniform7"] = ndi.uniform_filter(x, size=7).tolist()
    out["conventions"]["uniform16"] = ndi.uniform_filter(x, size=16).tolist()
    (HERE / "oracle_fixtures.json").write_text(json.dumps(out, indent=1))
    print("wrote", HERE / "oracle_fixtures.json")


if __name__ == "__main__":
    main()
