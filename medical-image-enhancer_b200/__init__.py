"""mdimg_b200 — B200-native (sm_100a) implementation of the MDIMG deterministic image hot path.

Drop-in modules mirroring the reference's ``pipeline`` package (same function names, argument
meaning, return types and error behaviour):

    from mdimg_b200.pipeline.metrics import compute_metrics, compute_validation, ...
    from mdimg_b200.pipeline.enhancement import apply_enhancements, apply_enhancements_from_params
    from mdimg_b200.pipeline.dicom_io import normalize_image
    from mdimg_b200.pipeline.schemas import PARAM_BOUNDS, EnhancementPlan, EnhancementParams

plus the stack API for batches of slices resident on the GPU (``mdimg_b200.batch``).

All pixel arithmetic runs in hand-written CUDA kernels behind the C ABI of ``libmdimg_b200.so``
(``include/mdimg_b200.h``).  There is no CPU fallback.
"""

__version__ = "0.1.0"

from ._lib import LIB_PATH, MdimgError  # noqa: F401


def install_as_pipeline() -> None:
    """Alias this package's drop-in modules over ``pipeline.metrics`` / ``pipeline.enhancement`` /
    ``pipeline.dicom_io.normalize_image`` so the reference's agents and runner pick them up
    unchanged."""
    import sys

    from .pipeline import enhancement, metrics

    sys.modules["pipeline.metrics"] = metrics
    sys.modules["pipeline.enhancement"] = enhancement
    pkg = sys.modules.get("pipeline")
    if pkg is not None:
        pkg.metrics = metrics
        pkg.enhancement = enhancement
