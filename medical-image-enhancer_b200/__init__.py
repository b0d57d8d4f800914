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


_DROP_IN_NAMES = {
    "metrics": ("compute_metrics", "detect_issues", "compute_validation", "compute_objective_score",
                "compute_niqe_approximation", "compute_edge_ratio", "THRESHOLDS"),
    "enhancement": ("apply_enhancements", "apply_enhancements_from_params", "ENHANCEMENT_PARAMS"),
    "dicom_io": ("normalize_image",),
}


def install_as_pipeline(reference_root=None) -> dict:
    """Put this package's drop-in functions behind the reference's import names so that its agents,
    tools and runner pick them up unchanged:

    * ``pipeline.metrics`` and ``pipeline.enhancement`` become the drop-in modules
      (``core_agents.py:16-17``, ``tools.py:21,106``, ``genai_agents.py:45-47`` import from them);
    * ``pipeline.dicom_io`` stays the reference's own module (DICOM parsing, report text) with its
      ``normalize_image`` (``dicom_io.py:84-91``, imported by ``runner.py:26``) replaced;
    * reference modules imported BEFORE this call hold their own bindings (``from ... import name``):
      those are rebound too;
    * safeguard warnings go to the logger ``pipeline.enhancement``, the name the reference logs
      them under (``enhancement.py:26``).

    ``reference_root``: directory holding the reference's ``pipeline`` package, appended to
    ``sys.path`` when given.  Returns ``{module name: [names aliased]}``."""
    import importlib
    import sys

    from . import engine
    from .pipeline import dicom_io, enhancement, metrics

    if reference_root is not None and str(reference_root) not in sys.path:
        sys.path.append(str(reference_root))
    done: dict = {}
    sys.modules["pipeline.metrics"] = metrics
    sys.modules["pipeline.enhancement"] = enhancement
    done["pipeline.metrics"] = list(_DROP_IN_NAMES["metrics"])
    done["pipeline.enhancement"] = list(_DROP_IN_NAMES["enhancement"])
    engine.set_logger_name("pipeline.enhancement")
    pkg = sys.modules.get("pipeline")
    if pkg is None:
        try:
            pkg = importlib.import_module("pipeline")
        except ImportError:
            pkg = None
    if pkg is not None:
        pkg.metrics = metrics
        pkg.enhancement = enhancement
        ref_dio = sys.modules.get("pipeline.dicom_io")
        if ref_dio is None:
            try:
                ref_dio = importlib.import_module("pipeline.dicom_io")   # needs pydicom: the host's business
            except Exception:  # noqa: BLE001
                ref_dio = None
        if ref_dio is not None and ref_dio is not dicom_io:
            ref_dio.normalize_image = dicom_io.normalize_image
            done["pipeline.dicom_io"] = ["normalize_image"]
    # modules that imported the names before this call
    ours = {"pipeline.metrics": metrics, "pipeline.enhancement": enhancement, "pipeline.dicom_io": dicom_io}
    for modname, mod in list(sys.modules.items()):
        if mod is None or not modname.startswith("pipeline.") or mod in ours.values() or modname == "pipeline.dicom_io":
            continue
        for short, names in _DROP_IN_NAMES.items():
            src = ours[f"pipeline.{short}"]
            for name in names:
                cur = getattr(mod, name, None)
                if cur is None or cur is getattr(src, name):
                    continue
                if getattr(cur, "__module__", None) == f"pipeline.{short}" or (
                        not callable(cur) and name in ("THRESHOLDS", "ENHANCEMENT_PARAMS")):
                    setattr(mod, name, getattr(src, name))
                    done.setdefault(modname, []).append(name)
    return done
