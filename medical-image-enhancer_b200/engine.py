"""Device-side execution of the reference's enhancement and validation control flow over a stack.

The step order, gating, clamping, safeguards and op-label strings follow
``pipeline/enhancement.py`` (``apply_enhancements`` :151-227, ``apply_enhancements_from_params``
:235-369) and ``pipeline/metrics.py`` (``compute_validation`` :225-329) of the reference; the pixels
never leave the GPU.  Each safeguard costs one small device->host read of per-slice scalars, after
which only the flagged slices are re-processed (``sel`` lists, no pixel gathers).
"""

from __future__ import annotations

import logging
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .stack import StackOps

logger = logging.getLogger("mdimg_b200.enhancement")


def set_logger_name(name: str) -> None:
    """Route the safeguard warnings to another logger: ``install_as_pipeline`` selects the reference's
    ``pipeline.enhancement`` (pipeline/enhancement.py:26) so a host filtering by that name keeps them."""
    global logger
    logger = logging.getLogger(name)

# pipeline/schemas.py:16-28
PARAM_BOUNDS: Dict[str, Tuple[float, float]] = {
    "clahe_clip_limit": (0.002, 0.08),
    "clahe_tile_size": (4, 48),
    "gamma": (0.6, 1.5),
    "unsharp_radius": (0.2, 3.0),
    "unsharp_amount": (0.03, 2.5),
    "post_denoise_strength": (0.0, 0.8),
    "bilateral_d": (0, 13),
    "bilateral_sigma_color": (0.005, 0.20),
    "bilateral_sigma_space": (0.005, 0.20),
    "tv_denoise_weight": (0.0, 0.15),
}

# pipeline/enhancement.py:32-42
ENHANCEMENT_PARAMS = {
    "clahe_clip_limit": 0.015,
    "clahe_tile_size": 16,
    "gamma_brighten": 0.95,
    "gamma_darken": 1.05,
    "unsharp_radius": 0.8,
    "unsharp_amount": 0.5,
    "denoise_sigma": None,
    "denoise_wavelet_mode": "soft",
    "post_denoise_strength": 0.3,
}

# pipeline/metrics.py:25-34
THRESHOLDS = {
    "noise_sigma": 0.08,
    "blur_lap_var": 0.001,
    "low_contrast_std": 0.12,
    "clip_pct": 0.01,
    "ssim": 0.70,
    "psnr": 22.0,
    "quality_improvement": 0.10,
}

METRIC_KEYS = (
    "sigma", "lap_var", "std", "pct_low", "pct_high", "entropy", "edge_density",
    "gradient_mag_mean", "gradient_mag_std", "snr_proxy", "cnr_proxy", "laplacian_energy",
    "histogram_spread", "local_contrast_std", "gradient_strength", "gradient_entropy",
)
MC_MEAN, MC_EDGE_RATIO, MC_NIQE = 16, 17, 18

HALO_MSG = "Halo detected (edge_ratio > 1.5) — re-applying with halved unsharp_amount."
NOISE_MSG = "Noise amplification detected — applying corrective denoise."
OVER_MSG = "Over-processing detected (NIQE degraded >0.5). Blending back."


@dataclass
class ClampedParams:
    """The ten PARAM_BOUNDS-clamped numbers + denoise mode (pipeline/enhancement.py:249-263)."""
    clip_limit: float
    tile_size: int
    gamma: float
    u_radius: float
    u_amount: float
    dn_mode: str
    post_str: float
    bilateral_d: int
    bilateral_sc: float
    bilateral_ss: float
    tv_weight: float

    @classmethod
    def from_params(cls, p) -> "ClampedParams":
        def c(name: str):
            lo, hi = PARAM_BOUNDS[name]
            return max(lo, min(hi, getattr(p, name)))

        return cls(
            clip_limit=c("clahe_clip_limit"), tile_size=int(c("clahe_tile_size")), gamma=c("gamma"),
            u_radius=c("unsharp_radius"), u_amount=c("unsharp_amount"),
            dn_mode=p.denoise_mode if p.denoise_mode in ("soft", "hard") else "soft",
            post_str=c("post_denoise_strength"), bilateral_d=int(c("bilateral_d")),
            bilateral_sc=c("bilateral_sigma_color"), bilateral_ss=c("bilateral_sigma_space"),
            tv_weight=c("tv_denoise_weight"),
        )


@dataclass
class EnhanceResult:
    image: torch.Tensor                      # [N, H, W] float32 in [0, 1]
    labels: List[List[str]]                  # applied-op labels per slice
    halo: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    noise_guard: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    over_processed: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    tv_iterations: Optional[np.ndarray] = None
    errors: Dict[int, str] = field(default_factory=dict)   # slice -> ValueError text (on_error='flag')
    sigma_before: Optional[torch.Tensor] = None     # device float64 [N], estimate_sigma(original)
    quality_before: Optional[torch.Tensor] = None   # device float64 [N, 2] (edge_ratio, niqe) of the original
    rows_after: Optional[torch.Tensor] = None       # device float64 [N, 24]: mdimg_metrics rows of `image`


_STEP_ORDER = ("denoise", "clahe", "gamma", "unsharp", "post_denoise", "bilateral", "tv_denoise")


class Engine:
    """Runs enhancement plans and validation on [N, H, W] float32 CUDA stacks."""

    def __init__(self, ops: StackOps):
        self.ops = ops

    # ---- helpers ----------------------------------------------------------------------------
    def _sel_tensor(self, mask: np.ndarray) -> Optional[torch.Tensor]:
        idx = np.flatnonzero(mask).astype(np.int32)
        if idx.size == 0:
            return None
        return torch.from_numpy(idx).to(self.ops.device)

    @staticmethod
    def _raise_if(flags: torch.Tensor, message: str, state: Optional[dict] = None) -> None:
        """ValueError parity with the reference.  With a `state`, the device flag is parked and
        checked at the next host synchronisation point (`_flush_checks`) instead of forcing one."""
        if state is not None:
            state.setdefault("checks", []).append((flags, message))
            return
        if bool(flags.any().item()):
            raise ValueError(message)

    @staticmethod
    def _flush_checks(state: dict) -> None:
        """Raise the reference's ValueError for the first failed check, or — in a stack call with
        ``on_error='flag'`` — remember which slices failed and carry on with the others."""
        for flags, message in state.pop("checks", []):
            bad = flags.cpu().numpy() != 0
            if not bad.any():
                continue
            if state.get("on_error", "raise") == "raise":
                raise ValueError(message)
            errs = state.setdefault("errors", {})
            for i in np.flatnonzero(bad):
                errs.setdefault(int(i), message)

    def _apply_step(self, name: str, q: ClampedParams, u_amount: float, cur: torch.Tensor,
                    tmp: torch.Tensor, sel: Optional[torch.Tensor], state: dict) -> Tuple[torch.Tensor, torch.Tensor]:
        """Apply one step to the slices in `sel` (all when None).  Returns (cur, tmp) — possibly
        swapped when the whole stack was written out of place."""
        ops = self.ops
        if name == "denoise":
            ops.wavelet_denoise(cur, cur, mode=q.dn_mode, sel=sel)
            state["nonneg"] = False
        elif name == "clahe":
            # an adjust_gamma that directly follows CLAHE is folded into CLAHE's final table pass
            g = q.gamma if state.get("fuse_gamma") else 1.0
            status = ops.clahe(cur, cur, q.clip_limit, q.tile_size, sel=sel, gamma=g)
            self._raise_if(status, "Images of type float must be between -1 and 1.", state)
            state["nonneg"] = True
            state["gamma_done"] = g != 1.0
        elif name == "gamma":
            if state.pop("gamma_done", False):      # already applied inside the CLAHE call
                return cur, tmp
            neg = ops.gamma(cur, cur, q.gamma, assume_nonneg=state["nonneg"], sel=sel)
            if not state["nonneg"]:
                self._raise_if(neg, "Image Correction methods work correctly only on images with "
                                    "non-negative values. Use skimage.exposure.rescale_intensity.", state)
            state["nonneg"] = True
        elif name == "unsharp":
            ops.unsharp(cur, tmp, q.u_radius, u_amount, assume_nonneg=state["nonneg"], sel=sel)
            if sel is None:
                cur, tmp = tmp, cur
            else:
                ops.copy(tmp, cur, sel=sel)
            # result is clipped to [0,1] when the input had no negative pixel, else to [-1,1]
        elif name == "post_denoise":
            ops.light_denoise(cur, cur, q.post_str, sel=sel)
            state["nonneg"] = False
        elif name == "bilateral":
            ops.bilateral(cur, tmp, q.bilateral_d, q.bilateral_sc, q.bilateral_ss, sel=sel)
            if sel is None:
                cur, tmp = tmp, cur
            else:
                ops.copy(tmp, cur, sel=sel)
        elif name == "tv_denoise":
            # out of place: the result pass writes the image directly while neighbouring strips still
            # read the input (an aliased output costs one more copy inside the library)
            iters = ops.tv_chambolle(cur, tmp, q.tv_weight, sel=sel)
            if sel is None:
                cur, tmp = tmp, cur
            else:
                ops.copy(tmp, cur, sel=sel)
            state["tv_iters"] = iters
            state["nonneg"] = False
        return cur, tmp

    @staticmethod
    def _enabled(name: str, q: ClampedParams) -> bool:
        if name == "gamma":
            return abs(q.gamma - 1.0) > 1e-4
        if name == "post_denoise":
            return q.post_str > 0
        if name == "bilateral":
            return q.bilateral_d > 0
        if name == "tv_denoise":
            return q.tv_weight > 0
        return name in _STEP_ORDER

    @staticmethod
    def _label(name: str, q: ClampedParams) -> str:
        if name == "denoise":
            return f"Wavelet denoise (pre, mode={q.dn_mode})"
        if name == "clahe":
            return f"CLAHE (clip={q.clip_limit:.4f}, tile={q.tile_size})"
        if name == "gamma":
            return f"Gamma {'brighten' if q.gamma < 1.0 else 'darken'} ({q.gamma:.3f})"
        if name == "unsharp":
            return f"Unsharp mask (r={q.u_radius:.2f}, a={q.u_amount:.2f})"
        if name == "post_denoise":
            return f"Light denoise (post, s={q.post_str:.2f})"
        if name == "bilateral":
            return f"Bilateral (d={q.bilateral_d}, sc={q.bilateral_sc:.3f}, ss={q.bilateral_ss:.3f})"
        return f"TV denoise (w={q.tv_weight:.4f})"

    def _noise_guard(self, image: torch.Tensor, cur: torch.Tensor, labels: List[List[str]],
                     sigma_before: Optional[torch.Tensor]) -> Tuple[np.ndarray, torch.Tensor]:
        """_check_noise_amplification + corrective light denoise (enhancement.py:55-63,221-225,356-360)."""
        ops = self.ops
        if sigma_before is None:
            sigma_before = ops.estimate_sigma(image)
        s1 = ops.estimate_sigma(cur)
        s0h = sigma_before.cpu().numpy()
        s1h = s1.cpu().numpy()
        with np.errstate(invalid="ignore"):
            fired = (~(s0h < 1e-8)) & (s1h > s0h * 1.3)
        sel = self._sel_tensor(fired)
        if sel is not None:
            logger.warning(NOISE_MSG)
            ops.light_denoise(cur, cur, 0.4, sel=sel)
            ops.clip01(cur, cur, sel=sel)
            for i in np.flatnonzero(fired):
                labels[i].append("Auto-corrective denoise (noise guard)")
        return fired, sigma_before

    # ---- apply_enhancements_from_params -------------------------------------------------------
    def enhance_from_params(self, image: torch.Tensor, plan, *, sigma_before: Optional[torch.Tensor] = None,
                            quality_before: Optional[torch.Tensor] = None,
                            rows_before: Optional[torch.Tensor] = None,
                            on_error: str = "raise") -> EnhanceResult:
        """on_error='raise' mirrors the reference (ValueError when CLAHE sees a pixel outside [-1, 1]
        or gamma a negative one); 'flag' is for stacks: the failing slices are returned unchanged
        (a copy of the input) with their error text in ``EnhanceResult.errors``."""
        ops = self.ops
        n = image.shape[0]
        q = ClampedParams.from_params(plan.params)
        plan_ops = [op.lower().strip() for op in plan.recommended_ops]
        cur = image.clone()
        tmp = torch.empty_like(image)
        state = {"nonneg": False, "tv_iters": None, "on_error": on_error,
                 # fixed step order: "gamma" always comes right after "clahe" when both are enabled
                 "fuse_gamma": "clahe" in plan_ops and "gamma" in plan_ops and self._enabled("gamma", q)}
        common: List[str] = []
        for name in _STEP_ORDER:   # fixed order, gated by membership
            if name in plan_ops and self._enabled(name, q):
                cur, tmp = self._apply_step(name, q, q.u_amount, cur, tmp, None, state)
                common.append(self._label(name, q))
        ops.clip01(cur, cur)
        self._flush_checks(state)
        labels = [list(common) for _ in range(n)]
        tv_iters = state["tv_iters"]

        # ---- safeguards -------------------------------------------------------------------------
        # One fused metrics pass over the current stack yields everything the three guards look at
        # (edge ratio, estimate_sigma, NIQE approximation) AND the compute_metrics rows of the final
        # image; only slices a guard actually modifies are measured again.
        if rows_before is not None:
            s0h = rows_before[:, 0].cpu().numpy()
            nb = rows_before[:, MC_NIQE].cpu().numpy()
        else:
            if sigma_before is None:
                sigma_before = ops.estimate_sigma(image)
            if quality_before is None:
                quality_before = ops.quality(image, niqe=True)
            s0h = sigma_before.cpu().numpy()
            nb = quality_before[:, 1].cpu().numpy()
        rows = ops.metrics(cur, with_niqe=True)

        def remeasure(mask: np.ndarray) -> None:
            sel_ = self._sel_tensor(mask)
            if sel_ is not None:
                ops.metrics(cur, with_niqe=True, sel=sel_, out=rows)

        # A slice a guard modifies is measured again -- but only for what the NEXT guard looks at
        # (estimate_sigma after the halo re-run, the NIQE approximation after that / the noise
        # correction); its full compute_metrics row is produced once, after the last guard.
        rows_h = rows.cpu().numpy()
        sig1 = rows_h[:, 0].copy()               # estimate_sigma(enhanced)
        niqe1 = rows_h[:, MC_NIQE].copy()        # compute_niqe_approximation(enhanced)
        dirty = np.zeros(n, bool)                # modified since `rows` was computed
        halo = np.zeros(n, bool)
        if "unsharp" in plan_ops:   # _check_halo -> re-run in the plan's own order with amount / 2
            halo = rows_h[:, MC_EDGE_RATIO] > 1.5
            sel = self._sel_tensor(halo)
            if sel is not None:
                logger.warning(HALO_MSG)
                reduced = q.u_amount * 0.5
                ops.copy(image, cur, sel=sel)
                st2 = {"nonneg": False, "tv_iters": None, "on_error": on_error}
                for op in plan_ops:
                    if op in _STEP_ORDER and self._enabled(op, q):
                        cur, tmp = self._apply_step(op, q, reduced, cur, tmp, sel, st2)
                ops.clip01(cur, cur, sel=sel)
                self._flush_checks(st2)
                state.setdefault("errors", {}).update(st2.get("errors", {}))
                for i in np.flatnonzero(halo):
                    labels[i].append(f"[safeguard] Unsharp reduced to {reduced:.2f}")
                if st2["tv_iters"] is not None and tv_iters is not None:
                    tv_iters = torch.where(torch.from_numpy(halo).to(ops.device), st2["tv_iters"], tv_iters)
                dirty |= halo
                sig1[halo] = ops.estimate_sigma(cur, sel=sel).cpu().numpy()[halo]

        # _check_noise_amplification -> corrective light denoise (enhancement.py:55-63,356-360)
        with np.errstate(invalid="ignore"):
            noise = (~(s0h < 1e-8)) & (sig1 > s0h * 1.3)
        sel = self._sel_tensor(noise)
        if sel is not None:
            logger.warning(NOISE_MSG)
            ops.light_denoise(cur, cur, 0.4, sel=sel)
            ops.clip01(cur, cur, sel=sel)
            for i in np.flatnonzero(noise):
                labels[i].append("Auto-corrective denoise (noise guard)")
            dirty |= noise

        # _check_over_processing: NIQE-approx degradation > 0.5 -> 0.6*enhanced + 0.4*original
        sel = self._sel_tensor(dirty)
        if sel is not None:
            niqe1[dirty] = ops.quality(cur, niqe=True, sel=sel)[:, 1].cpu().numpy()[dirty]
        with np.errstate(invalid="ignore"):
            over = (niqe1 - nb) > 0.5
        sel = self._sel_tensor(over)
        if sel is not None:
            logger.warning(OVER_MSG)
            ops.axpby(cur, image, cur, 0.6, 0.4, clip01=True, sel=sel)
            for i in np.flatnonzero(over):
                labels[i].append("Blend-back 40% original (over-processing guard)")
            dirty |= over
        remeasure(dirty)

        errors = state.get("errors", {})
        if errors:
            bad = np.zeros(n, bool)
            bad[list(errors)] = True
            ops.copy(image, cur, sel=self._sel_tensor(bad))
            remeasure(bad)
            for i, msg in errors.items():
                labels[i] = [f"ERROR: {msg}"]
        return EnhanceResult(
            image=cur, labels=labels, halo=halo, noise_guard=noise, over_processed=over, errors=errors,
            tv_iterations=None if tv_iters is None else tv_iters.cpu().numpy(),
            sigma_before=sigma_before, quality_before=quality_before, rows_after=rows,
        )

    def enhance_plan(self, image: torch.Tensor, plan, *, rows_before: Optional[torch.Tensor] = None,
                     on_error: str = "raise") -> EnhanceResult:
        """`enhance_from_params_native` (one library call) unless MDIMG_NATIVE_ENGINE=0 selects the torch-side
        control flow; the two are interchangeable (same pixels, flags, labels)."""
        from . import _lib
        known = sum(op.lower().strip() in _lib.STEP_NAMES for op in plan.recommended_ops)
        # a list longer than the native call's ops[] capacity (repeats count: the halo safeguard replays
        # the list as written) runs through the torch-side flow, which has no such limit
        if os.environ.get("MDIMG_NATIVE_ENGINE", "1") != "0" and known <= _lib.MAX_PLAN_OPS:
            return self.enhance_from_params_native(image, plan, rows_before=rows_before, on_error=on_error)
        return self.enhance_from_params(image, plan, rows_before=rows_before, on_error=on_error)

    def enhance_from_params_native(self, image: torch.Tensor, plan, *, rows_before: Optional[torch.Tensor] = None,
                                   on_error: str = "raise") -> EnhanceResult:
        """The same call through `mdimg_enhance`: the control flow above runs in C++ inside the library
        (one C-ABI call per stack, what a non-Python host binds); labels are rebuilt here from the
        returned flags.  Results equal `enhance_from_params` (tests/test_gpu_dropin.py)."""
        from . import _lib
        out, rows_after, flags, iters, qc = self.ops.enhance(image, plan, rows_before=rows_before)
        q = ClampedParams.from_params(plan.params)
        plan_ops = [op.lower().strip() for op in plan.recommended_ops]
        common = [self._label(nm, q) for nm in _STEP_ORDER if nm in plan_ops and self._enabled(nm, q)]
        n = image.shape[0]
        labels: List[List[str]] = []
        errors: Dict[int, str] = {}
        for i in range(n):
            f = int(flags[i])
            if f & (_lib.FLAG_ERR_CLAHE_RANGE | _lib.FLAG_ERR_GAMMA_NEG):
                msg = ("Images of type float must be between -1 and 1." if f & _lib.FLAG_ERR_CLAHE_RANGE else
                       "Image Correction methods work correctly only on images with non-negative values. "
                       "Use skimage.exposure.rescale_intensity.")
                if on_error == "raise":
                    raise ValueError(msg)
                errors[i] = msg
                labels.append([f"ERROR: {msg}"])
                continue
            lab = list(common)
            if f & _lib.FLAG_HALO:
                lab.append(f"[safeguard] Unsharp reduced to {q.u_amount * 0.5:.2f}")
            if f & _lib.FLAG_NOISE_GUARD:
                lab.append("Auto-corrective denoise (noise guard)")
            if f & _lib.FLAG_OVER_PROCESSED:
                lab.append("Blend-back 40% original (over-processing guard)")
            labels.append(lab)
        for bit, msg in ((_lib.FLAG_HALO, HALO_MSG), (_lib.FLAG_NOISE_GUARD, NOISE_MSG), (_lib.FLAG_OVER_PROCESSED, OVER_MSG)):
            if (flags & bit).any():
                logger.warning(msg)                 # the reference logs these (pipeline/enhancement.py:320,357,363)
        ran_tv = "tv_denoise" in plan_ops and self._enabled("tv_denoise", q)
        return EnhanceResult(image=out, labels=labels, halo=(flags & _lib.FLAG_HALO) != 0,
                             noise_guard=(flags & _lib.FLAG_NOISE_GUARD) != 0,
                             over_processed=(flags & _lib.FLAG_OVER_PROCESSED) != 0, errors=errors,
                             tv_iterations=iters.copy() if ran_tv else None, rows_after=rows_after)

    def enhance_issues(self, image: torch.Tensor, issues: Sequence[str], *,
                       sigma_before: Optional[torch.Tensor] = None) -> EnhanceResult:
        """apply_enhancements through `mdimg_enhance_issues` (one library call; MDIMG_NATIVE_ENGINE=0 selects
        `enhance_from_issues`, the same flow on torch tensors).  Raises the reference's ValueError when a
        slice would have raised it."""
        if os.environ.get("MDIMG_NATIVE_ENGINE", "1") == "0":
            return self.enhance_from_issues(image, issues, sigma_before=sigma_before)
        out, flags = self.ops.enhance_issues(image, issues, sigma_before=sigma_before)
        if (flags & _lib.FLAG_ERR_CLAHE_RANGE).any():
            raise ValueError("Images of type float must be between -1 and 1.")
        if (flags & _lib.FLAG_ERR_GAMMA_NEG).any():
            raise ValueError("Image Correction methods work correctly only on images with "
                             "non-negative values. Use skimage.exposure.rescale_intensity.")
        P = ENHANCEMENT_PARAMS
        has = set(issues).__contains__
        common: List[str] = []
        if has("noise"):
            common.append("Wavelet denoise (pre)")
        if has("low_contrast") or has("clipping_low") or has("clipping_high"):
            common.append(f"CLAHE (clip={P['clahe_clip_limit']}, tile={P['clahe_tile_size']})")
        if has("clipping_low") and not has("clipping_high"):
            common.append(f"Gamma brighten ({P['gamma_brighten']})")
        elif has("clipping_high") and not has("clipping_low"):
            common.append(f"Gamma darken ({P['gamma_darken']})")
        if has("blur"):
            common.append(f"Unsharp mask (r={P['unsharp_radius']}, a={P['unsharp_amount']})")
            if P["post_denoise_strength"] > 0:
                common.append(f"Light denoise (post, s={P['post_denoise_strength']})")
        noise = (flags & _lib.FLAG_NOISE_GUARD) != 0
        if noise.any():
            logger.warning(NOISE_MSG)
        labels = [list(common) + (["Auto-corrective denoise (noise guard)"] if noise[i] else [])
                  for i in range(image.shape[0])]
        n = image.shape[0]
        return EnhanceResult(image=out, labels=labels, noise_guard=noise, halo=np.zeros(n, bool),
                             over_processed=np.zeros(n, bool))

    # ---- apply_enhancements (issue-gated defaults) ------------------------------------------------
    def enhance_from_issues(self, image: torch.Tensor, issues: Sequence[str], *,
                            sigma_before: Optional[torch.Tensor] = None) -> EnhanceResult:
        ops = self.ops
        P = ENHANCEMENT_PARAMS
        n = image.shape[0]
        has = set(issues).__contains__
        cur = image.clone()
        tmp = torch.empty_like(image)
        common: List[str] = []
        nonneg = False
        if has("noise"):
            ops.wavelet_denoise(cur, cur, mode=P["denoise_wavelet_mode"])
            common.append("Wavelet denoise (pre)")
        if has("low_contrast") or has("clipping_low") or has("clipping_high"):
            k = P["clahe_tile_size"]
            status = ops.clahe(cur, cur, P["clahe_clip_limit"], k)
            self._raise_if(status, "Images of type float must be between -1 and 1.")
            common.append(f"CLAHE (clip={P['clahe_clip_limit']}, tile={k})")
            nonneg = True
        g = None
        if has("clipping_low") and not has("clipping_high"):
            g, lab = P["gamma_brighten"], f"Gamma brighten ({P['gamma_brighten']})"
        elif has("clipping_high") and not has("clipping_low"):
            g, lab = P["gamma_darken"], f"Gamma darken ({P['gamma_darken']})"
        if g is not None:
            neg = ops.gamma(cur, cur, g, assume_nonneg=nonneg)
            if not nonneg:
                self._raise_if(neg, "Image Correction methods work correctly only on images with "
                                    "non-negative values. Use skimage.exposure.rescale_intensity.")
            common.append(lab)
            nonneg = True
        if has("blur"):
            ops.unsharp(cur, tmp, P["unsharp_radius"], P["unsharp_amount"], assume_nonneg=nonneg)
            cur, tmp = tmp, cur
            common.append(f"Unsharp mask (r={P['unsharp_radius']}, a={P['unsharp_amount']})")
        if has("blur") and P["post_denoise_strength"] > 0:
            ops.light_denoise(cur, cur, P["post_denoise_strength"])
            common.append(f"Light denoise (post, s={P['post_denoise_strength']})")
        ops.clip01(cur, cur)
        labels = [list(common) for _ in range(n)]
        noise, sigma_before = self._noise_guard(image, cur, labels, sigma_before)
        return EnhanceResult(image=cur, labels=labels, noise_guard=noise, sigma_before=sigma_before,
                             halo=np.zeros(n, bool), over_processed=np.zeros(n, bool))

    # ---- metrics / validation -------------------------------------------------------------------
    def metrics_rows(self, image: torch.Tensor, with_niqe: bool = False) -> torch.Tensor:
        return self.ops.metrics(image, with_niqe=with_niqe)

    def validation_rows(self, original: torch.Tensor, enhanced: torch.Tensor,
                        rows_before: Optional[torch.Tensor] = None,
                        rows_after: Optional[torch.Tensor] = None):
        """Device rows needed by compute_validation: metrics (with NIQE / edge ratio) of both
        stacks and (ssim, psnr).  Metrics of the same stack are computed once (the reference
        recomputes identical values, pipeline/metrics.py:229-236,272)."""
        if rows_before is None and rows_after is None:      # one library call (mdimg_validation)
            v = self.ops.validation(original, enhanced)
            k = v.shape[1] // 2 - 1
            return v[:, :k], v[:, k:2 * k], v[:, 2 * k:]
        if rows_before is None:
            rows_before = self.ops.metrics(original, with_niqe=True)
        if rows_after is None:
            rows_after = self.ops.metrics(enhanced, with_niqe=True)
        fr = self.ops.fullref(original, enhanced)
        return rows_before, rows_after, fr


def metrics_dict(row: np.ndarray) -> Dict[str, float]:
    """One result row -> the reference's compute_metrics dict (python floats, same key order)."""
    return {k: float(row[i]) for i, k in enumerate(METRIC_KEYS)}


def validation_dict(mb: Dict[str, float], ma: Dict[str, float], ssim: float, psnr: float,
                    niqe_before: float, niqe_after: float, edge_ratio_after: float) -> Dict[str, object]:
    """Scalar part of compute_validation (pipeline/metrics.py:237-329), python-float arithmetic."""
    niqe_ok = niqe_after <= niqe_before
    eps = 1e-8
    contrast_gain = (ma["std"] - mb["std"]) / max(mb["std"], eps)
    sharpness_gain = (ma["lap_var"] - mb["lap_var"]) / max(mb["lap_var"], eps)
    noise_reduction = (mb["sigma"] - ma["sigma"]) / max(mb["sigma"], eps)
    qi = float(0.35 * contrast_gain + 0.35 * sharpness_gain + 0.30 * noise_reduction)
    ok_ssim = ssim >= THRESHOLDS["ssim"]
    ok_psnr = psnr >= THRESHOLDS["psnr"]
    ok_gain = qi >= THRESHOLDS["quality_improvement"]
    passes = (ok_ssim and ok_psnr) or (ok_ssim and ok_gain) or (ok_psnr and ok_gain and niqe_ok)
    out: Dict[str, object] = {
        "ssim": ssim, "psnr": psnr, "quality_improvement": qi,
        "meets_ssim": ok_ssim, "meets_psnr": ok_psnr, "meets_improvement": ok_gain, "passes": passes,
        "niqe_before": niqe_before, "niqe_after": niqe_after, "niqe_improved": niqe_ok,
        "contrast_gain": contrast_gain, "sharpness_gain": sharpness_gain, "noise_change": -noise_reduction,
    }
    for stem, key in (("entropy", "entropy"), ("snr", "snr_proxy"), ("cnr", "cnr_proxy")):
        out[f"{stem}_before"], out[f"{stem}_after"] = mb[key], ma[key]
        out[f"{stem}_change"] = ma[key] - mb[key]
    out["edge_density_change"] = ma["edge_density"] - mb["edge_density"]
    out["histogram_spread_change"] = ma["histogram_spread"] - mb["histogram_spread"]
    out["laplacian_energy_before"] = mb["laplacian_energy"]
    out["laplacian_energy_after"] = ma["laplacian_energy"]
    out["edge_ratio"] = edge_ratio_after
    for stem, key in (("local_contrast", "local_contrast_std"), ("gradient_strength", "gradient_strength"),
                      ("gradient_entropy", "gradient_entropy")):
        out[f"{stem}_before"], out[f"{stem}_after"] = mb[key], ma[key]
        out[f"{stem}_change"] = ma[key] - mb[key]
    out["metrics_before"] = mb
    out["metrics_after"] = ma
    return out


def objective_score(validation: dict) -> Tuple[float, dict]:
    """compute_objective_score (pipeline/metrics.py:337-408): scalar host arithmetic."""
    def f(key: str) -> float:
        return float(validation.get(key, 0))

    def capped(x: float, cap: float) -> float:
        return max(0.0, min(x, cap))

    passes = bool(validation.get("passes", False))
    parts = {
        "contrast_gain": f("contrast_gain"),
        "sharpness_gain": f("sharpness_gain"),
        "noise_penalty": max(0.0, f("noise_change")),
        "niqe_degradation": max(0.0, f("niqe_after") - f("niqe_before")),
        "halo_penalty": max(0.0, f("edge_ratio") - 1.0) * 5.0,
        "entropy_penalty": max(0.0, abs(f("entropy_change")) - 0.5) * 2.0,
        "snr_reward": capped(f("snr_change") * 0.1, 0.5),
        "hs_reward": capped(f("histogram_spread_change") * 0.5, 0.3),
        "local_contrast_reward": capped(f("local_contrast_change") * 0.3, 0.3),
        "gradient_strength_reward": capped(f("gradient_strength_change") * 0.2, 0.2),
        "gradient_entropy_penalty": max(0.0, abs(f("gradient_entropy_change")) - 0.3) * 1.5,
    }
    score = (
        0.35 * parts["contrast_gain"] + 0.35 * parts["sharpness_gain"] - 0.30 * parts["noise_penalty"]
        - 5.0 * parts["niqe_degradation"] - 10.0 * (0 if passes else 1) - parts["halo_penalty"]
        - parts["entropy_penalty"] + parts["snr_reward"] + parts["hs_reward"]
        + parts["local_contrast_reward"] + parts["gradient_strength_reward"]
        - parts["gradient_entropy_penalty"]
    )
    breakdown = {k: round(v, 4) for k, v in parts.items()}
    breakdown["passes"] = passes
    return round(float(score), 4), breakdown
