"""Parameter contract of the enhancement path — mirrors ``pipeline/schemas.py:16-116`` of the
reference (``PARAM_BOUNDS``, ``EnhancementParams``, ``EnhancementPlan``).  Only the fields the hot
path reads are modelled; field names, defaults and types are the reference's."""

from __future__ import annotations

from typing import List, Optional

from pydantic import BaseModel, Field

from ..engine import PARAM_BOUNDS  # noqa: F401  (re-exported: the safety clamps)


class EnhancementParams(BaseModel):
    """Tunable parameters; every value is clamped to ``PARAM_BOUNDS`` before execution."""

    clahe_clip_limit: float = Field(default=0.015, description="CLAHE clip limit, 0.002-0.08")
    clahe_tile_size: int = Field(default=16, description="CLAHE kernel size in pixels, 4-48")
    gamma: float = Field(default=1.0, description="gamma exponent, 0.6-1.5 (<1 brightens)")
    unsharp_radius: float = Field(default=0.8, description="unsharp Gaussian sigma, 0.2-3.0")
    unsharp_amount: float = Field(default=0.5, description="unsharp strength, 0.03-2.5")
    denoise_mode: str = Field(default="soft", description="wavelet shrinkage: 'soft' or 'hard'")
    post_denoise_strength: float = Field(default=0.3, description="light-denoise blend, 0-0.8 (0 = off)")
    bilateral_d: int = Field(default=0, description="bilateral diameter, 0 = off, up to 13")
    bilateral_sigma_color: float = Field(default=0.05, description="bilateral range sigma, 0.005-0.20")
    bilateral_sigma_space: float = Field(default=0.05, description="bilateral spatial sigma, 0.005-0.20")
    tv_denoise_weight: float = Field(default=0.0, description="TV-Chambolle weight, 0 = off, up to 0.15")


class EnhancementPlan(BaseModel):
    """Ordered operation list + parameters (the planner's structured output in the reference)."""

    recommended_ops: List[str] = Field(
        description="subset of: denoise, clahe, gamma, unsharp, post_denoise, bilateral, tv_denoise")
    params: EnhancementParams = Field(default_factory=EnhancementParams)
    risk_warnings: List[str] = Field(default_factory=list)
    rationale: str = Field(default="")
    safety: str = Field(default="")
    stop_reason: Optional[str] = Field(default=None)
