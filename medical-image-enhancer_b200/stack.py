"""Stack-level operators: one call = one hot-path step over an [N, H, W] CUDA tensor of slices.

PyTorch is the host container only (device memory, streams); every computation goes through the
C ABI into the hand-written sm_100a kernels.  Per-slice scalars (sigma, min/max, safeguard
inputs) come back as small device tensors; nothing here touches pixels on the host.
"""

from __future__ import annotations

import ctypes as C
import threading
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check, load_library

_tls = threading.local()


def percentile_plan(n: int, qs: Sequence[float] = (5, 25, 75, 95, 90)) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """numpy's ``np.percentile(a, q)`` plan ('linear' method) for a float32 array of ``n``
    elements: previous index, next index and float32 interpolation weight per q.

    Mirrors numpy >= 2 (``lib/_function_base_impl.py``): ``q / float32(100)``, virtual index
    ``(n - 1) * q`` in float32, ``floor``; indices at or above ``n - 1`` collapse to the last
    element (numpy writes -1)."""
    lo = np.zeros(len(qs), np.int32)
    hi = np.zeros(len(qs), np.int32)
    gamma = np.zeros(len(qs), np.float32)
    for k, q in enumerate(qs):
        # explicit float32 scalars: no dependence on NEP 50 promotion or on ufunc keyword forms
        quant = np.float32(q) / np.float32(100)
        virt = np.float32(n - 1) * quant
        prev = np.floor(virt)
        if virt >= n - 1:
            prev_i, next_i = n - 1, n - 1
        elif virt < 0:
            prev_i, next_i = 0, 0
        else:
            prev_i, next_i = int(prev), int(prev) + 1
        g = virt - np.float32(int(prev))
        lo[k], hi[k], gamma[k] = prev_i, next_i, np.float32(g)
    return lo, hi, gamma


def gaussian_taps(sigma: float, truncate: float = 4.0) -> np.ndarray:
    """scipy.ndimage ``_gaussian_kernel1d(sigma, 0, radius)``; returns w[0..radius] (centre first)."""
    radius = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x**2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64)


def bilateral_spatial(d: int, sigma_space: float) -> Tuple[int, np.ndarray]:
    """Effective odd diameter and the float64 spatial weights of the reference's
    ``_bilateral_filter`` (pipeline/enhancement.py:117-128)."""
    d = min(int(d), 9)
    if d % 2 == 0:
        d += 1
    r = d // 2
    yy, xx = np.mgrid[-r : r + 1, -r : r + 1]
    w = np.exp(-(xx**2 + yy**2) / (2 * sigma_space**2 * d**2))
    return d, np.ascontiguousarray(w, dtype=np.float64)


class StackOps:
    """Binds the C ABI to torch tensors on one CUDA device.  Thread-safe: the scratch workspace is
    per thread, and every call runs on the calling thread's current torch stream."""

    def __init__(self, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("mdimg_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.index is None:
            self.device = torch.device(f"cuda:{torch.cuda.current_device()}")
        self.lib = load_library()
        with torch.cuda.device(self.device):
            _lib.ensure_device(self.device.index)
        self.launches = 0     # number of C-ABI compute calls issued (each is >= 1 kernel launch)

    # ---- plumbing -------------------------------------------------------------------------
    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, nbytes: int) -> Tuple[C.c_void_p, int]:
        key = f"ws_{self.device.index}"
        buf = getattr(_tls, key, None)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                # kernels already queued on the stream may still use the old buffer
                torch.cuda.current_stream(self.device).synchronize()
            buf = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=self.device)
            setattr(_tls, key, buf)
        return C.c_void_p(buf.data_ptr()), buf.numel()

    def _ws_for(self, op: int, n: int, h: int, w: int, param: int = 0):
        return self._workspace(self.lib.mdimg_workspace_bytes(op, n, h, w, param))

    def _img(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        if t.device != self.device:
            raise ValueError(f"tensor on {t.device}, ops bound to {self.device}")
        if t.dtype != dtype or t.dim() != 3 or not t.is_contiguous():
            raise ValueError(f"expected a contiguous [N, H, W] {dtype} tensor, got {tuple(t.shape)} {t.dtype}")
        return t

    def _sel(self, sel: Optional[torch.Tensor]):
        if sel is None:
            return C.c_void_p(0), 0
        if sel.dtype != torch.int32 or sel.device != self.device or not sel.is_contiguous():
            raise ValueError("sel must be a contiguous int32 CUDA tensor")
        return C.c_void_p(sel.data_ptr()), int(sel.numel())

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
        return C.c_void_p(0 if t is None else t.data_ptr())

    def _call(self, fn, *args):
        with torch.cuda.device(self.device):
            rc = fn(*args)
        self.launches += 1
        check(rc)

    # ---- ingestion --------------------------------------------------------------------------
    def normalize(self, raw: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """normalize_image per slice (pipeline/dicom_io.py:84-91); uint16 or float32 input."""
        n, h, w = raw.shape
        if raw.dtype == torch.uint16 or raw.dtype == torch.int16:
            # int16 storage is accepted as a uint16 bit pattern container
            src = self._img(raw, raw.dtype)
            fn = self.lib.mdimg_normalize_u16
        else:
            src = self._img(raw)
            fn = self.lib.mdimg_normalize_f32
        if out is None:
            out = torch.empty((n, h, w), dtype=torch.float32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_NORMALIZE, n, h, w)
        self._call(fn, self._ptr(src), self._ptr(out), n, h, w, C.c_void_p(0), 0, ws, wsb, self._stream())
        return out

    def ingest(self, raw: torch.Tensor, slope: Optional[float] = None, intercept: Optional[float] = None,
               monochrome1: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """load_dicom's pixel path fused with normalize_image for a stack of 16-bit frames
        (pipeline/dicom_io.py:44-49,84-91): modality rescale (when slope/intercept are given),
        MONOCHROME1 inversion against the stack maximum, per-slice normalisation."""
        if raw.dtype not in (torch.int16, torch.uint16):
            raise ValueError("ingest expects 16-bit raw samples (int16 or uint16)")
        src = self._img(raw, raw.dtype)
        n, h, w = src.shape
        if out is None:
            out = torch.empty((n, h, w), dtype=torch.float32, device=self.device)
        has = slope is not None and intercept is not None
        ws, wsb = self._ws_for(_lib.OP_NORMALIZE, n, h, w)
        self._call(self.lib.mdimg_ingest_u16, self._ptr(src), self._ptr(out), n, h, w, C.c_void_p(0), 0,
                   float(slope) if has else 1.0, float(intercept) if has else 0.0, 1 if has else 0,
                   1 if monochrome1 else 0, 1 if raw.dtype == torch.int16 else 0, ws, wsb, self._stream())
        return out

    def mosaic(self, before: torch.Tensor, after: torch.Tensor, gap: int = 8, gap_level: int = 255) -> torch.Tensor:
        """[N, H, 2W + gap] uint8 before | after panels (matplotlib gray quantisation, per-panel autoscale)."""
        n, h, w = self._img(before).shape
        if tuple(self._img(after).shape) != (n, h, w):
            raise ValueError("before / after stacks must have the same shape")
        out = torch.empty((n, h, 2 * w + gap), dtype=torch.uint8, device=self.device)
        ws, wsb = self._workspace(2 * (n * 8 + 256))
        self._call(self.lib.mdimg_mosaic_u8, self._ptr(before), self._ptr(after), self._ptr(out), n, h, w,
                   C.c_void_p(0), 0, int(gap), int(gap_level), ws, wsb, self._stream())
        return out

    def minmax(self, img: torch.Tensor, sel=None) -> torch.Tensor:
        n, h, w = self._img(img).shape
        out = torch.zeros((n, 2), dtype=torch.float32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_MINMAX, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_minmax_f32, self._ptr(img), n, h, w, sp, ns, self._ptr(out), ws, wsb, self._stream())
        return out

    # ---- metrics ------------------------------------------------------------------------------
    def metrics(self, img: torch.Tensor, with_niqe: bool = False, sel=None,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[N, 24] float64 rows: the 16 compute_metrics values + mean, edge ratio, NIQE, ..."""
        n, h, w = self._img(img).shape
        lo, hi, gamma = percentile_plan(h * w)
        if out is None:
            out = torch.full((n, _lib.METRIC_COLS), float("nan"), dtype=torch.float64, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_METRICS, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_metrics, self._ptr(img), n, h, w, sp, ns, 1 if with_niqe else 0,
                   lo.ctypes.data_as(C.POINTER(C.c_int32)), hi.ctypes.data_as(C.POINTER(C.c_int32)),
                   gamma.ctypes.data_as(C.POINTER(C.c_float)), self._ptr(out), ws, wsb, self._stream())
        return out

    def estimate_sigma(self, img: torch.Tensor, sel=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, h, w = self._img(img).shape
        if out is None:
            out = torch.full((n,), float("nan"), dtype=torch.float64, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_SIGMA, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_estimate_sigma, self._ptr(img), n, h, w, sp, ns, self._ptr(out), ws, wsb, self._stream())
        return out

    def quality(self, img: torch.Tensor, niqe: bool = True, sel=None) -> torch.Tensor:
        """[N, 2] float64: (edge_ratio, niqe_approx)."""
        n, h, w = self._img(img).shape
        out = torch.full((n, 2), float("nan"), dtype=torch.float64, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_QUALITY, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_quality, self._ptr(img), n, h, w, sp, ns, 1 if niqe else 0, self._ptr(out), ws, wsb, self._stream())
        return out

    def fullref(self, original: torch.Tensor, enhanced: torch.Tensor, sel=None) -> torch.Tensor:
        """[N, 2] float64: (ssim, psnr)."""
        n, h, w = self._img(original).shape
        if tuple(self._img(enhanced).shape) != (n, h, w):
            raise ValueError("Input images must have the same dimensions.")
        out = torch.full((n, 2), float("nan"), dtype=torch.float64, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_FULLREF, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_fullref, self._ptr(original), self._ptr(enhanced), n, h, w, sp, ns,
                   self._ptr(out), ws, wsb, self._stream())
        return out

    def validation(self, original: torch.Tensor, enhanced: torch.Tensor, sel=None) -> torch.Tensor:
        """[N, 50] float64 in one library call: metrics row (with NIQE) of the original | of the
        enhanced image | ssim | psnr -- the image-sized work of compute_validation."""
        n, h, w = self._img(original).shape
        if tuple(self._img(enhanced).shape) != (n, h, w):
            raise ValueError("Input images must have the same dimensions.")
        lo, hi, gamma = percentile_plan(h * w)
        out = torch.full((n, _lib.VALIDATION_COLS), float("nan"), dtype=torch.float64, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_VALIDATION, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_validation, self._ptr(original), self._ptr(enhanced), n, h, w, sp, ns,
                   lo.ctypes.data_as(C.POINTER(C.c_int32)), hi.ctypes.data_as(C.POINTER(C.c_int32)),
                   gamma.ctypes.data_as(C.POINTER(C.c_float)), self._ptr(out), ws, wsb, self._stream())
        return out

    # ---- enhancement steps --------------------------------------------------------------------
    def wavelet_denoise(self, src, dst, mode: str = "soft", sigma: Optional[torch.Tensor] = None,
                        sigma_scale: float = 1.0, sel=None):
        n, h, w = self._img(src).shape
        self._img(dst)
        ws, wsb = self._ws_for(_lib.OP_WAVELET, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_wavelet_denoise, self._ptr(src), self._ptr(dst), n, h, w, sp, ns,
                   1 if mode == "hard" else 0, self._ptr(sigma), float(sigma_scale), C.c_void_p(0),
                   ws, wsb, self._stream())
        return dst

    def clahe(self, src, dst, clip_limit: float, kernel_size: int, sel=None, gamma: float = 1.0) -> torch.Tensor:
        """Returns the per-slice status tensor (1 = input outside [-1, 1]).  gamma != 1 folds an
        adjust_gamma that directly follows CLAHE into its final pass."""
        n, h, w = self._img(src).shape
        self._img(dst)
        status = torch.zeros((n,), dtype=torch.int32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_CLAHE, n, h, w, int(kernel_size))
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_clahe_gamma, self._ptr(src), self._ptr(dst), n, h, w, sp, ns,
                   float(clip_limit), int(kernel_size), float(gamma), self._ptr(status), ws, wsb, self._stream())
        return status

    def gamma(self, src, dst, gamma: float, assume_nonneg: bool = False, sel=None) -> torch.Tensor:
        n, h, w = self._img(src).shape
        self._img(dst)
        neg = torch.zeros((n,), dtype=torch.int32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_GAMMA, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_gamma, self._ptr(src), self._ptr(dst), n, h, w, sp, ns, float(gamma),
                   1 if assume_nonneg else 0, self._ptr(neg), ws, wsb, self._stream())
        return neg

    def unsharp(self, src, dst, radius: float, amount: float, assume_nonneg: bool = False, sel=None):
        n, h, w = self._img(src).shape
        self._img(dst)
        taps = gaussian_taps(radius)
        ws, wsb = self._ws_for(_lib.OP_UNSHARP, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_unsharp, self._ptr(src), self._ptr(dst), n, h, w, sp, ns,
                   taps.ctypes.data_as(C.POINTER(C.c_double)), len(taps) - 1, float(amount),
                   1 if assume_nonneg else 0, ws, wsb, self._stream())
        return dst

    def light_denoise(self, src, dst, strength: float, sel=None) -> torch.Tensor:
        """Returns per-slice `skipped` flags (sigma < 0.001 left the slice unchanged)."""
        n, h, w = self._img(src).shape
        self._img(dst)
        skipped = torch.zeros((n,), dtype=torch.int32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_LIGHT_DENOISE, n, h, w)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_light_denoise, self._ptr(src), self._ptr(dst), n, h, w, sp, ns,
                   float(strength), self._ptr(skipped), ws, wsb, self._stream())
        return skipped

    def bilateral(self, src, dst, d: int, sigma_color: float, sigma_space: float, sel=None):
        n, h, w = self._img(src).shape
        self._img(dst)
        deff, spatial = bilateral_spatial(d, sigma_space)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_bilateral, self._ptr(src), self._ptr(dst), n, h, w, sp, ns, deff,
                   spatial.ctypes.data_as(C.POINTER(C.c_double)), float(sigma_color), self._stream())
        return dst

    def tv_chambolle(self, src, dst, weight: float, eps: float = 2.0e-4, max_iter: int = 200, sel=None) -> torch.Tensor:
        """Returns the per-slice number of executed iterations."""
        n, h, w = self._img(src).shape
        self._img(dst)
        iters = torch.zeros((n,), dtype=torch.int32, device=self.device)
        ws, wsb = self._ws_for(_lib.OP_TV, n, h, w, int(max_iter))
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_tv_chambolle, self._ptr(src), self._ptr(dst), n, h, w, sp, ns,
                   float(weight), float(eps), int(max_iter), self._ptr(iters), ws, wsb, self._stream())
        return iters

    def axpby(self, a, b, dst, c0: float, c1: float, clip01: bool = False, sel=None):
        n, h, w = self._img(a).shape
        self._img(b)
        self._img(dst)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_axpby, self._ptr(a), self._ptr(b), self._ptr(dst), n, h, w, sp, ns,
                   float(c0), float(c1), 1 if clip01 else 0, self._stream())
        return dst

    def clip01(self, src, dst, sel=None):
        n, h, w = self._img(src).shape
        self._img(dst)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_clip01, self._ptr(src), self._ptr(dst), n, h, w, sp, ns, self._stream())
        return dst

    def copy(self, src, dst, sel=None):
        n, h, w = self._img(src).shape
        self._img(dst)
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_copy, self._ptr(src), self._ptr(dst), n, h, w, sp, ns, self._stream())
        return dst

    def enhance(self, image: torch.Tensor, plan, rows_before: Optional[torch.Tensor] = None):
        """apply_enhancements_from_params for a whole stack in ONE library call (`mdimg_enhance`: clamping,
        step gating, safeguards and their per-slice decisions run in C++).  `plan` has `recommended_ops`
        and `params` like the reference's EnhancementPlan.  Returns (enhanced [N, H, W] float32,
        rows_after [N, 24] float64 device, flags int32 numpy [N], tv_iterations int32 numpy [N],
        clamped plan struct)."""
        n, h, w = self._img(image).shape
        q = _lib.EnhancePlan()
        names = [op.lower().strip() for op in plan.recommended_ops]
        steps = [_lib.STEP_NAMES.index(nm) for nm in names if nm in _lib.STEP_NAMES]
        if len(steps) > _lib.MAX_PLAN_OPS:       # Engine.enhance_plan routes such plans to the torch-side flow
            raise ValueError(f"plan lists {len(steps)} recognised operations; the native call holds {_lib.MAX_PLAN_OPS}")
        q.n_ops = len(steps)
        for i, st in enumerate(steps):
            q.ops[i] = st
        p = plan.params
        q.clahe_clip_limit = float(p.clahe_clip_limit)
        # int() of the clamped value, as the reference does (pipeline/enhancement.py:250)
        q.clahe_tile_size = int(max(4, min(48, p.clahe_tile_size)))
        q.gamma = float(p.gamma)
        q.unsharp_radius = float(p.unsharp_radius)
        q.unsharp_amount = float(p.unsharp_amount)
        q.denoise_hard = 1 if p.denoise_mode == "hard" else 0
        q.post_denoise_strength = float(p.post_denoise_strength)
        q.bilateral_d = int(max(0, min(13, p.bilateral_d)))
        q.bilateral_sigma_color = float(p.bilateral_sigma_color)
        q.bilateral_sigma_space = float(p.bilateral_sigma_space)
        q.tv_denoise_weight = float(p.tv_denoise_weight)
        check(self.lib.mdimg_plan_clamp(C.byref(q)))
        # tables from numpy / scipy's own arithmetic (bit-for-bit the reference's weights)
        t = _lib.EnhanceTables()
        taps = gaussian_taps(q.unsharp_radius)
        t.gauss_radius = len(taps) - 1
        for i, v in enumerate(taps):
            t.gauss_taps[i] = float(v)
        if q.bilateral_d > 0:
            deff, spatial = bilateral_spatial(q.bilateral_d, q.bilateral_sigma_space)
            t.bilateral_d_eff = deff
            for i, v in enumerate(spatial.ravel()):
                t.bilateral_spatial[i] = float(v)
        lo, hi, gamma = percentile_plan(h * w)
        for i in range(5):
            t.pct_lo[i], t.pct_hi[i], t.pct_gamma[i] = int(lo[i]), int(hi[i]), float(gamma[i])
        out = torch.empty_like(image)
        rows_after = torch.full((n, _lib.METRIC_COLS), float("nan"), dtype=torch.float64, device=self.device)
        flags = np.zeros(n, np.int32)
        iters = np.zeros(n, np.int32)
        if rows_before is not None and (rows_before.dtype != torch.float64 or tuple(rows_before.shape) != (n, _lib.METRIC_COLS)
                                        or not rows_before.is_contiguous()):
            raise ValueError("rows_before must be a contiguous [N, 24] float64 tensor")
        ws, wsb = self._ws_for(_lib.OP_ENHANCE, n, h, w, int(q.clahe_tile_size))
        self._call(self.lib.mdimg_enhance, self._ptr(image), self._ptr(out), n, h, w, C.byref(q), C.byref(t),
                   self._ptr(rows_before), self._ptr(rows_after), flags.ctypes.data_as(C.POINTER(C.c_int32)),
                   iters.ctypes.data_as(C.POINTER(C.c_int32)), ws, wsb, self._stream())
        return out, rows_after, flags, iters, q

    def enhance_issues(self, image: torch.Tensor, issues, sigma_before: Optional[torch.Tensor] = None):
        """apply_enhancements(image, issues) for a whole stack in one library call (`mdimg_enhance_issues`).
        Returns (enhanced [N, H, W] float32, flags int32 numpy [N])."""
        n, h, w = self._img(image).shape
        mask = 0
        for name in issues:
            mask |= _lib.ISSUE_BITS.get(name, 0)
        t = _lib.EnhanceTables()
        taps = gaussian_taps(0.8)                   # ENHANCEMENT_PARAMS["unsharp_radius"], scipy's own weights
        t.gauss_radius = len(taps) - 1
        for i, v in enumerate(taps):
            t.gauss_taps[i] = float(v)
        out = torch.empty_like(image)
        flags = np.zeros(n, np.int32)
        if sigma_before is not None and (sigma_before.dtype != torch.float64 or tuple(sigma_before.shape) != (n,)
                                         or not sigma_before.is_contiguous()):
            raise ValueError("sigma_before must be a contiguous [N] float64 tensor")
        ws, wsb = self._ws_for(_lib.OP_ENHANCE, n, h, w, 16)
        self._call(self.lib.mdimg_enhance_issues, self._ptr(image), self._ptr(out), n, h, w, int(mask), C.byref(t),
                   self._ptr(sigma_before), flags.ctypes.data_as(C.POINTER(C.c_int32)), ws, wsb, self._stream())
        return out, flags

    def export_u16(self, src, dst=None, sel=None):
        """uint16(clip(rint(x * 65535), 0, 65535)) of a float32 stack, as an int16-typed tensor holding
        the uint16 bit pattern (torch has no arithmetic on uint16; view it with numpy)."""
        n, h, w = self._img(src).shape
        if dst is None:
            dst = torch.empty((n, h, w), dtype=torch.int16, device=self.device)
        if dst.dtype not in (torch.int16, torch.uint16) or tuple(dst.shape) != (n, h, w) or not dst.is_contiguous():
            raise ValueError("export_u16: dst must be a contiguous [N, H, W] 16-bit tensor")
        sp, ns = self._sel(sel)
        self._call(self.lib.mdimg_export_u16, self._ptr(src), self._ptr(dst), n, h, w, sp, ns, self._stream())
        return dst


_ops_lock = threading.Lock()
_ops_by_device: dict = {}


def get_ops(device=None) -> StackOps:
    """Process-wide StackOps per device."""
    if device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("mdimg_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback")
        device = torch.device(f"cuda:{torch.cuda.current_device()}")
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _ops_lock:
        ops = _ops_by_device.get(key)
        if ops is None:
            ops = StackOps(torch.device(f"cuda:{key}"))
            _ops_by_device[key] = ops
    return ops
