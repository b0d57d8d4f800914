/* mdimg_b200 — C ABI of the B200-native MDIMG image hot path (libmdimg_b200.so).
 *
 * The reference (Hiresh444/medical-image-enhancer) has no FFI: its hot path is a set of plain
 * Python functions in pipeline/enhancement.py, pipeline/metrics.py and pipeline/dicom_io.py that
 * call scikit-image / PyWavelets / scipy.  Each entry point below names the reference function
 * (file:line under /root/reference) whose arithmetic it replaces; the Python shim in
 * medical-image-enhancer_b200/ binds them with ctypes and re-exposes the reference's own
 * function signatures.
 *
 * Conventions
 *  - every image argument is a DEVICE pointer to an [n][h][w] contiguous stack of 2-D slices
 *    (float32 normalised to [0,1] unless stated otherwise); slices are independent;
 *  - `sel` (device int32[n_sel], may be NULL) restricts the call to a subset of slices; outputs
 *    indexed by slice keep their position in the full stack;
 *  - the caller owns all buffers; scratch memory is a caller-provided device workspace whose
 *    size is queried with mdimg_workspace_bytes(); nothing is allocated behind the caller's back
 *    except one pinned int per host thread that calls mdimg_tv_chambolle;
 *  - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream unless
 *    noted; results in device memory are valid after the stream is synchronised;
 *  - every function returns MDIMG_OK (0) or an error code; mdimg_last_error() returns a
 *    thread-local message.  There is NO CPU fallback: without an sm_100 device mdimg_init fails.
 */
#ifndef MDIMG_B200_H
#define MDIMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDIMG_OK 0
#define MDIMG_ERR_INVALID 1
#define MDIMG_ERR_CUDA 2
#define MDIMG_ERR_WORKSPACE 3
#define MDIMG_ERR_NO_DEVICE 4

/* Result-row layout of mdimg_metrics (doubles per slice). Columns 0..15 are the 16 metrics in the
 * key order of compute_metrics' dict (pipeline/metrics.py:90-109). */
#define MDIMG_METRIC_COLS 24
#define MDIMG_MC_MEAN 16
#define MDIMG_MC_EDGE_RATIO 17
#define MDIMG_MC_NIQE 18
#define MDIMG_MC_VAR_OF_VAR 19
#define MDIMG_MC_GMAX 20
#define MDIMG_MC_P05 21
#define MDIMG_MC_P95 22

/* op codes for mdimg_workspace_bytes */
enum mdimg_op {
    MDIMG_OP_NORMALIZE = 1,
    MDIMG_OP_METRICS = 2,
    MDIMG_OP_SIGMA = 3,
    MDIMG_OP_QUALITY = 4,
    MDIMG_OP_FULLREF = 5,
    MDIMG_OP_WAVELET = 6,
    MDIMG_OP_CLAHE = 7,      /* param = kernel_size */
    MDIMG_OP_GAMMA = 8,
    MDIMG_OP_UNSHARP = 9,
    MDIMG_OP_LIGHT_DENOISE = 10,
    MDIMG_OP_BILATERAL = 11,
    MDIMG_OP_TV = 12,        /* param = max_iter */
    MDIMG_OP_MINMAX = 13,
    MDIMG_OP_VALIDATION = 14,
    MDIMG_OP_ENHANCE = 15    /* param = clahe kernel size (clamped plan value) */
};

const char* mdimg_last_error(void);
int mdimg_version(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
unsigned long long mdimg_launch_count(void);
/* Device self-test of mdimg_normalize_u16's quotient: (a - min) / (max - min) is evaluated with the slice's
 * correctly rounded reciprocal and one exact-residual correction instead of an IEEE division per pixel; this
 * runs both over EVERY integer operand pair 0 <= a <= denom <= 65535 and writes the number of differing
 * results (must be 0) to *mismatches (host pointer).  Synchronises `stream`. */
int mdimg_selftest_div16(unsigned long long* mismatches, void* stream);

/* Select the device and verify it is compute capability 10.x (B200). */
int mdimg_init(int device);
int mdimg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes,
                      size_t* total_mem);

/* Device workspace needed by `op` for a stack of n slices of h x w (sized for n_sel == n). */
size_t mdimg_workspace_bytes(int op, int n, int h, int w, int param);

/* ---- ingestion ------------------------------------------------------------------------- */
/* out_minmax: device float[n][2]. */
int mdimg_minmax_f32(const float* img, int n, int h, int w, const int32_t* sel, int n_sel,
                     float* out_minmax, void* ws, size_t ws_bytes, void* stream);
/* Report mosaic of save_visuals (pipeline/dicom_io.py:99-126: imshow(original, cmap="gray") next to
 * imshow(enhanced, cmap="gray")): out = device uint8 [n][h][2w + gap], each panel autoscaled to its
 * own min / max and quantised to matplotlib's 256 gray levels, `gap` columns of `gap_level` between
 * them.  Workspace: MDIMG_OP_MINMAX twice (2 * n * 8 bytes). */
int mdimg_mosaic_u8(const float* before, const float* after, uint8_t* out, int n, int h, int w,
                    const int32_t* sel, int n_sel, int gap, int gap_level, void* ws, size_t ws_bytes,
                    void* stream);
/* normalize_image (pipeline/dicom_io.py:84-91): per-slice (x - min) / (max - min), zeros when
 * max - min < 1e-8. */
int mdimg_normalize_u16(const uint16_t* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, void* ws, size_t ws_bytes, void* stream);
int mdimg_normalize_f32(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, void* ws, size_t ws_bytes, void* stream);
/* The pixel path of load_dicom (pipeline/dicom_io.py:44-49) fused with normalize_image: pydicom's
 * modality rescale float32(float64(raw) * slope + intercept) when has_rescale, `image.max() - image`
 * over the whole stack when monochrome1, then per-slice normalisation.  raw: 16-bit samples
 * (is_signed selects int16).  Workspace: MDIMG_OP_NORMALIZE. */
int mdimg_ingest_u16(const uint16_t* raw, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                     double slope, double intercept, int has_rescale, int monochrome1, int is_signed,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- metrics ----------------------------------------------------------------------------- */
/* compute_metrics (pipeline/metrics.py:42-109) for every slice; flags bit0 additionally fills
 * MDIMG_MC_NIQE / MDIMG_MC_VAR_OF_VAR (compute_niqe_approximation, metrics.py:187-210).
 * MDIMG_MC_EDGE_RATIO (compute_edge_ratio, metrics.py:213-217) is always filled.
 * pct_lo / pct_hi / pct_gamma: HOST arrays of 5 entries — numpy's 'linear' percentile plan
 * (previous index, next index, float32 weight) for q = 5, 25, 75, 95 (of the image) and 90 (of the
 * gradient magnitude) at h*w elements.  out: device double[n][MDIMG_METRIC_COLS]. */
int mdimg_metrics(const float* img, int n, int h, int w, const int32_t* sel, int n_sel, int flags,
                  const int32_t* pct_lo, const int32_t* pct_hi, const float* pct_gamma,
                  double* out, void* ws, size_t ws_bytes, void* stream);
/* skimage.restoration.estimate_sigma(image, channel_axis=None, average_sigmas=True)
 * (called at pipeline/metrics.py:47, pipeline/enhancement.py:59-60,82). sigma: device double[n]. */
int mdimg_estimate_sigma(const float* img, int n, int h, int w, const int32_t* sel, int n_sel,
                         double* sigma, void* ws, size_t ws_bytes, void* stream);
/* out: device double[n][2] = (compute_edge_ratio, compute_niqe_approximation); flags bit0 enables
 * the NIQE column (pipeline/metrics.py:187-217; guards at pipeline/enhancement.py:50-72). */
int mdimg_quality(const float* img, int n, int h, int w, const int32_t* sel, int n_sel, int flags,
                  double* out, void* ws, size_t ws_bytes, void* stream);
/* skimage.metrics.structural_similarity / peak_signal_noise_ratio with data_range=1.0
 * (pipeline/metrics.py:232-233). out: device double[n][2] = (ssim, psnr). */
int mdimg_fullref(const float* original, const float* enhanced, int n, int h, int w,
                  const int32_t* sel, int n_sel, double* out, void* ws, size_t ws_bytes,
                  void* stream);

/* The image-sized work of compute_validation (pipeline/metrics.py:225-259) in one call:
 * compute_metrics + NIQE approximation + edge ratio of both images, SSIM and PSNR.  out: device
 * double[n][MDIMG_VALIDATION_COLS] = metrics row of the original [MDIMG_METRIC_COLS] | of the
 * enhanced image [MDIMG_METRIC_COLS] | ssim | psnr; the scalar gains and pass logic
 * (metrics.py:261-329) are host arithmetic on that row.  Percentile plan as for mdimg_metrics.
 * Workspace: MDIMG_OP_VALIDATION. */
#define MDIMG_VALIDATION_COLS (2 * MDIMG_METRIC_COLS + 2)
int mdimg_validation(const float* original, const float* enhanced, int n, int h, int w,
                     const int32_t* sel, int n_sel, const int32_t* pct_lo, const int32_t* pct_hi,
                     const float* pct_gamma, double* out, void* ws, size_t ws_bytes, void* stream);

/* ---- enhancement steps ------------------------------------------------------------------ */
/* denoise_wavelet(image, channel_axis=None, rescale_sigma=True, mode=...[, sigma=...])
 * (pipeline/enhancement.py:86,169,270,328).  sigma_in: device double[n] or NULL (estimate);
 * the sigma used is sigma_in[s] * sigma_scale.  skip: device int32[n] or NULL. */
int mdimg_wavelet_denoise(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                          int n_sel, int mode_hard, const double* sigma_in, double sigma_scale,
                          const int32_t* skip, void* ws, size_t ws_bytes, void* stream);
/* exposure.equalize_adapthist(image, clip_limit, kernel_size) (pipeline/enhancement.py:183,277,332).
 * status: device int32[n]; 1 where the slice leaves [-1, 1] (the reference raises ValueError). */
int mdimg_clahe(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                double clip_limit, int kernel_size, int32_t* status, void* ws, size_t ws_bytes,
                void* stream);
/* equalize_adapthist followed directly by adjust_gamma (pipeline/enhancement.py:277-286 when the
 * plan holds both "clahe" and "gamma"): the power is folded into CLAHE's final 16384-level
 * stretch table, so the step costs no pass of its own.  gamma == 1 is mdimg_clahe.  Same workspace
 * as mdimg_clahe (MDIMG_OP_CLAHE). */
int mdimg_clahe_gamma(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                      double clip_limit, int kernel_size, double gamma, int32_t* status, void* ws,
                      size_t ws_bytes, void* stream);
/* exposure.adjust_gamma(image, gamma) (pipeline/enhancement.py:194,197,284,336).
 * neg_flag: device int32[n]; 1 where a slice holds a negative pixel (ValueError in the reference).
 * assume_nonneg != 0 skips the min/max pre-pass (caller guarantees non-negative input). */
int mdimg_gamma(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                double gamma, int assume_nonneg, int32_t* neg_flag, void* ws, size_t ws_bytes,
                void* stream);
/* filters.unsharp_mask(image, radius, amount) (pipeline/enhancement.py:202,290,338).
 * weights: HOST double[gauss_radius + 1], scipy's normalised Gaussian taps w[0..R]. */
int mdimg_unsharp(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                  const double* weights, int gauss_radius, double amount, int assume_nonneg,
                  void* ws, size_t ws_bytes, void* stream);
/* _light_denoise(image, strength) (pipeline/enhancement.py:80-94). skipped: device int32[n] or
 * NULL, set to 1 where sigma < 0.001 returned the slice unchanged. */
int mdimg_light_denoise(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, double strength, int32_t* skipped, void* ws, size_t ws_bytes,
                        void* stream);
/* _bilateral_filter (pipeline/enhancement.py:102-143). d: odd effective diameter (1..9);
 * spatial: HOST double[d*d] spatial weights. */
int mdimg_bilateral(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                    int n_sel, int d, const double* spatial, double sigma_color, void* stream);
/* denoise_tv_chambolle(image, weight, channel_axis=None) (pipeline/enhancement.py:311,349).
 * iters: device int32[n] or NULL.  The call never drains `stream`: the launch sequence is cut short from live-slice
 * counts the kernels report through host-mapped memory, the host waits only on events a few launches back. */
int mdimg_tv_chambolle(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                       int n_sel, double weight, double eps, int max_iter, int32_t* iters,
                       void* ws, size_t ws_bytes, void* stream);
/* out = c0*a + c1*b in float32 (optionally clipped to [0,1]): the blends at
 * pipeline/enhancement.py:93,365. */
int mdimg_axpby(const float* a, const float* b, float* out, int n, int h, int w, const int32_t* sel,
                int n_sel, double c0, double c1, int clip01, void* stream);
/* np.clip(x, 0, 1) (pipeline/enhancement.py:218,225,314,352,360,367). */
int mdimg_clip01(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                 void* stream);
/* 16-bit export of an enhanced [0, 1] stack: uint16(clip(rint(float32(x) * 65535), 0, 65535)), i.e.
 * skimage's img_as_uint (the conversion equalize_adapthist applies to its input,
 * pipeline/enhancement.py:183).  The reference itself returns float32 (pipeline/enhancement.py:226,368);
 * this halves the device-to-host bytes of a stack for callers that store 16-bit pixels. */
int mdimg_export_u16(const float* in, uint16_t* out, int n, int h, int w, const int32_t* sel, int n_sel,
                     void* stream);
int mdimg_copy(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
               void* stream);

/* ---- the whole enhancement call ----------------------------------------------------------- */
/* apply_enhancements_from_params(image, plan) (pipeline/enhancement.py:235-369) for every slice of a
 * stack in ONE call: PARAM_BOUNDS clamping (pipeline/schemas.py:16-28), the seven gated steps in their
 * fixed order, the final clip, and the three safeguards (halo re-run in the plan's own order with half
 * the unsharp amount, corrective denoise, 40 % blend-back) with the reference's decisions taken per
 * slice.  The host arithmetic between the kernels (which slices a safeguard fires on) runs inside the
 * call, which therefore synchronises `stream` a few times. */
enum mdimg_step {            /* index into the reference's fixed step order (enhancement.py:266-315) */
    MDIMG_STEP_DENOISE = 0, MDIMG_STEP_CLAHE = 1, MDIMG_STEP_GAMMA = 2, MDIMG_STEP_UNSHARP = 3,
    MDIMG_STEP_POST_DENOISE = 4, MDIMG_STEP_BILATERAL = 5, MDIMG_STEP_TV_DENOISE = 6
};
#define MDIMG_MAX_PLAN_OPS 64    /* capacity of mdimg_enhance_plan.ops[]; a longer list is an error, never truncated */
typedef struct mdimg_enhance_plan {
    int32_t n_ops;               /* entries of ops[] (<= MDIMG_MAX_PLAN_OPS) */
    int32_t ops[MDIMG_MAX_PLAN_OPS]; /* plan.recommended_ops in the plan's order, repeats included (the halo
                                  * safeguard replays the list as written), as MDIMG_STEP_*; unknown names dropped */
    double clahe_clip_limit;     /* EnhancementParams (pipeline/schemas.py:36-100); clamped by mdimg_plan_clamp */
    int32_t clahe_tile_size;
    double gamma;
    double unsharp_radius;
    double unsharp_amount;
    int32_t denoise_hard;        /* denoise_mode: 0 'soft' (also for unknown strings), 1 'hard' */
    double post_denoise_strength;
    int32_t bilateral_d;
    double bilateral_sigma_color;
    double bilateral_sigma_space;
    double tv_denoise_weight;
} mdimg_enhance_plan;
/* Host tables that depend on the CLAMPED plan and the image size.  A Python caller fills them with
 * numpy / scipy's own values (bit-for-bit the reference's); mdimg_enhance_tables_default computes them
 * with the C library's exp (which may differ from numpy's in the last bit of a weight). */
typedef struct mdimg_enhance_tables {
    int32_t gauss_radius;        /* int(4 * unsharp_radius + 0.5) */
    double gauss_taps[13];       /* scipy.ndimage _gaussian_kernel1d(unsharp_radius): centre, then +1 ... +radius */
    int32_t bilateral_d_eff;     /* min(bilateral_d, 9), made odd (pipeline/enhancement.py:117-121); 0 = off */
    double bilateral_spatial[81];/* exp(-(dx^2 + dy^2) / (2 sigma_space^2 d^2)), row-major dy, dx */
    int32_t pct_lo[5], pct_hi[5];/* numpy percentile plan for h*w elements, q = 5, 25, 75, 95, 90 (see mdimg_metrics) */
    float pct_gamma[5];
} mdimg_enhance_tables;
/* Result flags per slice. */
#define MDIMG_FLAG_HALO 1            /* "[safeguard] Unsharp reduced to ..." */
#define MDIMG_FLAG_NOISE_GUARD 2     /* "Auto-corrective denoise (noise guard)" */
#define MDIMG_FLAG_OVER_PROCESSED 4  /* "Blend-back 40% original (over-processing guard)" */
#define MDIMG_FLAG_ERR_CLAHE_RANGE 8 /* the reference raises ValueError("Images of type float must be between -1 and 1.") */
#define MDIMG_FLAG_ERR_GAMMA_NEG 16  /* the reference raises ValueError("Image Correction methods work correctly only on ...") */

/* Clamps every parameter to PARAM_BOUNDS in place (idempotent). */
int mdimg_plan_clamp(mdimg_enhance_plan* plan);
/* Fills `tables` for a clamped plan and h x w images with the C library's arithmetic. */
int mdimg_enhance_tables_default(const mdimg_enhance_plan* plan, int h, int w, mdimg_enhance_tables* tables);
/* in / out: device float32 [n][h][w] (out must not alias in).  plan, tables: host.  rows_before: device
 * double[n][MDIMG_METRIC_COLS] rows of mdimg_metrics(in, flags = 1), or NULL (computed here).
 * rows_after: device double[n][MDIMG_METRIC_COLS], receives mdimg_metrics(out, flags = 1), or NULL.
 * flags_out: HOST int32[n] (MDIMG_FLAG_*); a slice with an error flag is returned unchanged.
 * tv_iters_out: HOST int32[n] or NULL.  Workspace: MDIMG_OP_ENHANCE with param = clamped clahe_tile_size. */
int mdimg_enhance(const float* in, float* out, int n, int h, int w, const mdimg_enhance_plan* plan,
                  const mdimg_enhance_tables* tables, const double* rows_before, double* rows_after,
                  int32_t* flags_out, int32_t* tv_iters_out, void* ws, size_t ws_bytes, void* stream);

/* apply_enhancements(image, issues) (pipeline/enhancement.py:151-227): the issue-gated chain with the fixed
 * ENHANCEMENT_PARAMS (enhancement.py:32-42) -- noise -> wavelet denoise; low_contrast / clipping -> CLAHE;
 * one-sided clipping -> gamma 0.95 / 1.05; blur -> unsharp + light denoise -- the final clip and the
 * noise-amplification guard.  issues: OR of MDIMG_ISSUE_*.  tables: gauss_taps / gauss_radius for
 * unsharp_radius = 0.8 (mdimg_enhance_tables_default on a default plan).  sigma_before: device double[n]
 * estimate_sigma of the input, or NULL.  flags_out: HOST int32[n] (MDIMG_FLAG_NOISE_GUARD, MDIMG_FLAG_ERR_*).
 * Workspace: MDIMG_OP_ENHANCE with param = 16. */
#define MDIMG_ISSUE_NOISE 1
#define MDIMG_ISSUE_BLUR 2
#define MDIMG_ISSUE_LOW_CONTRAST 4
#define MDIMG_ISSUE_CLIPPING_LOW 8
#define MDIMG_ISSUE_CLIPPING_HIGH 16
int mdimg_enhance_issues(const float* in, float* out, int n, int h, int w, int issues,
                         const mdimg_enhance_tables* tables, const double* sigma_before, int32_t* flags_out,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- host scalar logic of the path (no device work) -------------------------------------------- */
/* detect_issues (pipeline/metrics.py:166-179) on one HOST metrics row: OR of MDIMG_ISSUE_* in the reference's
 * order noise, blur, low_contrast, clipping_low, clipping_high (THRESHOLDS, metrics.py:25-34). */
int mdimg_detect_issues(const double* metrics_row);
/* The scalar part of compute_validation (pipeline/metrics.py:237-329) from one HOST row of mdimg_validation:
 * gains, quality_improvement, threshold tests and the pass rule, in the reference's python-float arithmetic. */
typedef struct mdimg_validation_scalars {
    double ssim, psnr, quality_improvement;
    int32_t meets_ssim, meets_psnr, meets_improvement, passes;
    double niqe_before, niqe_after;
    int32_t niqe_improved;
    double contrast_gain, sharpness_gain, noise_change;
    double entropy_change, snr_change, cnr_change, edge_density_change, histogram_spread_change;
    double edge_ratio, local_contrast_change, gradient_strength_change, gradient_entropy_change;
} mdimg_validation_scalars;
int mdimg_validation_scalars_of(const double* validation_row, mdimg_validation_scalars* out);
/* compute_objective_score (pipeline/metrics.py:337-408): the score BEFORE its round(., 4) and the eleven
 * penalty / reward parts in the order of the reference's breakdown dict (contrast_gain, sharpness_gain,
 * noise_penalty, niqe_degradation, halo_penalty, entropy_penalty, snr_reward, hs_reward,
 * local_contrast_reward, gradient_strength_reward, gradient_entropy_penalty), also unrounded. */
int mdimg_objective_score(const mdimg_validation_scalars* v, double* score, double parts[11]);

#ifdef __cplusplus
}
#endif
#endif /* MDIMG_B200_H */
