// Internal interface of the enhancement kernels (pointwise.cu, unsharp.cu, clahe.cu,
// wavelet.cu, bilateral.cu, tv.cu).  Every image argument is a device pointer to
// [n][h][w] float32 unless stated otherwise; `d.sel` restricts the work to a subset of slices.
#pragma once
#include "common.cuh"

namespace mdimg {

// ---- pointwise.cu ------------------------------------------------------------------------
// mm: device [n] uint2 = (f2key(min), f2key(max)) per slice.
int minmax_f32_run(const float* img, const Dims& d, uint2* mm, cudaStream_t stream);
int minmax_u16_run(const uint16_t* img, const Dims& d, uint2* mm, cudaStream_t stream);
int minmax_decode_run(const uint2* mm, const Dims& d, float* out2, cudaStream_t stream);
// normalize_image (pipeline/dicom_io.py:84-91)
int normalize_u16_run(const uint16_t* in, float* out, const Dims& d, const uint2* mm, cudaStream_t stream);
int normalize_f32_run(const float* in, float* out, const Dims& d, const uint2* mm, cudaStream_t stream);
// Exhaustive device check of the normalise quotient (see div16 in pointwise.cu); *mismatches_dev receives the count.
int selftest_div16_run(unsigned long long* mismatches_dev, cudaStream_t stream);
// load_dicom's pixel path (modality rescale, MONOCHROME1 inversion) fused with normalize_image.
// mm: device [n] uint2 scratch, gmm: device [1] uint2 scratch.
int ingest_run(const uint16_t* in, float* out, const Dims& d, double slope, double intercept,
               int has_rescale, int invert, int is_signed, uint2* mm, uint2* gmm, cudaStream_t stream);
// adjust_gamma on float input; neg_flag: device [n] int, set to 1 where a slice has a negative pixel
// (the reference raises ValueError there; the slice is left untouched).
int gamma_run(const float* in, float* out, const Dims& d, double gamma, const uint2* mm,
              int* neg_flag, cudaStream_t stream);
// out = clip(c0*a + c1*b, 0, 1) (clip optional); float32 products and sum, numpy order.
int axpby_run(const float* a, const float* b, float* out, const Dims& d, float c0, float c1,
              int clip01, cudaStream_t stream);
int clip01_run(const float* in, float* out, const Dims& d, cudaStream_t stream);
// uint16(clip(rint(x * 65535), 0, 65535)) of the selected slices (16-bit export of an enhanced stack).
int export_u16_run(const float* in, uint16_t* out, const Dims& d, cudaStream_t stream);
// skip[s] = sigma[s] < thresh (device int[n]);  blend: out = skip ? a : c0*a + c1*b
int skip_flags_run(const Dims& d, const double* sigma, double thresh, int* skip, cudaStream_t stream,
                   int* skipped_out = nullptr);
int blend_skip_run(const float* a, const float* b, float* out, const Dims& d, float c0, float c1,
                   const int* skip, int* skipped_out, cudaStream_t stream);
int copy_run(const float* in, float* out, const Dims& d, cudaStream_t stream);
// Before | after report mosaic in 8-bit gray (save_visuals, pipeline/dicom_io.py:99-126): each panel
// autoscaled to its own min / max (mm_b, mm_a) and quantised like matplotlib's gray colormap.
// out: device uint8 [n][h][2w + gap].
int mosaic_u8_run(const float* before, const float* after, uint8_t* out, const Dims& d, int gap,
                  int gap_level, const uint2* mm_b, const uint2* mm_a, cudaStream_t stream);

// ---- unsharp.cu ----------------------------------------------------------------------------
// weights: host array of radius+1 doubles (w[0] centre ... w[radius]); radius <= 12.
int unsharp_run(const float* in, float* out, const Dims& d, const double* weights, int radius,
                float amount, const uint2* mm, cudaStream_t stream);

// ---- clahe.cu ------------------------------------------------------------------------------
size_t clahe_workspace_bytes(int n, int n_sel, int h, int w, int kernel_size);
// status: device [n] int; 1 = input outside [-1, 1] (the reference raises ValueError).
// gamma != 1: exposure.adjust_gamma applied to the CLAHE result in the same final pass.
int clahe_run(const float* in, float* out, const Dims& d, double clip_limit, int kernel_size, double gamma,
              const uint2* mm, int* status, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---- wavelet.cu ----------------------------------------------------------------------------
size_t wavelet_workspace_bytes(int n, int n_sel, int h, int w);
// mode_hard: 0 soft / 1 hard.  sigma_in: device [n] doubles or nullptr (estimate from the finest
// Haar 'dd' band).  sigma_scale multiplies sigma_in (light denoise passes 0.5).
// skip: device [n] int or nullptr; slices with skip[s] != 0 are copied through unchanged.
// blend_on: the result is blend_c0 * in + blend_c1 * denoised (float32, numpy order) instead of the denoised
// image -- _light_denoise's blend fused into the last inverse level; skipped slices are copied through.
int wavelet_denoise_run(const float* in, float* out, const Dims& d, int mode_hard,
                        const double* sigma_in, double sigma_scale, const int* skip,
                        void* ws, size_t ws_bytes, cudaStream_t stream,
                        float blend_c0 = 0.0f, float blend_c1 = 0.0f, int blend_on = 0);

// ---- bilateral.cu --------------------------------------------------------------------------
// spatial: host array of dd*dd doubles (row-major dy, dx), dd odd <= 9.
int bilateral_run(const float* in, float* out, const Dims& d, int dd, const double* spatial,
                  double sigma_color, cudaStream_t stream);

// ---- tv.cu ---------------------------------------------------------------------------------
size_t tv_workspace_bytes(int n, int n_sel, int h, int w, int max_iter);
// iters_out: device [n] int (number of loop bodies executed per slice) or nullptr.
int tv_chambolle_run(const float* in, float* out, const Dims& d, double weight, double eps,
                     int max_iter, int* iters_out, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---- engine.cu (mdimg_enhance) ---------------------------------------------------------------
size_t enhance_workspace_bytes(int n, int h, int w, int clahe_kernel_size);

}  // namespace mdimg
