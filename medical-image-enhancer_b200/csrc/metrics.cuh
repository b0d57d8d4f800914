// Internal interface of the quality-metrics kernels (metrics.cu, fullref.cu).
#pragma once
#include "common.cuh"
#include "select.cuh"

namespace mdimg {

// Column layout of one per-slice result row (double) written by metrics_run.
// 0..15 follow the key order of the reference's compute_metrics dict (pipeline/metrics.py:90-109).
enum MetricCol {
    MC_SIGMA = 0, MC_LAP_VAR, MC_STD, MC_PCT_LOW, MC_PCT_HIGH, MC_ENTROPY, MC_EDGE_DENSITY,
    MC_GRAD_MEAN, MC_GRAD_STD, MC_SNR, MC_CNR, MC_LAP_ENERGY, MC_HIST_SPREAD, MC_LOCAL_CONTRAST,
    MC_GRAD_STRENGTH, MC_GRAD_ENTROPY,
    MC_MEAN = 16, MC_EDGE_RATIO = 17, MC_NIQE = 18, MC_VAR_OF_VAR = 19, MC_GMAX = 20,
    MC_P05 = 21, MC_P95 = 22, MC_RESERVED = 23,
    MC_COLS = 24
};

// Host-computed np.percentile(..., q) plan for arrays of `len` elements: numpy's 'linear'
// method with float32 virtual indices (numpy >= 2: q / float32(100)).
struct PctPlan {
    int lo[5];        // previous_indexes for q = 5, 25, 75, 95, 90
    int hi[5];        // next_indexes
    float gamma[5];   // float32 interpolation weight
};

size_t metrics_workspace_bytes(int n_sel, int h, int w);
// flags bit0: also compute the NIQE approximation (box-16 pass).
int metrics_run(const float* img, const Dims& d, const PctPlan& plan, int flags, double* out,
                void* ws, size_t ws_bytes, cudaStream_t stream);

size_t sigma_workspace_bytes(int n_sel, int h, int w);
// sigma_out: device [n] doubles (estimate_sigma); slices not in sel are left untouched.
int sigma_run(const float* img, const Dims& d, double* sigma_out, void* ws, size_t ws_bytes,
              cudaStream_t stream);

size_t quality_workspace_bytes(int n_sel, int h, int w);
// out: device [n][2] doubles = (edge_ratio, niqe_approx); flags bit0: compute niqe too.
int quality_run(const float* img, const Dims& d, int flags, double* out, void* ws, size_t ws_bytes,
                cudaStream_t stream);

size_t fullref_workspace_bytes(int n_sel, int h, int w);
// out: device [n][2] doubles = (ssim, psnr).
int fullref_run(const float* a, const float* b, const Dims& d, double* out, void* ws,
                size_t ws_bytes, cudaStream_t stream);

// out[s] = rows_a[s] (MC_COLS) | rows_b[s] (MC_COLS) | fr[s] (2) for the selected slices.
int validation_pack_run(const Dims& d, const double* rows_a, const double* rows_b, const double* fr,
                        double* out, cudaStream_t stream);

// Shared by metrics.cu / fullref.cu: box-16 local-variance statistics (NIQE, metrics.py:195-200).
// acc2: device [n_sel][2] doubles (sum lv, sum lv^2), must be zeroed by the caller.
void launch_box16_stats(const float* img, const Dims& d, double* acc2, cudaStream_t stream);

}  // namespace mdimg
