/* A host in plain C on top of libmdimg_b200.so: no Python, no torch -- the CUDA runtime for device memory
 * and the C ABI of include/mdimg_b200.h, one call per function of the reference's hot path:
 *
 *   normalize_image                   -> mdimg_normalize_u16        (pipeline/dicom_io.py:84-91)
 *   compute_metrics + detect_issues   -> mdimg_metrics, mdimg_detect_issues   (pipeline/metrics.py:42-179)
 *   apply_enhancements_from_params    -> mdimg_enhance              (pipeline/enhancement.py:235-369)
 *   compute_validation                -> mdimg_validation, mdimg_validation_scalars_of (metrics.py:225-329)
 *   compute_objective_score           -> mdimg_objective_score      (metrics.py:337-408)
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_host.c -o c_host \
 *       -L medical-image-enhancer_b200 -lmdimg_b200 -L /usr/local/cuda/lib64 -lcudart -lm
 *   ./c_host stack.u16 n h w enhanced.f32
 *
 * Reads n*h*w uint16 pixels, writes the enhanced float32 stack and prints one line per slice.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "mdimg_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != MDIMG_OK) {                                                           \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mdimg_last_error());     \
            return 1;                                                                    \
        }                                                                                \
    } while (0)
#define CUDA(call)                                                                       \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                  \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

int main(int argc, char** argv) {
    if (argc != 6) {
        fprintf(stderr, "usage: %s stack.u16 n h w enhanced.f32\n", argv[0]);
        return 2;
    }
    const int n = atoi(argv[2]), h = atoi(argv[3]), w = atoi(argv[4]);
    const size_t px = (size_t)n * h * w;
    uint16_t* raw = (uint16_t*)malloc(px * sizeof(uint16_t));
    FILE* f = fopen(argv[1], "rb");
    if (!raw || !f || fread(raw, sizeof(uint16_t), px, f) != px) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);

    CHECK(mdimg_init(0));                              /* fails without an sm_100 device: there is no CPU path */

    /* the plan: P_full of the benchmark */
    mdimg_enhance_plan plan = {0};
    plan.n_ops = 7;
    for (int i = 0; i < 7; ++i) plan.ops[i] = i;       /* denoise, clahe, gamma, unsharp, post_denoise, bilateral, tv_denoise */
    plan.clahe_clip_limit = 0.015; plan.clahe_tile_size = 16; plan.gamma = 0.95;
    plan.unsharp_radius = 0.8; plan.unsharp_amount = 0.5; plan.denoise_hard = 0;
    plan.post_denoise_strength = 0.3; plan.bilateral_d = 5; plan.bilateral_sigma_color = 0.05;
    plan.bilateral_sigma_space = 0.05; plan.tv_denoise_weight = 0.05;
    mdimg_enhance_tables tables;
    CHECK(mdimg_plan_clamp(&plan));
    CHECK(mdimg_enhance_tables_default(&plan, h, w, &tables));

    size_t wsb = mdimg_workspace_bytes(MDIMG_OP_ENHANCE, n, h, w, plan.clahe_tile_size);
    wsb = max_sz(wsb, mdimg_workspace_bytes(MDIMG_OP_VALIDATION, n, h, w, 0));
    wsb = max_sz(wsb, mdimg_workspace_bytes(MDIMG_OP_NORMALIZE, n, h, w, 0));
    wsb = max_sz(wsb, mdimg_workspace_bytes(MDIMG_OP_METRICS, n, h, w, 0));

    uint16_t* d_raw; float *d_x, *d_y; double *d_rows, *d_val; void* d_ws;
    cudaStream_t stream;
    CUDA(cudaStreamCreate(&stream));
    CUDA(cudaMalloc((void**)&d_raw, px * sizeof(uint16_t)));
    CUDA(cudaMalloc((void**)&d_x, px * sizeof(float)));
    CUDA(cudaMalloc((void**)&d_y, px * sizeof(float)));
    CUDA(cudaMalloc((void**)&d_rows, (size_t)n * MDIMG_METRIC_COLS * sizeof(double)));
    CUDA(cudaMalloc((void**)&d_val, (size_t)n * MDIMG_VALIDATION_COLS * sizeof(double)));
    CUDA(cudaMalloc(&d_ws, wsb));
    CUDA(cudaMemcpyAsync(d_raw, raw, px * sizeof(uint16_t), cudaMemcpyHostToDevice, stream));

    CHECK(mdimg_normalize_u16(d_raw, d_x, n, h, w, NULL, 0, d_ws, wsb, stream));
    CHECK(mdimg_metrics(d_x, n, h, w, NULL, 0, 1, tables.pct_lo, tables.pct_hi, tables.pct_gamma, d_rows, d_ws, wsb, stream));

    int32_t* flags = (int32_t*)calloc((size_t)n, sizeof(int32_t));
    int32_t* iters = (int32_t*)calloc((size_t)n, sizeof(int32_t));
    CHECK(mdimg_enhance(d_x, d_y, n, h, w, &plan, &tables, d_rows, NULL, flags, iters, d_ws, wsb, stream));
    CHECK(mdimg_validation(d_x, d_y, n, h, w, NULL, 0, tables.pct_lo, tables.pct_hi, tables.pct_gamma, d_val, d_ws, wsb, stream));

    double* val = (double*)malloc((size_t)n * MDIMG_VALIDATION_COLS * sizeof(double));
    float* enh = (float*)malloc(px * sizeof(float));
    CUDA(cudaMemcpyAsync(val, d_val, (size_t)n * MDIMG_VALIDATION_COLS * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CUDA(cudaMemcpyAsync(enh, d_y, px * sizeof(float), cudaMemcpyDeviceToHost, stream));
    CUDA(cudaStreamSynchronize(stream));

    for (int i = 0; i < n; ++i) {
        const double* row = val + (size_t)i * MDIMG_VALIDATION_COLS;
        mdimg_validation_scalars v;
        double score, parts[11];
        CHECK(mdimg_validation_scalars_of(row, &v));
        CHECK(mdimg_objective_score(&v, &score, parts));
        printf("slice %d issues %d flags %d tv_iters %d sigma_before %.17g sigma_after %.17g entropy_after %.17g "
               "ssim %.17g psnr %.17g quality_improvement %.17g passes %d score %.17g\n",
               i, mdimg_detect_issues(row), (int)flags[i], (int)iters[i], row[0], row[MDIMG_METRIC_COLS],
               row[MDIMG_METRIC_COLS + 5], v.ssim, v.psnr, v.quality_improvement, (int)v.passes, score);
    }
    f = fopen(argv[5], "wb");
    if (!f || fwrite(enh, sizeof(float), px, f) != px) { fprintf(stderr, "cannot write %s\n", argv[5]); return 2; }
    fclose(f);
    printf("launches %llu\n", mdimg_launch_count());
    return 0;
}
