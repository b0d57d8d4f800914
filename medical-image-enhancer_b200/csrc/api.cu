// extern "C" entry points of libmdimg_b200.so (declared in include/mdimg_b200.h).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/mdimg_b200.h"
#include "common.cuh"
#include "enhance.cuh"
#include "metrics.cuh"
#include "select.cuh"

namespace mdimg {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return MDIMG_OK;
}

namespace {

bool bad_dims(int n, int h, int w, const int32_t* sel, int n_sel) {
    if (n < 0 || h < 1 || w < 1 || (sel && n_sel < 0) || (long long)h * w > 0x7fffffffLL) {
        set_error(MDIMG_ERR_INVALID, "invalid stack shape n=%d h=%d w=%d n_sel=%d", n, h, w, n_sel);
        return true;
    }
    const int ns = sel ? n_sel : n;
    if (ns > 65535) {
        set_error(MDIMG_ERR_INVALID, "at most 65535 slices per call (got %d); split the stack", ns);
        return true;
    }
    return false;
}

size_t mm_bytes(int n) { Arena a(nullptr, 0); a.take<uint2>(n); return a.off; }

struct ValidationBufs { double* rows_a; double* rows_b; double* fr; void* sub; size_t sub_bytes; };
void carve_validation(Arena& a, int n, int h, int w, ValidationBufs& b) {
    b.rows_a = a.take<double>((size_t)n * MC_COLS);
    b.rows_b = a.take<double>((size_t)n * MC_COLS);
    b.fr = a.take<double>((size_t)n * 2);
    const size_t m = metrics_workspace_bytes(n, h, w), f = fullref_workspace_bytes(n, h, w);
    b.sub_bytes = m > f ? m : f;                  // the three calls run one after another on the stream
    b.sub = a.take<char>(b.sub_bytes);
}

struct LightBufs { double* sigma; int* skip; float* tmp; void* sws; size_t sws_bytes; void* wws; size_t wws_bytes; };
void carve_light(Arena& a, int n, int n_sel, int h, int w, LightBufs& b) {
    b.sigma = a.take<double>(n);
    b.skip = a.take<int>(n);
    b.tmp = a.take<float>((size_t)n * h * w);
    b.sws_bytes = sigma_workspace_bytes(n, h, w);
    b.sws = a.take<char>(b.sws_bytes);
    b.wws_bytes = wavelet_workspace_bytes(n, n_sel, h, w);
    b.wws = a.take<char>(b.wws_bytes);
}

}  // namespace
}  // namespace mdimg

using namespace mdimg;

extern "C" {

const char* mdimg_last_error(void) { return g_err; }

int mdimg_version(void) { return 100; }

unsigned long long mdimg_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int mdimg_selftest_div16(unsigned long long* mismatches, void* stream) {
    if (!mismatches) return set_error(MDIMG_ERR_INVALID, "selftest: null result pointer");
    unsigned long long* dev = nullptr;
    if (cudaMalloc(&dev, sizeof(*dev)) != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "selftest: cudaMalloc failed");
    int rc = selftest_div16_run(dev, (cudaStream_t)stream);
    if (!rc && cudaMemcpyAsync(mismatches, dev, sizeof(*dev), cudaMemcpyDeviceToHost, (cudaStream_t)stream) != cudaSuccess)
        rc = set_error(MDIMG_ERR_CUDA, "selftest: copy failed");
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dev);
    return rc;
}

int mdimg_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return set_error(MDIMG_ERR_NO_DEVICE, "no CUDA device available (%s); mdimg_b200 has no CPU fallback",
                         e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= count) return set_error(MDIMG_ERR_INVALID, "device %d out of range [0, %d)", device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return set_error(MDIMG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                         device, prop.major, prop.minor);
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return MDIMG_OK;
}

int mdimg_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes, size_t* total_mem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return MDIMG_OK;
}

size_t mdimg_workspace_bytes(int op, int n, int h, int w, int param) {
    if (n < 1) n = 1;
    switch (op) {
        case MDIMG_OP_NORMALIZE:
        case MDIMG_OP_MINMAX: return mm_bytes(n) + 256;
        case MDIMG_OP_METRICS: return metrics_workspace_bytes(n, h, w);
        case MDIMG_OP_SIGMA: return sigma_workspace_bytes(n, h, w);
        case MDIMG_OP_QUALITY: return quality_workspace_bytes(n, h, w);
        case MDIMG_OP_FULLREF: return fullref_workspace_bytes(n, h, w);
        case MDIMG_OP_VALIDATION: {
            Arena a(nullptr, 0);
            ValidationBufs b;
            carve_validation(a, n, h, w, b);
            return a.off;
        }
        case MDIMG_OP_WAVELET: return wavelet_workspace_bytes(n, n, h, w);
        case MDIMG_OP_CLAHE: return mm_bytes(n) + clahe_workspace_bytes(n, n, h, w, param);
        case MDIMG_OP_GAMMA:
        case MDIMG_OP_UNSHARP: return mm_bytes(n);
        case MDIMG_OP_LIGHT_DENOISE: {
            Arena a(nullptr, 0);
            LightBufs b;
            carve_light(a, n, n, h, w, b);
            return a.off;
        }
        case MDIMG_OP_BILATERAL: return 0;
        case MDIMG_OP_ENHANCE: return enhance_workspace_bytes(n, h, w, param);
        case MDIMG_OP_TV: return tv_workspace_bytes(n, n, h, w, param > 0 ? param : 200);
        default: return 0;
    }
}

int mdimg_minmax_f32(const float* img, int n, int h, int w, const int32_t* sel, int n_sel,
                     float* out_minmax, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm = a.take<uint2>(n);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "minmax: workspace too small");
    int rc = minmax_f32_run(img, d, mm, (cudaStream_t)stream);
    if (rc) return rc;
    return minmax_decode_run(mm, d, out_minmax, (cudaStream_t)stream);
}

int mdimg_mosaic_u8(const float* before, const float* after, uint8_t* out, int n, int h, int w,
                    const int32_t* sel, int n_sel, int gap, int gap_level, void* ws, size_t ws_bytes,
                    void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm_b = a.take<uint2>(n);
    uint2* mm_a = a.take<uint2>(n);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "mosaic: workspace too small");
    int rc = minmax_f32_run(before, d, mm_b, (cudaStream_t)stream);
    if (rc) return rc;
    rc = minmax_f32_run(after, d, mm_a, (cudaStream_t)stream);
    if (rc) return rc;
    return mosaic_u8_run(before, after, out, d, gap, gap_level, mm_b, mm_a, (cudaStream_t)stream);
}

int mdimg_normalize_u16(const uint16_t* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm = a.take<uint2>(n);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "normalize: workspace too small");
    int rc = minmax_u16_run(in, d, mm, (cudaStream_t)stream);
    if (rc) return rc;
    return normalize_u16_run(in, out, d, mm, (cudaStream_t)stream);
}

int mdimg_ingest_u16(const uint16_t* raw, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                     double slope, double intercept, int has_rescale, int monochrome1, int is_signed,
                     void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm = a.take<uint2>(n);
    uint2* gmm = a.take<uint2>(1);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "ingest: workspace too small");
    return ingest_run(raw, out, d, slope, intercept, has_rescale, monochrome1, is_signed, mm, gmm, (cudaStream_t)stream);
}

int mdimg_normalize_f32(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm = a.take<uint2>(n);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "normalize: workspace too small");
    int rc = minmax_f32_run(in, d, mm, (cudaStream_t)stream);
    if (rc) return rc;
    return normalize_f32_run(in, out, d, mm, (cudaStream_t)stream);
}

int mdimg_metrics(const float* img, int n, int h, int w, const int32_t* sel, int n_sel, int flags,
                  const int32_t* pct_lo, const int32_t* pct_hi, const float* pct_gamma,
                  double* out, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    if (!pct_lo || !pct_hi || !pct_gamma) return set_error(MDIMG_ERR_INVALID, "metrics: percentile plan missing");
    PctPlan plan;
    const int len = h * w;
    for (int i = 0; i < 5; ++i) {
        plan.lo[i] = pct_lo[i] < 0 ? len + pct_lo[i] : pct_lo[i];   // numpy uses -1 for "last"
        plan.hi[i] = pct_hi[i] < 0 ? len + pct_hi[i] : pct_hi[i];
        if (plan.lo[i] >= len || plan.hi[i] >= len) return set_error(MDIMG_ERR_INVALID, "metrics: percentile index out of range");
        plan.gamma[i] = pct_gamma[i];
    }
    Dims d = make_dims(n, h, w, sel, n_sel);
    return metrics_run(img, d, plan, flags, out, ws, ws_bytes, (cudaStream_t)stream);
}

int mdimg_estimate_sigma(const float* img, int n, int h, int w, const int32_t* sel, int n_sel,
                         double* sigma, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return sigma_run(img, d, sigma, ws, ws_bytes, (cudaStream_t)stream);
}

int mdimg_quality(const float* img, int n, int h, int w, const int32_t* sel, int n_sel, int flags,
                  double* out, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return quality_run(img, d, flags, out, ws, ws_bytes, (cudaStream_t)stream);
}

int mdimg_fullref(const float* original, const float* enhanced, int n, int h, int w,
                  const int32_t* sel, int n_sel, double* out, void* ws, size_t ws_bytes,
                  void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return fullref_run(original, enhanced, d, out, ws, ws_bytes, (cudaStream_t)stream);
}

int mdimg_validation(const float* original, const float* enhanced, int n, int h, int w,
                     const int32_t* sel, int n_sel, const int32_t* pct_lo, const int32_t* pct_hi,
                     const float* pct_gamma, double* out, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Arena a(ws, ws_bytes);
    ValidationBufs b;
    carve_validation(a, n, h, w, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "validation: workspace too small (%zu > %zu)", a.off, ws_bytes);
    int rc = mdimg_metrics(original, n, h, w, sel, n_sel, 1, pct_lo, pct_hi, pct_gamma, b.rows_a, b.sub, b.sub_bytes, stream);
    if (rc) return rc;
    rc = mdimg_metrics(enhanced, n, h, w, sel, n_sel, 1, pct_lo, pct_hi, pct_gamma, b.rows_b, b.sub, b.sub_bytes, stream);
    if (rc) return rc;
    rc = mdimg_fullref(original, enhanced, n, h, w, sel, n_sel, b.fr, b.sub, b.sub_bytes, stream);
    if (rc) return rc;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return validation_pack_run(d, b.rows_a, b.rows_b, b.fr, out, (cudaStream_t)stream);
}

int mdimg_wavelet_denoise(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                          int n_sel, int mode_hard, const double* sigma_in, double sigma_scale,
                          const int32_t* skip, void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return wavelet_denoise_run(in, out, d, mode_hard, sigma_in, sigma_scale, skip, ws, ws_bytes,
                               (cudaStream_t)stream);
}

int mdimg_clahe_gamma(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                      double clip_limit, int kernel_size, double gamma, int32_t* status, void* ws,
                      size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    Arena a(ws, ws_bytes);
    uint2* mm = a.take<uint2>(n);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "clahe: workspace too small");
    int rc = minmax_f32_run(in, d, mm, (cudaStream_t)stream);
    if (rc) return rc;
    return clahe_run(in, out, d, clip_limit, kernel_size, gamma, mm, status, (char*)ws + a.off,
                     ws_bytes - a.off, (cudaStream_t)stream);
}

int mdimg_clahe(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                double clip_limit, int kernel_size, int32_t* status, void* ws, size_t ws_bytes,
                void* stream) {
    return mdimg_clahe_gamma(in, out, n, h, w, sel, n_sel, clip_limit, kernel_size, 1.0, status, ws,
                             ws_bytes, stream);
}

int mdimg_gamma(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                double gamma, int assume_nonneg, int32_t* neg_flag, void* ws, size_t ws_bytes,
                void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    uint2* mm = nullptr;
    if (!assume_nonneg) {
        Arena a(ws, ws_bytes);
        mm = a.take<uint2>(n);
        if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "gamma: workspace too small");
        int rc = minmax_f32_run(in, d, mm, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return gamma_run(in, out, d, gamma, mm, neg_flag, (cudaStream_t)stream);
}

int mdimg_unsharp(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                  const double* weights, int gauss_radius, double amount, int assume_nonneg,
                  void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    if (!weights) return set_error(MDIMG_ERR_INVALID, "unsharp: weights missing");
    Dims d = make_dims(n, h, w, sel, n_sel);
    uint2* mm = nullptr;
    if (!assume_nonneg) {
        Arena a(ws, ws_bytes);
        mm = a.take<uint2>(n);
        if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "unsharp: workspace too small");
        int rc = minmax_f32_run(in, d, mm, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return unsharp_run(in, out, d, weights, gauss_radius, (float)amount, mm, (cudaStream_t)stream);
}

int mdimg_light_denoise(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                        int n_sel, double strength, int32_t* skipped, void* ws, size_t ws_bytes,
                        void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    cudaStream_t st = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    LightBufs b;
    carve_light(a, n, d.n_sel, h, w, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "light_denoise: workspace too small (%zu > %zu)", a.off, ws_bytes);
    int rc = sigma_run(in, d, b.sigma, b.sws, b.sws_bytes, st);
    if (rc) return rc;
    rc = skip_flags_run(d, b.sigma, 0.001, b.skip, st, skipped);
    if (rc) return rc;
    // (1 - strength) * image + strength * denoised, python-float scalars acting as float32: the blend is the
    // epilogue of the last inverse wavelet level (skipped slices are copied through)
    return wavelet_denoise_run(in, out, d, 0, b.sigma, 0.5, b.skip, b.wws, b.wws_bytes, st,
                               (float)(1.0 - strength), (float)strength, 1);
}

int mdimg_bilateral(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                    int n_sel, int dd, const double* spatial, double sigma_color, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    if (!spatial) return set_error(MDIMG_ERR_INVALID, "bilateral: spatial weights missing");
    Dims d = make_dims(n, h, w, sel, n_sel);
    return bilateral_run(in, out, d, dd, spatial, sigma_color, (cudaStream_t)stream);
}

int mdimg_tv_chambolle(const float* in, float* out, int n, int h, int w, const int32_t* sel,
                       int n_sel, double weight, double eps, int max_iter, int32_t* iters,
                       void* ws, size_t ws_bytes, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return tv_chambolle_run(in, out, d, weight, eps, max_iter, iters, ws, ws_bytes, (cudaStream_t)stream);
}

int mdimg_axpby(const float* a, const float* b, float* out, int n, int h, int w, const int32_t* sel,
                int n_sel, double c0, double c1, int clip01, void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return axpby_run(a, b, out, d, (float)c0, (float)c1, clip01, (cudaStream_t)stream);
}

int mdimg_clip01(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
                 void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return clip01_run(in, out, d, (cudaStream_t)stream);
}

int mdimg_export_u16(const float* in, uint16_t* out, int n, int h, int w, const int32_t* sel, int n_sel,
                     void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return export_u16_run(in, out, d, (cudaStream_t)stream);
}

int mdimg_copy(const float* in, float* out, int n, int h, int w, const int32_t* sel, int n_sel,
               void* stream) {
    if (bad_dims(n, h, w, sel, n_sel)) return MDIMG_ERR_INVALID;
    Dims d = make_dims(n, h, w, sel, n_sel);
    return copy_run(in, out, d, (cudaStream_t)stream);
}

}  // extern "C"
