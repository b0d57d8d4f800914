// Full-reference validation kernels: SSIM + PSNR (skimage.metrics, called at
// pipeline/metrics.py:232-233) and the 16x16 local-variance statistics of the NIQE
// approximation (pipeline/metrics.py:195-200).
//
// Box means follow scipy.ndimage.uniform_filter on float32: axis 0 first, each pass summed in
// double, scaled, rounded to float32 before the next pass; window [i-3, i+3] for size 7 and
// [i-8, i+7] for size 16; half-sample symmetric border.
#include "metrics.cuh"
#include "boxfilter.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int TW = 64, TH = 32;

// ---------------- NIQE box-16 ----------------
typedef BoxTile<16> B16;
constexpr int B16L = 8;   // window [i-8, i+7]

struct Box16Smem {
    float X[B16::XH * B16::XP];
    float VS[B16::TH * B16::XP];
    float VQ[B16::TH * B16::XP];
    double red[2 * 32];
};

__global__ void __launch_bounds__(NT)
k_box16_stats(const float* __restrict__ img, Dims d, double* __restrict__ acc2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Box16Smem& sm = *reinterpret_cast<Box16Smem*>(smem_raw);
    constexpr int XW = B16::XW, XH = B16::XH, XP = B16::XP;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = img + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x;
    load_tile<XW, XH, B16L, 0>(src, d.h, d.w, x0, y0, [&](int r, int c, float v) { sm.X[r * XP + c] = v; });
    __syncthreads();
    box_vertical_xq<16>(sm.X, sm.VS, sm.VQ, 0.0625);
    __syncthreads();
    double v[2] = {0.0, 0.0};
    {
        float* const vin[2] = {sm.VS, sm.VQ};
        box_horizontal_f<16, 2>(vin, 0.0625, [&](int r, int c, const float (&m)[2]) {
            if (y0 + r < d.h && x0 + c < d.w) {
                const double lv = (double)fmaxf(__fsub_rn(m[1], __fmul_rn(m[0], m[0])), 0.0f);
                v[0] += lv;
                v[1] += lv * lv;
            }
        });
    }
    block_sum<2>(v, sm.red);
    if (tid == 0) {
        atomicAdd(&acc2[(size_t)si * 2 + 0], v[0]);
        atomicAdd(&acc2[(size_t)si * 2 + 1], v[1]);
    }
}

// ---------------- SSIM + PSNR ----------------
typedef BoxTile<7> B7;
constexpr int HALO = 3;

struct SsimSmem {
    float A[B7::XH * B7::XP];
    float B[B7::XH * B7::XP];
    float V[5][B7::TH * B7::XP];    // axis-0 means of a, b, a*a, b*b, a*b (float32, as scipy stores them)
    double red[2 * 32];
};

__global__ void __launch_bounds__(NT)
k_ssim_psnr(const float* __restrict__ ia, const float* __restrict__ ib, Dims d,
            double* __restrict__ acc2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SsimSmem& sm = *reinterpret_cast<SsimSmem*>(smem_raw);
    constexpr int XW = B7::XW, XH = B7::XH, XP = B7::XP;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* pa = ia + (size_t)s * d.h * d.w;
    const float* pb = ib + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x;
    load_tile<XW, XH, HALO, 0>(pa, d.h, d.w, x0, y0, [&](int r, int c, float v) { sm.A[r * XP + c] = v; });
    load_tile<XW, XH, HALO, 0>(pb, d.h, d.w, x0, y0, [&](int r, int c, float v) { sm.B[r * XP + c] = v; });
    __syncthreads();
    const double inv7 = 1.0 / 7.0;
    // axis-0 pass: sliding window; the five operands are formed (and rounded to float32, as
    // numpy's im1 * im2 is) when a row enters or leaves the window
    if (tid < XW * B7::NSEG) {
        const int sg = tid / XW, c = tid - sg * XW;
        const int r0 = sg * B7::RS, r1 = min(r0 + B7::RS, TH);
        double sum[5] = {0, 0, 0, 0, 0};
        auto add = [&](int r, double sign) {
            const float a = sm.A[r * XP + c], b = sm.B[r * XP + c];
            sum[0] += sign * (double)a;
            sum[1] += sign * (double)b;
            sum[2] += sign * (double)__fmul_rn(a, a);
            sum[3] += sign * (double)__fmul_rn(b, b);
            sum[4] += sign * (double)__fmul_rn(a, b);
        };
#pragma unroll
        for (int k = 0; k < 7; ++k) add(r0 + k, 1.0);
        for (int r = r0; r < r1; ++r) {
#pragma unroll
            for (int q = 0; q < 5; ++q) sm.V[q][r * XP + c] = (float)(sum[q] * inv7);
            if (r + 1 < r1) { add(r + 7, 1.0); add(r, -1.0); }
        }
    }
    __syncthreads();
    const float cn = (float)(49.0 / 48.0);
    const float C1 = (float)1.0e-4, C2 = (float)9.0e-4;
    double v[2] = {0.0, 0.0};   // sum S over the crop, sum (a-b)^2 over the image
    {
        float* const vin[5] = {sm.V[0], sm.V[1], sm.V[2], sm.V[3], sm.V[4]};
        box_horizontal_f<7, 5>(vin, inv7, [&](int r, int c, const float (&m)[5]) {
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) {
                const float a = sm.A[(r + HALO) * XP + c + HALO], b = sm.B[(r + HALO) * XP + c + HALO];
                const float df = __fsub_rn(a, b);
                v[1] += (double)__fmul_rn(df, df);
                if (gy >= HALO && gy < d.h - HALO && gx >= HALO && gx < d.w - HALO) {
                    const float ux = m[0], uy = m[1], uxx = m[2], uyy = m[3], uxy = m[4];
                    const float vx = __fmul_rn(cn, __fsub_rn(uxx, __fmul_rn(ux, ux)));
                    const float vy = __fmul_rn(cn, __fsub_rn(uyy, __fmul_rn(uy, uy)));
                    const float vxy = __fmul_rn(cn, __fsub_rn(uxy, __fmul_rn(ux, uy)));
                    const float A1 = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, ux), uy), C1);
                    const float A2 = __fadd_rn(__fmul_rn(2.0f, vxy), C2);
                    const float B1 = __fadd_rn(__fadd_rn(__fmul_rn(ux, ux), __fmul_rn(uy, uy)), C1);
                    const float B2 = __fadd_rn(__fadd_rn(vx, vy), C2);
                    const float S = __fdiv_rn(__fmul_rn(A1, A2), __fmul_rn(B1, B2));
                    v[0] += (double)S;
                }
            }
        });
    }
    block_sum<2>(v, sm.red);
    if (tid == 0) {
        atomicAdd(&acc2[(size_t)si * 2 + 0], v[0]);
        atomicAdd(&acc2[(size_t)si * 2 + 1], v[1]);
    }
}

__global__ void k_fullref_out(Dims d, const double* __restrict__ acc2, double* __restrict__ out) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    const double ncrop = (double)(d.h - 2 * HALO) * (double)(d.w - 2 * HALO);
    const double N = (double)d.h * (double)d.w;
    out[s * 2] = acc2[si * 2] / ncrop;
    const double mse = acc2[si * 2 + 1] / N;
    out[s * 2 + 1] = 10.0 * log10(1.0 / mse);   // +inf when the images are identical
}

__global__ void k_validation_pack(Dims d, const double* __restrict__ ra, const double* __restrict__ rb,
                                  const double* __restrict__ fr, double* __restrict__ out) {
    constexpr int W = 2 * MC_COLS + 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_sel * W) return;
    const int si = i / W, c = i - si * W;
    const int s = slice_of(d.sel, si);
    double v;
    if (c < MC_COLS) v = ra[(size_t)s * MC_COLS + c];
    else if (c < 2 * MC_COLS) v = rb[(size_t)s * MC_COLS + c - MC_COLS];
    else v = fr[(size_t)s * 2 + c - 2 * MC_COLS];
    out[(size_t)s * W + c] = v;
}

}  // namespace

int validation_pack_run(const Dims& d, const double* rows_a, const double* rows_b, const double* fr,
                        double* out, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    const int total = d.n_sel * (2 * MC_COLS + 2);
    MDIMG_LAUNCH k_validation_pack<<<(total + 255) / 256, 256, 0, stream>>>(d, rows_a, rows_b, fr, out);
    return check_launch("validation");
}

void launch_box16_stats(const float* img, const Dims& d, double* acc2, cudaStream_t stream) {
    static unsigned long long devices_done = 0;
    opt_in_shared_memory(k_box16_stats, sizeof(Box16Smem), devices_done);
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
    MDIMG_LAUNCH k_box16_stats<<<grid, NT, sizeof(Box16Smem), stream>>>(img, d, acc2);
}

size_t fullref_workspace_bytes(int n_sel, int h, int w) {
    (void)h; (void)w;
    Arena a(nullptr, 0);
    a.take<double>((size_t)n_sel * 2);
    return a.off;
}

int fullref_run(const float* ia, const float* ib, const Dims& d, double* out, void* ws,
                size_t ws_bytes, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (d.h < 7 || d.w < 7) return set_error(MDIMG_ERR_INVALID, "win_size exceeds image extent.");
    Arena a(ws, ws_bytes);
    double* acc2 = a.take<double>((size_t)d.n_sel * 2);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "fullref: workspace too small");
    static unsigned long long devices_done = 0;
    opt_in_shared_memory(k_ssim_psnr, sizeof(SsimSmem), devices_done);
    cudaMemsetAsync(acc2, 0, sizeof(double) * 2 * d.n_sel, stream);
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
    MDIMG_LAUNCH k_ssim_psnr<<<grid, NT, sizeof(SsimSmem), stream>>>(ia, ib, d, acc2);
    MDIMG_LAUNCH k_fullref_out<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, acc2, out);
    return check_launch("fullref");
}

}  // namespace mdimg
