// Full-reference validation kernels: SSIM + PSNR (skimage.metrics, called at
// pipeline/metrics.py:232-233) and the 16x16 local-variance statistics of the NIQE
// approximation (pipeline/metrics.py:195-200).
//
// Box means follow scipy.ndimage.uniform_filter on float32: axis 0 first, each pass summed in
// double, scaled, rounded to float32 before the next pass; window [i-3, i+3] for size 7 and
// [i-8, i+7] for size 16; half-sample symmetric border.
#include "metrics.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int TW = 64, TH = 32;

// ---------------- NIQE box-16 ----------------
constexpr int B16L = 8, B16R = 7;
constexpr int B16W = TW + B16L + B16R;   // 79
constexpr int B16H = TH + B16L + B16R;   // 47
constexpr int B16P = B16W + 1;

struct Box16Smem {
    float X[B16H][B16P];
    float VS[TH][B16P];
    float VQ[TH][B16P];
    double red[2 * 32];
};

__global__ void __launch_bounds__(NT)
k_box16_stats(const float* __restrict__ img, Dims d, double* __restrict__ acc2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Box16Smem& sm = *reinterpret_cast<Box16Smem*>(smem_raw);
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = img + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < B16H * B16W; i += NT) {
        int r = i / B16W, c = i - r * B16W;
        int gy = refl_sym(y0 + r - B16L, d.h), gx = refl_sym(x0 + c - B16L, d.w);
        sm.X[r][c] = src[(size_t)gy * d.w + gx];
    }
    __syncthreads();
    for (int i = tid; i < TH * B16W; i += NT) {
        int r = i / B16W, c = i - r * B16W;
        double a = 0.0, q = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float v = sm.X[r + k][c];
            a += (double)v;
            q += (double)__fmul_rn(v, v);
        }
        sm.VS[r][c] = (float)(a * 0.0625);
        sm.VQ[r][c] = (float)(q * 0.0625);
    }
    __syncthreads();
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < TH / 8; ++j)
#pragma unroll
        for (int i = 0; i < TW / 32; ++i) {
            const int r = wid + 8 * j, c = lane + 32 * i;
            if (y0 + r < d.h && x0 + c < d.w) {
                double ms = 0.0, mq = 0.0;
#pragma unroll
                for (int k = 0; k < 16; ++k) { ms += (double)sm.VS[r][c + k]; mq += (double)sm.VQ[r][c + k]; }
                const float m = (float)(ms * 0.0625), q = (float)(mq * 0.0625);
                const float lv = fmaxf(__fsub_rn(q, __fmul_rn(m, m)), 0.0f);
                v[0] += (double)lv;
                v[1] += (double)lv * (double)lv;
            }
        }
    block_sum<2>(v, sm.red);
    if (tid == 0) {
        atomicAdd(&acc2[(size_t)si * 2 + 0], v[0]);
        atomicAdd(&acc2[(size_t)si * 2 + 1], v[1]);
    }
}

// ---------------- SSIM + PSNR ----------------
constexpr int HALO = 3;
constexpr int XW = TW + 2 * HALO, XH = TH + 2 * HALO, XP = XW + 1;

struct SsimSmem {
    float A[XH][XP];
    float B[XH][XP];
    float V[5][TH][XP];   // axis-0 means of a, b, a*a, b*b, a*b
    double red[2 * 32];
};

__global__ void __launch_bounds__(NT)
k_ssim_psnr(const float* __restrict__ ia, const float* __restrict__ ib, Dims d,
            double* __restrict__ acc2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SsimSmem& sm = *reinterpret_cast<SsimSmem*>(smem_raw);
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* pa = ia + (size_t)s * d.h * d.w;
    const float* pb = ib + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < XH * XW; i += NT) {
        int r = i / XW, c = i - r * XW;
        int gy = refl_sym(y0 + r - HALO, d.h), gx = refl_sym(x0 + c - HALO, d.w);
        size_t o = (size_t)gy * d.w + gx;
        sm.A[r][c] = pa[o];
        sm.B[r][c] = pb[o];
    }
    __syncthreads();
    const double inv7 = 1.0 / 7.0;
    for (int i = tid; i < TH * XW; i += NT) {
        int r = i / XW, c = i - r * XW;
        double sa = 0, sb = 0, saa = 0, sbb = 0, sab = 0;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            float a = sm.A[r + k][c], b = sm.B[r + k][c];
            sa += (double)a; sb += (double)b;
            saa += (double)__fmul_rn(a, a); sbb += (double)__fmul_rn(b, b); sab += (double)__fmul_rn(a, b);
        }
        sm.V[0][r][c] = (float)(sa * inv7); sm.V[1][r][c] = (float)(sb * inv7);
        sm.V[2][r][c] = (float)(saa * inv7); sm.V[3][r][c] = (float)(sbb * inv7);
        sm.V[4][r][c] = (float)(sab * inv7);
    }
    __syncthreads();
    const float cn = (float)(49.0 / 48.0);
    const float C1 = (float)1.0e-4, C2 = (float)9.0e-4;
    double v[2] = {0.0, 0.0};   // sum S over the crop, sum (a-b)^2 over the image
#pragma unroll
    for (int j = 0; j < TH / 8; ++j)
#pragma unroll
        for (int i = 0; i < TW / 32; ++i) {
            const int r = wid + 8 * j, c = lane + 32 * i;
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) {
                const float a = sm.A[r + HALO][c + HALO], b = sm.B[r + HALO][c + HALO];
                const float df = __fsub_rn(a, b);
                v[1] += (double)__fmul_rn(df, df);
                if (gy >= HALO && gy < d.h - HALO && gx >= HALO && gx < d.w - HALO) {
                    double m[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        double t = 0.0;
#pragma unroll
                        for (int k = 0; k < 7; ++k) t += (double)sm.V[q][r][c + k];
                        m[q] = t * inv7;
                    }
                    const float ux = (float)m[0], uy = (float)m[1], uxx = (float)m[2],
                                uyy = (float)m[3], uxy = (float)m[4];
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

}  // namespace

void launch_box16_stats(const float* img, const Dims& d, double* acc2, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_box16_stats, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(Box16Smem));
        attr_set = true;
    }
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
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_ssim_psnr, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(SsimSmem));
        attr_set = true;
    }
    cudaMemsetAsync(acc2, 0, sizeof(double) * 2 * d.n_sel, stream);
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
    MDIMG_LAUNCH k_ssim_psnr<<<grid, NT, sizeof(SsimSmem), stream>>>(ia, ib, d, acc2);
    MDIMG_LAUNCH k_fullref_out<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, acc2, out);
    return check_launch("fullref");
}

}  // namespace mdimg
