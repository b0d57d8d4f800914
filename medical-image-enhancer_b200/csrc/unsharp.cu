// Unsharp mask: skimage.filters.unsharp_mask(image, radius, amount) as called at
// pipeline/enhancement.py:202,290,338 — result = clip(x + (x - G_sigma(x)) * amount, lo, 1).
//
// G_sigma is scipy.ndimage.gaussian_filter(mode='reflect', truncate=4): two correlate1d passes
// (axis 0 then axis 1), each evaluated in double in scipy's symmetric-kernel order
//   acc = x[0]*w[0];  for j = R..1: acc += (x[-j] + x[+j]) * w[j]
// and rounded to float32 when stored.  Both passes run from one shared-memory tile, so the slice
// is read once and written once.
#include "enhance.cuh"
#include "boxfilter.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int TW = 64, TH = 32;
constexpr int MAXR = 12;

struct GaussW { double w[MAXR + 1]; };   // passed by value: re-entrant across streams

template <int R>
__global__ void __launch_bounds__(NT)
k_unsharp(const float* __restrict__ in, float* __restrict__ out, Dims d, float amount,
          const uint2* __restrict__ mm, const GaussW gw) {
    constexpr int XW = TW + 2 * R, XH = TH + 2 * R, XP = XW + 1;
    __shared__ float X[XH][XP];
    __shared__ float V[TH][XP];
    const int s = slice_of(d.sel, blockIdx.y);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = in + (size_t)s * d.h * d.w;
    float* dst = out + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float lo = (mm && key2f(mm[s].x) < 0.0f) ? -1.0f : 0.0f;   // vrange when any pixel is negative

    load_tile<XW, XH, R, 0>(src, d.h, d.w, x0, y0, [&](int r, int c, float v) { X[r][c] = v; });
    __syncthreads();
    for (int i = tid; i < TH * XW; i += NT) {
        int r = i / XW, c = i - r * XW;
        double acc = __dmul_rn((double)X[r + R][c], gw.w[0]);
#pragma unroll
        for (int j = R; j >= 1; --j)
            acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)X[r + R - j][c], (double)X[r + R + j][c]), gw.w[j]));
        V[r][c] = (float)acc;
    }
    __syncthreads();
#pragma unroll
    for (int j2 = 0; j2 < TH / 8; ++j2)
#pragma unroll
        for (int i2 = 0; i2 < TW / 32; ++i2) {
            const int r = wid + 8 * j2, c = lane + 32 * i2;
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) {
                double acc = __dmul_rn((double)V[r][c + R], gw.w[0]);
#pragma unroll
                for (int j = R; j >= 1; --j)
                    acc = __dadd_rn(acc, __dmul_rn(__dadd_rn((double)V[r][c + R - j], (double)V[r][c + R + j]), gw.w[j]));
                const float blurred = (float)acc;
                const float x = X[r + R][c + R];
                float res = __fadd_rn(x, __fmul_rn(__fsub_rn(x, blurred), amount));
                res = fminf(fmaxf(res, lo), 1.0f);
                dst[(size_t)gy * d.w + gx] = res;
            }
        }
}

template <int R>
void launch(const float* in, float* out, const Dims& d, float amount, const uint2* mm,
            const GaussW& gw, cudaStream_t st) {
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
    MDIMG_LAUNCH k_unsharp<R><<<grid, NT, 0, st>>>(in, out, d, amount, mm, gw);
}

}  // namespace

int unsharp_run(const float* in, float* out, const Dims& d, const double* weights, int radius,
                float amount, const uint2* mm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (radius < 1 || radius > MAXR)
        return set_error(MDIMG_ERR_INVALID, "unsharp: gaussian radius %d outside [1, %d]", radius, MAXR);
    if (in == out) return set_error(MDIMG_ERR_INVALID, "unsharp: in-place operation is not supported");
    GaussW gw;
    for (int i = 0; i <= MAXR; ++i) gw.w[i] = i <= radius ? weights[i] : 0.0;
    switch (radius) {
#define CASE(R) case R: launch<R>(in, out, d, amount, mm, gw, stream); break;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11) CASE(12)
#undef CASE
    }
    return check_launch("unsharp");
}

}  // namespace mdimg
