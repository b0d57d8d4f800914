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

// scipy's symmetric correlate1d on a window of 2R + 1 doubles (centre at win[R]).
template <int R>
__device__ __forceinline__ double gauss_window(const double (&win)[2 * R + 1], const GaussW& gw) {
    double acc = __dmul_rn(win[R], gw.w[0]);
#pragma unroll
    for (int j = R; j >= 1; --j)
        acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(win[R - j], win[R + j]), gw.w[j]));
    return acc;
}

// Both passes slide a register window of 2R + 1 doubles along the filter axis, so every tile element
// is converted float32 -> float64 once per pass instead of once per tap (the conversion pipe, not
// the float64 pipe, was this kernel's limiter: 2 (2R + 1) + 2 conversions per pixel).
//   axis 0: thread = (column, row segment), walks down its segment;
//   axis 1: lane = row, warp = 8-column segment, walks right; the result is parked in the input
//           tile (each thread overwrites only the pixel it alone reads) and written out coalesced.
template <int R>
__global__ void __launch_bounds__(NT)
k_unsharp(const float* __restrict__ in, float* __restrict__ out, Dims d, float amount,
          const uint2* __restrict__ mm, const GaussW gw) {
    constexpr int XW = TW + 2 * R, XH = TH + 2 * R, XP = XW | 1, W = 2 * R + 1;
    constexpr int NSEG = NT / XW, RS = (TH + NSEG - 1) / NSEG;
    static_assert(NSEG >= 1 && TH == 32 && TW == 8 * (NT / 32), "pass layout");
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
    if (tid < XW * NSEG) {
        const int sg = tid / XW, c = tid - sg * XW;
        const int r0 = sg * RS;
        double win[W];
#pragma unroll
        for (int k = 0; k < W - 1; ++k) win[k + 1] = (double)X[min(r0 + k, XH - 1)][c];
#pragma unroll
        for (int t = 0; t < RS; ++t) {
            const int r = r0 + t;
            if (r < TH) {
#pragma unroll
                for (int k = 0; k < W - 1; ++k) win[k] = win[k + 1];
                win[W - 1] = (double)X[r + 2 * R][c];
                V[r][c] = (float)gauss_window<R>(win, gw);
            }
        }
    }
    __syncthreads();
    {
        const int r = lane, c0 = wid * 8;
        double win[W];
#pragma unroll
        for (int k = 0; k < W - 1; ++k) win[k + 1] = (double)V[r][c0 + k];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int c = c0 + t;
#pragma unroll
            for (int k = 0; k < W - 1; ++k) win[k] = win[k + 1];
            win[W - 1] = (double)V[r][c + 2 * R];
            const float blurred = (float)gauss_window<R>(win, gw);
            const float x = X[r + R][c + R];
            float res = __fadd_rn(x, __fmul_rn(__fsub_rn(x, blurred), amount));
            X[r + R][c + R] = fminf(fmaxf(res, lo), 1.0f);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j2 = 0; j2 < TH / 8; ++j2)
#pragma unroll
        for (int i2 = 0; i2 < TW / 32; ++i2) {
            const int r = wid + 8 * j2, c = lane + 32 * i2;
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) dst[(size_t)gy * d.w + gx] = X[r + R][c + R];
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
