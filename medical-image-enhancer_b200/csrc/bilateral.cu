// Bilateral filter: the reference's own numpy implementation, pipeline/enhancement.py:102-143.
//
//   padded = np.pad(image, r, 'reflect')                      whole-sample mirror
//   for dy, dx (dy-major):  diff = image - shifted            float32
//       iw = exp(-(diff**2) / (2*sc**2))                      float32 (python-float divisor)
//       w  = spatial[dy, dx] * iw                             float64 (np.float64 scalar * float32)
//       result += w * shifted ; weight_sum += w               float64 add, stored back as float32
//   out = result / (weight_sum + 1e-10)                       float32
//
// One shared-memory tile with a (d//2)-pixel halo per CTA; the slice is read once and written
// once.  d*d exponentials per pixel make this kernel MUFU-bound rather than HBM-bound for d >= 5
// (stated in DESIGN.md).
//
// Precision: the reference's float32 exp is numpy's SIMD routine (not correctly rounded), so the
// weights are not reproducible to the bit on any other implementation; the accumulation is
// therefore done in float32 FMAs (tap order kept) instead of emulating numpy's float64-add /
// float32-store per tap.  Measured difference to the oracle: a few float32 ulps (tests allow 16).
#include "enhance.cuh"
#include "boxfilter.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int TW = 64, TH = 32;
constexpr int MAXD = 9;

struct SpatialW { float w[MAXD * MAXD]; };

template <int R>
__global__ void __launch_bounds__(NT)
k_bilateral(const float* __restrict__ in, float* __restrict__ out, Dims d, const SpatialW sw,
            float neg_k) {
    constexpr int D = 2 * R + 1;
    constexpr int XW = TW + 2 * R, XH = TH + 2 * R, XP = XW + 1;
    __shared__ float X[XH][XP];
    const int s = slice_of(d.sel, blockIdx.y);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = in + (size_t)s * d.h * d.w;
    float* dst = out + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    load_tile<XW, XH, R, 1>(src, d.h, d.w, x0, y0, [&](int r, int c, float v) { X[r][c] = v; });
    __syncthreads();
#pragma unroll
    for (int j2 = 0; j2 < TH / 8; ++j2)
#pragma unroll
        for (int i2 = 0; i2 < TW / 32; ++i2) {
            const int r = wid + 8 * j2, c = lane + 32 * i2;
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) {
                const float xc = X[r + R][c + R];
                float res = 0.0f, wsum = 0.0f;
#pragma unroll
                for (int dy = 0; dy < D; ++dy)
#pragma unroll
                    for (int dx = 0; dx < D; ++dx) {
                        const float nb = X[r + dy][c + dx];
                        const float diff = xc - nb;
                        // exp(-diff^2 / (2 sc^2)) = 2^(diff^2 * neg_k), neg_k = -log2(e) / (2 sc^2)
                        const float iw = exp2f(diff * diff * neg_k);
                        const float w = sw.w[dy * D + dx] * iw;
                        res = fmaf(w, nb, res);
                        wsum += w;
                    }
                dst[(size_t)gy * d.w + gx] = __fdiv_rn(res, __fadd_rn(wsum, 1e-10f));
            }
        }
}

template <int R>
void launch(const float* in, float* out, const Dims& d, const SpatialW& sw, float neg_k, cudaStream_t st) {
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
    MDIMG_LAUNCH k_bilateral<R><<<grid, NT, 0, st>>>(in, out, d, sw, neg_k);
}

}  // namespace

int bilateral_run(const float* in, float* out, const Dims& d, int dd, const double* spatial,
                  double sigma_color, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (dd < 1 || dd > MAXD || (dd & 1) == 0)
        return set_error(MDIMG_ERR_INVALID, "bilateral: diameter %d must be odd and in [1, 9]", dd);
    if (in == out) return set_error(MDIMG_ERR_INVALID, "bilateral: in-place operation is not supported");
    SpatialW sw;
    for (int i = 0; i < MAXD * MAXD; ++i) sw.w[i] = i < dd * dd ? (float)spatial[i] : 0.0f;
    const float neg_k = (float)(-1.4426950408889634 / (2.0 * sigma_color * sigma_color));
    switch (dd / 2) {
        case 0: launch<0>(in, out, d, sw, neg_k, stream); break;
        case 1: launch<1>(in, out, d, sw, neg_k, stream); break;
        case 2: launch<2>(in, out, d, sw, neg_k, stream); break;
        case 3: launch<3>(in, out, d, sw, neg_k, stream); break;
        case 4: launch<4>(in, out, d, sw, neg_k, stream); break;
    }
    return check_launch("bilateral");
}

}  // namespace mdimg
