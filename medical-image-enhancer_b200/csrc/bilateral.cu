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
#include "packed.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int TW = 64, TH = 32;
constexpr int MAXD = 9;

struct SpatialW { float w[MAXD * MAXD]; };

// Two adjacent pixels per thread on the packed float32x2 pipe (FADD2 / FMUL2 / FFMA2): per tap and
// pixel pair 6 packed instructions + 2 MUFU.EX2, and the d + 1 neighbours of a row are loaded once
// for both pixels (64-bit shared loads).  The per-pixel operations and their order are unchanged.
template <int R>
__global__ void __launch_bounds__(NT)
k_bilateral(const float* __restrict__ in, float* __restrict__ out, Dims d, const SpatialW sw,
            float neg_k) {
    constexpr int D = 2 * R + 1;
    constexpr int XW = TW + 2 * R, XH = TH + 2 * R, XP = XW;      // even pitch: pixel pairs stay 8-byte aligned
    __shared__ __align__(8) float X[XH][XP];
    const int s = slice_of(d.sel, blockIdx.y);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = in + (size_t)s * d.h * d.w;
    float* dst = out + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    load_tile<XW, XH, R, 1>(src, d.h, d.w, x0, y0, [&](int r, int c, float v) { X[r][c] = v; });
    __syncthreads();
    const int c = 2 * lane;                                       // this thread's pair: columns c, c + 1
    const int gx = x0 + c;
    const float2 k2 = bc2(neg_k);
#pragma unroll
    for (int j2 = 0; j2 < TH / 8; ++j2) {
        const int r = wid + 8 * j2;
        const int gy = y0 + r;
        if (gy < d.h && gx < d.w) {
            const float2 xc = *reinterpret_cast<const float2*>(&X[r + R][c + R - (R & 1)]);
            const float2 ctr = (R & 1) ? make_float2(xc.y, X[r + R][c + R + 1]) : xc;
            float2 res = make_float2(0.0f, 0.0f), wsum = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int dy = 0; dy < D; ++dy) {
                float v[D + 1];
#pragma unroll
                for (int q = 0; q < (D + 1) / 2; ++q) {
                    const float2 t = *reinterpret_cast<const float2*>(&X[r + dy][c + 2 * q]);
                    v[2 * q] = t.x;
                    v[2 * q + 1] = t.y;
                }
#pragma unroll
                for (int dx = 0; dx < D; ++dx) {
                    const float2 nb = make_float2(v[dx], v[dx + 1]);
                    const float2 diff = sub2(ctr, nb);
                    // exp(-diff^2 / (2 sc^2)) = 2^(diff^2 * neg_k), neg_k = -log2(e) / (2 sc^2)
                    const float2 e = mul2(mul2(diff, diff), k2);
                    float2 iw;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(iw.x) : "f"(e.x));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(iw.y) : "f"(e.y));
                    const float2 w = mul2(bc2(sw.w[dy * D + dx]), iw);
                    res = fma2(w, nb, res);
                    wsum = add2(wsum, w);
                }
            }
            const float o0 = __fdiv_rn(res.x, __fadd_rn(wsum.x, 1e-10f));
            const float o1 = __fdiv_rn(res.y, __fadd_rn(wsum.y, 1e-10f));
            float* o = dst + (size_t)gy * d.w + gx;
            if (gx + 1 < d.w && ((d.w & 1) == 0)) {
                *reinterpret_cast<float2*>(o) = make_float2(o0, o1);
            } else {
                o[0] = o0;
                if (gx + 1 < d.w) o[1] = o1;
            }
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
