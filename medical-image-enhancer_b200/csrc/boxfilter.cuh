// scipy.ndimage.uniform_filter on float32 tiles, the way scipy evaluates it: axis 0 then axis 1,
// each pass a running window sum in double, scaled by 1/size and ROUNDED TO FLOAT32 before the
// next pass.  Both passes slide a running sum (2 adds per output instead
// of `size`), operands are kept in shared memory as doubles that already hold float32-rounded
// values, so no conversion sits in the inner loops.
//
// Block = 256 threads (8 warps), output tile TH = 32 rows x TW = 64 columns.
//   vertical   : thread = (column, row segment)           -> V[r][c]
//   horizontal : lane = row, warp = 8-column segment      -> consume(r, c, mean...)
// With an odd pitch (in doubles) both walks are shared-memory bank-conflict free.
#pragma once
#include "common.cuh"

namespace mdimg {

template <int K>
struct BoxTile {
    static constexpr int TW = 64, TH = 32;
    static constexpr int XW = TW + K - 1;        // input columns (halo included)
    static constexpr int XH = TH + K - 1;        // input rows
    static constexpr int XP = XW | 1;            // odd pitch (doubles)
    static constexpr int NSEG = 256 / XW;        // row segments per column in the vertical pass
    static constexpr int RS = (TH + NSEG - 1) / NSEG;
};

// Single-step half-sample symmetric reflection (valid while the halo is smaller than n);
// falls back to the general form for tiny images.
__device__ __forceinline__ int refl_sym_fast(int i, int n) {
    if (i < 0) i = -1 - i;
    else if (i >= n) i = 2 * n - 1 - i;
    if (i < 0 || i >= n) i = refl_sym(i, n);
    return i;
}

__device__ __forceinline__ double round32(double v) { return (double)(float)v; }

// Cooperative tile load, one warp per tile row: the (<= 3) column indices of a lane are resolved
// once, the row index once per row, so the inner loop is one add + one load + the caller's store.
// REFLECT: 0 = half-sample symmetric (scipy 'reflect'), 1 = whole-sample mirror (np.pad 'reflect').
template <int XW, int XH, int HL, int REFLECT, typename Store>
__device__ __forceinline__ void load_tile(const float* __restrict__ src, int h, int w, int x0, int y0,
                                          Store&& store) {
    static_assert(XW <= 96, "at most three columns per lane");
    constexpr int NC = (XW + 31) / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int gx[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int x = x0 + lane + 32 * k - HL;
        gx[k] = REFLECT == 0 ? refl_sym_fast(x, w) : refl_mirror(x, w);
    }
    auto row_of = [&](int r) {
        const int y = y0 + r - HL;
        const int gy = REFLECT == 0 ? refl_sym_fast(y, h) : refl_mirror(y, h);
        return src + (size_t)gy * w;
    };
    // four rows per trip: their (up to 12) loads are issued before the first shared-memory store, so the
    // tile load is not one dependent load -> store round trip per row
    int r = wid;
    for (; r + 3 * nw < XH; r += 4 * nw) {
        float v[4][NC];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float* row = row_of(r + u * nw);
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (lane + 32 * k < XW) v[u][k] = row[gx[k]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (lane + 32 * k < XW) store(r + u * nw, lane + 32 * k, v[u][k]);
    }
    for (; r < XH; r += nw) {
        const float* row = row_of(r);
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (lane + 32 * k < XW) store(r, lane + 32 * k, row[gx[k]]);
    }
}

// Vertical pass for NQ quantities.  X[q]: [XH][XP] doubles, V[q]: [TH][XP] doubles.
template <int K, int NQ>
__device__ __forceinline__ void box_vertical(double* const (&X)[NQ], double* const (&V)[NQ], double inv) {
    typedef BoxTile<K> T;
    const int item = threadIdx.x;
    if (item < T::XW * T::NSEG) {
        const int sg = item / T::XW, c = item - sg * T::XW;
        const int r0 = sg * T::RS;
        const int r1 = min(r0 + T::RS, T::TH);
        double s[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            s[q] = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) s[q] += X[q][(r0 + k) * T::XP + c];
        }
        for (int r = r0; r < r1; ++r) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                V[q][r * T::XP + c] = round32(s[q] * inv);
                if (r + 1 < r1) s[q] += X[q][(r + K) * T::XP + c] - X[q][r * T::XP + c];
            }
        }
    }
}

// Horizontal pass: calls f(r, c, m[NQ]) for the 8 pixels (r = lane, c = 8*warp .. 8*warp+7) with
// m[q] = float32 box mean of quantity q.
template <int K, int NQ, typename F>
__device__ __forceinline__ void box_horizontal(double* const (&V)[NQ], double inv, F&& f) {
    typedef BoxTile<K> T;
    const int r = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int c0 = seg * 8;
    double s[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) s[q] += V[q][r * T::XP + c0 + k];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float m[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) m[q] = (float)(s[q] * inv);
        f(r, c0 + j, m);
        if (j < 7) {
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                s[q] += V[q][r * T::XP + c0 + j + K] - V[q][r * T::XP + c0 + j];
        }
    }
}

// ---- float32-storage variants -------------------------------------------------------------------
// Same arithmetic (double running sums, float32 rounding between the passes), but the tiles are
// kept as float32 in shared memory and converted when they enter / leave a window: a quarter of
// the shared memory of the double tiles, i.e. 4-6 CTAs per SM instead of 2 for these
// latency-bound kernels, for ~50 % more conversion instructions.

// Axis-0 pass of x and x*x (float32 product, as numpy's image**2) from a float tile.
template <int K>
__device__ __forceinline__ void box_vertical_xq(const float* X, float* VS, float* VQ, double inv) {
    typedef BoxTile<K> T;
    const int item = threadIdx.x;
    if (item < T::XW * T::NSEG) {
        const int sg = item / T::XW, c = item - sg * T::XW;
        const int r0 = sg * T::RS;
        const int r1 = min(r0 + T::RS, T::TH);
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float v = X[(r0 + k) * T::XP + c];
            s += (double)v;
            q += (double)__fmul_rn(v, v);
        }
        for (int r = r0; r < r1; ++r) {
            VS[r * T::XP + c] = (float)(s * inv);
            VQ[r * T::XP + c] = (float)(q * inv);
            if (r + 1 < r1) {
                const float vn = X[(r + K) * T::XP + c], vo = X[r * T::XP + c];
                s += (double)vn - (double)vo;
                q += (double)__fmul_rn(vn, vn) - (double)__fmul_rn(vo, vo);
            }
        }
    }
}

// Axis-1 pass over float32 V tiles: f(r, c, m[NQ]) for the 8 pixels of (lane = row, warp = segment).
template <int K, int NQ, typename F>
__device__ __forceinline__ void box_horizontal_f(float* const (&V)[NQ], double inv, F&& f) {
    typedef BoxTile<K> T;
    const int r = threadIdx.x & 31, seg = threadIdx.x >> 5;
    const int c0 = seg * 8;
    double s[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        s[q] = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) s[q] += (double)V[q][r * T::XP + c0 + k];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float m[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) m[q] = (float)(s[q] * inv);
        f(r, c0 + j, m);
        if (j < 7) {
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                s[q] += (double)V[q][r * T::XP + c0 + j + K] - (double)V[q][r * T::XP + c0 + j];
        }
    }
}

}  // namespace mdimg
