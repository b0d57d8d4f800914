// CLAHE: skimage.exposure.equalize_adapthist(image, clip_limit, kernel_size) as called at
// pipeline/enhancement.py:183,277,332 (skimage/exposure/_adapthist.py, NR_OF_GRAY = 16384,
// nbins = 256).  Integer / indexing work, reproduced bit-exactly:
//
//   u   = uint16(clip(rint(float32(x) * 65535), 0, 65535))                 img_as_uint
//   q   = uint16(rint_half_even((u - umin) / (umax - umin) * 16383))       float64 rescale
//   pad = np.pad(q, (k//2, (k - s%k)%k + ceil(k/2)), 'reflect')            whole-sample mirror
//   b   = q // 65                                                           256 gray bins (0..252)
//   per k x k contextual region: bincount -> clip_histogram (sequential redistribute loop)
//        -> map = int(min(cumsum * (16383 / k^2), 16383))
//   per pixel: float32 accumulation, in corner order (0,0),(0,1),(1,0),(1,1), of
//        float32(float64(map[b]) * (wx * wy)); truncate to uint16
//   out = (v - vmin) / (vmax - vmin) in float32                            final rescale_intensity
//
// Kernels: k_clahe_hist (one warp per contextual region: quantise, histogram in shared memory,
// clip/redistribute, CDF -> uint16 LUT; also stores the 8-bit bin image), k_clahe_blend (bilinear
// LUT blend -> uint16 + per-slice min/max), k_clahe_lut + k_clahe_final (float32 stretch, optionally
// followed by adjust_gamma, through a 16384-level table).
#include "enhance.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int WARPS = NT / 32;
constexpr int NBINS = 256;

struct ClaheGeom {
    int k;          // kernel_size (same on both axes)
    int nty, ntx;   // contextual regions per axis
    int clim;
    unsigned ntx_magic;   // floor(2^32 / ntx) + 1: tile / ntx == umulhi(tile, magic) while tile * ntx < 2^32; 0: divide
    double scale;   // 16383 / k^2
};

struct SliceRange {   // per slice: umin, umax of the uint16 image; vmin/vmax of the blended image
    unsigned umin, umax;
    unsigned vmin, vmax;
};

__device__ __forceinline__ unsigned to_u16(float x) {
    float t = rintf(__fmul_rn(x, 65535.0f));
    t = fminf(fmaxf(t, 0.0f), 65535.0f);
    return (unsigned)t;
}

// to_u16 without FRND / F2I: clamp, then one add with 2^23 leaves rint(t) (ties to even) in the mantissa
__device__ __forceinline__ unsigned to_u16_fast(float x) {
    const float t = fminf(fmaxf(__fmul_rn(x, 65535.0f), 0.0f), 65535.0f);
    return __float_as_uint(__fadd_rn(t, 8388608.0f)) & 0xffffu;
}

__global__ void k_clahe_prep(Dims d, const uint2* __restrict__ mm, SliceRange* __restrict__ rng,
                             int* __restrict__ status) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    float mn = key2f(mm[s].x), mx = key2f(mm[s].y);
    status[s] = (mn < -1.0f || mx > 1.0f) ? 1 : 0;
    SliceRange r;
    r.umin = to_u16(mn);          // rint(x*65535) is monotone, so min/max commute with it
    r.umax = to_u16(mx);
    r.vmin = 0xFFFFFFFFu;
    r.vmax = 0u;
    rng[si] = r;
}

// rescale of the uint16 image to 14 bits (float64, round half to even), skimage _clahe step (ii)
__device__ __forceinline__ unsigned quantise_u16(unsigned u, unsigned umin, unsigned umax) {
    if (umin == umax) return u > 16383u ? 16383u : u;
    double t = __ddiv_rn((double)(int)(u - umin), (double)(int)(umax - umin));
    t = __dadd_rn(__dmul_rn(t, 16383.0), 0.0);
    return (unsigned)rint(t);
}

// The float64 quantisation depends only on the 16-bit level: one table entry per level and slice
// (gray bin = quantised level / 65) instead of a float64 division per pixel.
constexpr int NU16 = 65536;

__global__ void __launch_bounds__(NT)
k_clahe_binlut(Dims d, const SliceRange* __restrict__ rng, const int* __restrict__ status,
               uint8_t* __restrict__ lut) {
    const int si = blockIdx.y;
    if (status[slice_of(d.sel, si)]) return;
    const unsigned u = blockIdx.x * NT + threadIdx.x;
    const unsigned umin = rng[si].umin, umax = rng[si].umax;
    unsigned b = 0;
    if (u >= umin && u <= umax) b = quantise_u16(u, umin, umax) / 65u;
    lut[(size_t)si * NU16 + u] = (uint8_t)b;
}

// np.pad 'reflect' (whole-sample mirror) for an index that overshoots by less than the extent;
// falls back to the general form otherwise (tiny images).
__device__ __forceinline__ int mirror_fast(int i, int n) {
    int r = i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i);
    if (r < 0 || r >= n) r = refl_mirror(i, n);
    return r;
}

// Pixel phase of an interior K x K region (K = 8, 16, 32) of a 4-pixel-aligned image; p / bq point at the
// region's first pixel / gray bin.  All 32 lanes call.
template <int K>
__device__ __forceinline__ void hist_fast(const float* __restrict__ p, uint8_t* __restrict__ bq,
                                          const uint8_t* __restrict__ lut, int* h, int w, int lane) {
    constexpr int LPR = K / 4, RPT = 32 / LPR;          // lanes per region row, region rows per trip
    constexpr int TRIPS = (K + RPT - 1) / RPT;          // K = 8: one trip on half of the lanes
    constexpr int UB = TRIPS < 2 ? 1 : 2;               // trips of loads in flight
    const int lr = lane / LPR, lc = (lane % LPR) * 4;
    const bool act = lr < K;
    const size_t o = (size_t)lr * w + lc, tstride = (size_t)RPT * w;
    const unsigned m_act = K == 8 ? 0x0000ffffu : 0xffffffffu;
#pragma unroll 1
    for (int t0 = 0; t0 < TRIPS; t0 += UB) {
        float4 v[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u)
            v[u] = act ? *reinterpret_cast<const float4*>(p + o + (size_t)(t0 + u) * tstride) : make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned b[UB][4];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            b[u][0] = lut[to_u16_fast(v[u].x)]; b[u][1] = lut[to_u16_fast(v[u].y)];
            b[u][2] = lut[to_u16_fast(v[u].z)]; b[u][3] = lut[to_u16_fast(v[u].w)];
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const unsigned b0 = b[u][0], b1 = b[u][1], b2 = b[u][2], b3 = b[u][3];
            const unsigned q = b0 | (b1 << 8) | (b2 << 16) | (b3 << 24);
            if (act) *reinterpret_cast<unsigned*>(bq + o + (size_t)(t0 + u) * tstride) = q;
            const unsigned lead = __shfl_sync(0xffffffffu, q, 0);
            const bool same = q == lead && q == b0 * 0x01010101u;
            if (__all_sync(0xffffffffu, !act || same)) {
                if (lane == 0) atomicAdd(&h[b0], 4 * __popc(m_act));
            } else if (act) {
                // runs of equal bins inside the lane's four pixels: one atomic per run
                const int e1 = b1 == b0, e2 = b2 == b1, e3 = b3 == b2;
                const int c1 = 1 + e1, c2 = 1 + e2 * c1, c3 = 1 + e3 * c2;
                if (!e1) atomicAdd(&h[b0], 1);
                if (!e2) atomicAdd(&h[b1], c1);
                if (!e3) atomicAdd(&h[b2], c2);
                atomicAdd(&h[b3], c3);
            }
        }
    }
}

// One warp per contextual region.
__global__ void __launch_bounds__(NT, 4)
k_clahe_hist(const float* __restrict__ in, Dims d, ClaheGeom g, const uint8_t* __restrict__ binlut,
             const int* __restrict__ status, uint8_t* __restrict__ bins, uint16_t* __restrict__ maps) {
    __shared__ __align__(16) int hist[WARPS][NBINS];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (status[s]) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tile = blockIdx.x * WARPS + wid;
    if (tile >= g.nty * g.ntx) return;
    const int ty = g.ntx_magic ? (int)__umulhi((unsigned)tile, g.ntx_magic) : tile / g.ntx;
    const int tx = tile - ty * g.ntx;
    const int k = g.k;
    int* h = hist[wid];
    *reinterpret_cast<int4*>(h + 8 * lane) = make_int4(0, 0, 0, 0);
    *reinterpret_cast<int4*>(h + 8 * lane + 4) = make_int4(0, 0, 0, 0);
    __syncwarp();
    const float* src = in + (size_t)s * d.h * d.w;
    uint8_t* bdst = bins + (size_t)si * d.h * d.w;
    const uint8_t* lut = binlut + (size_t)si * NU16;
    const int kk = k * k;
    const bool interior = (tx + 1) * k <= d.w && (ty + 1) * k <= d.h;
    if ((k == 8 || k == 16 || k == 32) && interior && (d.w & 3) == 0 && ((uintptr_t)src & 15) == 0) {
        // Interior region of a 4-pixel-aligned image: a lane owns FOUR adjacent pixels of a region row (one
        // 128-bit load, one 32-bit store of the four gray bins), 32 / (k / 4) rows per trip, two trips of
        // loads in flight.  rint + clamp + integer conversion of img_as_uint is one FP32 add with 2^23
        // (round-to-nearest-even into the mantissa) after the clamp: clamp and rint commute because the
        // clamp bounds are integers.  Histogram: a warp whose 128 pixels share one bin (CT air) issues ONE
        // shared-memory atomic; otherwise a lane adds its four pixels as runs of equal bins.
        const size_t o0 = (size_t)(ty * k) * d.w + tx * k;
        if (k == 16) hist_fast<16>(src + o0, bdst + o0, lut, h, d.w, lane);
        else if (k == 32) hist_fast<32>(src + o0, bdst + o0, lut, h, d.w, lane);
        else hist_fast<8>(src + o0, bdst + o0, lut, h, d.w, lane);
    } else if (32 % k == 0 && kk >= 32) {
        const int ry = lane / k, rx = lane - ry * k, dry = 32 / k;
        // k = 8, 16, 32: a trip covers whole rows, so a lane keeps its column (index, mirror, bounds
        // test) for the whole region, and the trips are independent of each other: eight pixel loads,
        // then eight table gathers are in flight at a time instead of one dependent load -> gather ->
        // atomic chain per trip
        const int ox = tx * k + rx, gx = mirror_fast(ox, d.w);
        const bool in_w = ox < d.w;
        const int trips = kk >> 5;
        for (int t0 = 0; t0 < trips; t0 += 8) {
            float v[8];
            int b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < trips) {
                    const int gy = mirror_fast(ty * k + ry + (t0 + u) * dry, d.h);
                    v[u] = src[(size_t)gy * d.w + gx];
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < trips) b[u] = lut[to_u16(v[u])];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (t0 + u < trips) {
                    const int oy = ty * k + ry + (t0 + u) * dry;
                    atomicAdd(&h[b[u]], 1);
                    if (oy < d.h && in_w) bdst[(size_t)oy * d.w + ox] = (uint8_t)b[u];
                }
        }
    } else {
    // lane -> (row, column) inside the region, advanced by 32 elements per trip without divisions
    int ry = lane / k, rx = lane - ry * k;
    const int dry = 32 / k, drx = 32 - dry * k;
    for (int i = lane; i < kk; i += 32) {
        const int oy = ty * k + ry, ox = tx * k + rx;        // padded index minus pad_start
        const int gy = mirror_fast(oy, d.h), gx = mirror_fast(ox, d.w);
        const int b = lut[to_u16(src[(size_t)gy * d.w + gx])];
        atomicAdd(&h[b], 1);
        if (oy < d.h && ox < d.w) bdst[(size_t)oy * d.w + ox] = (uint8_t)b;
        rx += drx; ry += dry;
        if (rx >= k) { rx -= k; ry += 1; }
    }
    }
    __syncwarp();

    // ---- clip_histogram: a lane holds the eight consecutive bins j = 8 * lane + m in registers ----
    int hv[8];
    {
        const int4 q0 = *reinterpret_cast<const int4*>(h + 8 * lane), q1 = *reinterpret_cast<const int4*>(h + 8 * lane + 4);
        hv[0] = q0.x; hv[1] = q0.y; hv[2] = q0.z; hv[3] = q0.w;
        hv[4] = q1.x; hv[5] = q1.y; hv[6] = q1.z; hv[7] = q1.w;
    }
    const int clim = g.clim;
    int part = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        if (hv[m] > clim) { part += hv[m] - clim; hv[m] = clim; }
    }
    int n_excess = __reduce_add_sync(0xffffffffu, part);
    if (n_excess > 0) {            // warp-uniform; a region without a clipped bin skips all of it (its steps add nothing)
        const int bin_incr = n_excess / NBINS;
        const int upper = clim - bin_incr;
        part = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (hv[m] < upper) { hv[m] += bin_incr; part += 1; }
        }
        n_excess -= __reduce_add_sync(0xffffffffu, part) * bin_incr;
        part = 0;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (hv[m] >= upper && hv[m] < clim) { part += hv[m] - clim; hv[m] = clim; }
        }
        n_excess += __reduce_add_sync(0xffffffffu, part);
    }

    while (n_excess > 0) {
        const int prev = n_excess;
        for (int index = 0; index < NBINS; ++index) {
            int under = 0;
#pragma unroll
            for (int m = 0; m < 8; ++m) under += (hv[m] < clim);
            under = __reduce_add_sync(0xffffffffu, under);
            // step = max(1, under // n_excess): no division in the usual case under heavy clipping
            const int step = under > n_excess ? under / n_excess : 1;
            int cnt = 0;
            if (step == 1) {                   // warp-uniform
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int j = 8 * lane + m;
                    if (j >= index && hv[m] < clim) { hv[m] += 1; cnt += 1; }
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int j = 8 * lane + m;
                    if (j >= index && ((j - index) % step) == 0 && hv[m] < clim) { hv[m] += 1; cnt += 1; }
                }
            }
            n_excess -= __reduce_add_sync(0xffffffffu, cnt);
            if (n_excess <= 0) break;
        }
        if (prev == n_excess) break;
    }

    // ---- map_histogram: cumulative sum in bin order (eight local sums + one warp scan), scaled, truncated ----
    int cum[8];
    cum[0] = hv[0];
#pragma unroll
    for (int m = 1; m < 8; ++m) cum[m] = cum[m - 1] + hv[m];
    int inc = cum[7];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int excl = inc - cum[7];
    unsigned lv[8];
    const int kk2 = k * k;
    if ((kk2 & (kk2 - 1)) == 0) {
        // k^2 a power of two: 16383 / k^2 is exact in float64 and so is its product with the count, hence
        // int(min(cum * scale, 16383)) == min((cum * 16383) >> log2(k^2), 16383) in integers
        const int sh = 31 - __clz(kk2);
#pragma unroll
        for (int m = 0; m < 8; ++m) lv[m] = min((unsigned)((excl + cum[m]) * 16383) >> sh, 16383u);
    } else {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            double v = __dadd_rn(__dmul_rn((double)(excl + cum[m]), g.scale), 0.0);
            if (v > 16383.0) v = 16383.0;
            lv[m] = (unsigned)(int)v;
        }
    }
    uint16_t* mdst = maps + ((size_t)si * g.nty * g.ntx + tile) * NBINS;
    *reinterpret_cast<uint4*>(mdst + 8 * lane) =
        make_uint4(lv[0] | (lv[1] << 16), lv[2] | (lv[3] << 16), lv[4] | (lv[5] << 16), lv[6] | (lv[7] << 16));
}

// Blend tile: 128 columns x 32 rows per block of 256 threads.  A thread owns four adjacent columns
// and walks four rows (warp w: rows w, w + 8, ...): everything that depends on the column only --
// block index, position inside the block, the two interpolation weights, the offsets of the two
// neighbouring regions' LUTs -- is computed once per thread, what depends on the row only once per
// warp and row, and the per-block setup (k float64 divisions for np.arange(k) / k) is shared by
// 4096 pixels.  Per pixel remain the four LUT gathers and the float64 products in skimage's order.
constexpr int BPX = 4, BW = 32 * BPX, BROWS = 4, BH = WARPS * BROWS;

// uint16 LUT entry -> float64 without an integer conversion: (2^52 + m) - 2^52, exact for m < 2^32
__device__ __forceinline__ double u2d(unsigned m) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)m), 4503599627370496.0);
}

__global__ void __launch_bounds__(NT)
k_clahe_blend(Dims d, ClaheGeom g, SliceRange* __restrict__ rng, const int* __restrict__ status,
              const uint8_t* __restrict__ bins, const uint16_t* __restrict__ maps,
              uint16_t* __restrict__ vout) {
    __shared__ double coef[48], icoef[48];
    __shared__ unsigned smin[WARPS], smax[WARPS];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (status[s]) return;
    const int k = g.k;
    if (threadIdx.x < k) {
        double c = __ddiv_rn((double)threadIdx.x, (double)k);   // np.arange(k) / k
        coef[threadIdx.x] = c;
        icoef[threadIdx.x] = __dsub_rn(1.0, c);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tiles_x = (d.w + BW - 1) / BW;
    const int bx = blockIdx.x % tiles_x, by = blockIdx.x / tiles_x;
    const int x0 = bx * BW + lane * BPX;
    // ---- per column: map_array is the LUT grid edge-padded by one: entry j -> region clamp(j - 1) ----
    double wx0[BPX], wx1[BPX];
    int o0[BPX], o1[BPX];
#pragma unroll
    for (int c = 0; c < BPX; ++c) {
        const int px = x0 + c + k / 2;
        const int bxk = px / k, ix = px - bxk * k;
        o0[c] = min(max(bxk - 1, 0), g.ntx - 1) * NBINS;
        o1[c] = min(bxk, g.ntx - 1) * NBINS;
        wx0[c] = icoef[ix];
        wx1[c] = coef[ix];
    }
    const size_t img = (size_t)si * d.h * d.w;
    const uint16_t* mbase = maps + (size_t)si * g.nty * g.ntx * NBINS;
    const bool vec = (d.w % BPX) == 0 && x0 + BPX <= d.w;       // 4 bins in one 32-bit load, 4 levels in one 64-bit store
    unsigned vmin = 0xFFFFFFFFu, vmax = 0u;
#pragma unroll
    for (int j = 0; j < BROWS; ++j) {
        const int y = by * BH + wid + WARPS * j;
        if (y >= d.h || x0 >= d.w) continue;
        // ---- per row (warp-uniform) ----
        const int py = y + k / 2;
        const int byk = py / k, iy = py - byk * k;
        const uint16_t* m0 = mbase + (size_t)min(max(byk - 1, 0), g.nty - 1) * g.ntx * NBINS;
        const uint16_t* m1 = mbase + (size_t)min(byk, g.nty - 1) * g.ntx * NBINS;
        const double wy0 = icoef[iy], wy1 = coef[iy];
        const size_t o = img + (size_t)y * d.w + x0;
        unsigned b4;
        if (vec) {
            b4 = *reinterpret_cast<const unsigned*>(bins + o);
        } else {
            b4 = 0;
#pragma unroll
            for (int c = 0; c < BPX; ++c)
                if (x0 + c < d.w) b4 |= (unsigned)bins[o + c] << (8 * c);
        }
        unsigned v4[BPX];
#pragma unroll
        for (int c = 0; c < BPX; ++c) {
            const unsigned b = (b4 >> (8 * c)) & 0xffu;
            float acc = 0.0f;
            acc = __fadd_rn(acc, (float)__dmul_rn(u2d(m0[o0[c] + b]), __dmul_rn(wx0[c], wy0)));
            acc = __fadd_rn(acc, (float)__dmul_rn(u2d(m0[o1[c] + b]), __dmul_rn(wx1[c], wy0)));
            acc = __fadd_rn(acc, (float)__dmul_rn(u2d(m1[o0[c] + b]), __dmul_rn(wx0[c], wy1)));
            acc = __fadd_rn(acc, (float)__dmul_rn(u2d(m1[o1[c] + b]), __dmul_rn(wx1[c], wy1)));
            v4[c] = (unsigned)acc;             // astype(uint16): truncation
            if (x0 + c < d.w) { vmin = min(vmin, v4[c]); vmax = max(vmax, v4[c]); }
        }
        if (vec) {
            *reinterpret_cast<uint2*>(vout + o) = make_uint2(v4[0] | (v4[1] << 16), v4[2] | (v4[3] << 16));
        } else {
#pragma unroll
            for (int c = 0; c < BPX; ++c)
                if (x0 + c < d.w) vout[o + c] = (uint16_t)v4[c];
        }
    }
    vmin = __reduce_min_sync(0xffffffffu, vmin);
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if (lane == 0) { smin[wid] = vmin; smax[wid] = vmax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < WARPS; ++w) { vmin = min(vmin, smin[w]); vmax = max(vmax, smax[w]); }
        if (vmin <= vmax) {
            atomicMin(&rng[si].vmin, vmin);
            atomicMax(&rng[si].vmax, vmax);
        }
    }
}

// The blended image holds at most 16384 distinct levels, so the final float32 stretch
// (v - vmin) / (vmax - vmin) -- and an adjust_gamma that directly follows CLAHE in the plan
// (pipeline/enhancement.py:277-286) -- is a per-slice table: one IEEE division (and one correctly
// rounded power) per LEVEL instead of per pixel, then a pure gather.
constexpr int NLEVELS = 16384;

__global__ void __launch_bounds__(NT)
k_clahe_lut(Dims d, const SliceRange* __restrict__ rng, const int* __restrict__ status, double gamma,
            float* __restrict__ lut) {
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (status[s]) return;
    const unsigned vmin = rng[si].vmin, vmax = rng[si].vmax;
    const int v = blockIdx.x * NT + threadIdx.x;
    if (v >= NLEVELS) return;
    const float lo = (float)vmin;
    const float den = (float)((double)vmax - (double)vmin);
    const float f = (float)v;
    float r;
    if (vmin != vmax) r = __fdiv_rn(__fsub_rn(f, lo), den);
    else r = fminf(fmaxf(f, 0.0f), 1.0f);
    if (gamma != 1.0 && v >= (int)vmin && v <= (int)vmax) r = (float)pow((double)r, gamma);   // as GammaF (pointwise.cu)
    lut[(size_t)si * NLEVELS + v] = r;
}

__global__ void __launch_bounds__(NT)
k_clahe_final(Dims d, const int* __restrict__ status, const uint16_t* __restrict__ vin,
              const float* __restrict__ lut, float* __restrict__ out) {
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (status[s]) return;
    const long long len = d.px();
    const uint16_t* v = vin + (size_t)si * len;
    float* o = out + (size_t)s * len;
    const float* L = lut + (size_t)si * NLEVELS;
    const long long tid = (long long)blockIdx.x * NT + threadIdx.x, nthr = (long long)gridDim.x * NT;
    if (((((uintptr_t)v) | ((uintptr_t)o)) & 15) == 0) {
        const long long n8 = len >> 3;
        const uint4* v8 = reinterpret_cast<const uint4*>(v);
        float4* o4 = reinterpret_cast<float4*>(o);
        for (long long i = tid; i < n8; i += nthr) {
            const uint4 q = v8[i];
            float4 a, b;
            // levels are <= 16383 by construction; the clamp only keeps a corrupted level in the table
            a.x = __ldg(L + min(q.x & 0xffffu, NLEVELS - 1u)); a.y = __ldg(L + min(q.x >> 16, NLEVELS - 1u));
            a.z = __ldg(L + min(q.y & 0xffffu, NLEVELS - 1u)); a.w = __ldg(L + min(q.y >> 16, NLEVELS - 1u));
            b.x = __ldg(L + min(q.z & 0xffffu, NLEVELS - 1u)); b.y = __ldg(L + min(q.z >> 16, NLEVELS - 1u));
            b.z = __ldg(L + min(q.w & 0xffffu, NLEVELS - 1u)); b.w = __ldg(L + min(q.w >> 16, NLEVELS - 1u));
            o4[2 * i] = a;
            o4[2 * i + 1] = b;
        }
        for (long long i = (n8 << 3) + tid; i < len; i += nthr) o[i] = __ldg(L + min((unsigned)v[i], NLEVELS - 1u));
    } else {
        for (long long i = tid; i < len; i += nthr) o[i] = __ldg(L + min((unsigned)v[i], NLEVELS - 1u));
    }
}

struct ClaheBufs { SliceRange* rng; uint8_t* bins; uint16_t* maps; uint16_t* v; float* lut; uint8_t* binlut; };

void carve(Arena& a, int n_sel, int h, int w, int nty, int ntx, ClaheBufs& b) {
    b.rng = a.take<SliceRange>(n_sel);
    b.bins = a.take<uint8_t>((size_t)n_sel * h * w);
    b.maps = a.take<uint16_t>((size_t)n_sel * nty * ntx * NBINS);
    b.v = a.take<uint16_t>((size_t)n_sel * h * w);
    b.lut = a.take<float>((size_t)n_sel * NLEVELS);
    b.binlut = a.take<uint8_t>((size_t)n_sel * NU16);
}

inline void geom(int h, int w, int k, double clip_limit, ClaheGeom& g) {
    g.k = k;
    g.nty = (h + k - 1) / k;      // int(padded / k) - 1 with padded = ceil(s/k)*k + k
    g.ntx = (w + k - 1) / k;
    const int kk = k * k;
    if (clip_limit > 0.0) {
        double c = clip_limit * (double)kk;
        if (c < 1.0) c = 1.0;
        g.clim = (int)c;
    } else {
        g.clim = 65535;           // np.iinfo(uint16).max: no clipping (AHE)
    }
    g.scale = 16383.0 / (double)kk;
    const unsigned long long reach = (unsigned long long)g.nty * g.ntx * g.ntx;      // largest tile index times ntx
    g.ntx_magic = (g.ntx > 1 && reach < (1ull << 32)) ? (unsigned)((1ull << 32) / (unsigned)g.ntx) + 1u : 0u;
}

}  // namespace

size_t clahe_workspace_bytes(int n, int n_sel, int h, int w, int kernel_size) {
    (void)n;
    if (kernel_size < 1) return 0;
    Arena a(nullptr, 0);
    ClaheBufs b;
    ClaheGeom g;
    geom(h, w, kernel_size, 0.01, g);
    carve(a, n_sel, h, w, g.nty, g.ntx, b);
    return a.off;
}

int clahe_run(const float* in, float* out, const Dims& d, double clip_limit, int kernel_size, double gamma,
              const uint2* mm, int* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (kernel_size < 1 || kernel_size > 48)
        return set_error(MDIMG_ERR_INVALID, "clahe: kernel_size %d outside [1, 48]", kernel_size);
    if (!(gamma >= 0.0)) return set_error(MDIMG_ERR_INVALID, "Gamma should be a non-negative real number.");
    ClaheGeom g;
    geom(d.h, d.w, kernel_size, clip_limit, g);
    Arena a(ws, ws_bytes);
    ClaheBufs b;
    carve(a, d.n_sel, d.h, d.w, g.nty, g.ntx, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "clahe: workspace too small (%zu > %zu)", a.off, ws_bytes);
    MDIMG_LAUNCH k_clahe_prep<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, mm, b.rng, status);
    const int ntiles = g.nty * g.ntx;
    MDIMG_LAUNCH k_clahe_binlut<<<dim3(NU16 / NT, d.n_sel), NT, 0, stream>>>(d, b.rng, status, b.binlut);
    MDIMG_LAUNCH k_clahe_hist<<<dim3((ntiles + WARPS - 1) / WARPS, d.n_sel), NT, 0, stream>>>(in, d, g, b.binlut, status, b.bins, b.maps);
    dim3 bgrid(((d.w + BW - 1) / BW) * ((d.h + BH - 1) / BH), d.n_sel);
    MDIMG_LAUNCH k_clahe_blend<<<bgrid, NT, 0, stream>>>(d, g, b.rng, status, b.bins, b.maps, b.v);
    long long len = d.px();
    int fb = (int)((len + NT * 8 - 1) / (NT * 8));
    if (fb > 4096) fb = 4096;
    MDIMG_LAUNCH k_clahe_lut<<<dim3(NLEVELS / NT, d.n_sel), NT, 0, stream>>>(d, b.rng, status, gamma, b.lut);
    MDIMG_LAUNCH k_clahe_final<<<dim3(fb, d.n_sel), NT, 0, stream>>>(d, status, b.v, b.lut, out);
    return check_launch("clahe");
}

}  // namespace mdimg
