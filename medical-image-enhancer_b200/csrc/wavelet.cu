// Wavelet denoise: skimage.restoration.denoise_wavelet(image, channel_axis=None,
// rescale_sigma=True, mode=soft|hard[, sigma=...]) as called at pipeline/enhancement.py:86,169,
// 270,328 — PyWavelets 'db1' (Haar) pyramid, L = max(floor(log2(min(h, w))) - 3, 1) levels,
// BayesShrink threshold per detail band, soft / hard shrink, inverse transform.
//
// Arithmetic follows pywt: the forward transform accumulates in float32
// (out[o] = f[0]*x[2o+1] + f[1]*x[2o], axis 0 then axis 1, symmetric extension on odd lengths);
// the inverse runs axis 1 then axis 0 as lo*a + hi*d.  numpy >= 2 promotion decides the
// precision of the inverse: with an estimated sigma (a float64 scalar) and soft shrinkage the
// shrunk details and therefore the whole inverse are float64; with a caller-supplied python-float
// sigma (the _light_denoise path) or hard shrinkage everything stays float32.
//
// This is the general per-level implementation (any size, odd lengths included): one kernel per
// level on the way down (also accumulates the band energies and the finest 'dd' histogram), a tiny
// threshold kernel, one kernel per level on the way up with the shrinkage fused into the load.
#include "enhance.cuh"
#include "select.cuh"

#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

namespace mdimg {

namespace {

constexpr int NT = 256;
constexpr int MAXL = 14;
constexpr double SQ = 0.7071067811865476;

struct Pyramid {
    int L;
    int H[MAXL + 1], W[MAXL + 1];
    long long off[MAXL + 1];    // float offset of level l's three detail bands inside one slice
    long long det_per_slice;    // floats
    long long a_cap;            // floats per approximation ping-pong buffer
    long long r_cap;            // elements per reconstruction ping-pong buffer
};

Pyramid make_pyramid(int h, int w) {
    Pyramid p;
    int m = h < w ? h : w;
    int lg = 0;
    while ((1 << (lg + 1)) <= m) ++lg;
    p.L = lg - 3 > 1 ? lg - 3 : 1;
    if (p.L > MAXL) p.L = MAXL;
    p.H[0] = h; p.W[0] = w;
    long long o = 0;
    for (int l = 1; l <= p.L; ++l) {
        p.H[l] = (p.H[l - 1] + 1) / 2;
        p.W[l] = (p.W[l - 1] + 1) / 2;
        p.off[l] = o;
        o += 3LL * p.H[l] * p.W[l];
    }
    p.det_per_slice = (o + 3) & ~3LL;          // slices start 16-byte aligned (128-bit coefficient accesses)
    p.a_cap = (long long)p.H[1] * p.W[1];
    p.r_cap = (long long)(p.H[1] + 2) * (p.W[1] + 2);
    return p;
}

struct WaveAcc {                 // per slice (position in sel)
    float energy[MAXL + 1][3];   // np.sum(d*d) per detail band (ad, da, dd): float32, numpy's pairwise order
    double thr[MAXL + 1][3];     // BayesShrink thresholds
    double sigma;
    unsigned dd_zero;
    unsigned pad;
};

// ---- forward level -------------------------------------------------------------------------
// in: level l-1 approximation (h0 x w0, pitch w0).  First level reads the image (slice id `s`),
// deeper levels read the compact ping-pong buffer (index `si`).
template <bool FIRST>
__global__ void __launch_bounds__(NT)
k_haar_fwd(const float* __restrict__ in, long long in_stride, int h0, int w0, Dims d,
           const int* __restrict__ skip, int level, float* __restrict__ aout, long long a_stride,
           float* __restrict__ det, long long det_stride, long long det_off,
           WaveAcc* __restrict__ acc, unsigned* __restrict__ l1, int want_hist) {
    __shared__ unsigned hh[SEL_L1_BINS];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (skip && skip[s]) return;
    const int h1 = (h0 + 1) / 2, w1 = (w0 + 1) / 2;
    const float* src = in + (size_t)(FIRST ? s : si) * in_stride;
    float* ao = aout + (size_t)si * a_stride;
    float* db = det + (size_t)si * det_stride + det_off;
    const long long band = (long long)h1 * w1;
    const float S = (float)SQ, NS = -(float)SQ;
    const bool hist = FIRST && want_hist;
    if (hist) for (int i = threadIdx.x; i < SEL_L1_BINS; i += NT) hh[i] = 0;
    if (hist) __syncthreads();
    unsigned nz = 0;
    const long long total = band;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int y = (int)(i / w1), x = (int)(i - (long long)y * w1);
        const int r0 = 2 * y, r1 = min(2 * y + 1, h0 - 1);
        const int c0 = 2 * x, c1 = min(2 * x + 1, w0 - 1);
        const float v00 = src[(size_t)r0 * w0 + c0], v01 = src[(size_t)r0 * w0 + c1];
        const float v10 = src[(size_t)r1 * w0 + c0], v11 = src[(size_t)r1 * w0 + c1];
        // axis 0
        const float lo0 = __fadd_rn(__fmul_rn(S, v10), __fmul_rn(S, v00));
        const float lo1 = __fadd_rn(__fmul_rn(S, v11), __fmul_rn(S, v01));
        const float hi0 = __fadd_rn(__fmul_rn(NS, v10), __fmul_rn(S, v00));
        const float hi1 = __fadd_rn(__fmul_rn(NS, v11), __fmul_rn(S, v01));
        // axis 1
        const float aa = __fadd_rn(__fmul_rn(S, lo1), __fmul_rn(S, lo0));
        const float ad = __fadd_rn(__fmul_rn(NS, lo1), __fmul_rn(S, lo0));
        const float da = __fadd_rn(__fmul_rn(S, hi1), __fmul_rn(S, hi0));
        const float dd = __fadd_rn(__fmul_rn(NS, hi1), __fmul_rn(S, hi0));
        ao[i] = aa;
        db[i] = ad;
        db[band + i] = da;
        db[2 * band + i] = dd;
        if (hist) {
            const float a = fabsf(dd);
            nz += (a == 0.0f);
            atomicAdd(&hh[sel_bin1(a)], 1u);
        }
    }
    WaveAcc* A = acc + si;
    if (hist) {
        nz = warp_sum_u(nz);
        if ((threadIdx.x & 31) == 0 && nz) atomicAdd(&A->dd_zero, nz);
        __syncthreads();
        unsigned* g = l1 + (size_t)si * SEL_L1_BINS;
        for (int i = threadIdx.x; i < SEL_L1_BINS; i += NT) { unsigned v = hh[i]; if (v) atomicAdd(&g[i], v); }
    }
}

// ---- fused forward levels 1..3 ---------------------------------------------------------------
// For extents divisible by 8 every 8 x 8 pixel block is an independent 3-level Haar pyramid (SURVEY 8c
// item 5).  ONE THREAD owns a block: it loads its 64 pixels (128-bit loads), runs levels 1-3 entirely in
// registers with the arithmetic of k_haar_fwd, and stores the detail coefficients of the three levels
// (4 / 2 / 1 consecutive values per band row: 128- / 64- / 32-bit stores, contiguous across the warp) and
// its level-3 approximation, where the per-level kernels continue on 1/64 of the data.  No shared
// memory, no barrier, one read of the image and one write of the coefficients.
struct Haar4 { float aa, ad, da, dd; };
__device__ __forceinline__ Haar4 haar_fwd_2x2(float v00, float v01, float v10, float v11) {
    const float S = (float)SQ, NS = -(float)SQ;
    // axis 0, then axis 1 (pywt: out[o] = f[0] * x[2o + 1] + f[1] * x[2o], float32)
    const float lo0 = __fadd_rn(__fmul_rn(S, v10), __fmul_rn(S, v00));
    const float lo1 = __fadd_rn(__fmul_rn(S, v11), __fmul_rn(S, v01));
    const float hi0 = __fadd_rn(__fmul_rn(NS, v10), __fmul_rn(S, v00));
    const float hi1 = __fadd_rn(__fmul_rn(NS, v11), __fmul_rn(S, v01));
    Haar4 r;
    r.aa = __fadd_rn(__fmul_rn(S, lo1), __fmul_rn(S, lo0));
    r.ad = __fadd_rn(__fmul_rn(NS, lo1), __fmul_rn(S, lo0));
    r.da = __fadd_rn(__fmul_rn(S, hi1), __fmul_rn(S, hi0));
    r.dd = __fadd_rn(__fmul_rn(NS, hi1), __fmul_rn(S, hi0));
    return r;
}

__global__ void __launch_bounds__(NT)
k_haar_fwd_reg3(const float* __restrict__ in, Dims d, const int* __restrict__ skip, Pyramid p,
                float* __restrict__ aout, float* __restrict__ det, WaveAcc* __restrict__ acc,
                unsigned* __restrict__ l1, int want_hist) {
    __shared__ unsigned hh[SEL_L1_BINS];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (skip && skip[s]) return;
    const int tid = threadIdx.x;
    if (want_hist) {
        for (int i = tid; i < SEL_L1_BINS; i += NT) hh[i] = 0;
        __syncthreads();
    }
    const int bw = d.w >> 3, nblk = (d.h >> 3) * bw;
    const float* src = in + (size_t)s * d.h * d.w;
    float* db = det + (size_t)si * p.det_per_slice;
    const int W1 = p.W[1], W2 = p.W[2], W3 = p.W[3];
    const long long band1 = (long long)p.H[1] * W1, band2 = (long long)p.H[2] * W2, band3 = (long long)p.H[3] * W3;
    float* d1 = db + p.off[1];
    float* d2 = db + p.off[2];
    float* d3 = db + p.off[3];
    float* a3 = aout + (size_t)si * p.a_cap;
    unsigned nz = 0;
    for (int b = blockIdx.x * NT + tid; b < nblk; b += gridDim.x * NT) {
        const int by = b / bw, bx = b - by * bw;
        const float* blk = src + (size_t)(8 * by) * d.w + 8 * bx;
        float a1[4][4];                                   // level-1 approximations of the block
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(blk + (size_t)(2 * yy) * d.w));
            const float4 t1 = __ldg(reinterpret_cast<const float4*>(blk + (size_t)(2 * yy) * d.w) + 1);
            const float4 u0 = __ldg(reinterpret_cast<const float4*>(blk + (size_t)(2 * yy + 1) * d.w));
            const float4 u1 = __ldg(reinterpret_cast<const float4*>(blk + (size_t)(2 * yy + 1) * d.w) + 1);
            const Haar4 c0 = haar_fwd_2x2(t0.x, t0.y, u0.x, u0.y), c1 = haar_fwd_2x2(t0.z, t0.w, u0.z, u0.w);
            const Haar4 c2 = haar_fwd_2x2(t1.x, t1.y, u1.x, u1.y), c3 = haar_fwd_2x2(t1.z, t1.w, u1.z, u1.w);
            a1[yy][0] = c0.aa; a1[yy][1] = c1.aa; a1[yy][2] = c2.aa; a1[yy][3] = c3.aa;
            const size_t o = (size_t)(4 * by + yy) * W1 + 4 * bx;
            *reinterpret_cast<float4*>(d1 + o) = make_float4(c0.ad, c1.ad, c2.ad, c3.ad);
            *reinterpret_cast<float4*>(d1 + band1 + o) = make_float4(c0.da, c1.da, c2.da, c3.da);
            *reinterpret_cast<float4*>(d1 + 2 * band1 + o) = make_float4(c0.dd, c1.dd, c2.dd, c3.dd);
            if (want_hist) {
                const float e[4] = {fabsf(c0.dd), fabsf(c1.dd), fabsf(c2.dd), fabsf(c3.dd)};
#pragma unroll
                for (int k = 0; k < 4; ++k) { nz += (e[k] == 0.0f); atomicAdd(&hh[sel_bin1(e[k])], 1u); }
            }
        }
        float a2[2][2];
#pragma unroll
        for (int yy = 0; yy < 2; ++yy) {
            const Haar4 c0 = haar_fwd_2x2(a1[2 * yy][0], a1[2 * yy][1], a1[2 * yy + 1][0], a1[2 * yy + 1][1]);
            const Haar4 c1 = haar_fwd_2x2(a1[2 * yy][2], a1[2 * yy][3], a1[2 * yy + 1][2], a1[2 * yy + 1][3]);
            a2[yy][0] = c0.aa; a2[yy][1] = c1.aa;
            const size_t o = (size_t)(2 * by + yy) * W2 + 2 * bx;
            *reinterpret_cast<float2*>(d2 + o) = make_float2(c0.ad, c1.ad);
            *reinterpret_cast<float2*>(d2 + band2 + o) = make_float2(c0.da, c1.da);
            *reinterpret_cast<float2*>(d2 + 2 * band2 + o) = make_float2(c0.dd, c1.dd);
        }
        const Haar4 c = haar_fwd_2x2(a2[0][0], a2[0][1], a2[1][0], a2[1][1]);
        const size_t o3 = (size_t)by * W3 + bx;
        d3[o3] = c.ad;
        d3[band3 + o3] = c.da;
        d3[2 * band3 + o3] = c.dd;
        a3[o3] = c.aa;
    }
    if (want_hist) {
        nz = warp_sum_u(nz);
        if ((tid & 31) == 0 && nz) atomicAdd(&acc[si].dd_zero, nz);
        __syncthreads();
        unsigned* g = l1 + (size_t)si * SEL_L1_BINS;
        for (int i = tid; i < SEL_L1_BINS; i += NT) { unsigned v = hh[i]; if (v) atomicAdd(&g[i], v); }
    }
}

// ---- np.sum(d*d) in numpy's exact float32 pairwise order ---------------------------------------
// numpy reduces a contiguous float32 array with pairwise_sum: blocks of <= 128 elements are summed
// with 8 interleaved accumulators combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a
// sequential tail; larger ranges split at n/2 rounded down to a multiple of 8 and add the halves.
// The BayesShrink threshold depends on these float32 sums, and a 1-ulp change of a threshold can
// move a pixel across a CLAHE bin edge downstream, so the order is reproduced exactly: a leaf
// table and a post-order combine program are built on the host for each band length.
struct PwLeaf { int level; int off; int len; };          // off: element offset inside the band
// An internal node adds the sum of its right subtree (first leaf m) to that of its left subtree
// (first leaf a); a node's value lives in the slot of its first leaf, so the combine is in place.
struct PwNode { int a; int m; };
constexpr int PW_MAX_ROUNDS = 40;
struct PwLevel {
    int leaf0; int n_leaf;
    int node0;                              // first node of this level in the node table (sorted by height)
    int n_rounds;                           // tree height
    int round_end[PW_MAX_ROUNDS];           // nodes of height <= r + 1 end here (relative to node0)
    long long det_off; long long band;
};

struct PwPlan {
    std::vector<PwLeaf> leaves;
    std::vector<PwNode> nodes;
    std::vector<PwLevel> levels;           // index 0 unused
    int max_depth = 0;
};

struct PwTmpNode { int a, m, height; };

// returns the height of the subtree; leaf indices are relative to the level's first leaf
int pw_build(int off, int n, int level, int leaf0, PwPlan& p, std::vector<PwTmpNode>& tmp, int depth) {
    if (depth > p.max_depth) p.max_depth = depth;
    if (n <= 128) {
        p.leaves.push_back({level, off, n});
        return 0;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const int a = (int)p.leaves.size() - leaf0;
    const int hl = pw_build(off, n2, level, leaf0, p, tmp, depth + 1);
    const int m = (int)p.leaves.size() - leaf0;
    const int hr = pw_build(off + n2, n - n2, level, leaf0, p, tmp, depth + 1);
    const int h = 1 + (hl > hr ? hl : hr);
    tmp.push_back({a, m, h});
    return h;
}

PwPlan make_pw_plan(const Pyramid& py) {
    PwPlan p;
    p.levels.resize(py.L + 1);
    for (int l = 1; l <= py.L; ++l) {
        PwLevel& lv = p.levels[l];
        lv.leaf0 = (int)p.leaves.size();
        lv.node0 = (int)p.nodes.size();
        lv.band = (long long)py.H[l] * py.W[l];
        lv.det_off = py.off[l];
        std::vector<PwTmpNode> tmp;
        const int height = pw_build(0, (int)lv.band, l, lv.leaf0, p, tmp, 1);
        lv.n_leaf = (int)p.leaves.size() - lv.leaf0;
        lv.n_rounds = height;
        for (int r = 1; r <= height && r <= PW_MAX_ROUNDS; ++r) {
            for (const PwTmpNode& t : tmp) if (t.height == r) p.nodes.push_back({t.a, t.m});
            lv.round_end[r - 1] = (int)p.nodes.size() - lv.node0;
        }
    }
    return p;
}

const PwPlan& cached_pw_plan(const Pyramid& py) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::unique_ptr<PwPlan>> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_pair(py.H[0], py.W[0]);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, std::make_unique<PwPlan>(make_pw_plan(py))).first;
    return *it->second;
}

// 8 lanes per leaf (lane j = numpy's accumulator r[j]); grid.y = slice, grid.z = band.
__global__ void __launch_bounds__(NT)
k_pw_leaves(const float* __restrict__ det, long long det_stride, const PwLeaf* __restrict__ leaves,
            const PwLevel* __restrict__ levels, int n_leaves, Dims d, const int* __restrict__ skip,
            float* __restrict__ leaf_sums) {
    const int si = blockIdx.y, b = blockIdx.z;
    if (skip && skip[slice_of(d.sel, si)]) return;
    const int g = (blockIdx.x * NT + threadIdx.x) >> 3, j = threadIdx.x & 7;
    if (g >= n_leaves) return;                      // whole 8-lane groups leave together
    const PwLeaf lf = leaves[g];
    const PwLevel lv = levels[lf.level];
    const float* a = det + (size_t)si * det_stride + lv.det_off + (size_t)b * lv.band + lf.off;
    const int n = lf.len;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    float res;
    if (n < 8) {
        res = 0.0f;
        if (j == 0) for (int i = 0; i < n; ++i) { const float v = a[i]; res = __fadd_rn(res, __fmul_rn(v, v)); }
    } else {
        float v = a[j];
        float r = __fmul_rn(v, v);
        const int n8 = n - (n & 7);
        for (int i = 8; i < n8; i += 8) { v = a[i + j]; r = __fadd_rn(r, __fmul_rn(v, v)); }
        r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 1));     // r0+r1, r2+r3, ...
        r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 2));     // (r0+r1)+(r2+r3), ...
        r = __fadd_rn(r, __shfl_xor_sync(gmask, r, 4));
        res = r;
        if (j == 0) for (int i = n8; i < n; ++i) { v = a[i]; res = __fadd_rn(res, __fmul_rn(v, v)); }
    }
    if (j == 0) leaf_sums[((size_t)si * 3 + b) * n_leaves + g] = res;
}

// One block per (slice, level, band): the pairwise tree is reduced in place, one round per tree
// height (nodes of equal height are independent), exactly the additions numpy performs.
__global__ void __launch_bounds__(NT)
k_pw_combine(const PwNode* __restrict__ nodes, const PwLevel* __restrict__ levels, int L, int n_leaves,
             Dims d, const int* __restrict__ skip, float* __restrict__ leaf_sums, WaveAcc* __restrict__ acc) {
    const int si = blockIdx.y;
    const int l = blockIdx.x / 3 + 1, b = blockIdx.x - (l - 1) * 3;
    if (skip && skip[slice_of(d.sel, si)]) return;
    __shared__ PwLevel lv;
    if (threadIdx.x == 0) lv = levels[l];
    __syncthreads();
    float* ls = leaf_sums + ((size_t)si * 3 + b) * n_leaves + lv.leaf0;
    const PwNode* nd = nodes + lv.node0;
    int begin = 0;
    for (int r = 0; r < lv.n_rounds; ++r) {
        const int end = lv.round_end[r];
        for (int i = begin + threadIdx.x; i < end; i += NT) {
            const PwNode n = nd[i];
            ls[n.a] = __fadd_rn(ls[n.a], ls[n.m]);
        }
        begin = end;
        __syncthreads();
    }
    if (threadIdx.x == 0) acc[si].energy[l][b] = ls[0];
}

__global__ void k_wave_ranks(Dims d, int len, const WaveAcc* __restrict__ acc, int* __restrict__ ranks) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    int nz = (int)acc[si].dd_zero;
    int m = len - nz;
    if (m <= 0) { ranks[s * 2] = -1; ranks[s * 2 + 1] = -1; return; }
    ranks[s * 2] = nz + (m - 1) / 2;
    ranks[s * 2 + 1] = nz + m / 2;
}

// BayesShrink thresholds (skimage _bayes_thresh) for every level / band of every slice.
__global__ void k_wave_thresholds(Dims d, Pyramid p, WaveAcc* __restrict__ acc,
                                  const float* __restrict__ med_pair, const double* __restrict__ sigma_in,
                                  double sigma_scale) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    WaveAcc& A = acc[si];
    const float eps = 1.1920928955078125e-07f;   // np.finfo(np.float32).eps
    if (sigma_in == nullptr) {
        // sigma = np.float64: float32 median / norm.ppf(0.75)
        const float med = __fdiv_rn(__fadd_rn(med_pair[s * 2], med_pair[s * 2 + 1]), 2.0f);
        const double sigma = (double)med / 0.6744897501960817;
        A.sigma = sigma;
        const double var = sigma * sigma;
        for (int l = 1; l <= p.L; ++l) {
            const double cnt = (double)p.H[l] * (double)p.W[l];
            for (int b = 0; b < 3; ++b) {
                const float dvar = __fdiv_rn(A.energy[l][b], (float)cnt);   // np.mean: float32 sum / count
                const double diff = (double)dvar - var;               // float32 - float64 -> float64
                double root;
                if (diff >= (double)eps) root = sqrt(diff);           // max(diff, eps) keeps `diff` on ties
                else if (diff != diff) root = diff;                   // NaN propagates through max()
                else root = (double)sqrtf(eps);
                A.thr[l][b] = var / root;
            }
        }
    } else {
        // sigma is a python float: everything around it stays float32 (NEP 50 weak scalar)
        const double sigma = sigma_in[s] * sigma_scale;
        A.sigma = sigma;
        const double var = sigma * sigma;
        const float varf = (float)var;
        for (int l = 1; l <= p.L; ++l) {
            const double cnt = (double)p.H[l] * (double)p.W[l];
            for (int b = 0; b < 3; ++b) {
                const float dvar = __fdiv_rn(A.energy[l][b], (float)cnt);
                float diff = __fsub_rn(dvar, varf);
                float m = diff >= eps ? diff : (diff != diff ? diff : eps);
                A.thr[l][b] = (double)__fdiv_rn(varf, __fsqrt_rn(m));
            }
        }
    }
}

// Optional epilogue of the last inverse level: out = c0 * x + c1 * denoised in float32 (products and sum
// rounded separately, numpy's order) -- the blend of _light_denoise (pipeline/enhancement.py:93) fused into
// the store, so the denoised image never makes a round trip through HBM.
struct Blend { float c0, c1; int on; };
__device__ __forceinline__ float blend_px(const Blend& bl, float x, float v) {
    return bl.on ? __fadd_rn(__fmul_rn(bl.c0, x), __fmul_rn(bl.c1, v)) : v;
}

template <typename T> struct Ops;
template <> struct Ops<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct Ops<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

// pywt.threshold on one coefficient.  MODE 0: soft with a float64 threshold (result float64);
// 1: soft with a float32 threshold; 2: hard (data stays float32; the compare is exact).
template <typename T, int MODE>
__device__ __forceinline__ T shrink(float dcoef, double thr) {
    if (MODE == 0) {
        const double dv = (double)dcoef;
        double f = __dsub_rn(1.0, __ddiv_rn(thr, fabs(dv)));
        f = f < 0.0 ? 0.0 : f;           // ndarray.clip(min=0): NaN stays NaN
        return (T)__dmul_rn(dv, f);
    } else if (MODE == 1) {
        float f = __fsub_rn(1.0f, __fdiv_rn((float)thr, fabsf(dcoef)));
        f = f < 0.0f ? 0.0f : f;
        return (T)__fmul_rn(dcoef, f);
    } else {
        return (T)(((double)fabsf(dcoef) < thr) ? 0.0f : dcoef);
    }
}

// ---- inverse level -------------------------------------------------------------------------
// ain: approximation at this level (float32 for the coarsest level, else T), read on [hd x wd]
// with pitch a_pitch.  Writes the 2hd x 2wd reconstruction (pitch out_pitch), or, for the last
// level, the cropped float32 image.
template <typename T, typename TA, int MODE, bool LAST>
__global__ void __launch_bounds__(NT)
k_haar_inv(const TA* __restrict__ ain, long long a_stride, int a_pitch, int hd, int wd, Dims d,
           const int* __restrict__ skip, int level, const float* __restrict__ det,
           long long det_stride, long long det_off, const WaveAcc* __restrict__ acc,
           T* __restrict__ rout, long long r_stride, int out_pitch,
           const float* __restrict__ img_in, float* __restrict__ img_out, Blend bl) {
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (skip && skip[s]) {
        if (LAST) {   // untouched slice: copy through
            const long long len = d.px();
            const float* a = img_in + (size_t)s * len;
            float* o = img_out + (size_t)s * len;
            if (a != o)
                for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < len; i += (long long)gridDim.x * NT) o[i] = a[i];
        }
        return;
    }
    const TA* ap = ain + (size_t)si * a_stride;
    const float* db = det + (size_t)si * det_stride + det_off;
    const long long band = (long long)hd * wd;
    const WaveAcc& A = acc[si];
    const double t_ad = A.thr[level][0], t_da = A.thr[level][1], t_dd = A.thr[level][2];
    const T S = (T)SQ, NS = -(T)SQ;
    typedef Ops<T> O;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < band; i += (long long)gridDim.x * NT) {
        const int y = (int)(i / wd), x = (int)(i - (long long)y * wd);
        const T aa = (T)ap[(size_t)y * a_pitch + x];
        const T ad = shrink<T, MODE>(db[i], t_ad);
        const T da = shrink<T, MODE>(db[band + i], t_da);
        const T dd = shrink<T, MODE>(db[2 * band + i], t_dd);
        // axis 1: 'a' = idwt(aa, ad), 'd' = idwt(da, dd);  out[2x] = lo0*a + hi0*d, out[2x+1] = lo1*a + hi1*d
        const T a_e = O::add(O::mul(S, aa), O::mul(S, ad));
        const T a_o = O::add(O::mul(S, aa), O::mul(NS, ad));
        const T d_e = O::add(O::mul(S, da), O::mul(S, dd));
        const T d_o = O::add(O::mul(S, da), O::mul(NS, dd));
        // axis 0
        const T o00 = O::add(O::mul(S, a_e), O::mul(S, d_e));
        const T o10 = O::add(O::mul(S, a_e), O::mul(NS, d_e));
        const T o01 = O::add(O::mul(S, a_o), O::mul(S, d_o));
        const T o11 = O::add(O::mul(S, a_o), O::mul(NS, d_o));
        if (LAST) {
            float* o = img_out + (size_t)s * d.h * d.w;
            const float* xi = img_in + (size_t)s * d.h * d.w;      // read only when blending (same pixel this thread writes)
            const int Y = 2 * y, X = 2 * x;
            auto put = [&](size_t at, T v) { o[at] = blend_px(bl, bl.on ? xi[at] : 0.0f, (float)v); };
            if (Y < d.h && X < d.w) put((size_t)Y * d.w + X, o00);
            if (Y < d.h && X + 1 < d.w) put((size_t)Y * d.w + X + 1, o01);
            if (Y + 1 < d.h && X < d.w) put((size_t)(Y + 1) * d.w + X, o10);
            if (Y + 1 < d.h && X + 1 < d.w) put((size_t)(Y + 1) * d.w + X + 1, o11);
        } else {
            T* o = rout + (size_t)si * r_stride;
            const size_t b0 = (size_t)(2 * y) * out_pitch + 2 * x;
            o[b0] = o00; o[b0 + 1] = o01;
            o[b0 + out_pitch] = o10; o[b0 + out_pitch + 1] = o11;
        }
    }
}

// ---- fused inverse levels 3..1 ---------------------------------------------------------------
// The counterpart of k_haar_fwd_reg3: a thread takes its level-3 approximation and the 63 detail
// coefficients of its block (shrinkage fused into the loads), rebuilds levels 2 and 1 in registers with the
// arithmetic of k_haar_inv and stores its 8 x 8 pixels of the float32 result (128-bit stores).
template <typename T> struct Quad { T o00, o01, o10, o11; };
template <typename T>
__device__ __forceinline__ Quad<T> haar_inv_2x2(T aa, T ad, T da, T dd) {
    typedef Ops<T> O;
    const T S = (T)SQ, NS = -(T)SQ;
    // axis 1: 'a' = idwt(aa, ad), 'd' = idwt(da, dd); then axis 0
    const T a_e = O::add(O::mul(S, aa), O::mul(S, ad));
    const T a_o = O::add(O::mul(S, aa), O::mul(NS, ad));
    const T d_e = O::add(O::mul(S, da), O::mul(S, dd));
    const T d_o = O::add(O::mul(S, da), O::mul(NS, dd));
    Quad<T> q;
    q.o00 = O::add(O::mul(S, a_e), O::mul(S, d_e));
    q.o10 = O::add(O::mul(S, a_e), O::mul(NS, d_e));
    q.o01 = O::add(O::mul(S, a_o), O::mul(S, d_o));
    q.o11 = O::add(O::mul(S, a_o), O::mul(NS, d_o));
    return q;
}

template <typename T, typename TA, int MODE>
__global__ void __launch_bounds__(NT)
k_haar_inv_reg3(const TA* __restrict__ ain, long long a_stride, int a_pitch, Dims d, const int* __restrict__ skip,
                Pyramid p, const float* __restrict__ det, const WaveAcc* __restrict__ acc,
                const float* __restrict__ img_in, float* __restrict__ img_out, Blend bl) {
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tid = threadIdx.x;
    float* dst = img_out + (size_t)s * d.h * d.w;
    const float* xin = img_in + (size_t)s * d.h * d.w;
    if (skip && skip[s]) {                          // untouched slice: copy through
        const float* a = img_in + (size_t)s * d.h * d.w;
        const long long len = d.px();
        if (a != dst)
            for (long long i = (long long)blockIdx.x * NT + tid; i < len; i += (long long)gridDim.x * NT) dst[i] = a[i];
        return;
    }
    const int bw = d.w >> 3, nblk = (d.h >> 3) * bw;
    const float* db = det + (size_t)si * p.det_per_slice;
    const int W1 = p.W[1], W2 = p.W[2], W3 = p.W[3];
    const long long band1 = (long long)p.H[1] * W1, band2 = (long long)p.H[2] * W2, band3 = (long long)p.H[3] * W3;
    const float* d1 = db + p.off[1];
    const float* d2 = db + p.off[2];
    const float* d3 = db + p.off[3];
    const WaveAcc& A = acc[si];
    const double t1a = A.thr[1][0], t1b = A.thr[1][1], t1c = A.thr[1][2];
    const double t2a = A.thr[2][0], t2b = A.thr[2][1], t2c = A.thr[2][2];
    const double t3a = A.thr[3][0], t3b = A.thr[3][1], t3c = A.thr[3][2];
    const TA* ap = ain + (size_t)si * a_stride;
    for (int b = blockIdx.x * NT + tid; b < nblk; b += gridDim.x * NT) {
        const int by = b / bw, bx = b - by * bw;
        const size_t o3 = (size_t)by * W3 + bx;
        const Quad<T> q3 = haar_inv_2x2<T>((T)ap[(size_t)by * a_pitch + bx], shrink<T, MODE>(d3[o3], t3a),
                                           shrink<T, MODE>(d3[band3 + o3], t3b), shrink<T, MODE>(d3[2 * band3 + o3], t3c));
        const T a2[2][2] = {{q3.o00, q3.o01}, {q3.o10, q3.o11}};
        T a1[4][4];
#pragma unroll
        for (int yy = 0; yy < 2; ++yy) {
            const size_t o = (size_t)(2 * by + yy) * W2 + 2 * bx;
            const float2 ad = *reinterpret_cast<const float2*>(d2 + o);
            const float2 da = *reinterpret_cast<const float2*>(d2 + band2 + o);
            const float2 dd = *reinterpret_cast<const float2*>(d2 + 2 * band2 + o);
            const Quad<T> qa = haar_inv_2x2<T>(a2[yy][0], shrink<T, MODE>(ad.x, t2a), shrink<T, MODE>(da.x, t2b), shrink<T, MODE>(dd.x, t2c));
            const Quad<T> qb = haar_inv_2x2<T>(a2[yy][1], shrink<T, MODE>(ad.y, t2a), shrink<T, MODE>(da.y, t2b), shrink<T, MODE>(dd.y, t2c));
            a1[2 * yy][0] = qa.o00; a1[2 * yy][1] = qa.o01; a1[2 * yy][2] = qb.o00; a1[2 * yy][3] = qb.o01;
            a1[2 * yy + 1][0] = qa.o10; a1[2 * yy + 1][1] = qa.o11; a1[2 * yy + 1][2] = qb.o10; a1[2 * yy + 1][3] = qb.o11;
        }
        float* blk = dst + (size_t)(8 * by) * d.w + 8 * bx;
#pragma unroll
        for (int yy = 0; yy < 4; ++yy) {
            const size_t o = (size_t)(4 * by + yy) * W1 + 4 * bx;
            const float4 ad = *reinterpret_cast<const float4*>(d1 + o);
            const float4 da = *reinterpret_cast<const float4*>(d1 + band1 + o);
            const float4 dd = *reinterpret_cast<const float4*>(d1 + 2 * band1 + o);
            const Quad<T> q0 = haar_inv_2x2<T>(a1[yy][0], shrink<T, MODE>(ad.x, t1a), shrink<T, MODE>(da.x, t1b), shrink<T, MODE>(dd.x, t1c));
            const Quad<T> q1 = haar_inv_2x2<T>(a1[yy][1], shrink<T, MODE>(ad.y, t1a), shrink<T, MODE>(da.y, t1b), shrink<T, MODE>(dd.y, t1c));
            const Quad<T> q2 = haar_inv_2x2<T>(a1[yy][2], shrink<T, MODE>(ad.z, t1a), shrink<T, MODE>(da.z, t1b), shrink<T, MODE>(dd.z, t1c));
            const Quad<T> q3r = haar_inv_2x2<T>(a1[yy][3], shrink<T, MODE>(ad.w, t1a), shrink<T, MODE>(da.w, t1b), shrink<T, MODE>(dd.w, t1c));
            float4* r0 = reinterpret_cast<float4*>(blk + (size_t)(2 * yy) * d.w);
            float4* r1 = reinterpret_cast<float4*>(blk + (size_t)(2 * yy + 1) * d.w);
            float4 e0 = make_float4((float)q0.o00, (float)q0.o01, (float)q1.o00, (float)q1.o01);
            float4 e1 = make_float4((float)q2.o00, (float)q2.o01, (float)q3r.o00, (float)q3r.o01);
            float4 f0 = make_float4((float)q0.o10, (float)q0.o11, (float)q1.o10, (float)q1.o11);
            float4 f1 = make_float4((float)q2.o10, (float)q2.o11, (float)q3r.o10, (float)q3r.o11);
            if (bl.on) {                            // the same 16 pixels of the input image, read before they are written
                const float* xb = xin + (size_t)(8 * by) * d.w + 8 * bx;
                const float4* x0 = reinterpret_cast<const float4*>(xb + (size_t)(2 * yy) * d.w);
                const float4* x1 = reinterpret_cast<const float4*>(xb + (size_t)(2 * yy + 1) * d.w);
                auto mix = [&](float4& v, const float4 x) {
                    v.x = blend_px(bl, x.x, v.x); v.y = blend_px(bl, x.y, v.y);
                    v.z = blend_px(bl, x.z, v.z); v.w = blend_px(bl, x.w, v.w);
                };
                const float4 xa = x0[0], xb4 = x0[1], xc = x1[0], xd = x1[1];
                mix(e0, xa); mix(e1, xb4); mix(f0, xc); mix(f1, xd);
            }
            r0[0] = e0; r0[1] = e1;
            r1[0] = f0; r1[1] = f1;
        }
    }
}

struct WaveBufs {
    WaveAcc* acc; float* det; float* a0; float* a1; double* r0; double* r1;
    unsigned* l1; int* ranks; float* med; void* sel_ws; size_t sel_ws_bytes;
    PwLeaf* pw_leaves; PwLevel* pw_levels; PwNode* pw_nodes; float* pw_sums;
};

// upper bounds that depend only on the pyramid (the plan itself is built per call)
inline size_t pw_max_leaves(const Pyramid& p) {
    size_t n = 0;
    for (int l = 1; l <= p.L; ++l) n += (size_t)p.H[l] * p.W[l] / 32 + 2;   // leaves hold > 64 elements, except tiny bands
    return n;
}

void carve(Arena& a, int n, int n_sel, const Pyramid& p, WaveBufs& b) {
    b.acc = a.take<WaveAcc>(n_sel);
    b.det = a.take<float>((size_t)n_sel * p.det_per_slice);
    b.a0 = a.take<float>((size_t)n_sel * p.a_cap);
    b.a1 = a.take<float>((size_t)n_sel * p.a_cap);
    b.r0 = a.take<double>((size_t)n_sel * p.r_cap);
    b.r1 = a.take<double>((size_t)n_sel * p.r_cap);
    b.l1 = a.take<unsigned>((size_t)n_sel * SEL_L1_BINS);
    b.ranks = a.take<int>((size_t)n * 2);
    b.med = a.take<float>((size_t)n * 2);
    b.sel_ws_bytes = select_workspace_bytes(n_sel);
    b.sel_ws = a.take<char>(b.sel_ws_bytes);
    const size_t ml = pw_max_leaves(p);
    b.pw_leaves = a.take<PwLeaf>(ml);
    b.pw_levels = a.take<PwLevel>(MAXL + 2);
    b.pw_nodes = a.take<PwNode>(ml + 16);
    b.pw_sums = a.take<float>((size_t)n_sel * 3 * ml);
}

inline int grid_x(long long items) {
    long long g = (items + NT * 4 - 1) / (NT * 4);
    if (g < 1) g = 1;
    if (g > 2048) g = 2048;
    return (int)g;
}

// Levels 1..3 run in the fused register kernels when both extents are multiples of 8 and the pyramid has
// at least 3 levels (and the 128-bit accesses are aligned); 0 selects the per-level kernels throughout
// (odd extents, images below 64 pixels, MDIMG_HAAR_FUSED=0).
int fused_levels(const Pyramid& p, int h, int w, const void* a, const void* b) {
    static const bool off = [] { const char* e = getenv("MDIMG_HAAR_FUSED"); return e && e[0] == '0'; }();
    if (off || p.L < 3 || (h & 7) || (w & 7)) return 0;
    if ((((uintptr_t)a) | ((uintptr_t)b)) & 15) return 0;
    return 3;
}

template <typename T, int MODE>
void run_inverse(const Pyramid& p, const Dims& d, const int* skip, const WaveBufs& b,
                 const float* coarse, const float* img_in, float* img_out, int K, Blend bl, cudaStream_t st) {
    T* r[2] = {reinterpret_cast<T*>(b.r0), reinterpret_cast<T*>(b.r1)};
    int cur = 0;
    if (K > 0) {
        for (int l = p.L; l > K; --l) {              // deeper levels: per-level kernels, never the last one
            const int hd = p.H[l], wd = p.W[l];
            dim3 grid(grid_x((long long)hd * wd), d.n_sel);
            T* rout = r[cur];
            if (l == p.L)
                MDIMG_LAUNCH k_haar_inv<T, float, MODE, false><<<grid, NT, 0, st>>>(coarse, p.a_cap, wd, hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, 2 * wd, img_in, img_out, bl);
            else
                MDIMG_LAUNCH k_haar_inv<T, T, MODE, false><<<grid, NT, 0, st>>>(r[cur ^ 1], p.r_cap, 2 * p.W[l + 1], hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, 2 * wd, img_in, img_out, bl);
            cur ^= 1;
        }
        dim3 grid(grid_x((long long)(d.h >> 3) * (d.w >> 3) * 2), d.n_sel);
        if (K == p.L)
            MDIMG_LAUNCH k_haar_inv_reg3<T, float, MODE><<<grid, NT, 0, st>>>(coarse, p.a_cap, p.W[K], d, skip, p, b.det, b.acc,
                                                                                 img_in, img_out, bl);
        else
            MDIMG_LAUNCH k_haar_inv_reg3<T, T, MODE><<<grid, NT, 0, st>>>(r[cur ^ 1], p.r_cap, 2 * p.W[K + 1], d, skip, p, b.det,
                                                                             b.acc, img_in, img_out, bl);
        return;
    }
    for (int l = p.L; l >= 1; --l) {
        const int hd = p.H[l], wd = p.W[l];
        dim3 grid(grid_x((long long)hd * wd), d.n_sel);
        const bool last = (l == 1);
        const bool first = (l == p.L);
        const int out_pitch = 2 * wd;
        T* rout = r[cur];
        if (first) {
            if (last)
                MDIMG_LAUNCH k_haar_inv<T, float, MODE, true><<<grid, NT, 0, st>>>(coarse, p.a_cap, wd, hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, out_pitch, img_in, img_out, bl);
            else
                MDIMG_LAUNCH k_haar_inv<T, float, MODE, false><<<grid, NT, 0, st>>>(coarse, p.a_cap, wd, hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, out_pitch, img_in, img_out, bl);
        } else {
            const T* ain = r[cur ^ 1];
            const int a_pitch = 2 * p.W[l + 1];
            if (last)
                MDIMG_LAUNCH k_haar_inv<T, T, MODE, true><<<grid, NT, 0, st>>>(ain, p.r_cap, a_pitch, hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, out_pitch, img_in, img_out, bl);
            else
                MDIMG_LAUNCH k_haar_inv<T, T, MODE, false><<<grid, NT, 0, st>>>(ain, p.r_cap, a_pitch, hd, wd, d, skip, l,
                    b.det, p.det_per_slice, p.off[l], b.acc, rout, p.r_cap, out_pitch, img_in, img_out, bl);
        }
        cur ^= 1;
    }
}

}  // namespace

size_t wavelet_workspace_bytes(int n, int n_sel, int h, int w) {
    Pyramid p = make_pyramid(h, w);
    Arena a(nullptr, 0);
    WaveBufs b;
    carve(a, n, n_sel, p, b);
    return a.off;
}

int wavelet_denoise_run(const float* in, float* out, const Dims& d, int mode_hard,
                        const double* sigma_in, double sigma_scale, const int* skip,
                        void* ws, size_t ws_bytes, cudaStream_t stream, float blend_c0, float blend_c1, int blend_on) {
    const Blend bl = {blend_c0, blend_c1, blend_on};
    if (d.n_sel == 0) return MDIMG_OK;
    Pyramid p = make_pyramid(d.h, d.w);
    Arena a(ws, ws_bytes);
    WaveBufs b;
    carve(a, d.n, d.n_sel, p, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "wavelet: workspace too small (%zu > %zu)", a.off, ws_bytes);
    cudaMemsetAsync(b.acc, 0, sizeof(WaveAcc) * d.n_sel, stream);
    const int want_hist = sigma_in == nullptr;
    if (want_hist) cudaMemsetAsync(b.l1, 0, (size_t)d.n_sel * SEL_L1_BINS * sizeof(unsigned), stream);

    // ---- forward ----
    float* abuf[2] = {b.a0, b.a1};
    const int K = fused_levels(p, d.h, d.w, in, out);
    if (K > 0) {
        dim3 grid(grid_x((long long)(d.h >> 3) * (d.w >> 3) * 2), d.n_sel);
        MDIMG_LAUNCH k_haar_fwd_reg3<<<grid, NT, 0, stream>>>(in, d, skip, p, abuf[(K - 1) & 1], b.det, b.acc, b.l1, want_hist);
    }
    for (int l = K + 1; l <= p.L; ++l) {
        const int h0 = p.H[l - 1], w0 = p.W[l - 1];
        dim3 grid(grid_x((long long)p.H[l] * p.W[l]), d.n_sel);
        float* aout = abuf[(l - 1) & 1];
        if (l == 1)
            MDIMG_LAUNCH k_haar_fwd<true><<<grid, NT, 0, stream>>>(in, (long long)d.h * d.w, h0, w0, d, skip, l, aout, p.a_cap,
                                                      b.det, p.det_per_slice, p.off[l], b.acc, b.l1, want_hist);
        else
            MDIMG_LAUNCH k_haar_fwd<false><<<grid, NT, 0, stream>>>(abuf[l & 1], p.a_cap, h0, w0, d, skip, l, aout, p.a_cap,
                                                       b.det, p.det_per_slice, p.off[l], b.acc, b.l1, 0);
    }
    const float* coarse = abuf[(p.L - 1) & 1];

    // ---- band energies np.sum(d*d) in numpy's pairwise order ----
    {
        const PwPlan& plan = cached_pw_plan(p);      // process-lifetime host copy: safe source for async copies
        const int n_leaves = (int)plan.leaves.size();
        if ((size_t)n_leaves > pw_max_leaves(p) || plan.max_depth > PW_MAX_ROUNDS - 2)
            return set_error(MDIMG_ERR_INVALID, "wavelet: pairwise plan exceeds its bounds (%d leaves, depth %d)", n_leaves, plan.max_depth);
        cudaMemcpyAsync(b.pw_leaves, plan.leaves.data(), sizeof(PwLeaf) * n_leaves, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(b.pw_levels, plan.levels.data(), sizeof(PwLevel) * plan.levels.size(), cudaMemcpyHostToDevice, stream);
        if (!plan.nodes.empty())
            cudaMemcpyAsync(b.pw_nodes, plan.nodes.data(), sizeof(PwNode) * plan.nodes.size(), cudaMemcpyHostToDevice, stream);
        dim3 lgrid((n_leaves * 8 + NT - 1) / NT, d.n_sel, 3);
        MDIMG_LAUNCH k_pw_leaves<<<lgrid, NT, 0, stream>>>(b.det, p.det_per_slice, b.pw_leaves, b.pw_levels, n_leaves, d, skip, b.pw_sums);
        MDIMG_LAUNCH k_pw_combine<<<dim3(p.L * 3, d.n_sel), NT, 0, stream>>>(b.pw_nodes, b.pw_levels, p.L, n_leaves, d, skip, b.pw_sums, b.acc);
    }

    // ---- sigma (finest 'dd' band) and thresholds ----
    if (want_hist) {
        const int len = p.H[1] * p.W[1];
        MDIMG_LAUNCH k_wave_ranks<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, len, b.acc, b.ranks);
        int rc = select_run(b.det + p.off[1] + 2LL * len, p.det_per_slice, len, d, 2, b.ranks, b.l1, b.med,
                            b.sel_ws, b.sel_ws_bytes, stream, SEL_COMPACT | SEL_ABS);
        if (rc) return rc;
    }
    MDIMG_LAUNCH k_wave_thresholds<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, p, b.acc, b.med, sigma_in, sigma_scale);

    // ---- inverse ----
    if (mode_hard) run_inverse<float, 2>(p, d, skip, b, coarse, in, out, K, bl, stream);
    else if (sigma_in == nullptr) run_inverse<double, 0>(p, d, skip, b, coarse, in, out, K, bl, stream);
    else run_inverse<float, 1>(p, d, skip, b, coarse, in, out, K, bl, stream);
    return check_launch("wavelet_denoise");
}

}  // namespace mdimg
