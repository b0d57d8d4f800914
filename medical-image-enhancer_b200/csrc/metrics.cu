// Quality metrics on device: the 16 no-reference metrics of the reference's compute_metrics
// (pipeline/metrics.py:42-109), estimate_sigma (skimage, called at metrics.py:47), the edge ratio
// (metrics.py:213-217) and the NIQE approximation (metrics.py:187-210).
//
// Data flow for one batch of slices (all per-slice scalars stay on the device):
//   k_stencil_stats : one read of the image -> Laplacian / Sobel / 7x7 box statistics, 256-bin
//                     histogram, level-1 radix histograms of x and |grad|, max|grad|; writes |grad|
//   k_db2_dd        : 1-level db2 'dd' band (float32 accumulation order of pywt) -> |dd|, zero count
//   select_run x3   : exact order statistics (P5/P25/P75/P95 of x, P90 of |grad|, median |dd|)
//   k_grad_prep/k_grad_pass : the three |grad| statistics that depend on max / P90 of |grad|
//   k_finalize      : the 16 numbers (+ mean, edge ratio, NIQE) per slice
#include "metrics.cuh"
#include "boxfilter.cuh"
#include <cstdlib>

namespace mdimg {

namespace {

constexpr int TW = 64, TH = 32, HALO = 3;
constexpr int NT = 256;

struct MetAcc {               // per slice, zero-initialised
    double sum_x, sum_x2, sum_lap, sum_lap2, sum_abslap, sum_g, sum_g2, sum_ls, sum_ls2;
    double sum_strong;
    double box16[2];          // sum lv, sum lv^2 (NIQE)
    unsigned long long cnt_low, cnt_high, cnt_edge, cnt_strong;
    unsigned gmax_key;
    unsigned dd_zero;         // number of exactly-zero 'dd' coefficients
    unsigned hist256[256];
    unsigned hist128[128];
};

struct GradPrep {             // per slice, derived from max|grad| and P90
    float edges[129];
    float denom;              // float32(last_edge)
    float last;               // float32(last_edge) for the keep test
    float thr_edge;           // 0.1f * gmax
    float t90;                // np.percentile(grad, 90)
};

__device__ __forceinline__ float lerp_np(float a, float b, float t) {
    // numpy _lerp in float32: a + (b-a)*t, or b - (b-a)*(1-t) when t >= 0.5
    float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, t));
    if (t >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    return r;
}

// ---------------------------------------------------------------------------------------
// Fused stencil statistics (one read of the image).
//
// Phases per 64x32 tile (256 threads):
//   A  load tile + 3-pixel halo (half-sample symmetric border) into shared memory as doubles
//      (x and the float32-rounded x*x): every later phase works on exact doubles, conversions
//      happen once per loaded pixel;
//   B  7-tap box mean along axis 0 (sliding window, boxfilter.cuh);
//   C  7-tap box mean along axis 1 -> local std -> sum / sum of squares (nothing is stored);
//   D  Laplacian + Sobel pair from a 3x3 register window sliding down 4 rows per thread,
//      |grad| written to HBM (coalesced), clip counters, 256-bin histogram and the level-1 radix
//      histograms of x and |grad| (warp-uniform fast path for flat regions such as CT air).
// ---------------------------------------------------------------------------------------
typedef BoxTile<7> B7;

struct StencilSmem {
    float X[B7::XH * B7::XP];
    float VS[B7::TH * B7::XP];
    float VQ[B7::TH * B7::XP];
    unsigned h256[256];
    unsigned hx[SEL_L1_BINS];
    unsigned hg[SEL_L1_BINS];
    double red[9 * 32];
};

// Shared-memory histogram increment; all 32 lanes must call.  Flat regions (every valid lane in
// the same bin) cost one atomic per warp instead of a 32-way same-address serialisation.
__device__ __forceinline__ void hist_add(unsigned* h, int bin, bool valid, int lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
    if (mask == 0) return;
    const int leader = __ffs(mask) - 1;
    const int b0 = __shfl_sync(0xffffffffu, bin, leader);
    const bool same = !valid || bin == b0;
    if (__all_sync(0xffffffffu, same)) {
        if (lane == leader) atomicAdd(&h[b0], (unsigned)__popc(mask));
    } else if (valid) {
        atomicAdd(&h[bin], 1u);
    }
}

// Three histograms of the same pixel with ONE uniformity vote: in flat regions (every valid lane in
// the same three bins) a warp issues three atomics in total; otherwise each lane adds its own.
// v0 / v1 / v2: per-histogram validity (v1 and v2 are the pixel's validity, v0 adds the range test).
__device__ __forceinline__ void hist_add3(unsigned* h0, int b0, bool v0, unsigned* h1, int b1, unsigned* h2, int b2,
                                          bool valid, int lane) {
    const unsigned mask = __ballot_sync(0xffffffffu, valid);
    if (mask == 0) return;
    const int leader = __ffs(mask) - 1;
    // pack (b0 | invalid flag, b1, b2) so that one shuffle broadcasts the leader's bins
    const unsigned packed = ((v0 ? (unsigned)b0 : 0x3ffu) << 20) | ((unsigned)b1 << 10) | (unsigned)b2;
    const unsigned lead = __shfl_sync(0xffffffffu, packed, leader);
    if (__all_sync(0xffffffffu, !valid || packed == lead)) {
        if (lane == leader) {
            const unsigned n = (unsigned)__popc(mask);
            if ((lead >> 20) != 0x3ffu) atomicAdd(&h0[lead >> 20], n);
            atomicAdd(&h1[(lead >> 10) & 0x3ffu], n);
            atomicAdd(&h2[lead & 0x3ffu], n);
        }
    } else if (valid) {
        if (v0) atomicAdd(&h0[b0], 1u);
        atomicAdd(&h1[b1], 1u);
        atomicAdd(&h2[b2], 1u);
    }
}

// The same for FOUR pixels per lane (the strip kernel's thread owns four rows of a column): one uniformity vote
// for the 128 pixels of the warp.  pk[k]: the three bins of pixel k packed as in hist_add3 (b0 field 0x3ff when the
// value is outside the 256-bin range); valid[k]: pixel k is inside the image.  Flat regions (CT air) cost three
// atomics per 128 pixels; a warp with any partly valid or non-uniform lane adds its pixels one by one.
__device__ __forceinline__ void hist_add3x4(unsigned* h0, unsigned* h1, unsigned* h2, const unsigned (&pk)[4],
                                            const bool (&valid)[4], int lane) {
    const bool any = valid[0] || valid[1] || valid[2] || valid[3];
    const bool full = valid[0] && valid[1] && valid[2] && valid[3] && pk[0] == pk[1] && pk[1] == pk[2] && pk[2] == pk[3];
    const unsigned m_any = __ballot_sync(0xffffffffu, any);
    if (m_any == 0) return;
    const unsigned m_full = __ballot_sync(0xffffffffu, full);
    bool uniform = m_full == m_any;
    unsigned lead = 0;
    if (uniform) {
        const int leader = __ffs(m_full) - 1;
        lead = __shfl_sync(0xffffffffu, pk[0], leader);
        uniform = __all_sync(0xffffffffu, !full || pk[0] == lead);
        if (uniform) {
            if (lane == leader) {
                const unsigned n = 4u * (unsigned)__popc(m_full);
                if ((lead >> 20) != 0x3ffu) atomicAdd(&h0[lead >> 20], n);
                atomicAdd(&h1[(lead >> 10) & 0x3ffu], n);
                atomicAdd(&h2[lead & 0x3ffu], n);
            }
            return;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (valid[k]) {
            if ((pk[k] >> 20) != 0x3ffu) atomicAdd(&h0[pk[k] >> 20], 1u);
            atomicAdd(&h1[(pk[k] >> 10) & 0x3ffu], 1u);
            atomicAdd(&h2[pk[k] & 0x3ffu], 1u);
        }
}

__global__ void __launch_bounds__(NT, 4)
k_stencil_stats(const float* __restrict__ img, Dims d, MetAcc* __restrict__ acc,
                float* __restrict__ gout, unsigned* __restrict__ l1x, unsigned* __restrict__ l1g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StencilSmem& sm = *reinterpret_cast<StencilSmem*>(smem_raw);
    constexpr int XW = B7::XW, XH = B7::XH, XP = B7::XP;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const float* src = img + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < 256; i += NT) sm.h256[i] = 0;
    for (int i = tid; i < SEL_L1_BINS; i += NT) { sm.hx[i] = 0; sm.hg[i] = 0; }

    // ---- A: tile + halo ----
    load_tile<XW, XH, HALO, 0>(src, d.h, d.w, x0, y0, [&](int r, int c, float v) { sm.X[r * XP + c] = v; });
    __syncthreads();

    // ---- B: axis-0 box means ----
    const double inv7 = 1.0 / 7.0;
    box_vertical_xq<7>(sm.X, sm.VS, sm.VQ, inv7);
    __syncthreads();

    double acc_v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};

    // ---- C: axis-1 box means -> local std ----
    {
        float* const vin[2] = {sm.VS, sm.VQ};
        double ls1 = 0.0, ls2 = 0.0;
        box_horizontal_f<7, 2>(vin, inv7, [&](int r, int c, const float (&m)[2]) {
            if (y0 + r < d.h && x0 + c < d.w) {
                const float lv = fmaxf(__fsub_rn(m[1], __fmul_rn(m[0], m[0])), 0.0f);
                const double ls = (double)sqrt_rn_fast(lv);
                ls1 += ls;
                ls2 += ls * ls;
            }
        });
        acc_v[7] = ls1;
        acc_v[8] = ls2;
    }

    // ---- D: 3x3 stencils ----
    unsigned c_low = 0, c_high = 0;
    float gmax = 0.0f;
    float f_lap = 0.0f, f_lap2 = 0.0f, f_abs = 0.0f, f_g = 0.0f, f_g2 = 0.0f;
    double d_x = 0.0, d_x2 = 0.0;
    float* gdst = gout + (size_t)s * d.h * d.w;
#pragma unroll
    for (int i = 0; i < TW / 32; ++i) {
        const int c = lane + 32 * i;
        const int cc = c + HALO;
        const int gx = x0 + c;
        const int rbase = wid * 4;
        const float* col = sm.X + (rbase + HALO - 1) * XP + cc;
        double u0 = col[-1], u1 = col[0], u2 = col[1];
        double m0 = col[XP - 1], m1 = col[XP], m2 = col[XP + 1];
        float xc = col[XP];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = rbase + j;
            const int gy = y0 + r;
            const float* dn = sm.X + (r + HALO + 1) * XP + cc;
            const float xn = dn[0];
            const double n0 = dn[-1], n1 = xn, n2 = dn[1];
            const bool valid = gy < d.h && gx < d.w;
            // scipy.ndimage.convolve: exact double accumulation, one rounding to float32
            const float lap = (float)(4.0 * m1 - u1 - m0 - m2 - n1);
            const float sh = (float)(0.25 * (u0 - n0) + 0.5 * (u1 - n1) + 0.25 * (u2 - n2));
            const float sv = (float)(0.25 * (u0 - u2) + 0.5 * (m0 - m2) + 0.25 * (n0 - n2));
            const float g = sqrt_rn_fast(__fadd_rn(__fmul_rn(sh, sh), __fmul_rn(sv, sv)));
            int b256 = (int)(xc * 256.0f);
            b256 = b256 > 255 ? 255 : b256;
            if (valid) {
                gdst[(size_t)gy * d.w + gx] = g;
                gmax = fmaxf(gmax, g);
                d_x += m1;
                d_x2 += m1 * m1;
                f_lap += lap;
                f_lap2 = fmaf(lap, lap, f_lap2);
                f_abs += fabsf(lap);
                f_g += g;
                f_g2 = fmaf(g, g, f_g2);
                c_low += (xc <= 0.01f);
                c_high += (xc >= 0.99f);
            }
            // np.histogram(bins=256, range=(0,1)): exact power-of-two edges, out-of-range dropped
            static_assert(SEL_L1_BINS <= 1024, "bins are packed into 10-bit fields");
            hist_add3(sm.h256, b256, valid && xc >= 0.0f && xc <= 1.0f, sm.hx, sel_bin1(xc), sm.hg, sel_bin1(g),
                      valid, lane);
            u0 = m0; u1 = m1; u2 = m2;
            m0 = n0; m1 = n1; m2 = n2;
            xc = xn;
        }
    }
    acc_v[0] = d_x; acc_v[1] = d_x2;
    acc_v[2] = (double)f_lap; acc_v[3] = (double)f_lap2; acc_v[4] = (double)f_abs;
    acc_v[5] = (double)f_g; acc_v[6] = (double)f_g2;
    __syncthreads();

    block_sum<9>(acc_v, sm.red);
    MetAcc* A = acc + si;
    unsigned cl = warp_sum_u(c_low), ch = warp_sum_u(c_high);
    float gm = warp_max(gmax);
    if (lane == 0) {
        if (cl) atomicAdd(&A->cnt_low, (unsigned long long)cl);
        if (ch) atomicAdd(&A->cnt_high, (unsigned long long)ch);
        atomicMax(&A->gmax_key, f2key(gm));
    }
    if (tid == 0) {
        atomicAdd(&A->sum_x, acc_v[0]); atomicAdd(&A->sum_x2, acc_v[1]);
        atomicAdd(&A->sum_lap, acc_v[2]); atomicAdd(&A->sum_lap2, acc_v[3]);
        atomicAdd(&A->sum_abslap, acc_v[4]); atomicAdd(&A->sum_g, acc_v[5]);
        atomicAdd(&A->sum_g2, acc_v[6]); atomicAdd(&A->sum_ls, acc_v[7]);
        atomicAdd(&A->sum_ls2, acc_v[8]);
    }
    for (int i = tid; i < 256; i += NT) { unsigned v = sm.h256[i]; if (v) atomicAdd(&A->hist256[i], v); }
    unsigned* gx_ = l1x + (size_t)si * SEL_L1_BINS;
    unsigned* gg_ = l1g + (size_t)si * SEL_L1_BINS;
    for (int i = tid; i < SEL_L1_BINS; i += NT) {
        unsigned v = sm.hx[i]; if (v) atomicAdd(&gx_[i], v);
        unsigned u = sm.hg[i]; if (u) atomicAdd(&gg_[i], u);
    }
}

// ---------------------------------------------------------------------------------------
// Strip kernel: ONE read of the image for the Laplacian / Sobel statistics, |grad| (written), the
// 7x7 box local std (metrics.py:120-129), the 16x16 box local variance of the NIQE approximation
// (metrics.py:195-200), clip counters, the 256-bin histogram and the level-1 select histograms.
//
// A block owns a strip of SW = 128 output columns (+ 8 / 7 halo columns: the 16-wide window
// [i-8, i+7] bounds every other halo) and marches down a band of rows, SR = 8 rows per step:
//   load   rows enter a 24-row ring in shared memory as DOUBLES (x and the float32-rounded x*x):
//          one conversion per loaded pixel, none in any window below; the next step's rows are
//          prefetched into registers while this step computes;
//   V      axis-0 box sums: 143 threads slide the 16-row window, 134 threads the 7-row window down
//          their column; the running sums live in registers for the whole band, so no window is ever
//          warmed up twice.  scipy stores the axis-0 result as float32: the double is rounded to
//          float32 precision ON THE FP64 PIPE ((v + C) - C with C = 1.5 * 2^(exponent + 29)), not
//          through two conversions, and stays a double in shared memory;
//   H      axis-1 box sums: (row, 8-column segment) threads slide along their segment; only the final
//          means become float32 (scipy's output dtype), followed by the float32 variance arithmetic;
//   S      3x3 stencils from the ring in double (scipy.ndimage.convolve accumulates in double and
//          rounds once), histograms, counters, |grad| store.
// Conversions issue on the 16-lane XU pipe, which bounded the previous tile kernels (XU 60-80 %
// busy, ~50 conversions per pixel for the same work); here ~13 per pixel remain.
// ---------------------------------------------------------------------------------------
namespace strip {

constexpr int NT = 320;               // 10 warps: axis-0 tasks on 5 + 5 warps (143 + 134 columns), axis-1 and stencils on 8
constexpr int SW = 128;               // output columns per strip
constexpr int HL = 8, HR = 7;         // halo of the 16-wide window [i-8, i+7]
constexpr int SWP = SW + HL + HR;     // 143 input columns
constexpr int P = 145;                // row pitch in doubles, = 1 (mod 16): conflict-free (row, segment) walks
constexpr int SR = 8;                 // rows per step
constexpr int RING = 3 * SR;          // ring rows: load blocks B, B+1, B+2 serve output block B
constexpr int BLK = SR * P;           // doubles per ring block
constexpr int NV16 = SWP;             // axis-0 tasks of the 16-window: every input column (threads 0..142)
constexpr int NV7 = SW + 6;           // axis-0 tasks of the 7-window: input columns 5 .. SW+10 (threads 160..293)
constexpr int SEG = 8;                // axis-1: outputs per (row, segment) task
constexpr int NACC = 11;
constexpr int STAGE_P = 144;            // floats per staged row (SWP rounded up to a 16-byte multiple)

struct Smem {
    double Xs[RING * P];              // x
    double Xq[RING * P];              // float32(x * x)
    double V[4][SR * P];              // axis-0 means: 16-window of x, of x*x; 7-window of x, of x*x
    unsigned h256[256];
    unsigned hx[SEL_L1_BINS];
    unsigned hg[SEL_L1_BINS];
    double red[NACC * 32];
    float stage[SR * STAGE_P];        // bulk-copy landing zone: the 8 float32 rows of the load block in flight
    unsigned long long mbar;          // mbarrier the bulk copies complete on
};

// ---- bulk asynchronous copies (TMA engine, cp.async.bulk) of whole row segments into shared memory ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    const unsigned a = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}

// v rounded to float32 precision, as a double.  FAST: v is 0 or a normal float32-range magnitude
// (guaranteed by the caller); the sum lands in the binade whose double spacing is the float32
// spacing of v, where the FP64 adder's round-to-nearest-even is float32's.
template <bool FAST>
__device__ __forceinline__ double round_to_f32(double v) {
    if (FAST) {
        const int e = __double2hiint(v) & 0x7ff00000;
        const double c = __hiloint2double(e + 0x01d80000, 0);     // 1.5 * 2^(exponent + 29)
        return __dsub_rn(__dadd_rn(v, c), c);
    }
    return (double)(float)v;
}

// float32 bits of a double that holds a float32 value (0 or normal float32 range, any sign)
__device__ __forceinline__ float f32_bits_of(double v) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mag = hi & 0x7fffffffu;
    const unsigned b = __funnelshift_l(lo, mag - 0x38000000u, 3) | (hi & 0x80000000u);
    return __uint_as_float(mag ? b : (hi & 0x80000000u));
}

// level-1 select bin and 256-bin index without a float->int conversion (same values as sel_bin1 /
// int(x * 256): the add with 2^23 in round-up / round-down mode leaves ceil / floor in the mantissa)
__device__ __forceinline__ int bin1_fp(float v) {
    const float t = fminf(fmaxf(__fmul_rn(v, 1021.0f), 0.0f), 2048.0f);
    const int c = __float_as_int(__fadd_ru(t, 8388608.0f)) & 0x7fffff;
    const int b = 1 + min(c, 1022);
    return v < 0.0f ? 0 : b;
}
__device__ __forceinline__ int bin256_fp(float x) {
    const float t = fminf(fmaxf(__fmul_rn(x, 256.0f), 0.0f), 1024.0f);
    const int c = __float_as_int(__fadd_rd(t, 8388608.0f)) & 0x7fffff;
    return min(c, 255);
}

// Offsets (in doubles) of ring rows relative to the first row of output block B: row r in [0, 24)
// lives in load block B + r/8, whose ring block offset is o[r/8].
struct RingOff { int o0, o1, o2; };
__device__ __forceinline__ int ring_at(const RingOff& o, int r) {      // r is a compile-time constant after unrolling
    return (r < SR ? o.o0 : (r < 2 * SR ? o.o1 : o.o2)) + (r & (SR - 1)) * P;
}

// ---- V: one step of the axis-0 window K of this thread's column vc ----
template <int K, bool FAST>
__device__ __forceinline__ void vertical(Smem& sm, const RingOff& o, bool first, int vc, double& vs, double& vq) {
    constexpr int lead = K == 16 ? 15 : 11;      // row entering the window, relative to the output row
    constexpr int tail = K == 16 ? 0 : 5;        // row leaving it after the output
    const double inv = K == 16 ? 0.0625 : 1.0 / 7.0;
    double* vS = sm.V[K == 16 ? 0 : 2] + vc;
    double* vQ = sm.V[K == 16 ? 1 : 3] + vc;
    const double* xs = sm.Xs + vc;
    const double* xq = sm.Xq + vc;
    if (first) {                                 // fresh window sums at the top of the band
        vs = 0.0; vq = 0.0;
#pragma unroll
        for (int r = tail; r < lead; ++r) {
            vs += xs[ring_at(o, r)];
            vq += xq[ring_at(o, r)];
        }
    }
#pragma unroll
    for (int j = 0; j < SR; ++j) {
        const int ra = ring_at(o, j + lead);
        vs += xs[ra];
        vq += xq[ra];
        vS[j * P] = round_to_f32<FAST>(vs * inv);
        vQ[j * P] = round_to_f32<FAST>(vq * inv);
        const int rt = ring_at(o, j + tail);
        vs -= xs[rt];
        vq -= xq[rt];
    }
}

// ---- H: the SEG outputs (row j, columns SEG*g ..) of window K; f(i, mean of x, mean of x*x), float32 means ----
template <int K, typename F>
__device__ __forceinline__ void horizontal(const double* vS, const double* vQ, int j, int g, F&& f) {
    constexpr int lead = K == 16 ? 15 : 11;
    constexpr int tail = K == 16 ? 0 : 5;
    const double inv = K == 16 ? 0.0625 : 1.0 / 7.0;
    const double* rs = vS + j * P + SEG * g;
    const double* rq = vQ + j * P + SEG * g;
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = tail; k < lead; ++k) { s += rs[k]; q += rq[k]; }
#pragma unroll
    for (int i = 0; i < SEG; ++i) {
        s += rs[i + lead];
        q += rq[i + lead];
        f(i, (float)(s * inv), (float)(q * inv));
        s -= rs[i + tail];
        q -= rq[i + tail];
    }
}

template <bool FULL>
__global__ void __launch_bounds__(NT, 2)
k_strip_stats(const float* __restrict__ img, Dims d, int band_h, int want16, int bulk, MetAcc* __restrict__ acc,
              float* __restrict__ gout, unsigned* __restrict__ l1x, unsigned* __restrict__ l1g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    __shared__ int slow_s;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int strips = (d.w + SW - 1) / SW;
    const int strip_i = blockIdx.x % strips, band_i = blockIdx.x / strips;
    const int x0 = strip_i * SW, yb0 = band_i * band_h;
    const int bh = min(band_h, d.h - yb0);
    const int nsteps = (bh + SR - 1) / SR;
    const float* src = img + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    if (FULL) {
        for (int i = tid; i < 256; i += NT) sm.h256[i] = 0;
        for (int i = tid; i < SEL_L1_BINS; i += NT) { sm.hx[i] = 0; sm.hg[i] = 0; }
    }
    if (tid == 0) {
        slow_s = 0;
        if (bulk) mbar_init(&sm.mbar, SR);        // one arrival (+ its bytes) per loader warp and load block
    }
    __syncthreads();

    // ---- loader: warp w < 8 owns row w of a load block, a lane its columns lane + 32 k ----
    // bulk path (row pitch and bases 16-byte aligned): lane 0 of the warp hands the row's in-image column
    // range [cb, ce) to the TMA engine as ONE cp.async.bulk into the staging row; only the <= 8 reflected
    // border columns of the first / last strip are fetched with ordinary loads.  The copy of load block
    // B + 3 is in flight during the whole step B; stash() waits for it on the mbarrier.
    const int cb = x0 == 0 ? HL : 0;                                   // first staged input column
    const int ce = min(STAGE_P, d.w - (x0 - HL));                      // end of the staged range (multiple of 4)
    unsigned phase = 0;
    constexpr int NC = (SWP + 31) / 32;           // 5
    int gx[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) gx[k] = refl_sym_fast(x0 - HL + lane + 32 * k, d.w);
    float pf[NC];
    bool bad = false;
    auto fetch = [&](int lb) {                    // load block lb -> staging row (bulk) / registers
        if (wid < SR) {
            const float* row = src + (size_t)refl_sym_fast(yb0 - HL + lb * SR + wid, d.h) * d.w;
            if (bulk) {
                if (lane == 0) {
                    const unsigned bytes = (unsigned)(ce - cb) * 4u;
                    mbar_expect_tx(&sm.mbar, bytes);
                    bulk_g2s(sm.stage + wid * STAGE_P + cb, row + (x0 - HL + cb), bytes, &sm.mbar);
                }
#pragma unroll
                for (int k = 0; k < NC; ++k) {
                    const int c = lane + 32 * k;
                    if (c < SWP && (c < cb || c >= ce)) pf[k] = row[gx[k]];
                }
            } else {
#pragma unroll
                for (int k = 0; k < NC; ++k)
                    if (lane + 32 * k < SWP) pf[k] = row[gx[k]];
            }
        }
    };
    auto stash = [&](int off) {                   // staging row / registers -> the ring block at offset `off`, as doubles
        if (wid < SR) {
            if (bulk) {
                mbar_wait(&sm.mbar, phase);
                phase ^= 1u;
            }
            const int r = off + wid * P + lane;
#pragma unroll
            for (int k = 0; k < NC; ++k)
                if (lane + 32 * k < SWP) {
                    const int c = lane + 32 * k;
                    const float v = (bulk && c >= cb && c < ce) ? sm.stage[wid * STAGE_P + c] : pf[k];
                    // the FP64-pipe rounding needs 0 or a normal magnitude whose square is normal too
                    const unsigned u = __float_as_uint(v);
                    bad |= (u != 0u) && (u - 0x21800000u >= 0x3c000000u);      // outside [2^-60, 2^60) or negative
                    sm.Xs[r + 32 * k] = (double)v;
                    sm.Xq[r + 32 * k] = (double)__fmul_rn(v, v);
                }
            __syncwarp();                         // the staging row is free again: this warp's next copy may land
        }
    };

    double vs = 0.0, vq = 0.0;                    // axis-0 window sums of this thread's column (whole band)
    double a_x = 0.0, a_x2 = 0.0, a_lap = 0.0, a_lap2 = 0.0, a_abs = 0.0, a_g = 0.0, a_g2 = 0.0;
    double a_b1 = 0.0, a_b2 = 0.0;                // axis-1 role's pair: (sum ls, sum ls^2) or (sum lv, sum lv^2)
    unsigned c_low = 0, c_high = 0;
    float gmax = 0.0f;
    float* gdst = FULL ? gout + (size_t)s * d.h * d.w : nullptr;

    RingOff o = {0, BLK, 2 * BLK};
    fetch(0); stash(o.o0);
    fetch(1); stash(o.o1);
    fetch(2);
    bool slow = false;
    for (int B = 0; B < nsteps; ++B) {
        stash(o.o2);
        if (B + 1 < nsteps) fetch(B + 3);
        if (bad && !slow) slow_s = 1;             // sticky for the rest of the band
        __syncthreads();
        slow = slow_s != 0;

        // ---- V: warps 0-4 the 16-window (143 columns), warps 5-9 the 7-window (134 columns) ----
        if (wid < 5) {
            if (want16 && tid < NV16) {
                if (!slow) vertical<16, true>(sm, o, B == 0, tid, vs, vq);
                else vertical<16, false>(sm, o, B == 0, tid, vs, vq);
            }
        } else if (FULL && tid - 160 < NV7) {
            if (!slow) vertical<7, true>(sm, o, B == 0, 5 + tid - 160, vs, vq);
            else vertical<7, false>(sm, o, B == 0, 5 + tid - 160, vs, vq);
        }
        __syncthreads();

        if (wid < 8) {
            // ---- H: (row, 8-column segment) tasks; warps 0-3 the 16-window, warps 4-7 the 7-window ----
            {
                const int u = tid & 127;
                const int j = u & 7, g = u >> 3;
                const bool rowok = B * SR + j < bh;
                const int xlim = d.w - x0 - SEG * g;           // outputs i < xlim are inside the image
                if (wid < 4) {
                    if (want16)
                        horizontal<16>(sm.V[0], sm.V[1], j, g, [&](int i, float m, float q) {
                            const float lv = fmaxf(__fsub_rn(q, __fmul_rn(m, m)), 0.0f);
                            const double dl = (rowok && i < xlim) ? (double)lv : 0.0;
                            a_b1 += dl;
                            a_b2 = fma(dl, dl, a_b2);
                        });
                } else if (FULL) {
                    horizontal<7>(sm.V[2], sm.V[3], j, g, [&](int i, float m, float q) {
                        const float lv = fmaxf(__fsub_rn(q, __fmul_rn(m, m)), 0.0f);
                        const double dl = (rowok && i < xlim) ? (double)sqrt_rn_fast(lv) : 0.0;
                        a_b1 += dl;
                        a_b2 = fma(dl, dl, a_b2);
                    });
                }
            }
            // ---- S: 3x3 stencils, a thread owns 4 rows of one column ----
            const int col = tid & 127, rg = tid >> 7;
            const int cc = col + HL;
            const int gxo = x0 + col;
            const bool colok = gxo < d.w;
            const double* xs = sm.Xs + cc;
            // ring rows 7 + 4 rg .. : the row above this thread's first output row, ...
            double u0, u1, u2, m0, m1, m2;
            if (rg == 0) {
                u0 = xs[ring_at(o, 7) - 1]; u1 = xs[ring_at(o, 7)]; u2 = xs[ring_at(o, 7) + 1];
                m0 = xs[ring_at(o, 8) - 1]; m1 = xs[ring_at(o, 8)]; m2 = xs[ring_at(o, 8) + 1];
            } else {
                u0 = xs[ring_at(o, 11) - 1]; u1 = xs[ring_at(o, 11)]; u2 = xs[ring_at(o, 11) + 1];
                m0 = xs[ring_at(o, 12) - 1]; m1 = xs[ring_at(o, 12)]; m2 = xs[ring_at(o, 12) + 1];
            }
            float f_lap = 0.0f, f_lap2 = 0.0f, f_abs = 0.0f, f_g = 0.0f, f_g2 = 0.0f;
            unsigned pk[4];
            bool pv[4];
            const int yr0 = B * SR + 4 * rg;
            float* gp = FULL ? gdst + (size_t)(yb0 + yr0) * d.w + gxo : nullptr;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int rn = rg == 0 ? ring_at(o, 9 + jj) : ring_at(o, 13 + jj);
                const double n0 = xs[rn - 1], n1 = xs[rn], n2 = xs[rn + 1];
                const bool valid = colok && yr0 + jj < bh;
                // scipy.ndimage.convolve: exact double accumulation, one rounding to float32
                const float lap = (float)(4.0 * m1 - u1 - m0 - m2 - n1);
                const float sh = (float)(0.25 * (u0 - n0) + 0.5 * (u1 - n1) + 0.25 * (u2 - n2));
                const float sv = (float)(0.25 * (u0 - u2) + 0.5 * (m0 - m2) + 0.25 * (n0 - n2));
                const float g = sqrt_rn_fast(__fadd_rn(__fmul_rn(sh, sh), __fmul_rn(sv, sv)));
                const float lapv = valid ? lap : 0.0f, gv = valid ? g : 0.0f;
                f_lap += lapv;
                f_lap2 = fmaf(lapv, lapv, f_lap2);
                f_abs += fabsf(lapv);
                f_g += gv;
                f_g2 = fmaf(gv, gv, f_g2);
                if (FULL) {
                    const float xc = slow ? (float)m1 : f32_bits_of(m1);
                    const double xv = valid ? m1 : 0.0;
                    if (valid) gp[(size_t)jj * d.w] = g;
                    gmax = fmaxf(gmax, gv);
                    a_x += xv;
                    a_x2 = fma(xv, xv, a_x2);
                    c_low += (valid && xc <= 0.01f);
                    c_high += (valid && xc >= 0.99f);
                    const bool in01 = xc >= 0.0f && xc <= 1.0f;       // np.histogram drops values outside [0, 1]
                    pk[jj] = ((in01 ? (unsigned)bin256_fp(xc) : 0x3ffu) << 20) | ((unsigned)bin1_fp(xc) << 10) |
                             (unsigned)bin1_fp(g);
                    pv[jj] = valid;
                }
                u0 = m0; u1 = m1; u2 = m2;
                m0 = n0; m1 = n1; m2 = n2;
            }
            if (FULL) hist_add3x4(sm.h256, sm.hx, sm.hg, pk, pv, lane);
            a_lap += (double)f_lap; a_lap2 += (double)f_lap2; a_abs += (double)f_abs;
            a_g += (double)f_g; a_g2 += (double)f_g2;
        }
        __syncthreads();
        const int t = o.o0; o.o0 = o.o1; o.o1 = o.o2; o.o2 = t;
    }

    const bool k16 = wid < 4;
    double v[NACC] = {a_x, a_x2, a_lap, a_lap2, a_abs, a_g, a_g2, k16 ? 0.0 : a_b1, k16 ? 0.0 : a_b2,
                      k16 ? a_b1 : 0.0, k16 ? a_b2 : 0.0};
    block_sum<NACC>(v, sm.red);
    MetAcc* A = acc + si;
    if (FULL) {
        const unsigned cl = warp_sum_u(c_low), ch = warp_sum_u(c_high);
        const float gm = warp_max(gmax);
        if (lane == 0) {
            if (cl) atomicAdd(&A->cnt_low, (unsigned long long)cl);
            if (ch) atomicAdd(&A->cnt_high, (unsigned long long)ch);
            atomicMax(&A->gmax_key, f2key(gm));
        }
    }
    if (tid == 0) {
        if (FULL) {
            atomicAdd(&A->sum_x, v[0]); atomicAdd(&A->sum_x2, v[1]);
            atomicAdd(&A->sum_lap, v[2]); atomicAdd(&A->sum_lap2, v[3]);
            atomicAdd(&A->sum_g2, v[6]); atomicAdd(&A->sum_ls, v[7]); atomicAdd(&A->sum_ls2, v[8]);
        }
        atomicAdd(&A->sum_abslap, v[4]); atomicAdd(&A->sum_g, v[5]);
        if (want16) { atomicAdd(&A->box16[0], v[9]); atomicAdd(&A->box16[1], v[10]); }
    }
    if (FULL) {
        for (int i = tid; i < 256; i += NT) { unsigned c = sm.h256[i]; if (c) atomicAdd(&A->hist256[i], c); }
        unsigned* gx_ = l1x + (size_t)si * SEL_L1_BINS;
        unsigned* gg_ = l1g + (size_t)si * SEL_L1_BINS;
        for (int i = tid; i < SEL_L1_BINS; i += NT) {
            unsigned c = sm.hx[i]; if (c) atomicAdd(&gx_[i], c);
            unsigned u = sm.hg[i]; if (u) atomicAdd(&gg_[i], u);
        }
    }
}

// rows per band: whole strips when the batch alone fills the machine, shorter bands (15 warm-up
// rows each) for small batches
inline int band_rows(int n_sel, int h, int w) {
    const int strips = (w + SW - 1) / SW;
    const long long want = 4LL * 148 * 2;                  // ~4 waves of 2 blocks per SM
    long long bands = (want + (long long)strips * n_sel - 1) / ((long long)strips * n_sel);
    const int max_bands = (h + 63) / 64;                   // at least 64 rows per band
    if (bands > max_bands) bands = max_bands;
    if (bands < 1) bands = 1;
    int bh = (int)((h + bands - 1) / bands);
    bh = (bh + SR - 1) / SR * SR;
    return bh;
}

template <bool FULL>
void launch(const float* img, const Dims& d, int want16, MetAcc* acc, float* gout, unsigned* l1x, unsigned* l1g,
            cudaStream_t stream) {
    static unsigned long long devices_done = 0;
    opt_in_shared_memory(k_strip_stats<FULL>, sizeof(Smem), devices_done);
    const int bh = band_rows(d.n_sel, d.h, d.w);
    const int strips = (d.w + SW - 1) / SW, bands = (d.h + bh - 1) / bh;
    // Bulk (TMA engine) row copies need 16-byte aligned row segments: row pitch a multiple of 4 pixels, aligned
    // base.  Measured on B200 (512 slices): mdimg_quality 1.25 ms with them against 1.09 ms with the register
    // prefetch, mdimg_metrics 3.69 against 3.55 -- a 576-byte copy per row is too small a unit for the engine and
    // the staging row adds a shared-memory round trip -- so the path is opt-in (MDIMG_STRIP_BULK=1).
    static const bool bulk_on = [] { const char* e = getenv("MDIMG_STRIP_BULK"); return e && e[0] == '1'; }();
    const int bulk = (bulk_on && (d.w & 3) == 0 && (((uintptr_t)img) & 15) == 0 && d.w >= 16) ? 1 : 0;
    MDIMG_LAUNCH k_strip_stats<FULL><<<dim3(strips * bands, d.n_sel), NT, sizeof(Smem), stream>>>(
        img, d, bh, want16, bulk, acc, gout, l1x, l1g);
}

}  // namespace strip

// Light variant for the halo guard / NIQE: only sum|laplace| and sum|grad|.
constexpr int EW = 64, EH = 32, EXW = EW + 2, EXH = EH + 2, EXP = EXW + 1;

__global__ void __launch_bounds__(NT)
k_edge_stats(const float* __restrict__ img, Dims d, double* __restrict__ acc2) {
    __shared__ float X[EXH][EXP];
    __shared__ double red[2 * 32];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tiles_x = (d.w + EW - 1) / EW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * EW, y0 = ty * EH;
    const float* src = img + (size_t)s * d.h * d.w;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < EXH * EXW; i += NT) {
        int r = i / EXW, c = i - r * EXW;
        int gy = refl_sym(y0 + r - 1, d.h), gx = refl_sym(x0 + c - 1, d.w);
        X[r][c] = src[(size_t)gy * d.w + gx];
    }
    __syncthreads();
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < EH / 8; ++j)
#pragma unroll
        for (int i = 0; i < EW / 32; ++i) {
            const int r = wid + 8 * j, c = lane + 32 * i;
            if (y0 + r < d.h && x0 + c < d.w) {
                const int rr = r + 1, cc = c + 1;
                const double xc = X[rr][cc];
                const double n00 = X[rr - 1][cc - 1], n01 = X[rr - 1][cc], n02 = X[rr - 1][cc + 1];
                const double n10 = X[rr][cc - 1], n12 = X[rr][cc + 1];
                const double n20 = X[rr + 1][cc - 1], n21 = X[rr + 1][cc], n22 = X[rr + 1][cc + 1];
                const float lap = (float)(4.0 * xc - n01 - n10 - n12 - n21);
                const float sh = (float)(0.25 * (n00 - n20) + 0.5 * (n01 - n21) + 0.25 * (n02 - n22));
                const float sv = (float)(0.25 * (n00 - n02) + 0.5 * (n10 - n12) + 0.25 * (n20 - n22));
                const float g = sqrt_rn_fast(__fadd_rn(__fmul_rn(sh, sh), __fmul_rn(sv, sv)));
                v[0] += (double)fabsf(lap);
                v[1] += (double)g;
            }
        }
    block_sum<2>(v, red);
    if (tid == 0) {
        atomicAdd(&acc2[(size_t)si * 2 + 0], v[0]);
        atomicAdd(&acc2[(size_t)si * 2 + 1], v[1]);
    }
}

// ---------------------------------------------------------------------------------------
// estimate_sigma: pywt.dwtn(x, 'db2')['dd'] with float32 accumulation, then |dd|.
//
// out[o] = ((f0*x[2o+1] + f1*x[2o]) + f2*x[2o-1]) + f3*x[2o-2] along axis 0, then the same along axis 1
// (half-sample symmetric border), every product and sum rounded to float32 as pywt's float32 convolution
// does.  A warp owns 31 coefficient columns x DB_ROWS coefficient rows: lane L holds the input column pair
// (2o-2+2L, 2o-1+2L) -- one coalesced 64-bit load per input row -- and slides the four axis-0 taps down its two
// columns in registers; the axis-1 taps of coefficient L-1 need the pair of lane L-1, one shuffle each.
// No shared-memory tile, no per-element index arithmetic: ~12 instructions per pixel (the two-pass tile
// kernel it replaces spent 72), i.e. a streaming kernel.
// ---------------------------------------------------------------------------------------
constexpr int DB_ROWS = 16;                    // coefficient rows per warp task
constexpr int DB_COLS = 31;                    // coefficient columns per warp task (lane 0 only feeds lane 1)

__global__ void __launch_bounds__(NT)
k_db2_dd(const float* __restrict__ img, Dims d, int hd, int wd, float* __restrict__ absdd,
         unsigned* __restrict__ l1, MetAcc* __restrict__ acc) {
    __shared__ unsigned hh[SEL_L1_BINS];
    const float f0 = (float)-0.48296291314453416, f1 = (float)0.8365163037378079,
                f2 = (float)-0.2241438680420134, f3 = (float)-0.12940952255126037;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const float* src = img + (size_t)s * d.h * d.w;
    float* dst = absdd + (size_t)s * hd * wd;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < SEL_L1_BINS; i += NT) hh[i] = 0;
    __syncthreads();
    const int n_cg = (wd + DB_COLS - 1) / DB_COLS, n_rb = (hd + DB_ROWS - 1) / DB_ROWS;
    const int warps = gridDim.x * (NT / 32);
    const bool aligned = (d.w & 1) == 0 && ((((uintptr_t)src) & 7) == 0);
    unsigned nz = 0;
    for (int task = blockIdx.x * (NT / 32) + (tid >> 5); task < n_cg * n_rb; task += warps) {
        const int rb = task / n_cg, cg = task - rb * n_cg;
        const int cx = 2 * DB_COLS * cg - 2 + 2 * lane;          // this lane's first input column (even, may be outside)
        const int c0 = refl_sym_fast(cx, d.w), c1 = refl_sym_fast(cx + 1, d.w);
        const bool vec = aligned && cx >= 0 && cx + 1 < d.w;     // both columns real and adjacent: one 64-bit load
        const int ox = DB_COLS * cg + lane - 1;                  // coefficient column of lanes 1 .. 31
        const bool col_ok = lane >= 1 && ox < wd;
        auto load = [&](int iy, float& a, float& b) {
            const float* row = src + (size_t)refl_sym_fast(iy, d.h) * d.w;
            if (vec) { const float2 v = __ldg(reinterpret_cast<const float2*>(row + cx)); a = v.x; b = v.y; }
            else { a = __ldg(row + c0); b = __ldg(row + c1); }
        };
        const int oy0 = rb * DB_ROWS, oy1 = min(hd, oy0 + DB_ROWS);
        float a0, b0, a1, b1;                                     // input rows 2oy-2, 2oy-1 of the two columns
        load(2 * oy0 - 2, a0, b0);
        load(2 * oy0 - 1, a1, b1);
#pragma unroll 4
        for (int oy = oy0; oy < oy1; ++oy) {
            float a2, b2, a3, b3;
            load(2 * oy, a2, b2);
            load(2 * oy + 1, a3, b3);
            float ta = __fmul_rn(f0, a3);
            ta = __fadd_rn(ta, __fmul_rn(f1, a2));
            ta = __fadd_rn(ta, __fmul_rn(f2, a1));
            ta = __fadd_rn(ta, __fmul_rn(f3, a0));
            float tb = __fmul_rn(f0, b3);
            tb = __fadd_rn(tb, __fmul_rn(f1, b2));
            tb = __fadd_rn(tb, __fmul_rn(f2, b1));
            tb = __fadd_rn(tb, __fmul_rn(f3, b0));
            const float na = __shfl_up_sync(0xffffffffu, ta, 1), nb = __shfl_up_sync(0xffffffffu, tb, 1);
            float v = __fmul_rn(f0, tb);                          // t[2ox+1]
            v = __fadd_rn(v, __fmul_rn(f1, ta));                  // t[2ox]
            v = __fadd_rn(v, __fmul_rn(f2, nb));                  // t[2ox-1]
            v = __fadd_rn(v, __fmul_rn(f3, na));                  // t[2ox-2]
            v = fabsf(v);
            if (col_ok) {
                dst[(size_t)oy * wd + ox] = v;
                nz += (v == 0.0f);
                atomicAdd(&hh[sel_bin1(v)], 1u);
            }
            a0 = a2; b0 = b2; a1 = a3; b1 = b3;
        }
    }
    __syncthreads();
    nz = warp_sum_u(nz);
    if (lane == 0 && nz) atomicAdd(&acc[si].dd_zero, nz);
    unsigned* g = l1 + (size_t)si * SEL_L1_BINS;
    for (int i = tid; i < SEL_L1_BINS; i += NT) { unsigned v = hh[i]; if (v) atomicAdd(&g[i], v); }
}

// ranks of the two middle non-zero |dd| values (np.median after dropping exact zeros)
__global__ void k_sigma_ranks(Dims d, int len, const MetAcc* __restrict__ acc, int* __restrict__ ranks) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    int nz = (int)acc[si].dd_zero;
    int m = len - nz;
    if (m <= 0) { ranks[s * 2] = -1; ranks[s * 2 + 1] = -1; return; }
    ranks[s * 2] = nz + (m - 1) / 2;
    ranks[s * 2 + 1] = nz + m / 2;
}

__device__ __forceinline__ double sigma_from_pair(float a, float b) {
    float med = __fdiv_rn(__fadd_rn(a, b), 2.0f);
    return (double)med / 0.6744897501960817;
}

__global__ void k_sigma_out(Dims d, const float* __restrict__ sel_out, double* __restrict__ sigma) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    sigma[s] = sigma_from_pair(sel_out[s * 2], sel_out[s * 2 + 1]);
}

__global__ void k_fill_ranks(Dims d, int Q, const int* __restrict__ src, int* __restrict__ ranks) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n_sel * Q) return;
    int si = i / Q, q = i - si * Q;
    int s = slice_of(d.sel, si);
    ranks[s * Q + q] = src[q];
}

// ---------------------------------------------------------------------------------------
// |grad| statistics that need max|grad| and P90(|grad|) first.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160)
k_grad_prep(Dims d, const MetAcc* __restrict__ acc, const float* __restrict__ g90, float gamma90,
            GradPrep* __restrict__ prep) {
    const int si = blockIdx.x;
    const int s = slice_of(d.sel, si);
    const float gmax = key2f(acc[si].gmax_key);
    const double last = (double)gmax + 1e-8;      // np.histogram range upper edge (python float)
    const int i = threadIdx.x;
    if (i < 128) {
        const double step = last / 128.0;         // np.linspace in float64, then cast to float32
        prep[si].edges[i] = (float)((double)i * step + 0.0);
    } else if (i == 128) {
        prep[si].edges[128] = (float)last;
    } else if (i == 129) {
        prep[si].denom = (float)last;
        prep[si].last = (float)last;
        prep[si].thr_edge = gmax > 0.0f ? __fmul_rn(0.1f, gmax) : 0.0f;
        prep[si].t90 = lerp_np(g90[s * 2], g90[s * 2 + 1], gamma90);
    }
}

// Bin of one |grad| value in numpy's 128-bin histogram over [0, last]: the bin is DEFINED by the float32 edge
// array (edges[b] <= v < edges[b+1], last bin closed); numpy reaches it from the tentative index
// int((v / denom) * 128) by at most one correction step either way, and so does any tentative index within one
// bin of it -- a multiply by the reciprocal replaces the IEEE division.
__device__ __forceinline__ int grad_bin(float v, float scale, const float* edges) {
    int b = (int)__fmul_rn(v, scale);
    b = min(max(b, 0), 127);
    if (v < edges[b]) b -= 1;
    else if (b != 127 && v >= edges[b + 1]) b += 1;
    return min(max(b, 0), 127);
}

__global__ void __launch_bounds__(NT)
k_grad_pass(const float* __restrict__ gbuf, Dims d, const GradPrep* __restrict__ prep,
            MetAcc* __restrict__ acc) {
    __shared__ float edges[129];
    __shared__ unsigned h[128];
    __shared__ double red[32];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const int tid = threadIdx.x, lane = tid & 31;
    const GradPrep& P = prep[si];
    for (int i = tid; i < 129; i += NT) edges[i] = P.edges[i];
    for (int i = tid; i < 128; i += NT) h[i] = 0;
    const float scale = __fdiv_rn(128.0f, P.denom), last = P.last, thr = P.thr_edge, t90 = P.t90;
    __syncthreads();
    const float* g = gbuf + (size_t)s * d.h * d.w;
    const int len = d.h * d.w;
    unsigned c_edge = 0, c_strong = 0;
    double s_strong[1] = {0.0};
    // one element (tails, unaligned slices); all 32 lanes call
    auto visit = [&](float v, bool ok) {
        c_edge += (ok && v > thr);
        if (ok && v >= t90) { c_strong++; s_strong[0] += (double)v; }
        hist_add(h, grad_bin(v, scale, edges), ok && v >= 0.0f && v <= last, lane);
    };
    // four elements per lane: counters and the strong-edge sum on the four values at once (their float32 partial
    // sum is exact to 2 ulp of the largest and enters the float64 accumulator with ONE conversion); one uniformity
    // vote for all 128 values of the warp (flat regions: a single atomic), per-element atomics otherwise
    auto visit4 = [&](const float4 q, bool ok) {
        const float v[4] = {q.x, q.y, q.z, q.w};
        float fs = 0.0f;
        int b[4];
        bool in[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            c_edge += (ok && v[k] > thr);
            const bool st = ok && v[k] >= t90;
            c_strong += st;
            fs += st ? v[k] : 0.0f;
            in[k] = ok && v[k] >= 0.0f && v[k] <= last;
            b[k] = grad_bin(v[k], scale, edges);
        }
        s_strong[0] += (double)fs;
        const bool all_in = in[0] && in[1] && in[2] && in[3];
        const bool same = all_in && b[0] == b[1] && b[1] == b[2] && b[2] == b[3];
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        if (act == 0) return;
        const int leader = __ffs(act) - 1;
        const int b0 = __shfl_sync(0xffffffffu, b[0], leader);
        if (__all_sync(0xffffffffu, !ok || (same && b[0] == b0))) {
            if (lane == leader) atomicAdd(&h[b0], 4u * (unsigned)__popc(act));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (in[k]) atomicAdd(&h[b[k]], 1u);
        }
    };
    const int tid0 = blockIdx.x * NT + tid, nthr = gridDim.x * NT;
    if ((((uintptr_t)g) & 15) == 0) {
        const int n4 = len >> 2;
        const float4* g4 = reinterpret_cast<const float4*>(g);
        int i = tid0;
        for (; i - lane + 31 + nthr < n4; i += 2 * nthr) {     // two 128-bit loads in flight while the whole warp is in range
            const float4 q0 = g4[i], q1 = g4[i + nthr];
            visit4(q0, true);
            visit4(q1, true);
        }
        for (; i - lane < n4; i += nthr) {                       // warp-uniform trip count (the votes need every lane)
            const bool ok = i < n4;
            visit4(ok ? g4[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f), ok);
        }
        for (int k = (n4 << 2) + tid0; k - lane < len; k += nthr) visit(k < len ? g[k] : 0.0f, k < len);
    } else {
        for (int i = tid0; i - lane < len; i += nthr) visit(i < len ? g[i] : 0.0f, i < len);
    }
    __syncthreads();
    block_sum<1>(s_strong, red);
    c_edge = warp_sum_u(c_edge);
    c_strong = warp_sum_u(c_strong);
    MetAcc* A = acc + si;
    if (lane == 0) {
        if (c_edge) atomicAdd(&A->cnt_edge, (unsigned long long)c_edge);
        if (c_strong) atomicAdd(&A->cnt_strong, (unsigned long long)c_strong);
    }
    if (tid == 0) atomicAdd(&A->sum_strong, s_strong[0]);
    for (int i = tid; i < 128; i += NT) { unsigned v = h[i]; if (v) atomicAdd(&A->hist128[i], v); }
}

// ---------------------------------------------------------------------------------------
// Finalisation: one block per slice.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double f32r(double v) { return (double)(float)v; }

__device__ double block_entropy(const unsigned* hist, int nb, double* red) {
    // -sum p log2 p over non-empty bins, p = c / sum(c)
    double tot[1] = {0.0};
    for (int i = threadIdx.x; i < nb; i += blockDim.x) tot[0] += (double)hist[i];
    block_sum<1>(tot, red);
    __shared__ double total_s;
    if (threadIdx.x == 0) total_s = tot[0];
    __syncthreads();
    const double total = total_s;
    double e[1] = {0.0};
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
        unsigned c = hist[i];
        if (c) { double p = (double)c / total; e[0] += p * log2(p); }
    }
    block_sum<1>(e, red);
    __shared__ double ent_s;
    if (threadIdx.x == 0) ent_s = total > 0.0 ? -e[0] : 0.0;
    __syncthreads();
    return ent_s;
}

__global__ void __launch_bounds__(NT)
k_finalize(Dims d, const MetAcc* __restrict__ acc, const float* __restrict__ xsel, PctPlan plan,
           const double* __restrict__ sigma_arr, const GradPrep* __restrict__ prep, int flags,
           double* __restrict__ out) {
    __shared__ double red[32];
    const int si = blockIdx.x;
    const int s = slice_of(d.sel, si);
    const MetAcc& A = acc[si];
    const double ent = block_entropy(A.hist256, 256, red);
    const double gent = block_entropy(A.hist128, 128, red);
    if (threadIdx.x != 0) return;
    const double N = (double)d.h * (double)d.w;
    double* o = out + (size_t)s * MC_COLS;
    const double sigma = sigma_arr[s];
    const double mean = A.sum_x / N;
    const double var = fmax(A.sum_x2 / N - mean * mean, 0.0);
    const double lmean = A.sum_lap / N;
    const double lap_var = fmax(A.sum_lap2 / N - lmean * lmean, 0.0);
    const double gmean = A.sum_g / N;
    const double gvar = fmax(A.sum_g2 / N - gmean * gmean, 0.0);
    const double lsm = A.sum_ls / N;
    const double lsv = fmax(A.sum_ls2 / N - lsm * lsm, 0.0);
    const float* xs = xsel + (size_t)s * 8;
    const float p05 = lerp_np(xs[0], xs[1], plan.gamma[0]);
    const float p25 = lerp_np(xs[2], xs[3], plan.gamma[1]);
    const float p75 = lerp_np(xs[4], xs[5], plan.gamma[2]);
    const float p95 = lerp_np(xs[6], xs[7], plan.gamma[3]);
    const double sden = (1e-8 > sigma) ? 1e-8 : sigma;      // python max(sigma, 1e-8): a NaN sigma stays NaN

    o[MC_SIGMA] = sigma;
    o[MC_LAP_VAR] = f32r(lap_var);
    o[MC_STD] = f32r(sqrt(var));
    o[MC_PCT_LOW] = (double)A.cnt_low / N;
    o[MC_PCT_HIGH] = (double)A.cnt_high / N;
    o[MC_ENTROPY] = ent;
    o[MC_EDGE_DENSITY] = (double)A.cnt_edge / N;
    o[MC_GRAD_MEAN] = f32r(gmean);
    o[MC_GRAD_STD] = f32r(sqrt(gvar));
    // np.float32 mean / python float -> float32 division
    o[MC_SNR] = (double)__fdiv_rn((float)mean, (float)sden);
    o[MC_CNR] = ((double)p95 - (double)p05) / sden;
    o[MC_LAP_ENERGY] = f32r(A.sum_lap2 / N);
    o[MC_HIST_SPREAD] = (double)p75 - (double)p25;
    o[MC_LOCAL_CONTRAST] = f32r(sqrt(lsv));
    o[MC_GRAD_STRENGTH] = A.cnt_strong ? f32r(A.sum_strong / (double)A.cnt_strong) : 0.0;
    o[MC_GRAD_ENTROPY] = gent;
    o[MC_MEAN] = f32r(mean);
    const float ml = (float)(A.sum_abslap / N), mg = (float)gmean;
    const float er = __fdiv_rn(ml, __fadd_rn(mg, 1e-8f));
    o[MC_EDGE_RATIO] = (double)er;
    if (flags & 1) {
        const double lvm = A.box16[0] / N;
        const double lvv = fmax(A.box16[1] / N - lvm * lvm, 0.0);
        const float vov = __fdiv_rn((float)sqrt(lvv), __fadd_rn((float)lvm, 1e-8f));
        o[MC_VAR_OF_VAR] = (double)vov;
        o[MC_NIQE] = (double)vov + fmax(0.0, (double)er - 1.0) * 10.0;
    } else {
        o[MC_VAR_OF_VAR] = nan("");
        o[MC_NIQE] = nan("");
    }
    o[MC_GMAX] = (double)key2f(A.gmax_key);
    o[MC_P05] = (double)p05;
    o[MC_P95] = (double)p95;
    o[MC_RESERVED] = (double)prep[si].t90;
}

__global__ void k_quality_out(Dims d, const double* __restrict__ edge2, const double* __restrict__ box2,
                              int flags, double* __restrict__ out) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    const double N = (double)d.h * (double)d.w;
    const float ml = (float)(edge2[si * 2] / N), mg = (float)(edge2[si * 2 + 1] / N);
    const float er = __fdiv_rn(ml, __fadd_rn(mg, 1e-8f));
    out[s * 2] = (double)er;
    if (flags & 1) {
        const double lvm = box2[si * 2] / N;
        const double lvv = fmax(box2[si * 2 + 1] / N - lvm * lvm, 0.0);
        const float vov = __fdiv_rn((float)sqrt(lvv), __fadd_rn((float)lvm, 1e-8f));
        out[s * 2 + 1] = (double)vov + fmax(0.0, (double)er - 1.0) * 10.0;
    } else {
        out[s * 2 + 1] = nan("");
    }
}

__global__ void k_quality_gather(int n_sel, const MetAcc* __restrict__ acc, double* __restrict__ edge2,
                                 double* __restrict__ box2) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n_sel) return;
    edge2[si * 2] = acc[si].sum_abslap;
    edge2[si * 2 + 1] = acc[si].sum_g;
    box2[si * 2] = acc[si].box16[0];
    box2[si * 2 + 1] = acc[si].box16[1];
}

__global__ void k_copy_box16(int n_sel, const double* __restrict__ box2, MetAcc* __restrict__ acc) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n_sel) return;
    acc[si].box16[0] = box2[si * 2];
    acc[si].box16[1] = box2[si * 2 + 1];
}

bool legacy_tiles() {
    static const bool v = [] { const char* e = getenv("MDIMG_METRICS_TILES"); return e && e[0] == '1'; }();
    return v;
}

struct SigmaBufs {
    float* absdd; unsigned* l1; int* ranks; float* sel_out; void* sel_ws; size_t sel_ws_bytes;
};

void carve_sigma(Arena& a, int n, int n_sel, int hd, int wd, SigmaBufs& b) {
    b.absdd = a.take<float>((size_t)n * hd * wd);
    b.l1 = a.take<unsigned>((size_t)n_sel * SEL_L1_BINS);
    b.ranks = a.take<int>((size_t)n * 2);
    b.sel_out = a.take<float>((size_t)n * 2);
    b.sel_ws_bytes = select_workspace_bytes(n_sel);
    b.sel_ws = a.take<char>(b.sel_ws_bytes);
}

// db2 'dd' band + level-1 histogram + median ranks (everything before the selection)
void sigma_produce(const float* img, const Dims& d, MetAcc* acc, SigmaBufs& b, cudaStream_t stream) {
    const int hd = (d.h + 3) / 2, wd = (d.w + 3) / 2;
    cudaMemsetAsync(b.l1, 0, (size_t)d.n_sel * SEL_L1_BINS * sizeof(unsigned), stream);
    const int tasks = ((wd + DB_COLS - 1) / DB_COLS) * ((hd + DB_ROWS - 1) / DB_ROWS);     // warp tasks per slice
    int gx = (tasks + NT / 32 - 1) / (NT / 32);
    // few long-lived blocks per slice (one histogram flush each) once the batch alone fills the machine: about two
    // waves of five resident blocks on 148 SMs in total, never fewer than four per slice
    int cap = (1480 + d.n_sel - 1) / d.n_sel;
    if (cap < 4) cap = 4;
    if (cap > 64) cap = 64;
    if (gx > cap) gx = cap;
    dim3 grid(gx, d.n_sel);
    MDIMG_LAUNCH k_db2_dd<<<grid, NT, 0, stream>>>(img, d, hd, wd, b.absdd, b.l1, acc);
    MDIMG_LAUNCH k_sigma_ranks<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, hd * wd, acc, b.ranks);
}

SelJob sigma_job(const Dims& d, const SigmaBufs& b) {
    const int hd = (d.h + 3) / 2, wd = (d.w + 3) / 2;
    SelJob j;
    j.vals = b.absdd; j.stride = (long long)hd * wd; j.len = hd * wd; j.Q = 2; j.opts = 0;
    j.ranks = b.ranks; j.l1_hist = b.l1; j.out = b.sel_out; j.ws = b.sel_ws;
    return j;
}

int sigma_core(const float* img, const Dims& d, MetAcc* acc, SigmaBufs& b, double* sigma_out,
               cudaStream_t stream) {
    sigma_produce(img, d, acc, b, stream);
    const SelJob j = sigma_job(d, b);
    int rc = select_run_multi(&j, 1, d, stream);
    if (rc) return rc;
    MDIMG_LAUNCH k_sigma_out<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, b.sel_out, sigma_out);
    return check_launch("estimate_sigma");
}

}  // namespace

// ---------------------------------------------------------------------------------------
// Host entry points (internal C++; the C-ABI wrappers live in api.cu)
// ---------------------------------------------------------------------------------------
size_t sigma_workspace_bytes(int n_sel, int h, int w) {
    // sized for n == n_sel is not enough when sel indexes a larger stack; api.cu passes n.
    Arena a(nullptr, 0);
    SigmaBufs b;
    a.take<MetAcc>(n_sel);
    carve_sigma(a, n_sel, n_sel, (h + 3) / 2, (w + 3) / 2, b);
    return a.off;
}

int sigma_run(const float* img, const Dims& d, double* sigma_out, void* ws, size_t ws_bytes,
              cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    Arena a(ws, ws_bytes);
    MetAcc* acc = a.take<MetAcc>(d.n_sel);
    SigmaBufs b;
    carve_sigma(a, d.n, d.n_sel, (d.h + 3) / 2, (d.w + 3) / 2, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "estimate_sigma: workspace too small (%zu > %zu)", a.off, ws_bytes);
    cudaMemsetAsync(acc, 0, sizeof(MetAcc) * d.n_sel, stream);
    return sigma_core(img, d, acc, b, sigma_out, stream);
}

size_t quality_workspace_bytes(int n_sel, int h, int w) {
    (void)h; (void)w;
    Arena a(nullptr, 0);
    a.take<double>((size_t)n_sel * 2);
    a.take<double>((size_t)n_sel * 2);
    a.take<MetAcc>(n_sel);
    return a.off;
}

int quality_run(const float* img, const Dims& d, int flags, double* out, void* ws, size_t ws_bytes,
                cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    Arena a(ws, ws_bytes);
    double* edge2 = a.take<double>((size_t)d.n_sel * 2);
    double* box2 = a.take<double>((size_t)d.n_sel * 2);
    MetAcc* acc = a.take<MetAcc>(d.n_sel);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "quality: workspace too small");
    if (legacy_tiles()) {
        cudaMemsetAsync(edge2, 0, sizeof(double) * 2 * d.n_sel, stream);
        cudaMemsetAsync(box2, 0, sizeof(double) * 2 * d.n_sel, stream);
        dim3 grid(((d.w + EW - 1) / EW) * ((d.h + EH - 1) / EH), d.n_sel);
        MDIMG_LAUNCH k_edge_stats<<<grid, NT, 0, stream>>>(img, d, edge2);
        if (flags & 1) launch_box16_stats(img, d, box2, stream);
    } else {
        // the strip kernel without histograms / 7-window / |grad| store: sum|laplace|, sum|grad|, 16-window stats
        cudaMemsetAsync(acc, 0, sizeof(MetAcc) * d.n_sel, stream);
        strip::launch<false>(img, d, flags & 1, acc, nullptr, nullptr, nullptr, stream);
        MDIMG_LAUNCH k_quality_gather<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d.n_sel, acc, edge2, box2);
    }
    MDIMG_LAUNCH k_quality_out<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, edge2, box2, flags, out);
    return check_launch("quality");
}

namespace {
struct MetBufs {
    MetAcc* acc; float* g; unsigned* l1x; unsigned* l1g; int* plan_ranks; int* xranks; int* granks;
    float* xsel; float* gsel; GradPrep* prep; double* sigma; double* box2;
    void* sel_ws; void* sel_ws_g; size_t sel_ws_bytes; SigmaBufs sb;
};
void carve_metrics(Arena& a, int n, int n_sel, int h, int w, MetBufs& m) {
    m.acc = a.take<MetAcc>(n_sel);
    m.g = a.take<float>((size_t)n * h * w);
    m.l1x = a.take<unsigned>((size_t)n_sel * SEL_L1_BINS);
    m.l1g = a.take<unsigned>((size_t)n_sel * SEL_L1_BINS);
    m.plan_ranks = a.take<int>(16);
    m.xranks = a.take<int>((size_t)n * 8);
    m.granks = a.take<int>((size_t)n * 2);
    m.xsel = a.take<float>((size_t)n * 8);
    m.gsel = a.take<float>((size_t)n * 2);
    m.prep = a.take<GradPrep>(n_sel);
    m.sigma = a.take<double>(n);
    m.box2 = a.take<double>((size_t)n_sel * 2);
    m.sel_ws_bytes = select_workspace_bytes(n_sel);
    m.sel_ws = a.take<char>(m.sel_ws_bytes);
    m.sel_ws_g = a.take<char>(m.sel_ws_bytes);
    carve_sigma(a, n, n_sel, (h + 3) / 2, (w + 3) / 2, m.sb);
}
}  // namespace

size_t metrics_workspace_bytes(int n_sel, int h, int w) {
    Arena a(nullptr, 0);
    MetBufs m;
    carve_metrics(a, n_sel, n_sel, h, w, m);
    return a.off;
}

static int metrics_run_batch(const float* img, const Dims& d, const PctPlan& plan, int flags, double* out,
                             void* ws, size_t ws_bytes, cudaStream_t stream);

// Slices per sub-batch of one mdimg_metrics call.  The passes after the strip kernel re-read x, |grad| and
// |dd| (level-1 scan, <= 3 refinement passes, the |grad| statistics): run over a whole 512-slice chunk they
// stream those arrays from HBM every time (35 B/px of DRAM traffic measured); run sub-batch by sub-batch,
// with every kernel of a sub-batch issued before the next sub-batch starts, the re-reads hit the 126 MB L2.
// MDIMG_METRICS_SUB=<slices> overrides (0 = one batch).
static int metrics_sub_slices(int h, int w) {
    static const int env = [] { const char* e = getenv("MDIMG_METRICS_SUB"); return e ? atoi(e) : -1; }();
    if (env >= 0) return env;
    return 0;
}

int metrics_run(const float* img, const Dims& d, const PctPlan& plan, int flags, double* out,
                void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    const int sub = metrics_sub_slices(d.h, d.w);
    if (sub <= 0 || d.n_sel <= sub) return metrics_run_batch(img, d, plan, flags, out, ws, ws_bytes, stream);
    for (int a0 = 0; a0 < d.n_sel; a0 += sub) {
        const int cnt = d.n_sel - a0 < sub ? d.n_sel - a0 : sub;
        Dims ds = d;
        int rc;
        if (d.sel) {                   // positions a0 .. a0+cnt of the selection; arrays stay indexed by slice id
            ds.sel = d.sel + a0;
            ds.n_sel = cnt;
            rc = metrics_run_batch(img, ds, plan, flags, out, ws, ws_bytes, stream);
        } else {                       // contiguous slices: shift the bases
            ds.n = cnt;
            ds.n_sel = cnt;
            rc = metrics_run_batch(img + (size_t)a0 * d.h * d.w, ds, plan, flags, out + (size_t)a0 * MC_COLS, ws, ws_bytes, stream);
        }
        if (rc) return rc;
    }
    return MDIMG_OK;
}

static int metrics_run_batch(const float* img, const Dims& d, const PctPlan& plan, int flags, double* out,
                             void* ws, size_t ws_bytes, cudaStream_t stream) {
    Arena a(ws, ws_bytes);
    MetBufs m;
    carve_metrics(a, d.n, d.n_sel, d.h, d.w, m);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "metrics: workspace too small (%zu > %zu)", a.off, ws_bytes);
    const int len = d.h * d.w;

    cudaMemsetAsync(m.acc, 0, sizeof(MetAcc) * d.n_sel, stream);
    cudaMemsetAsync(m.l1x, 0, (size_t)d.n_sel * SEL_L1_BINS * sizeof(unsigned), stream);
    cudaMemsetAsync(m.l1g, 0, (size_t)d.n_sel * SEL_L1_BINS * sizeof(unsigned), stream);

    // one read of the image: stencil statistics, |grad|, both box filters, histograms (flags bit0: the
    // 16x16 NIQE window too).  MDIMG_METRICS_TILES=1 selects the previous tile kernels (A/B measurements).
    const bool tiles = legacy_tiles();
    if (tiles) {
        static unsigned long long devices_done = 0;
        opt_in_shared_memory(k_stencil_stats, sizeof(StencilSmem), devices_done);
        dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
        MDIMG_LAUNCH k_stencil_stats<<<grid, NT, sizeof(StencilSmem), stream>>>(img, d, m.acc, m.g, m.l1x, m.l1g);
    } else {
        strip::launch<true>(img, d, flags & 1, m.acc, m.g, m.l1x, m.l1g, stream);
    }
    int rc = check_launch("stencil_stats");
    if (rc) return rc;

    // db2 'dd' band of estimate_sigma (+ its level-1 histogram and median ranks)
    sigma_produce(img, d, m.acc, m.sb, stream);

    // percentiles of x: ranks (lo,hi) for 5, 25, 75, 95; of |grad|: 90; median of |dd|:
    // three selections refined by the same launches
    int host_ranks[10];
    for (int k = 0; k < 4; ++k) { host_ranks[2 * k] = plan.lo[k]; host_ranks[2 * k + 1] = plan.hi[k]; }
    host_ranks[8] = plan.lo[4]; host_ranks[9] = plan.hi[4];
    cudaMemcpyAsync(m.plan_ranks, host_ranks, sizeof(host_ranks), cudaMemcpyHostToDevice, stream);
    MDIMG_LAUNCH k_fill_ranks<<<(d.n_sel * 8 + 127) / 128, 128, 0, stream>>>(d, 8, m.plan_ranks, m.xranks);
    MDIMG_LAUNCH k_fill_ranks<<<(d.n_sel * 2 + 127) / 128, 128, 0, stream>>>(d, 2, m.plan_ranks + 8, m.granks);
    SelJob jobs[3];
    jobs[0].vals = img; jobs[0].stride = len; jobs[0].len = len; jobs[0].Q = 8; jobs[0].opts = 0;
    jobs[0].ranks = m.xranks; jobs[0].l1_hist = m.l1x; jobs[0].out = m.xsel; jobs[0].ws = m.sel_ws;
    jobs[1].vals = m.g; jobs[1].stride = len; jobs[1].len = len; jobs[1].Q = 2; jobs[1].opts = 0;
    jobs[1].ranks = m.granks; jobs[1].l1_hist = m.l1g; jobs[1].out = m.gsel; jobs[1].ws = m.sel_ws_g;
    jobs[2] = sigma_job(d, m.sb);
    rc = select_run_multi(jobs, 3, d, stream);
    if (rc) return rc;
    MDIMG_LAUNCH k_sigma_out<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, m.sb.sel_out, m.sigma);

    MDIMG_LAUNCH k_grad_prep<<<d.n_sel, 160, 0, stream>>>(d, m.acc, m.gsel, plan.gamma[4], m.prep);
    int bx = (len + NT * 8 - 1) / (NT * 8);
    if (bx > 512) bx = 512;
    MDIMG_LAUNCH k_grad_pass<<<dim3(bx, d.n_sel), NT, 0, stream>>>(m.g, d, m.prep, m.acc);

    if ((flags & 1) && tiles) {
        cudaMemsetAsync(m.box2, 0, sizeof(double) * 2 * d.n_sel, stream);
        launch_box16_stats(img, d, m.box2, stream);
        MDIMG_LAUNCH k_copy_box16<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d.n_sel, m.box2, m.acc);
    }
    MDIMG_LAUNCH k_finalize<<<d.n_sel, NT, 0, stream>>>(d, m.acc, m.xsel, plan, m.sigma, m.prep, flags, out);
    return check_launch("metrics");
}

}  // namespace mdimg
