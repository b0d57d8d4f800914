// Elementwise and reduction kernels: per-slice min/max, normalize_image
// (pipeline/dicom_io.py:84-91), adjust_gamma (skimage, called at pipeline/enhancement.py:194,197,
// 284,336), the blends of _light_denoise / the over-processing guard (enhancement.py:93,365) and
// np.clip(., 0, 1) (enhancement.py:218,314,352).  All are one read + one write of the slice,
// 128-bit vectorised when the slice length is a multiple of 4.
#include "enhance.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;

inline int blocks_for(long long len, int per_thread) {
    long long b = (len + (long long)NT * per_thread - 1) / ((long long)NT * per_thread);
    if (b < 1) b = 1;
    if (b > 4096) b = 4096;
    return (int)b;
}

// ---------------- min / max ----------------
__global__ void k_mm_init(Dims d, uint2* __restrict__ mm) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    mm[slice_of(d.sel, si)] = make_uint2(0xFFFFFFFFu, 0u);
}

template <typename T>
__global__ void __launch_bounds__(NT)
k_minmax(const T* __restrict__ img, Dims d, uint2* __restrict__ mm) {
    __shared__ float smin[NT / 32], smax[NT / 32];
    const int s = slice_of(d.sel, blockIdx.y);
    const long long len = d.px();
    const T* p = img + (size_t)s * len;
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < len; i += (long long)gridDim.x * NT) {
        float v = (float)p[i];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { smin[wid] = lo; smax[wid] = hi; }
    __syncthreads();
    if (wid == 0) {
        lo = lane < NT / 32 ? smin[lane] : INFINITY;
        hi = lane < NT / 32 ? smax[lane] : -INFINITY;
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) {
            atomicMin(&mm[s].x, f2key(lo));
            atomicMax(&mm[s].y, f2key(hi));
        }
    }
}

// uint16 samples: 128-bit loads (8 samples), four in flight per thread, packed 16-bit min / max; the
// float keys are formed once per block.  Streams 2 B/px at HBM speed (the scalar version above moved
// 64 bytes per warp instruction).
__global__ void __launch_bounds__(NT)
k_minmax_u16v(const uint16_t* __restrict__ img, Dims d, uint2* __restrict__ mm) {
    __shared__ unsigned smin[NT / 32], smax[NT / 32];
    const int s = slice_of(d.sel, blockIdx.y);
    const long long len = d.px();
    const uint16_t* p = img + (size_t)s * len;
    unsigned lo2 = 0xFFFFFFFFu, hi2 = 0u;                 // two 16-bit lanes each
    auto take = [&](const uint4& v) {
        lo2 = __vminu2(__vminu2(lo2, v.x), __vminu2(v.y, __vminu2(v.z, v.w)));
        hi2 = __vmaxu2(__vmaxu2(hi2, v.x), __vmaxu2(v.y, __vmaxu2(v.z, v.w)));
    };
    const long long tid0 = (long long)blockIdx.x * NT + threadIdx.x, nthr = (long long)gridDim.x * NT;
    long long done = 0;
    if ((((uintptr_t)p) & 15) == 0) {
        const long long n8 = len >> 3;
        const uint4* p8 = reinterpret_cast<const uint4*>(p);
        long long i = tid0;
        for (; i + 3 * nthr < n8; i += 4 * nthr) {
            const uint4 a = __ldg(p8 + i), b = __ldg(p8 + i + nthr), c = __ldg(p8 + i + 2 * nthr), e = __ldg(p8 + i + 3 * nthr);
            take(a); take(b); take(c); take(e);
        }
        for (; i < n8; i += nthr) take(__ldg(p8 + i));
        done = n8 << 3;
    }
    unsigned lo = min(lo2 & 0xffffu, lo2 >> 16), hi = max(hi2 & 0xffffu, hi2 >> 16);
    for (long long i = done + tid0; i < len; i += nthr) {
        const unsigned v = p[i];
        lo = min(lo, v);
        hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { smin[wid] = lo; smax[wid] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < NT / 32; ++w) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
        atomicMin(&mm[s].x, f2key((float)lo));
        atomicMax(&mm[s].y, f2key((float)hi));
    }
}

__global__ void k_mm_decode(Dims d, const uint2* __restrict__ mm, float* __restrict__ out2) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    out2[s * 2] = key2f(mm[s].x);
    out2[s * 2 + 1] = key2f(mm[s].y);
}

// ---------------- generic elementwise driver ----------------
// F: struct with `__device__ bool prepare(int s)` (returns false to skip the slice) and
// `__device__ float apply(float a, float b)`.
template <typename F, bool TWO>
__global__ void __launch_bounds__(NT)
k_map(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, Dims d, F f) {
    const int s = slice_of(d.sel, blockIdx.y);
    if (!f.prepare(s)) return;
    const long long len = d.px();
    const size_t base = (size_t)s * len;
    const float* pa = a + base;
    const float* pb = TWO ? b + base : nullptr;
    float* po = out + base;
    const bool aligned = (((uintptr_t)pa | (uintptr_t)po | (TWO ? (uintptr_t)pb : 0)) & 15) == 0;
    if ((len & 3) == 0 && aligned) {
        const long long n4 = len >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(pa);
        const float4* b4 = reinterpret_cast<const float4*>(pb);
        float4* o4 = reinterpret_cast<float4*>(po);
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
            float4 va = a4[i];
            float4 vb = TWO ? b4[i] : make_float4(0, 0, 0, 0);
            float4 r;
            r.x = f.apply(va.x, vb.x); r.y = f.apply(va.y, vb.y);
            r.z = f.apply(va.z, vb.z); r.w = f.apply(va.w, vb.w);
            o4[i] = r;
        }
    } else {
        for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < len; i += (long long)gridDim.x * NT)
            po[i] = f.apply(pa[i], TWO ? pb[i] : 0.0f);
    }
}

struct NormF {
    const uint2* mm; float lo, denom; bool zero;
    __device__ bool prepare(int s) {
        float mn = key2f(mm[s].x), mx = key2f(mm[s].y);
        zero = ((double)mx - (double)mn) < 1e-8;
        lo = mn;
        denom = (float)((double)mx - (double)mn);   // python-float difference used as a float32 scalar
        return true;
    }
    __device__ float apply(float a, float) const {
        return zero ? 0.0f : __fdiv_rn(__fsub_rn(a, lo), denom);
    }
};

// (a - lo) / denom for INTEGER operands 0 <= a <= denom <= 65535 (uint16 samples minus the slice minimum
// over max - min): the IEEE quotient from the slice's correctly rounded reciprocal and one exact-residual
// correction -- q = a * r; q += fma(-denom, q, a) * r -- i.e. 3 FP32-pipe instructions per pixel instead
// of the ~10-instruction division sequence with its MUFU.  Markstein's correction step; that it equals
// __fdiv_rn for EVERY operand pair of this domain is checked exhaustively on the device
// (mdimg_selftest_div16, tests/test_gpu_parity.py).
__device__ __forceinline__ float div16(float a, float denom, float rcp) {
    const float q = __fmul_rn(a, rcp);
    return __fmaf_rn(__fmaf_rn(-denom, q, a), rcp, q);
}

__global__ void __launch_bounds__(NT)
k_normalize_u16(const uint16_t* __restrict__ in, float* __restrict__ out, Dims d,
                const uint2* __restrict__ mm) {
    const int s = slice_of(d.sel, blockIdx.y);
    NormF f; f.mm = mm; f.prepare(s);
    const float lo = f.lo, denom = f.denom;
    const bool zero = f.zero;
    const float rcp = zero ? 0.0f : __frcp_rn(denom);
    const long long len = d.px();
    const uint16_t* p = in + (size_t)s * len;
    float* o = out + (size_t)s * len;
    auto one = [&](unsigned u) { return zero ? 0.0f : div16(__fsub_rn((float)u, lo), denom, rcp); };
    const long long tid0 = (long long)blockIdx.x * NT + threadIdx.x, nthr = (long long)gridDim.x * NT;
    long long done = 0;
    if (((((uintptr_t)p) | ((uintptr_t)o)) & 15) == 0) {
        const long long n8 = len >> 3;
        const uint4* p8 = reinterpret_cast<const uint4*>(p);
        float4* o4 = reinterpret_cast<float4*>(o);
        auto emit = [&](long long i, const uint4& v) {
            float4 a, b;
            a.x = one(v.x & 0xffffu); a.y = one(v.x >> 16); a.z = one(v.y & 0xffffu); a.w = one(v.y >> 16);
            b.x = one(v.z & 0xffffu); b.y = one(v.z >> 16); b.z = one(v.w & 0xffffu); b.w = one(v.w >> 16);
            __stcs(o4 + 2 * i, a);
            __stcs(o4 + 2 * i + 1, b);
        };
        long long i = tid0;
        for (; i + nthr < n8; i += 2 * nthr) {           // two 128-bit loads in flight
            const uint4 v0 = __ldg(p8 + i), v1 = __ldg(p8 + i + nthr);
            emit(i, v0);
            emit(i + nthr, v1);
        }
        for (; i < n8; i += nthr) emit(i, __ldg(p8 + i));
        done = n8 << 3;
    }
    for (long long i = done + tid0; i < len; i += nthr) o[i] = one(p[i]);
}

// Exhaustive check of div16 against the IEEE division over its whole domain: denom = blockIdx-strided
// 1 .. 65535, a = 0 .. denom.  mismatches: device counter.
__global__ void __launch_bounds__(NT)
k_selftest_div16(unsigned long long* __restrict__ mismatches) {
    unsigned bad = 0;
    for (unsigned dn = 1 + blockIdx.x; dn <= 65535u; dn += gridDim.x) {
        const float denom = (float)dn, rcp = __frcp_rn(denom);
        for (unsigned a = threadIdx.x; a <= dn; a += NT) {
            const float fa = (float)a;
            bad += __float_as_uint(div16(fa, denom, rcp)) != __float_as_uint(__fdiv_rn(fa, denom));
        }
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(mismatches, (unsigned long long)bad);
}

// ---- ingestion: modality rescale + MONOCHROME1 inversion + normalize_image in one pass --------
// load_dicom (pipeline/dicom_io.py:44-49): v = float32(float64(raw) * slope + intercept) (pydicom's
// apply_modality_lut), then `image.max() - image` for MONOCHROME1 (max over the whole pixel array,
// i.e. all frames), then normalize_image per 2-D frame (dicom_io.py:84-91).  All three maps are
// monotone in the raw value, so every min / max follows from the raw per-slice and global extrema.
struct IngestParams { double slope, intercept; int has_rescale, invert, is_signed; };

__device__ __forceinline__ float ingest_value(float raw, const IngestParams& P) {
    if (!P.has_rescale) return raw;
    return (float)__dadd_rn(__dmul_rn((double)raw, P.slope), P.intercept);
}

__global__ void k_global_mm(int n_sel, Dims d, const uint2* __restrict__ mm, uint2* __restrict__ gmm) {
    // one block: extrema over the selected slices (keys are order preserving)
    __shared__ unsigned smin[32], smax[32];
    unsigned lo = 0xFFFFFFFFu, hi = 0u;
    for (int si = threadIdx.x; si < n_sel; si += blockDim.x) {
        const uint2 v = mm[slice_of(d.sel, si)];
        lo = min(lo, v.x);
        hi = max(hi, v.y);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { smin[wid] = lo; smax[wid] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = min(lo, smin[w]); hi = max(hi, smax[w]); }
        gmm[0] = make_uint2(lo, hi);
    }
}

__global__ void __launch_bounds__(NT)
k_ingest(const uint16_t* __restrict__ in, float* __restrict__ out, Dims d, const uint2* __restrict__ mm,
         const uint2* __restrict__ gmm, IngestParams P) {
    const int s = slice_of(d.sel, blockIdx.y);
    const float rmin = key2f(mm[s].x), rmax = key2f(mm[s].y);
    const bool inc = !P.has_rescale || P.slope >= 0.0;
    float vmin = ingest_value(inc ? rmin : rmax, P), vmax = ingest_value(inc ? rmax : rmin, P);
    float M = 0.0f;
    if (P.invert) {
        M = ingest_value(inc ? key2f(gmm[0].y) : key2f(gmm[0].x), P);     // image.max() over all frames
        const float nmin = __fsub_rn(M, vmax), nmax = __fsub_rn(M, vmin);
        vmin = nmin; vmax = nmax;
    }
    const double span = (double)vmax - (double)vmin;
    const bool zero = span < 1e-8;
    const float lo = vmin, denom = (float)span;
    const long long len = d.px();
    const uint16_t* p = in + (size_t)s * len;
    float* o = out + (size_t)s * len;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < len; i += (long long)gridDim.x * NT) {
        const uint16_t u = p[i];
        const float raw = P.is_signed ? (float)(short)u : (float)u;
        float v = ingest_value(raw, P);
        if (P.invert) v = __fsub_rn(M, v);
        o[i] = zero ? 0.0f : __fdiv_rn(__fsub_rn(v, lo), denom);
    }
}

struct GammaF {
    const uint2* mm; int* neg_flag; double gamma;
    __device__ bool prepare(int s) {
        if (mm && key2f(mm[s].x) < 0.0f) {     // _assert_non_negative -> ValueError in the reference
            if (threadIdx.x == 0 && blockIdx.x == 0) neg_flag[s] = 1;
            return false;
        }
        return true;
    }
    // float32 power: evaluated in double and rounded once (a correctly rounded powf)
    __device__ float apply(float a, float) const { return (float)pow((double)a, gamma); }
};

struct AxpbyF {
    float c0, c1; int clip;
    __device__ bool prepare(int) { return true; }
    __device__ float apply(float a, float b) const {
        float r = __fadd_rn(__fmul_rn(c0, a), __fmul_rn(c1, b));
        if (clip) r = fminf(fmaxf(r, 0.0f), 1.0f);
        return r;
    }
};

struct BlendSkipF {      // _light_denoise: (1-k)*x + k*den, or x itself where the slice was skipped
    float c0, c1; const int* skip; const double* sigma; int* skipped_out; bool ident;
    __device__ bool prepare(int s) {
        ident = skip[s] != 0;
        if (skipped_out && threadIdx.x == 0 && blockIdx.x == 0) skipped_out[s] = ident ? 1 : 0;
        return true;
    }
    __device__ float apply(float a, float b) const {
        return ident ? a : __fadd_rn(__fmul_rn(c0, a), __fmul_rn(c1, b));
    }
};

__global__ void k_skip_flags(Dims d, const double* __restrict__ sigma, double thresh, int* __restrict__ skip,
                             int* __restrict__ skipped_out) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= d.n_sel) return;
    int s = slice_of(d.sel, si);
    const int v = sigma[s] < thresh ? 1 : 0;  // NaN compares false, as in the reference
    skip[s] = v;
    if (skipped_out) skipped_out[s] = v;
}

struct ClipF {
    __device__ bool prepare(int) { return true; }
    __device__ float apply(float a, float) const { return fminf(fmaxf(a, 0.0f), 1.0f); }
};

struct CopyF {
    __device__ bool prepare(int) { return true; }
    __device__ float apply(float a, float) const { return a; }
};

}  // namespace

int minmax_f32_run(const float* img, const Dims& d, uint2* mm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_mm_init<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, mm);
    MDIMG_LAUNCH k_minmax<float><<<dim3(blocks_for(d.px(), 16), d.n_sel), NT, 0, stream>>>(img, d, mm);
    return check_launch("minmax_f32");
}

int minmax_u16_run(const uint16_t* img, const Dims& d, uint2* mm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_mm_init<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, mm);
    MDIMG_LAUNCH k_minmax_u16v<<<dim3(blocks_for(d.px(), 64), d.n_sel), NT, 0, stream>>>(img, d, mm);
    return check_launch("minmax_u16");
}

int ingest_run(const uint16_t* in, float* out, const Dims& d, double slope, double intercept,
               int has_rescale, int invert, int is_signed, uint2* mm, uint2* gmm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_mm_init<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, mm);
    if (is_signed)
        MDIMG_LAUNCH k_minmax<short><<<dim3(blocks_for(d.px(), 16), d.n_sel), NT, 0, stream>>>((const short*)in, d, mm);
    else
        MDIMG_LAUNCH k_minmax_u16v<<<dim3(blocks_for(d.px(), 64), d.n_sel), NT, 0, stream>>>(in, d, mm);
    MDIMG_LAUNCH k_global_mm<<<1, 256, 0, stream>>>(d.n_sel, d, mm, gmm);
    IngestParams P;
    P.slope = slope; P.intercept = intercept; P.has_rescale = has_rescale; P.invert = invert; P.is_signed = is_signed;
    MDIMG_LAUNCH k_ingest<<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(in, out, d, mm, gmm, P);
    return check_launch("ingest");
}

int minmax_decode_run(const uint2* mm, const Dims& d, float* out2, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_mm_decode<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, mm, out2);
    return check_launch("minmax_decode");
}

int normalize_u16_run(const uint16_t* in, float* out, const Dims& d, const uint2* mm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_normalize_u16<<<dim3(blocks_for(d.px(), 32), d.n_sel), NT, 0, stream>>>(in, out, d, mm);
    return check_launch("normalize_u16");
}

int selftest_div16_run(unsigned long long* mismatches_dev, cudaStream_t stream) {
    cudaMemsetAsync(mismatches_dev, 0, sizeof(unsigned long long), stream);
    MDIMG_LAUNCH k_selftest_div16<<<148 * 8, NT, 0, stream>>>(mismatches_dev);
    return check_launch("selftest_div16");
}

int normalize_f32_run(const float* in, float* out, const Dims& d, const uint2* mm, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    NormF f; f.mm = mm; f.lo = 0; f.denom = 1; f.zero = false;
    MDIMG_LAUNCH k_map<NormF, false><<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(in, nullptr, out, d, f);
    return check_launch("normalize_f32");
}

int gamma_run(const float* in, float* out, const Dims& d, double gamma, const uint2* mm,
              int* neg_flag, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (gamma < 0) return set_error(MDIMG_ERR_INVALID, "Gamma should be a non-negative real number.");
    GammaF f; f.mm = mm; f.neg_flag = neg_flag; f.gamma = gamma;
    MDIMG_LAUNCH k_map<GammaF, false><<<dim3(blocks_for(d.px(), 4), d.n_sel), NT, 0, stream>>>(in, nullptr, out, d, f);
    return check_launch("gamma");
}

int axpby_run(const float* a, const float* b, float* out, const Dims& d, float c0, float c1,
              int clip01, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    AxpbyF f; f.c0 = c0; f.c1 = c1; f.clip = clip01;
    MDIMG_LAUNCH k_map<AxpbyF, true><<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(a, b, out, d, f);
    return check_launch("axpby");
}

int blend_skip_run(const float* a, const float* b, float* out, const Dims& d, float c0, float c1,
                   const int* skip, int* skipped_out, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    BlendSkipF f; f.c0 = c0; f.c1 = c1; f.skip = skip; f.sigma = nullptr; f.skipped_out = skipped_out; f.ident = false;
    MDIMG_LAUNCH k_map<BlendSkipF, true><<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(a, b, out, d, f);
    return check_launch("blend_skip");
}

int skip_flags_run(const Dims& d, const double* sigma, double thresh, int* skip, cudaStream_t stream,
                   int* skipped_out) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_skip_flags<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d, sigma, thresh, skip, skipped_out);
    return check_launch("skip_flags");
}

namespace {
// 16-bit export of a [0, 1] float32 stack: skimage's img_as_uint on float input,
// uint16(clip(rint(float32(x) * 65535), 0, 65535)) -- round half to even, as numpy's rint.
__device__ __forceinline__ unsigned export_level(float x) {
    const float t = fminf(fmaxf(rintf(__fmul_rn(x, 65535.0f)), 0.0f), 65535.0f);
    return (unsigned)t;
}

__global__ void __launch_bounds__(NT)
k_export_u16(const float* __restrict__ in, uint16_t* __restrict__ out, Dims d) {
    const int s = slice_of(d.sel, blockIdx.y);
    const long long len = d.px();
    const float* p = in + (size_t)s * len;
    uint16_t* o = out + (size_t)s * len;
    const long long tid = (long long)blockIdx.x * NT + threadIdx.x, nthr = (long long)gridDim.x * NT;
    if ((len & 3) == 0 && (((uintptr_t)p & 15) == 0) && (((uintptr_t)o & 7) == 0)) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        uint2* o4 = reinterpret_cast<uint2*>(o);
        for (long long i = tid; i < (len >> 2); i += nthr) {
            const float4 v = p4[i];
            o4[i] = make_uint2(export_level(v.x) | (export_level(v.y) << 16),
                               export_level(v.z) | (export_level(v.w) << 16));
        }
    } else {
        for (long long i = tid; i < len; i += nthr) o[i] = (uint16_t)export_level(p[i]);
    }
}
}  // namespace

int export_u16_run(const float* in, uint16_t* out, const Dims& d, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    MDIMG_LAUNCH k_export_u16<<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(in, out, d);
    return check_launch("export_u16");
}

int clip01_run(const float* in, float* out, const Dims& d, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    ClipF f;
    MDIMG_LAUNCH k_map<ClipF, false><<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(in, nullptr, out, d, f);
    return check_launch("clip01");
}

namespace {
// matplotlib's gray colormap on an autoscaled panel: Normalize in float32, 256 levels, the
// level that lands on 256 (x == vmax) folded into 255; a flat panel maps to level 0.
__device__ __forceinline__ uint8_t gray_level(float x, float vmin, float den) {
    if (!(den > 0.0f)) return 0;
    const float t = __fmul_rn(__fdiv_rn(__fsub_rn(x, vmin), den), 256.0f);
    int l = (int)t;
    l = l > 255 ? 255 : (l < 0 ? 0 : l);
    return (uint8_t)l;
}

__global__ void __launch_bounds__(NT)
k_mosaic_u8(const float* __restrict__ before, const float* __restrict__ after, uint8_t* __restrict__ out,
            Dims d, int gap, int gap_level, const uint2* __restrict__ mm_b, const uint2* __restrict__ mm_a) {
    const int s = slice_of(d.sel, blockIdx.y);
    const float bmin = key2f(mm_b[s].x), bmax = key2f(mm_b[s].y);
    const float amin = key2f(mm_a[s].x), amax = key2f(mm_a[s].y);
    const float bden = (float)((double)bmax - (double)bmin), aden = (float)((double)amax - (double)amin);
    const int ow = 2 * d.w + gap;
    const long long total = (long long)d.h * ow;
    const float* pb = before + (size_t)s * d.h * d.w;
    const float* pa = after + (size_t)s * d.h * d.w;
    uint8_t* o = out + (size_t)s * total;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < total; i += (long long)gridDim.x * NT) {
        const int y = (int)(i / ow), x = (int)(i - (long long)y * ow);
        uint8_t v;
        if (x < d.w) v = gray_level(pb[(size_t)y * d.w + x], bmin, bden);
        else if (x < d.w + gap) v = (uint8_t)gap_level;
        else v = gray_level(pa[(size_t)y * d.w + x - d.w - gap], amin, aden);
        o[i] = v;
    }
}
}  // namespace

int mosaic_u8_run(const float* before, const float* after, uint8_t* out, const Dims& d, int gap,
                  int gap_level, const uint2* mm_b, const uint2* mm_a, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (gap < 0 || gap > 1024) return set_error(MDIMG_ERR_INVALID, "mosaic: gap %d outside [0, 1024]", gap);
    const long long total = (long long)d.h * (2LL * d.w + gap);
    MDIMG_LAUNCH k_mosaic_u8<<<dim3(blocks_for(total, 8), d.n_sel), NT, 0, stream>>>(before, after, out, d, gap,
                                                                               gap_level & 255, mm_b, mm_a);
    return check_launch("mosaic");
}

int copy_run(const float* in, float* out, const Dims& d, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    CopyF f;
    MDIMG_LAUNCH k_map<CopyF, false><<<dim3(blocks_for(d.px(), 8), d.n_sel), NT, 0, stream>>>(in, nullptr, out, d, f);
    return check_launch("copy");
}

}  // namespace mdimg
