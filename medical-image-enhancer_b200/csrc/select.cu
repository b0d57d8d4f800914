// Exact multi-query range select (see select.cuh).
#include "select.cuh"

namespace mdimg {

namespace {

constexpr int ST = 256;                       // threads per block (scan and refine)
constexpr unsigned KEY_NEG_INF = 0x007FFFFFu; // f2key(-inf)
constexpr unsigned KEY_POS_INF = 0xFF800000u; // f2key(+inf)

struct JobDev {
    const float* vals;
    long long stride;
    int len, Q, opts;
    const int* ranks;
    const unsigned* l1;
    float* out;
    SelState* states;
    unsigned* hist;        // [n_sel][SEL_MAX_Q][SEL_REFINE_BINS], zero between passes
    unsigned* counters;    // [n_sel] block tickets, zero between passes
};

struct Params {
    JobDev job[SEL_MAX_JOBS];
};

// Inclusive scan of 256 * K bins into smem `cum`; every thread of the block calls.
template <int K, typename Load>
__device__ void block_inclusive_scan(Load&& load, unsigned* cum, unsigned* warp_tot) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned loc[K];
    unsigned run = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        run += load(t * K + k);
        loc[k] = run;
    }
    unsigned inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();                 // previous users of cum / warp_tot are done
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < wid; ++w) base += warp_tot[w];
    unsigned excl = base + inc - run;
#pragma unroll
    for (int k = 0; k < K; ++k) cum[t * K + k] = excl + loc[k];
    __syncthreads();
}

// smallest b with cum[b] > rank
__device__ int upper_bin(const unsigned* cum, int nb, unsigned rank) {
    int lo = 0, hi = nb - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cum[mid] > rank) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// smallest key in [f2key(-inf), f2key(+inf) + 1] whose level-1 bin is >= b
__device__ unsigned first_key_of_bin(int b) {
    unsigned lo = KEY_NEG_INF, hi = KEY_POS_INF + 1u;
    while (lo < hi) {
        const unsigned mid = lo + ((hi - lo) >> 1);
        if (sel_bin1(key2f(mid)) >= b) hi = mid; else lo = mid + 1u;
    }
    return lo;
}

__device__ __forceinline__ int shift_for(unsigned span) {
    const int bits = 32 - __clz(span);            // span >= 1
    return bits > 11 ? bits - 11 : 0;
}

__device__ __forceinline__ float nan_f() { return __int_as_float(0x7fc00000); }

// Group the unresolved queries by key range (thread 0 only).
__device__ void rebuild_unique(SelState& st, int Q) {
    int nu = 0;
    for (int q = 0; q < Q; ++q) {
        if (st.uid[q] < 0) continue;
        int u = -1;
        for (int j = 0; j < nu; ++j)
            if (st.ulo[j] == st.lo[q] && st.uspan[j] == st.span[q]) { u = j; break; }
        if (u < 0) {
            u = nu++;
            st.ulo[u] = st.lo[q];
            st.uspan[u] = st.span[q];
            st.ushift[u] = shift_for(st.span[q]);
            st.ubin[u] = sel_bin1(key2f(st.lo[q]));
        }
        st.uid[q] = u;
    }
    st.nuniq = nu;
}

// Level 1: locate every query's bin in the producer's histogram.
__global__ void __launch_bounds__(ST)
k_sel_scan1(Params P, Dims d) {
    __shared__ unsigned cum[SEL_L1_BINS];
    __shared__ unsigned warp_tot[ST / 32];
    __shared__ SelState st;
    const JobDev& J = P.job[blockIdx.y];
    const int si = blockIdx.x;
    const int s = slice_of(d.sel, si);
    const unsigned* h = J.l1 + (size_t)si * SEL_L1_BINS;
    block_inclusive_scan<SEL_L1_BINS / ST>([&](int i) { return h[i]; }, cum, warp_tot);
    const int q = threadIdx.x;
    if (q < SEL_MAX_Q) {
        st.uid[q] = -1;
        st.rank[q] = -1;
        st.lo[q] = 0;
        st.span[q] = 0;
    }
    if (q < J.Q) {
        const int r = J.ranks[(size_t)s * J.Q + q];
        if (J.len <= 0 || r < 0 || (unsigned)r >= cum[SEL_L1_BINS - 1]) {
            J.out[(size_t)s * J.Q + q] = nan_f();
        } else {
            const int b = upper_bin(cum, SEL_L1_BINS, (unsigned)r);
            const unsigned below = b ? cum[b - 1] : 0u;
            unsigned lo, hi;
            if (b == 1) {
                lo = hi = 0x80000000u;                 // exact zero (sel_key folds -0.0 into +0.0)
            } else {
                lo = first_key_of_bin(b);
                hi = first_key_of_bin(b + 1) - 1u;
            }
            st.rank[q] = r - (int)below;
            st.lo[q] = lo;
            st.span[q] = hi - lo;
            if (hi == lo) J.out[(size_t)s * J.Q + q] = key2f(lo);
            else st.uid[q] = 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st.valid = J.len > 0;
        rebuild_unique(st, J.Q);
        J.states[si] = st;
    }
}

// Per-element work of a refinement pass: level-1 bin -> one byte lookup (12 instructions, branch
// free).  Only ~1 % of the elements fall into a bin that holds an unresolved query, but a warp sees
// at least one of them almost every other step, and running the range search + histogram update
// divergently for a single lane would cost more than the whole fast path.  Hits are therefore
// compacted into a per-warp queue (ballot + popc) and drained 32 at a time with all lanes busy.
constexpr int WQ = 64;                             // queue slots per warp: < 32 pending + one full ballot

struct RefineCtx {
    unsigned flag_sa;                              // shared-window address of the bin flag table
    unsigned* q;                                   // this warp's queue (shared memory)
    int qn;                                        // queued keys (warp-uniform)
    int nu;
    const unsigned* ulo; const unsigned* uspan; const int* ushift;
    unsigned* hbase;
    int lane;
    unsigned lo0, span0;                           // the slice's only unresolved range (ONE variants)
};

__device__ __forceinline__ void refine_account(const RefineCtx& c, unsigned key, unsigned mask) {
    for (int j = 0; j < c.nu; ++j) {
        const unsigned off = key - c.ulo[j];
        if (off <= c.uspan[j]) {
            const unsigned slot = (unsigned)j * SEL_REFINE_BINS + (off >> c.ushift[j]);
            const unsigned peers = __match_any_sync(__activemask(), slot);
            if (c.lane == __ffs(peers) - 1) atomicAdd(c.hbase + slot, (unsigned)__popc(peers));
            break;
        }
    }
    (void)mask;
}

__device__ __forceinline__ void refine_drain32(RefineCtx& c) {
    __syncwarp();
    refine_account(c, c.q[c.lane], 0xffffffffu);
    const int rest = c.qn - 32;
    __syncwarp();
    const unsigned t = c.lane < rest ? c.q[32 + c.lane] : 0u;
    __syncwarp();
    if (c.lane < rest) c.q[c.lane] = t;
    c.qn = rest;
}

// All 32 lanes call (valid = this lane holds an element).
// ONE: the slice has a single unresolved key range (always so for |grad| and |dd|, whose queries are one percentile):
// the element is tested against that range directly -- key, subtract, compare: 4 instructions instead of the 12 of the
// level-1 bin + flag lookup, and exact, so later passes queue nothing but the elements they must count.
template <bool ABS, bool ONE>
__device__ __forceinline__ void refine_visit(RefineCtx& c, float f, bool valid) {
    unsigned hit, key;
    if (ONE) {
        key = ABS ? (__float_as_uint(f) | 0x80000000u) : f2key(__fadd_rn(f, 0.0f));   // |f| >= 0: its key is bits | sign bit
        hit = (key - c.lo0) <= c.span0;
    } else {
        f = __fadd_rn(ABS ? fabsf(f) : f, 0.0f);
        // shared-window address kept in a register: ptxas otherwise rebuilds it per element
        asm("ld.shared.u8 %0, [%1];" : "=r"(hit) : "r"(c.flag_sa + (unsigned)sel_bin1(f)));
        key = f2key(f);
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid && hit);
    if (m) {                                       // warp-uniform
        if (valid && hit) c.q[c.qn + __popc(m & ((1u << c.lane) - 1u))] = key;
        c.qn += __popc(m);
        if (c.qn >= 32) refine_drain32(c);
    }
}

template <bool ABS, bool ONE>
__device__ __forceinline__ void refine_stream(RefineCtx& c, const float* __restrict__ v, int len, int tid, int nthr) {
    const int lane = c.lane;
    // head: the 0..3 elements in front of the first 16-byte boundary (slices of odd length -- the 257 x 257
    // 'dd' band of a 512 x 512 image -- start at any 4-byte offset), then 128-bit loads, then the tail
    int head = (int)(((16u - (unsigned)((uintptr_t)v & 15u)) & 15u) >> 2);
    if (head > len) head = len;
    if (tid - lane < head) {                       // first warp of the slice only (warp-uniform)
        const bool ok = tid < head;
        refine_visit<ABS, ONE>(c, ok ? v[tid] : 0.0f, ok);
    }
    const float* va = v + head;
    const int rem = len - head;
    const int n4 = rem >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(va);
    int i = tid;
    // four independent 128-bit loads in flight while the whole warp is in range (warp-uniform test)
    for (; i - lane + 31 + 3 * nthr < n4; i += 4 * nthr) {
        const float4 a = v4[i], b = v4[i + nthr], d = v4[i + 2 * nthr], e = v4[i + 3 * nthr];
        refine_visit<ABS, ONE>(c, a.x, true); refine_visit<ABS, ONE>(c, a.y, true); refine_visit<ABS, ONE>(c, a.z, true); refine_visit<ABS, ONE>(c, a.w, true);
        refine_visit<ABS, ONE>(c, b.x, true); refine_visit<ABS, ONE>(c, b.y, true); refine_visit<ABS, ONE>(c, b.z, true); refine_visit<ABS, ONE>(c, b.w, true);
        refine_visit<ABS, ONE>(c, d.x, true); refine_visit<ABS, ONE>(c, d.y, true); refine_visit<ABS, ONE>(c, d.z, true); refine_visit<ABS, ONE>(c, d.w, true);
        refine_visit<ABS, ONE>(c, e.x, true); refine_visit<ABS, ONE>(c, e.y, true); refine_visit<ABS, ONE>(c, e.z, true); refine_visit<ABS, ONE>(c, e.w, true);
    }
    for (; i - lane < n4; i += nthr) {
        const bool ok = i < n4;
        const float4 q = ok ? v4[i] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        refine_visit<ABS, ONE>(c, q.x, ok); refine_visit<ABS, ONE>(c, q.y, ok); refine_visit<ABS, ONE>(c, q.z, ok); refine_visit<ABS, ONE>(c, q.w, ok);
    }
    for (int k = (n4 << 2) + tid; k - lane < rem; k += nthr) {
        const bool ok = k < rem;
        refine_visit<ABS, ONE>(c, ok ? va[k] : 0.0f, ok);
    }
    __syncwarp();
    if (lane < c.qn) refine_account(c, c.q[lane], 0u);            // leftovers
    __syncwarp();
    c.qn = 0;
}

// One refinement pass (+ the slice's scan, done by the last block to finish).
__global__ void __launch_bounds__(ST)
k_sel_refine(Params P, Dims d) {
    __shared__ unsigned ulo[SEL_MAX_Q], uspan[SEL_MAX_Q];
    __shared__ int ushift[SEL_MAX_Q];
    __shared__ __align__(16) unsigned char flag[SEL_L1_BINS];
    __shared__ int nu_s, last_s;
    __shared__ unsigned cum[SEL_REFINE_BINS];
    __shared__ unsigned warp_tot[ST / 32];
    __shared__ unsigned queue[ST / 32][WQ];
    __shared__ SelState st;
    const JobDev& J = P.job[blockIdx.z];
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    const SelState* gst = J.states + si;
    reinterpret_cast<unsigned*>(flag)[threadIdx.x] = 0;           // 256 threads x 4 bytes = SEL_L1_BINS
    static_assert(SEL_L1_BINS == ST * 4, "flag table is cleared one word per thread");
    if (threadIdx.x == 0) nu_s = gst->nuniq;
    __syncthreads();
    const int nu = nu_s;
    if (nu == 0) return;                           // every query of this slice is resolved
    if (threadIdx.x < nu) {
        ulo[threadIdx.x] = gst->ulo[threadIdx.x];
        uspan[threadIdx.x] = gst->uspan[threadIdx.x];
        ushift[threadIdx.x] = gst->ushift[threadIdx.x];
        flag[gst->ubin[threadIdx.x]] = 1;
    }
    __syncthreads();
    const float* v = J.vals + (size_t)((J.opts & SEL_COMPACT) ? si : s) * J.stride;
    unsigned* hbase = J.hist + (size_t)si * SEL_MAX_Q * SEL_REFINE_BINS;
    const int lane = threadIdx.x & 31;
    const int tid = blockIdx.x * ST + threadIdx.x, nthr = gridDim.x * ST;
    RefineCtx c;
    // opaque copy: ptxas would otherwise rematerialise the shared window base per element
    asm volatile("mov.u32 %0, %1;" : "=r"(c.flag_sa) : "r"((unsigned)__cvta_generic_to_shared(flag)));
    c.q = queue[threadIdx.x >> 5];
    c.qn = 0;
    c.nu = nu; c.ulo = ulo; c.uspan = uspan; c.ushift = ushift; c.hbase = hbase; c.lane = lane;
    c.lo0 = ulo[0]; c.span0 = uspan[0];
    const bool one = nu == 1;                      // block-uniform
    if (J.opts & SEL_ABS) {
        if (one) refine_stream<true, true>(c, v, J.len, tid, nthr);
        else refine_stream<true, false>(c, v, J.len, tid, nthr);
    } else {
        if (one) refine_stream<false, true>(c, v, J.len, tid, nthr);
        else refine_stream<false, false>(c, v, J.len, tid, nthr);
    }

    // ---- last block of this slice: scan the digit histograms, narrow the ranges ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_s = (atomicAdd(J.counters + si, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last_s) return;
    __threadfence();
    if (threadIdx.x == 0) st = *gst;
    __syncthreads();
    for (int u = 0; u < nu; ++u) {
        unsigned* row = hbase + (size_t)u * SEL_REFINE_BINS;
        block_inclusive_scan<SEL_REFINE_BINS / ST>([&](int i) { return __ldcg(row + i); }, cum, warp_tot);
        const int q = threadIdx.x;
        if (q < J.Q && st.uid[q] == u) {
            const int r = st.rank[q];
            if (r < 0 || (unsigned)r >= cum[SEL_REFINE_BINS - 1]) {     // cannot happen for consistent histograms
                st.rank[q] = -1;
                st.uid[q] = -1;
                J.out[(size_t)s * J.Q + q] = nan_f();
            } else {
                const int b = upper_bin(cum, SEL_REFINE_BINS, (unsigned)r);
                const unsigned below = b ? cum[b - 1] : 0u;
                const int sh = st.ushift[u];
                const unsigned base = (unsigned)b << sh;
                const unsigned width = sh ? ((1u << sh) - 1u) : 0u;
                st.rank[q] = r - (int)below;
                st.lo[q] = st.ulo[u] + base;
                st.span[q] = min(st.uspan[u] - base, width);
                if (st.span[q] == 0) {
                    st.uid[q] = -1;
                    J.out[(size_t)s * J.Q + q] = key2f(st.lo[q]);
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < SEL_REFINE_BINS; i += ST) row[i] = 0;   // ready for the next pass
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        rebuild_unique(st, J.Q);
        J.states[si] = st;
        J.counters[si] = 0;
    }
}

struct Carve { SelState* states; unsigned* hist; unsigned* counters; size_t zero_off, zero_bytes; };

void carve(Arena& a, int n_sel, Carve& c) {
    c.states = a.take<SelState>(n_sel);
    c.zero_off = a.off;
    c.hist = a.take<unsigned>((size_t)n_sel * SEL_MAX_Q * SEL_REFINE_BINS);
    c.counters = a.take<unsigned>(n_sel);
    c.zero_bytes = a.off - c.zero_off;
}

}  // namespace

size_t select_workspace_bytes(int n_sel) {
    Arena a(nullptr, 0);
    Carve c;
    carve(a, n_sel, c);
    return a.off;
}

int select_run_multi(const SelJob* jobs, int njobs, const Dims& d, cudaStream_t stream) {
    if (njobs < 1 || njobs > SEL_MAX_JOBS) return set_error(MDIMG_ERR_INVALID, "select: %d jobs", njobs);
    if (d.n_sel == 0) return MDIMG_OK;
    Params P;
    int max_len = 1;
    const size_t need = select_workspace_bytes(d.n_sel);
    for (int j = 0; j < njobs; ++j) {
        const SelJob& in = jobs[j];
        if (in.Q < 1 || in.Q > SEL_MAX_Q) return set_error(MDIMG_ERR_INVALID, "select: Q=%d out of range", in.Q);
        Arena a(in.ws, need);
        Carve c;
        carve(a, d.n_sel, c);
        JobDev& o = P.job[j];
        o.vals = in.vals; o.stride = in.stride; o.len = in.len; o.Q = in.Q; o.opts = in.opts;
        o.ranks = in.ranks; o.l1 = in.l1_hist; o.out = in.out;
        o.states = c.states; o.hist = c.hist; o.counters = c.counters;
        cudaMemsetAsync((char*)in.ws + c.zero_off, 0, c.zero_bytes, stream);
        if (in.len > max_len) max_len = in.len;
    }
    for (int j = njobs; j < SEL_MAX_JOBS; ++j) P.job[j] = P.job[0];
    // few, long-lived blocks per slice: the per-block cost (state load, fence, ticket) is a few
    // microseconds of dependent latency
    int bx = (max_len + ST * 64 - 1) / (ST * 64);
    if (bx < 1) bx = 1;
    if (bx > 256) bx = 256;
    MDIMG_LAUNCH k_sel_scan1<<<dim3(d.n_sel, njobs), ST, 0, stream>>>(P, d);
    // key ranges of a level-1 bin span at most 2^31 keys: three 11-bit refinements always resolve
    for (int level = 0; level < 3; ++level)
        MDIMG_LAUNCH k_sel_refine<<<dim3(bx, d.n_sel, njobs), ST, 0, stream>>>(P, d);
    return check_launch("select");
}

int select_run(const float* vals, long long stride, int len, const Dims& d, int Q,
               const int* ranks, const unsigned* l1_hist, float* out,
               void* ws, size_t ws_bytes, cudaStream_t stream, int opts) {
    if (ws_bytes < select_workspace_bytes(d.n_sel)) return set_error(MDIMG_ERR_WORKSPACE, "select: workspace too small");
    SelJob j;
    j.vals = vals; j.stride = stride; j.len = len; j.Q = Q; j.opts = opts;
    j.ranks = ranks; j.l1_hist = l1_hist; j.out = out; j.ws = ws;
    return select_run_multi(&j, 1, d, stream);
}

}  // namespace mdimg
