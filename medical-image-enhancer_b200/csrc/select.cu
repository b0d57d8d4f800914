// Exact multi-query radix select (see select.cuh).
#include "select.cuh"

namespace mdimg {

namespace {

constexpr int SCAN_THREADS = 256;

// Inclusive scan of `nb` bins (nb = 256 * K) into smem `cum`; every thread of the block calls.
template <int K>
__device__ void block_inclusive_scan(const unsigned* __restrict__ hist, unsigned* cum,
                                     unsigned* warp_tot) {
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    unsigned loc[K];
    unsigned run = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        run += hist[t * K + k];
        loc[k] = run;
    }
    unsigned inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
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

__device__ void rebuild_unique(SelState& st, int Q, unsigned mask) {
    int nu = 0;
    for (int q = 0; q < Q; ++q) {
        unsigned p = st.prefix[q] & mask;
        int u = -1;
        for (int j = 0; j < nu; ++j) if (st.uprefix[j] == p) { u = j; break; }
        if (u < 0) { u = nu; st.uprefix[nu++] = p; }
        st.uid[q] = u;
    }
    st.nuniq = nu;
}

// LEVEL 1: hist = l1_hist[slice]; LEVEL 2/3: hist = lvl_hist[si][uid][...].
template <int LEVEL>
__global__ void __launch_bounds__(SCAN_THREADS)
k_sel_scan(Dims d, int len, int Q, const int* __restrict__ ranks,
           const unsigned* __restrict__ l1_hist, const unsigned* __restrict__ lvl_hist,
           SelState* __restrict__ states, float* __restrict__ out) {
    constexpr int NB = LEVEL == 1 ? SEL_L1_BINS : 2048;
    constexpr int SHIFT = LEVEL == 1 ? SEL_L1_SHIFT : (LEVEL == 2 ? 11 : 0);
    __shared__ unsigned cum[NB];
    __shared__ unsigned warp_tot[SCAN_THREADS / 32];
    __shared__ SelState st;
    const int si = blockIdx.x;
    const int s = slice_of(d.sel, si);
    if (threadIdx.x == 0) {
        if (LEVEL == 1) {
            st.valid = len > 0;
            for (int q = 0; q < Q; ++q) {
                st.prefix[q] = 0;
                st.rank[q] = ranks[(size_t)s * Q + q];
                st.uid[q] = 0;
            }
            st.nuniq = 1;
            st.uprefix[0] = 0;
        } else {
            st = states[si];
        }
    }
    __syncthreads();
    const int nu = st.nuniq;
    for (int u = 0; u < nu; ++u) {
        const unsigned* h = LEVEL == 1 ? l1_hist + (size_t)si * SEL_L1_BINS
                                       : lvl_hist + ((size_t)si * SEL_MAX_Q + u) * 2048;
        block_inclusive_scan<NB / SCAN_THREADS>(h, cum, warp_tot);
        if (threadIdx.x == 0) {
            for (int q = 0; q < Q; ++q) {
                if (st.uid[q] != u) continue;
                int r = st.rank[q];
                if (r < 0 || (unsigned)r >= cum[NB - 1]) { st.rank[q] = -1; continue; }
                int b = upper_bin(cum, NB, (unsigned)r);
                unsigned below = b ? cum[b - 1] : 0u;
                st.prefix[q] |= (unsigned)b << SHIFT;
                st.rank[q] = r - (int)below;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (LEVEL == 3) {
            for (int q = 0; q < Q; ++q)
                out[(size_t)s * Q + q] = (st.valid && st.rank[q] >= 0) ? key2f(st.prefix[q])
                                                                        : __int_as_float(0x7fc00000);
        } else {
            rebuild_unique(st, Q, LEVEL == 1 ? 0xFFC00000u : 0xFFFFF800u);
            states[si] = st;
        }
    }
}

// Histogram the next digit of every element whose resolved prefix matches a query.
// Almost no element matches: a 2048-bit map of the level-1 bins that hold a query is tested first
// (one shared-memory word + a shift), only hits walk the prefix list.  128-bit loads when aligned.
template <int LEVEL>
__global__ void __launch_bounds__(256)
k_sel_pass(const float* __restrict__ vals, long long stride, int len, Dims d, int opts,
           const SelState* __restrict__ states, unsigned* __restrict__ lvl_hist) {
    constexpr unsigned MASK = LEVEL == 2 ? 0xFFC00000u : 0xFFFFF800u;
    __shared__ unsigned up[SEL_MAX_Q];
    __shared__ unsigned bitmap[SEL_L1_BINS / 32];
    __shared__ int nu_s;
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    if (threadIdx.x < SEL_L1_BINS / 32) bitmap[threadIdx.x] = 0;
    if (threadIdx.x == 0) nu_s = states[si].nuniq;
    if (threadIdx.x < SEL_MAX_Q) up[threadIdx.x] = states[si].uprefix[threadIdx.x];
    __syncthreads();
    const int nu = nu_s;
    if (threadIdx.x < nu) {
        const unsigned bin = up[threadIdx.x] >> SEL_L1_SHIFT;
        atomicOr(&bitmap[bin >> 5], 1u << (bin & 31));
    }
    __syncthreads();
    const float* v = vals + (size_t)((opts & SEL_COMPACT) ? si : s) * stride;
    const bool use_abs = (opts & SEL_ABS) != 0;
    unsigned* hbase = lvl_hist + (size_t)si * SEL_MAX_Q * 2048;
    const int lane = threadIdx.x & 31;

    auto visit = [&](float f) {
        const unsigned key = f2key(use_abs ? fabsf(f) : f);
        const unsigned bin = key >> SEL_L1_SHIFT;
        if ((bitmap[bin >> 5] >> (bin & 31)) & 1u) {
            const unsigned pre = key & MASK;
            int u = -1;
            for (int j = 0; j < nu; ++j) if (up[j] == pre) u = j;
            if (u >= 0) {
                const unsigned digit = LEVEL == 2 ? ((key >> 11) & 0x7FFu) : (key & 0x7FFu);
                const unsigned slot = (unsigned)u * 2048u + digit;
                const unsigned am = __activemask();
                const unsigned peers = __match_any_sync(am, slot);
                if (lane == __ffs(peers) - 1) atomicAdd(hbase + slot, (unsigned)__popc(peers));
            }
        }
    };
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if ((((uintptr_t)v) & 15) == 0) {
        const int n4 = len >> 2;
        const float4* v4 = reinterpret_cast<const float4*>(v);
        for (int i = tid; i < n4; i += nthr) {
            const float4 q = v4[i];
            visit(q.x); visit(q.y); visit(q.z); visit(q.w);
        }
        for (int i = (n4 << 2) + tid; i < len; i += nthr) visit(v[i]);
    } else {
        for (int i = tid; i < len; i += nthr) visit(v[i]);
    }
}

}  // namespace

size_t select_workspace_bytes(int n_sel) {
    Arena a(nullptr, 0);
    a.take<SelState>(n_sel);
    a.take<unsigned>((size_t)n_sel * SEL_MAX_Q * 2048);
    return a.off;
}

int select_run(const float* vals, long long stride, int len, const Dims& d, int Q,
               const int* ranks, const unsigned* l1_hist, float* out,
               void* ws, size_t ws_bytes, cudaStream_t stream, int opts) {
    if (Q < 1 || Q > SEL_MAX_Q) return set_error(MDIMG_ERR_INVALID, "select: Q=%d out of range", Q);
    Arena a(ws, ws_bytes);
    SelState* states = a.take<SelState>(d.n_sel);
    size_t hist_elems = (size_t)d.n_sel * SEL_MAX_Q * 2048;
    unsigned* hist = a.take<unsigned>(hist_elems);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "select: workspace too small");
    if (d.n_sel == 0) return MDIMG_OK;

    int bx = (len + 256 * 8 - 1) / (256 * 8);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    dim3 pgrid(bx, d.n_sel);

    MDIMG_LAUNCH k_sel_scan<1><<<d.n_sel, SCAN_THREADS, 0, stream>>>(d, len, Q, ranks, l1_hist, nullptr, states, out);
    cudaMemsetAsync(hist, 0, hist_elems * sizeof(unsigned), stream);
    MDIMG_LAUNCH k_sel_pass<2><<<pgrid, 256, 0, stream>>>(vals, stride, len, d, opts, states, hist);
    MDIMG_LAUNCH k_sel_scan<2><<<d.n_sel, SCAN_THREADS, 0, stream>>>(d, len, Q, ranks, l1_hist, hist, states, out);
    cudaMemsetAsync(hist, 0, hist_elems * sizeof(unsigned), stream);
    MDIMG_LAUNCH k_sel_pass<3><<<pgrid, 256, 0, stream>>>(vals, stride, len, d, opts, states, hist);
    MDIMG_LAUNCH k_sel_scan<3><<<d.n_sel, SCAN_THREADS, 0, stream>>>(d, len, Q, ranks, l1_hist, hist, states, out);
    return check_launch("select");
}

}  // namespace mdimg
