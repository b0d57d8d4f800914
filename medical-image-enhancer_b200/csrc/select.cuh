// Exact multi-query order statistics on float32 arrays without sorting.
//
// Replaces np.percentile / np.median (numpy partition + lerp) at
// pipeline/metrics.py:70,77,134 and inside skimage's estimate_sigma (pipeline/metrics.py:47).
//
// Range select on the order-preserving uint32 image of the float bits:
//   level 1  a 1024-bin histogram over a monotone map of the value (sel_bin1: uniform bins over
//            (0, 1), where this pipeline's pixels, |gradients| and |wavelet coefficients| live; one
//            bin each for negatives, exact zero and values > 1).  It is accumulated by the kernel
//            that PRODUCES the data, so it costs no extra read.  A query whose bin holds a single key
//            (exact zero: CT air) is resolved here.
//   refine   up to three passes; each histograms, for the few elements that fall into the key range
//            [lo, lo + span] of an unresolved query, the digit (key - lo) >> shift with 2048 bins.
//            The last block of a slice to finish a pass does that slice's scan (no separate scan
//            launch) and narrows the ranges; a range of one key is resolved.  With uniform level-1
//            bins a range holds ~1 % of the elements, so the passes are plain streaming reads.
// Several arrays (jobs) are refined in the same launches (blockIdx.z = job).
#pragma once
#include "common.cuh"

namespace mdimg {

constexpr int SEL_L1_BINS = 1024;
constexpr int SEL_MAX_Q = 8;
constexpr int SEL_MAX_JOBS = 3;
constexpr int SEL_REFINE_BINS = 2048;
constexpr int SEL_COMPACT = 1;   // vals is indexed by position in `sel`, not by slice id
constexpr int SEL_ABS = 2;       // select on |v| (the level-1 histogram must be of |v| too)

#ifdef __CUDACC__
// Selection key: -0.0 is folded into +0.0 so that "exactly zero" is a single key.
__device__ __forceinline__ unsigned sel_key(float v) { return f2key(__fadd_rn(v, 0.0f)); }

// Level-1 bin, monotone non-decreasing in the value and branch-free: 0 negatives, 1 zero,
// 2..1022 uniform over (0, 1] (bin = 1 + ceil(1021 v)), 1023 above.  float->int conversion
// saturates; NaNs land in bin 1 and are never selected.
__device__ __forceinline__ int sel_bin1(float v) {
    const int c = __float2int_ru(__fmul_rn(v, 1021.0f));
    const int b = 1 + min(c, 1022);
    return v < 0.0f ? 0 : b;
}
#endif

struct SelState {
    int valid;                     // 0 => no elements (results are NaN)
    int nuniq;                     // unresolved unique key ranges
    int rank[SEL_MAX_Q];           // 0-based rank inside the query's current range; -1 => NaN result
    unsigned lo[SEL_MAX_Q];        // current key range [lo, lo + span] of the query
    unsigned span[SEL_MAX_Q];
    int uid[SEL_MAX_Q];            // index of the unique range the query shares; -1 once resolved
    unsigned ulo[SEL_MAX_Q];       // unique unresolved ranges
    unsigned uspan[SEL_MAX_Q];
    int ushift[SEL_MAX_Q];         // digit = (key - ulo) >> ushift
    int ubin[SEL_MAX_Q];           // level-1 bin of the range (prefilter bitmap)
};

struct SelJob {
    const float* vals;             // [slice][stride] floats, the first `len` of each slice are the data
    long long stride;
    int len;
    int Q;                         // queries per slice (<= SEL_MAX_Q)
    int opts;                      // SEL_COMPACT | SEL_ABS
    const int* ranks;              // device [n][Q] 0-based ranks into the sorted data (negative => NaN)
    const unsigned* l1_hist;       // device [n_sel][SEL_L1_BINS] histogram of sel_bin1, indexed by position in sel
    float* out;                    // device [n][Q] selected values
    void* ws;                      // select_workspace_bytes(n_sel) bytes
};

// Workspace bytes of one job for `n_sel` slices.
size_t select_workspace_bytes(int n_sel);

// Runs up to SEL_MAX_JOBS selections over the same slices in shared launches.
int select_run_multi(const SelJob* jobs, int njobs, const Dims& d, cudaStream_t stream);

// Single-array convenience wrapper.
int select_run(const float* vals, long long stride, int len, const Dims& d, int Q,
               const int* ranks, const unsigned* l1_hist, float* out,
               void* ws, size_t ws_bytes, cudaStream_t stream, int opts = 0);

}  // namespace mdimg
