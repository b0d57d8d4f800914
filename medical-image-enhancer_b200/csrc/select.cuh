// Exact multi-query order statistics on float32 arrays without sorting.
//
// Replaces np.percentile / np.median (numpy partition + lerp) at
// pipeline/metrics.py:70,77,134 and inside skimage's estimate_sigma (pipeline/metrics.py:47).
// Three-level radix select (10 + 11 + 11 bits, MSB first) on the order-preserving uint32
// image of the float bits.  Level 1 histograms are produced by the kernel that generates the
// data (one read of the image); levels 2 and 3 only touch the few elements whose prefix
// matches one of the queries.
#pragma once
#include "common.cuh"

namespace mdimg {

constexpr int SEL_L1_BINS = 1024;     // level 1: top 10 key bits (kept small: it lives in the producers' shared memory)
constexpr int SEL_L1_SHIFT = 22;       // levels 2 and 3 resolve 11 bits each
constexpr int SEL_MAX_Q = 8;
constexpr int SEL_COMPACT = 1;   // vals is indexed by position in `sel`, not by slice id
constexpr int SEL_ABS = 2;       // select on |v| (the level-1 histogram must be of |v| too)

struct SelState {
    unsigned prefix[SEL_MAX_Q];   // key bits resolved so far (low bits zero)
    int rank[SEL_MAX_Q];          // remaining 0-based rank inside the prefix bucket
    int uid[SEL_MAX_Q];           // index of the unique prefix this query shares
    int nuniq;
    unsigned uprefix[SEL_MAX_Q];  // unique prefixes
    int valid;                    // 0 => no elements (results are NaN)
};

// Workspace bytes for `n_sel` slices.
size_t select_workspace_bytes(int n_sel);

// vals: [slice][stride] floats, first `len` of each slice are the data.
// ranks: device [n][Q] 0-based ranks into the sorted data (negative => result NaN).
// l1_hist: device [n_sel][SEL_L1_BINS] level-1 histogram of (f2key(v) >> SEL_L1_SHIFT), already filled,
//          indexed by position in `sel`.
// out: device [n][Q] selected values.
int select_run(const float* vals, long long stride, int len, const Dims& d, int Q,
               const int* ranks, const unsigned* l1_hist, float* out,
               void* ws, size_t ws_bytes, cudaStream_t stream, int opts = 0);

}  // namespace mdimg
