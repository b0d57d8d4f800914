// Total-variation denoise: skimage.restoration.denoise_tv_chambolle(image, weight,
// channel_axis=None) as called at pipeline/enhancement.py:311,349 (eps = 2e-4, at most 200
// iterations, float32 state).  Chambolle's projection, one kernel launch per iteration:
//
//   d   = -(p0 + p1);  d[1:, :] += p0[:-1, :];  d[:, 1:] += p1[:, :-1];  out = x + d
//   g0  = diff(out, axis 0) (last row 0);  g1 = diff(out, axis 1) (last column 0)
//   nrm = sqrt(g0^2 + g1^2);   E = (sum d^2 + w * sum nrm) / size
//   p   = (p - tau * g) / (1 + nrm * tau / w),   tau = 1/4
//   stop when |E_prev - E| < eps * E_0 (i > 0); the result is `out` of the last executed body.
//
// All per-pixel operations are float32 with numpy's rounding (explicit _rn intrinsics), so the
// field matches the reference bit for bit for a given iteration count.  The energies are summed
// in float64 (numpy: float32 pairwise) and rounded where numpy rounds; a borderline stop test can
// therefore differ by one iteration (tolerance documented in DESIGN.md and tests).
//
// p is ping-ponged between two buffers, so the `out` of the stopping iteration can be rebuilt
// from the buffer that iteration read.  Each slice stops independently: every CTA re-derives the
// stop decision from the per-iteration energy table, so no extra launch or host round trip is
// needed per iteration; the host polls the number of live slices every few iterations only to
// cut the launch sequence short.
#include "enhance.cuh"

namespace mdimg {

namespace {

constexpr int NT = 256;

struct TvState {             // per slice (position in sel)
    int stop_iter;           // index of the last executed loop body, -1 while running
    int pad;
};

__device__ __forceinline__ float energy_f32(double sum_d2, double sum_n, float w, float size_f) {
    float e = (float)sum_d2;
    e = __fadd_rn(e, __fmul_rn(w, (float)sum_n));
    return __fdiv_rn(e, size_f);
}

// true if the loop breaks at body j (j >= 1)
__device__ __forceinline__ bool tv_stops(const double* E, int j, float w, float size_f, float eps) {
    const float e0 = energy_f32(E[0], E[1], w, size_f);
    const float ep = energy_f32(E[2 * (j - 1)], E[2 * (j - 1) + 1], w, size_f);
    const float ej = energy_f32(E[2 * j], E[2 * j + 1], w, size_f);
    return fabsf(__fsub_rn(ep, ej)) < __fmul_rn(eps, e0);
}

// One warp owns a strip of TV_COLS output columns x TV_ROWS rows and walks down the rows keeping
// the previous / current / next row in registers.  x-neighbours come from warp shuffles: the warp
// loads 32 consecutive columns starting one to the left of its strip, so lane 0 and lane 31 only
// feed their neighbours (p1 of the column to the left, `out` of the column to the right).  No
// shared memory, no block barrier, one pointer bump per row.  Out-of-image loads read as 0, which
// reproduces the reference's slicing (d[1:] += p0[:-1] etc.) because x + 0.0f == x.
constexpr int TV_COLS = 30;
constexpr int TV_ROWS = 64;

__global__ void __launch_bounds__(NT)
k_tv_iter(const float* __restrict__ img, Dims d, int iter, const float* __restrict__ pin,
          float* __restrict__ pout, long long p_stride, double* __restrict__ energy, int max_iter,
          TvState* __restrict__ state, float w, float tau_over_w, float eps) {
    const int si = blockIdx.y;
    if (state[si].stop_iter >= 0) return;
    const int s = slice_of(d.sel, si);
    double* E = energy + (size_t)si * max_iter * 2;
    const float size_f = (float)((double)d.h * (double)d.w);
    if (iter >= 2 && tv_stops(E, iter - 1, w, size_f, eps)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) state[si].stop_iter = iter - 1;
        return;
    }
    const int lane = threadIdx.x & 31;
    const int strips_x = (d.w + TV_COLS - 1) / TV_COLS;
    const int strips_y = (d.h + TV_ROWS - 1) / TV_ROWS;
    const int wg = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (wg >= strips_x * strips_y) return;
    const int sy = wg / strips_x, sx = wg - sy * strips_x;
    const int c = sx * TV_COLS - 1 + lane;               // image column held by this lane
    const int y0 = sy * TV_ROWS;
    const int y1 = min(y0 + TV_ROWS, d.h);
    const bool col_in = (c >= 0) && (c < d.w);
    const bool owner = col_in && lane >= 1 && lane <= TV_COLS;   // lanes that store results
    const bool has_right = c < d.w - 1;
    const size_t plane = (size_t)d.h * d.w;
    const int cc = col_in ? c : 0;
    const float* src = img + (size_t)s * plane + cc;
    const float* p0 = pin + (size_t)si * p_stride + cc;
    const float* p1 = p0 + plane;
    float* q0 = pout + (size_t)si * p_stride + cc;
    float* q1 = q0 + plane;

    // row y0 - 1 (only p0 is needed) and row y0
    float up0 = (col_in && y0 > 0) ? p0[(size_t)(y0 - 1) * d.w] : 0.0f;
    size_t off = (size_t)y0 * d.w;
    float a0 = col_in ? p0[off] : 0.0f;
    float a1 = col_in ? p1[off] : 0.0f;
    float im = col_in ? src[off] : 0.0f;
    float left = __shfl_up_sync(0xffffffffu, a1, 1);
    float dcur = -__fadd_rn(a0, a1);
    dcur = __fadd_rn(dcur, up0);
    dcur = __fadd_rn(dcur, left);
    float ocur = __fadd_rn(im, dcur);
    float e_d2 = 0.0f, e_n = 0.0f;

    // software pipeline: row y+1 is already in registers when row y is processed, and the loads
    // of row y+2 are issued before any arithmetic of the iteration
    bool nin = col_in && (y0 + 1 < d.h);
    size_t noff = off + d.w;
    float n0 = nin ? p0[noff] : 0.0f;
    float n1 = nin ? p1[noff] : 0.0f;
    float nim = nin ? src[noff] : 0.0f;
#pragma unroll 2
    for (int y = y0; y < y1; ++y) {
        const size_t moff = noff + d.w;
        const bool min_ = col_in && (y + 2 < d.h) && (y + 1 < y1);
        const float m0 = min_ ? p0[moff] : 0.0f;
        const float m1 = min_ ? p1[moff] : 0.0f;
        const float mim = min_ ? src[moff] : 0.0f;

        const float nleft = __shfl_up_sync(0xffffffffu, n1, 1);
        float dn = -__fadd_rn(n0, n1);
        dn = __fadd_rn(dn, a0);
        dn = __fadd_rn(dn, nleft);
        const float onext = __fadd_rn(nim, dn);
        const float oright = __shfl_down_sync(0xffffffffu, ocur, 1);
        const float g0 = (y < d.h - 1) ? __fsub_rn(onext, ocur) : 0.0f;
        const float g1 = has_right ? __fsub_rn(oright, ocur) : 0.0f;
        float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(g0, g0), __fmul_rn(g1, g1)));
        if (owner) {
            e_d2 += __fmul_rn(dcur, dcur);
            e_n += nrm;
        }
        nrm = __fadd_rn(__fmul_rn(nrm, tau_over_w), 1.0f);
        const float r0 = __fdiv_rn(__fsub_rn(a0, __fmul_rn(0.25f, g0)), nrm);
        const float r1 = __fdiv_rn(__fsub_rn(a1, __fmul_rn(0.25f, g1)), nrm);
        if (owner) {
            q0[off] = r0;
            q1[off] = r1;
        }
        a0 = n0; a1 = n1; dcur = dn; ocur = onext; off = noff;
        n0 = m0; n1 = m1; nim = mim; noff = moff;
    }
    // per-lane float32 partial sums (<= TV_ROWS terms) -> float64 across the warp
    double sd = warp_sum((double)e_d2);
    double sn = warp_sum((double)e_n);
    if (lane == 0) {
        atomicAdd(&E[2 * iter], sd);
        atomicAdd(&E[2 * iter + 1], sn);
    }
}

__global__ void k_tv_count(int n_sel, const TvState* __restrict__ state, int* __restrict__ live) {
    int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n_sel) return;
    if (state[si].stop_iter < 0) atomicAdd(live, 1);
}

// out = x + div(p) from the buffer the last executed iteration read.
__global__ void __launch_bounds__(NT)
k_tv_final(const float* __restrict__ img, float* __restrict__ out, Dims d, const float* __restrict__ pa,
           const float* __restrict__ pb, long long p_stride, const double* __restrict__ energy,
           int max_iter, int launched, TvState* __restrict__ state, int* __restrict__ iters_out,
           float w, float eps) {
    const int si = blockIdx.y;
    const int s = slice_of(d.sel, si);
    int stop = state[si].stop_iter;
    if (stop < 0) {
        // the break of body launched-1 is only visible now; either way that body was the last one
        stop = launched - 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && iters_out) iters_out[s] = stop + 1;
    const size_t plane = (size_t)d.h * d.w;
    const float* src = img + (size_t)s * plane;
    const float* p0 = ((stop & 1) ? pb : pa) + (size_t)si * p_stride;
    const float* p1 = p0 + plane;
    float* dst = out + (size_t)s * plane;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < (long long)plane; i += (long long)gridDim.x * NT) {
        const int y = (int)(i / d.w), x = (int)(i - (long long)y * d.w);
        float dv = -__fadd_rn(p0[i], p1[i]);
        if (y > 0) dv = __fadd_rn(dv, p0[i - d.w]);
        if (x > 0) dv = __fadd_rn(dv, p1[i - 1]);
        dst[i] = __fadd_rn(src[i], dv);
    }
}

struct TvBufs { TvState* state; double* energy; float* pa; float* pb; int* live; };

void carve(Arena& a, int n_sel, int h, int w, int max_iter, TvBufs& b) {
    b.state = a.take<TvState>(n_sel);
    b.energy = a.take<double>((size_t)n_sel * max_iter * 2);
    b.pa = a.take<float>((size_t)n_sel * 2 * h * w);
    b.pb = a.take<float>((size_t)n_sel * 2 * h * w);
    b.live = a.take<int>(64);
}

}  // namespace

size_t tv_workspace_bytes(int n, int n_sel, int h, int w, int max_iter) {
    (void)n;
    Arena a(nullptr, 0);
    TvBufs b;
    carve(a, n_sel, h, w, max_iter < 1 ? 1 : max_iter, b);
    return a.off;
}

int tv_chambolle_run(const float* in, float* out, const Dims& d, double weight, double eps,
                     int max_iter, int* iters_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (d.n_sel == 0) return MDIMG_OK;
    if (max_iter < 1) return set_error(MDIMG_ERR_INVALID, "tv: max_iter must be >= 1");
    if (!(weight > 0.0)) return set_error(MDIMG_ERR_INVALID, "tv: weight must be > 0");
    Arena a(ws, ws_bytes);
    TvBufs b;
    carve(a, d.n_sel, d.h, d.w, max_iter, b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "tv: workspace too small (%zu > %zu)", a.off, ws_bytes);
    const long long p_stride = 2LL * d.h * d.w;
    cudaMemsetAsync(b.state, 0xFF, sizeof(TvState) * d.n_sel, stream);       // stop_iter = -1
    cudaMemsetAsync(b.energy, 0, sizeof(double) * 2 * max_iter * d.n_sel, stream);
    cudaMemsetAsync(b.pa, 0, sizeof(float) * p_stride * d.n_sel, stream);      // p = 0

    const float w = (float)weight;
    const float tau_over_w = (float)(0.25 / weight);
    const float epsf = (float)eps;
    const int n_warps = ((d.w + TV_COLS - 1) / TV_COLS) * ((d.h + TV_ROWS - 1) / TV_ROWS);
    dim3 grid((n_warps + NT / 32 - 1) / (NT / 32), d.n_sel);
    int launched = 0;
    int* live_host = nullptr;
    cudaError_t herr = cudaMallocHost((void**)&live_host, sizeof(int));
    if (herr != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "tv: cudaMallocHost failed: %s", cudaGetErrorString(herr));
    const int POLL = 8;
    for (int i = 0; i < max_iter; ++i) {
        const float* pin = (i & 1) ? b.pb : b.pa;
        float* pout = (i & 1) ? b.pa : b.pb;
        MDIMG_LAUNCH k_tv_iter<<<grid, NT, 0, stream>>>(in, d, i, pin, pout, p_stride, b.energy, max_iter, b.state,
                                           w, tau_over_w, epsf);
        launched = i + 1;
        if (eps > 0.0 && i >= 2 && (i % POLL) == 0 && i + 1 < max_iter) {
            cudaMemsetAsync(b.live, 0, sizeof(int), stream);
            MDIMG_LAUNCH k_tv_count<<<(d.n_sel + 127) / 128, 128, 0, stream>>>(d.n_sel, b.state, b.live);
            cudaMemcpyAsync(live_host, b.live, sizeof(int), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
            if (*live_host == 0) break;
        }
    }
    cudaFreeHost(live_host);
    int fb = (int)(((long long)d.h * d.w + NT * 8 - 1) / (NT * 8));
    if (fb > 4096) fb = 4096;
    MDIMG_LAUNCH k_tv_final<<<dim3(fb, d.n_sel), NT, 0, stream>>>(in, out, d, b.pa, b.pb, p_stride, b.energy, max_iter,
                                                      launched, b.state, iters_out, w, epsf);
    return check_launch("tv_chambolle");
}

}  // namespace mdimg
