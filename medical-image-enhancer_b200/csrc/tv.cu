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
constexpr int TW = 64, TH = 32;
constexpr int PW = TW + 2, PH = TH + 2, PP = PW + 1;   // p tiles with a 1-pixel halo
constexpr int OW = TW + 1, OH = TH + 1, OP = OW + 1;   // out tile with +1 row / column

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

__device__ __forceinline__ float div_at(const float (*P0)[PP], const float (*P1)[PP], int r, int c,
                                        int gy, int gx) {
    // r, c index the halo tiles (tile origin at [1][1]); gy, gx are image coordinates
    float dv = -__fadd_rn(P0[r][c], P1[r][c]);
    if (gy > 0) dv = __fadd_rn(dv, P0[r - 1][c]);
    if (gx > 0) dv = __fadd_rn(dv, P1[r][c - 1]);
    return dv;
}

__global__ void __launch_bounds__(NT)
k_tv_iter(const float* __restrict__ img, Dims d, int iter, const float* __restrict__ pin,
          float* __restrict__ pout, long long p_stride, double* __restrict__ energy, int max_iter,
          TvState* __restrict__ state, float w, float tau_over_w, float eps) {
    __shared__ float P0[PH][PP], P1[PH][PP];
    __shared__ float O[OH][OP];
    __shared__ double red[2 * 32];
    const int si = blockIdx.y;
    if (state[si].stop_iter >= 0) return;
    const int s = slice_of(d.sel, si);
    double* E = energy + (size_t)si * max_iter * 2;
    const float size_f = (float)((double)d.h * (double)d.w);
    if (iter >= 2 && tv_stops(E, iter - 1, w, size_f, eps)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) state[si].stop_iter = iter - 1;
        return;
    }
    const int tiles_x = (d.w + TW - 1) / TW;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int x0 = tx * TW, y0 = ty * TH;
    const size_t plane = (size_t)d.h * d.w;
    const float* src = img + (size_t)s * plane;
    const float* p0 = pin + (size_t)si * p_stride;
    const float* p1 = p0 + plane;
    float* q0 = pout + (size_t)si * p_stride;
    float* q1 = q0 + plane;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    for (int i = tid; i < PH * PW; i += NT) {
        int r = i / PW, c = i - r * PW;
        int gy = y0 + r - 1, gx = x0 + c - 1;
        float a = 0.0f, b = 0.0f;
        if (gy >= 0 && gy < d.h && gx >= 0 && gx < d.w) {
            a = p0[(size_t)gy * d.w + gx];
            b = p1[(size_t)gy * d.w + gx];
        }
        P0[r][c] = a;
        P1[r][c] = b;
    }
    __syncthreads();
    // out = x + d on the tile plus one extra row / column
    double acc[2] = {0.0, 0.0};
    for (int i = tid; i < OH * OW; i += NT) {
        int r = i / OW, c = i - r * OW;
        int gy = y0 + r, gx = x0 + c;
        float o = 0.0f;
        if (gy < d.h && gx < d.w) {
            const float dv = div_at(P0, P1, r + 1, c + 1, gy, gx);
            o = __fadd_rn(src[(size_t)gy * d.w + gx], dv);
            if (r < TH && c < TW) acc[0] += (double)__fmul_rn(dv, dv);
        }
        O[r][c] = o;
    }
    __syncthreads();
#pragma unroll
    for (int j2 = 0; j2 < TH / 8; ++j2)
#pragma unroll
        for (int i2 = 0; i2 < TW / 32; ++i2) {
            const int r = wid + 8 * j2, c = lane + 32 * i2;
            const int gy = y0 + r, gx = x0 + c;
            if (gy < d.h && gx < d.w) {
                const float o = O[r][c];
                const float g0 = gy < d.h - 1 ? __fsub_rn(O[r + 1][c], o) : 0.0f;
                const float g1 = gx < d.w - 1 ? __fsub_rn(O[r][c + 1], o) : 0.0f;
                float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(g0, g0), __fmul_rn(g1, g1)));
                acc[1] += (double)nrm;
                nrm = __fadd_rn(__fmul_rn(nrm, tau_over_w), 1.0f);
                const float n0 = __fdiv_rn(__fsub_rn(P0[r + 1][c + 1], __fmul_rn(0.25f, g0)), nrm);
                const float n1 = __fdiv_rn(__fsub_rn(P1[r + 1][c + 1], __fmul_rn(0.25f, g1)), nrm);
                q0[(size_t)gy * d.w + gx] = n0;
                q1[(size_t)gy * d.w + gx] = n1;
            }
        }
    block_sum<2>(acc, red);
    if (tid == 0) {
        atomicAdd(&E[2 * iter], acc[0]);
        atomicAdd(&E[2 * iter + 1], acc[1]);
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
    dim3 grid(((d.w + TW - 1) / TW) * ((d.h + TH - 1) / TH), d.n_sel);
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
