// mdimg_enhance: apply_enhancements_from_params (pipeline/enhancement.py:235-369) for a whole stack in
// one C-ABI call.  Host-side control flow only -- every pixel is touched by the step entry points of
// api.cu; what lives here is what the reference does in Python between its library calls: clamping to
// PARAM_BOUNDS, step gating in the fixed order, the final clip, and the three safeguards with their
// per-slice decisions (which need a handful of small device-to-host reads, hence the stream syncs).
// The Python engine (engine.py: Engine.enhance_from_params) is the same logic on torch tensors; the
// GPU tests require the two to agree pixel for pixel and flag for flag.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/mdimg_b200.h"
#include "common.cuh"
#include "metrics.cuh"

namespace mdimg {

namespace {

constexpr int MAX_CHECKS = 40;        // flag arrays: <= 16 listed ops, twice (main pass + halo re-run), + slack

struct Bufs {
    float* tmp;
    double* rows;          // metrics rows of the current image (used when the caller passes no rows_after)
    double* rows_before;   // metrics rows of the input (used when the caller passes none)
    double* sigma;         // [n]
    double* quality;       // [n][2]
    int32_t* iters;        // [n] TV iterations, main pass
    int32_t* iters2;       // [n] TV iterations, halo re-run
    int32_t* skipped;      // [n] light-denoise skip flags (unused output)
    int32_t* sel;          // [n] slice list of the current safeguard
    int32_t* checks;       // [MAX_CHECKS][n] per-call ValueError flags
    void* sub;             // workspace of the step entry points
    size_t sub_bytes;
};

size_t sub_workspace(int n, int h, int w, int k) {
    const int ops[] = {MDIMG_OP_METRICS, MDIMG_OP_SIGMA, MDIMG_OP_QUALITY, MDIMG_OP_WAVELET, MDIMG_OP_GAMMA,
                       MDIMG_OP_UNSHARP, MDIMG_OP_LIGHT_DENOISE};
    size_t m = mdimg_workspace_bytes(MDIMG_OP_CLAHE, n, h, w, k);
    const size_t tv = mdimg_workspace_bytes(MDIMG_OP_TV, n, h, w, 200);
    if (tv > m) m = tv;
    for (int op : ops) {
        const size_t b = mdimg_workspace_bytes(op, n, h, w, 0);
        if (b > m) m = b;
    }
    return m;
}

void carve(Arena& a, int n, int h, int w, int k, Bufs& b) {
    b.tmp = a.take<float>((size_t)n * h * w);
    b.rows = a.take<double>((size_t)n * MC_COLS);
    b.rows_before = a.take<double>((size_t)n * MC_COLS);
    b.sigma = a.take<double>(n);
    b.quality = a.take<double>((size_t)n * 2);
    b.iters = a.take<int32_t>(n);
    b.iters2 = a.take<int32_t>(n);
    b.skipped = a.take<int32_t>(n);
    b.sel = a.take<int32_t>(n);
    b.checks = a.take<int32_t>((size_t)MAX_CHECKS * n);
    b.sub_bytes = sub_workspace(n, h, w, k);
    b.sub = a.take<char>(b.sub_bytes);
}

double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   // max(lo, min(hi, v))

bool enabled(int step, const mdimg_enhance_plan& q) {
    switch (step) {
        case MDIMG_STEP_GAMMA: return std::fabs(q.gamma - 1.0) > 1e-4;
        case MDIMG_STEP_POST_DENOISE: return q.post_denoise_strength > 0;
        case MDIMG_STEP_BILATERAL: return q.bilateral_d > 0;
        case MDIMG_STEP_TV_DENOISE: return q.tv_denoise_weight > 0;
        default: return step >= MDIMG_STEP_DENOISE && step <= MDIMG_STEP_TV_DENOISE;
    }
}

struct Check { const int32_t* flags; int bit; };

struct Pass {                 // what the Python engine keeps in its `state` dict
    bool nonneg = false;
    bool fuse_gamma = false;
    bool gamma_done = false;
    bool tv_ran = false;
    std::vector<Check> checks;
};

struct Ctx {
    int n, h, w;
    const float* in;
    mdimg_enhance_plan q;
    const mdimg_enhance_tables* t;
    Bufs b;
    void* stream;
    int next_check = 0;

    int32_t* check_array() {
        if (next_check >= MAX_CHECKS) return nullptr;
        int32_t* p = b.checks + (size_t)(next_check++) * n;
        cudaMemsetAsync(p, 0, sizeof(int32_t) * n, (cudaStream_t)stream);
        return p;
    }
    int fetch(void* host, const void* dev, size_t bytes) {
        cudaError_t e = cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "enhance: %s", cudaGetErrorString(e));
        return MDIMG_OK;
    }
    // device slice list of the set entries of `mask`; returns the count
    int select(const std::vector<char>& mask) {
        std::vector<int32_t> idx;
        for (int i = 0; i < n; ++i)
            if (mask[i]) idx.push_back(i);
        if (!idx.empty()) {
            // the list must outlive the asynchronous copy: synchronise before `idx` goes away
            cudaMemcpyAsync(b.sel, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream);
            cudaStreamSynchronize((cudaStream_t)stream);
        }
        return (int)idx.size();
    }
};

// One step on the slices in `sel` (all when null); `cur` / `tmp` may be swapped when the whole stack was
// written out of place.
int apply_step(Ctx& c, int step, double u_amount, float*& cur, float*& tmp, const int32_t* sel, int n_sel,
               Pass& st, int32_t* iters) {
    const mdimg_enhance_plan& q = c.q;
    const int n = c.n, h = c.h, w = c.w;
    void* ws = c.b.sub;
    const size_t wsb = c.b.sub_bytes;
    int rc = MDIMG_OK;
    switch (step) {
        case MDIMG_STEP_DENOISE:
            rc = mdimg_wavelet_denoise(cur, cur, n, h, w, sel, n_sel, q.denoise_hard ? 1 : 0, nullptr, 1.0, nullptr, ws, wsb, c.stream);
            st.nonneg = false;
            break;
        case MDIMG_STEP_CLAHE: {
            // an adjust_gamma that directly follows CLAHE is folded into CLAHE's final table pass
            const double g = st.fuse_gamma ? q.gamma : 1.0;
            int32_t* status = c.check_array();
            if (!status) return set_error(MDIMG_ERR_INVALID, "enhance: too many checked steps in one plan");
            rc = mdimg_clahe_gamma(cur, cur, n, h, w, sel, n_sel, q.clahe_clip_limit, q.clahe_tile_size, g, status, ws, wsb, c.stream);
            st.checks.push_back({status, MDIMG_FLAG_ERR_CLAHE_RANGE});
            st.nonneg = true;
            st.gamma_done = g != 1.0;
            break;
        }
        case MDIMG_STEP_GAMMA: {
            if (st.gamma_done) { st.gamma_done = false; break; }     // already applied inside the CLAHE call
            int32_t* neg = c.check_array();
            if (!neg) return set_error(MDIMG_ERR_INVALID, "enhance: too many checked steps in one plan");
            rc = mdimg_gamma(cur, cur, n, h, w, sel, n_sel, q.gamma, st.nonneg ? 1 : 0, neg, ws, wsb, c.stream);
            if (!st.nonneg) st.checks.push_back({neg, MDIMG_FLAG_ERR_GAMMA_NEG});
            st.nonneg = true;
            break;
        }
        case MDIMG_STEP_UNSHARP:
            rc = mdimg_unsharp(cur, tmp, n, h, w, sel, n_sel, c.t->gauss_taps, c.t->gauss_radius, u_amount,
                               st.nonneg ? 1 : 0, ws, wsb, c.stream);
            if (rc) return rc;
            if (!sel) std::swap(cur, tmp);
            else rc = mdimg_copy(tmp, cur, n, h, w, sel, n_sel, c.stream);
            break;
        case MDIMG_STEP_POST_DENOISE:
            rc = mdimg_light_denoise(cur, cur, n, h, w, sel, n_sel, q.post_denoise_strength, c.b.skipped, ws, wsb, c.stream);
            st.nonneg = false;
            break;
        case MDIMG_STEP_BILATERAL:
            rc = mdimg_bilateral(cur, tmp, n, h, w, sel, n_sel, c.t->bilateral_d_eff, c.t->bilateral_spatial,
                                 q.bilateral_sigma_color, c.stream);
            if (rc) return rc;
            if (!sel) std::swap(cur, tmp);
            else rc = mdimg_copy(tmp, cur, n, h, w, sel, n_sel, c.stream);
            break;
        case MDIMG_STEP_TV_DENOISE:
            cudaMemsetAsync(iters, 0, sizeof(int32_t) * n, (cudaStream_t)c.stream);
            rc = mdimg_tv_chambolle(cur, tmp, n, h, w, sel, n_sel, q.tv_denoise_weight, 2.0e-4, 200, iters, ws, wsb, c.stream);
            if (rc) return rc;
            if (!sel) std::swap(cur, tmp);
            else rc = mdimg_copy(tmp, cur, n, h, w, sel, n_sel, c.stream);
            st.tv_ran = true;
            st.nonneg = false;
            break;
        default: break;
    }
    return rc;
}

// First failed check of a pass wins per slice (the reference raises at the first failing call); a later
// pass (the halo re-run) replaces the entry of an earlier one, as the Python engine's dict update does.
int flush_checks(Ctx& c, Pass& st, std::vector<int32_t>& err) {
    std::vector<int32_t> host(c.n), first(c.n, 0);
    for (const Check& k : st.checks) {
        int rc = c.fetch(host.data(), k.flags, sizeof(int32_t) * c.n);
        if (rc) return rc;
        for (int i = 0; i < c.n; ++i)
            if (host[i] != 0 && first[i] == 0) first[i] = k.bit;
    }
    for (int i = 0; i < c.n; ++i)
        if (first[i]) err[i] = first[i];
    st.checks.clear();
    return MDIMG_OK;
}

}  // namespace

size_t enhance_workspace_bytes(int n, int h, int w, int k) {
    Arena a(nullptr, 0);
    Bufs b;
    carve(a, n, h, w, k < 1 ? 16 : k, b);
    return a.off;
}

}  // namespace mdimg

using namespace mdimg;

extern "C" {

int mdimg_plan_clamp(mdimg_enhance_plan* p) {
    if (!p) return set_error(MDIMG_ERR_INVALID, "plan missing");
    // PARAM_BOUNDS (pipeline/schemas.py:16-28); integer parameters are clamped, then truncated like int()
    p->clahe_clip_limit = clampd(p->clahe_clip_limit, 0.002, 0.08);
    p->clahe_tile_size = (int32_t)clampd((double)p->clahe_tile_size, 4, 48);
    p->gamma = clampd(p->gamma, 0.6, 1.5);
    p->unsharp_radius = clampd(p->unsharp_radius, 0.2, 3.0);
    p->unsharp_amount = clampd(p->unsharp_amount, 0.03, 2.5);
    p->denoise_hard = p->denoise_hard ? 1 : 0;
    p->post_denoise_strength = clampd(p->post_denoise_strength, 0.0, 0.8);
    p->bilateral_d = (int32_t)clampd((double)p->bilateral_d, 0, 13);
    p->bilateral_sigma_color = clampd(p->bilateral_sigma_color, 0.005, 0.20);
    p->bilateral_sigma_space = clampd(p->bilateral_sigma_space, 0.005, 0.20);
    p->tv_denoise_weight = clampd(p->tv_denoise_weight, 0.0, 0.15);
    if (p->n_ops < 0) p->n_ops = 0;
    if (p->n_ops > MDIMG_MAX_PLAN_OPS)      // never truncate: membership and the halo re-run depend on every entry
        return set_error(MDIMG_ERR_INVALID, "plan lists %d operations, capacity is %d", p->n_ops, MDIMG_MAX_PLAN_OPS);
    return MDIMG_OK;
}

int mdimg_enhance_tables_default(const mdimg_enhance_plan* plan, int h, int w, mdimg_enhance_tables* t) {
    if (!plan || !t || h < 1 || w < 1) return set_error(MDIMG_ERR_INVALID, "enhance tables: bad arguments");
    std::memset(t, 0, sizeof(*t));
    // scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius), radius = int(4 sigma + 0.5)
    const double sigma = plan->unsharp_radius;
    const int radius = (int)(4.0 * sigma + 0.5);
    if (radius < 1 || radius > 12) return set_error(MDIMG_ERR_INVALID, "unsharp radius %d outside [1, 12]", radius);
    double phi[25];
    const int np_ = 2 * radius + 1;
    for (int x = -radius; x <= radius; ++x) phi[x + radius] = std::exp(-0.5 / (sigma * sigma) * (double)(x * x));
    // phi.sum() in numpy's order: below 8 elements a plain loop, otherwise eight running sums over blocks
    // of eight, combined pairwise, then the tail
    double sum = 0.0;
    if (np_ < 8) {
        for (int i = 0; i < np_; ++i) sum += phi[i];
    } else {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = phi[j];
        int i = 8;
        for (; i < np_ - (np_ % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += phi[i + j];
        sum = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < np_; ++i) sum += phi[i];
    }
    t->gauss_radius = radius;
    for (int j = 0; j <= radius; ++j) t->gauss_taps[j] = phi[radius + j] / sum;
    // _bilateral_filter (pipeline/enhancement.py:117-128)
    int d = plan->bilateral_d < 9 ? plan->bilateral_d : 9;
    if (d > 0 && d % 2 == 0) d += 1;
    t->bilateral_d_eff = d > 0 ? d : 0;
    if (d > 0) {
        const int r = d / 2;
        const double ss = plan->bilateral_sigma_space;
        const double den = (2.0 * (ss * ss)) * (double)(d * d);          // 2 * sigma_space**2 * d**2, python's order
        for (int y = -r; y <= r; ++y)
            for (int x = -r; x <= r; ++x)
                t->bilateral_spatial[(y + r) * d + (x + r)] = std::exp((double)(-(x * x + y * y)) / den);
    }
    // np.percentile 'linear' plan in float32 (numpy >= 2): q / float32(100), (n - 1) * q, floor
    const int len = h * w;
    const int qs[5] = {5, 25, 75, 95, 90};
    for (int k = 0; k < 5; ++k) {
        const float quant = (float)qs[k] / 100.0f;
        const float virt = (float)(len - 1) * quant;
        const float prev = std::floor(virt);
        if (virt >= (float)(len - 1)) { t->pct_lo[k] = len - 1; t->pct_hi[k] = len - 1; }
        else if (virt < 0.0f) { t->pct_lo[k] = 0; t->pct_hi[k] = 0; }
        else { t->pct_lo[k] = (int32_t)prev; t->pct_hi[k] = (int32_t)prev + 1; }
        t->pct_gamma[k] = virt - prev;
    }
    return MDIMG_OK;
}

int mdimg_enhance(const float* in, float* out, int n, int h, int w, const mdimg_enhance_plan* plan,
                  const mdimg_enhance_tables* tables, const double* rows_before_in, double* rows_after,
                  int32_t* flags_out, int32_t* tv_iters_out, void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || h < 1 || w < 1 || !in || !out || !plan || !tables || !flags_out || in == out)
        return set_error(MDIMG_ERR_INVALID, "enhance: bad arguments");
    if (n == 0) return MDIMG_OK;
    Ctx c;
    c.n = n; c.h = h; c.w = w; c.in = in; c.q = *plan; c.t = tables; c.stream = stream;
    if (int rc0 = mdimg_plan_clamp(&c.q)) return rc0;
    const mdimg_enhance_plan& q = c.q;
    Arena a(ws, ws_bytes);
    carve(a, n, h, w, q.clahe_tile_size, c.b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "enhance: workspace too small (%zu > %zu)", a.off, ws_bytes);
    cudaStream_t st_ = (cudaStream_t)stream;
    const size_t img_bytes = sizeof(float) * (size_t)n * h * w;
    double* rows = rows_after ? rows_after : c.b.rows;
    auto listed = [&](int step) {
        for (int i = 0; i < q.n_ops; ++i)
            if (q.ops[i] == step) return true;
        return false;
    };
    auto metrics = [&](const float* img, double* dst, const int32_t* sel, int n_sel) {
        return mdimg_metrics(img, n, h, w, sel, n_sel, 1, tables->pct_lo, tables->pct_hi, tables->pct_gamma, dst,
                             c.b.sub, c.b.sub_bytes, stream);
    };

    // ---- the seven steps in their fixed order, gated by membership (enhancement.py:266-315) ----
    float* cur = out;
    float* tmp = c.b.tmp;
    cudaMemcpyAsync(cur, in, img_bytes, cudaMemcpyDeviceToDevice, st_);
    Pass main_pass;
    main_pass.fuse_gamma = listed(MDIMG_STEP_CLAHE) && listed(MDIMG_STEP_GAMMA) && enabled(MDIMG_STEP_GAMMA, q);
    int rc = MDIMG_OK;
    for (int step = MDIMG_STEP_DENOISE; step <= MDIMG_STEP_TV_DENOISE; ++step)
        if (listed(step) && enabled(step, q)) {
            rc = apply_step(c, step, q.unsharp_amount, cur, tmp, nullptr, 0, main_pass, c.b.iters);
            if (rc) return rc;
        }
    rc = mdimg_clip01(cur, cur, n, h, w, nullptr, 0, stream);
    if (rc) return rc;
    // ---- safeguards: one metrics pass yields what all three look at; only modified slices are re-measured ----
    // The metrics passes are enqueued BEFORE the first host read-back: the GPU works on them while the host waits for
    // the step flags, and the later read-backs find their data already there.
    const double* rows_b = rows_before_in;
    if (!rows_b) {
        rc = metrics(in, c.b.rows_before, nullptr, 0);
        if (rc) return rc;
        rows_b = c.b.rows_before;
    }
    rc = metrics(cur, rows, nullptr, 0);
    if (rc) return rc;
    std::vector<int32_t> err(n, 0);
    rc = flush_checks(c, main_pass, err);
    if (rc) return rc;
    std::vector<int32_t> tv_iters(n, 0);
    const bool tv_main = main_pass.tv_ran;
    if (tv_main) { rc = c.fetch(tv_iters.data(), c.b.iters, sizeof(int32_t) * n); if (rc) return rc; }
    std::vector<double> hb((size_t)n * MC_COLS), hr((size_t)n * MC_COLS);
    rc = c.fetch(hb.data(), rows_b, sizeof(double) * hb.size());
    if (rc) return rc;
    rc = c.fetch(hr.data(), rows, sizeof(double) * hr.size());
    if (rc) return rc;
    std::vector<double> s0(n), nb(n), sig1(n), niqe1(n), tmpd((size_t)n * 2);
    for (int i = 0; i < n; ++i) {
        s0[i] = hb[(size_t)i * MC_COLS + MC_SIGMA];
        nb[i] = hb[(size_t)i * MC_COLS + MC_NIQE];
        sig1[i] = hr[(size_t)i * MC_COLS + MC_SIGMA];
        niqe1[i] = hr[(size_t)i * MC_COLS + MC_NIQE];
    }
    std::vector<char> dirty(n, 0), halo(n, 0), noise(n, 0), over(n, 0);

    // _check_halo -> re-run in the plan's own order with half the unsharp amount (enhancement.py:318-353)
    if (listed(MDIMG_STEP_UNSHARP)) {
        for (int i = 0; i < n; ++i) halo[i] = hr[(size_t)i * MC_COLS + MC_EDGE_RATIO] > 1.5;
        const int ns = c.select(halo);
        if (ns > 0) {
            const double reduced = q.unsharp_amount * 0.5;
            rc = mdimg_copy(in, cur, n, h, w, c.b.sel, ns, stream);
            if (rc) return rc;
            Pass re;
            for (int i = 0; i < q.n_ops; ++i) {
                const int step = q.ops[i];
                if (step >= MDIMG_STEP_DENOISE && step <= MDIMG_STEP_TV_DENOISE && enabled(step, q)) {
                    rc = apply_step(c, step, reduced, cur, tmp, c.b.sel, ns, re, c.b.iters2);
                    if (rc) return rc;
                }
            }
            rc = mdimg_clip01(cur, cur, n, h, w, c.b.sel, ns, stream);
            if (rc) return rc;
            rc = flush_checks(c, re, err);
            if (rc) return rc;
            if (re.tv_ran && tv_main) {
                std::vector<int32_t> it2(n);
                rc = c.fetch(it2.data(), c.b.iters2, sizeof(int32_t) * n);
                if (rc) return rc;
                for (int i = 0; i < n; ++i)
                    if (halo[i]) tv_iters[i] = it2[i];
            }
            rc = mdimg_estimate_sigma(cur, n, h, w, c.b.sel, ns, c.b.sigma, c.b.sub, c.b.sub_bytes, stream);
            if (rc) return rc;
            rc = c.fetch(tmpd.data(), c.b.sigma, sizeof(double) * n);
            if (rc) return rc;
            for (int i = 0; i < n; ++i)
                if (halo[i]) { sig1[i] = tmpd[i]; dirty[i] = 1; }
        }
    }

    // _check_noise_amplification -> corrective light denoise (enhancement.py:55-63,356-360)
    for (int i = 0; i < n; ++i) noise[i] = !(s0[i] < 1e-8) && (sig1[i] > s0[i] * 1.3);
    {
        const int ns = c.select(noise);
        if (ns > 0) {
            rc = mdimg_light_denoise(cur, cur, n, h, w, c.b.sel, ns, 0.4, c.b.skipped, c.b.sub, c.b.sub_bytes, stream);
            if (rc) return rc;
            rc = mdimg_clip01(cur, cur, n, h, w, c.b.sel, ns, stream);
            if (rc) return rc;
            for (int i = 0; i < n; ++i)
                if (noise[i]) dirty[i] = 1;
        }
    }

    // _check_over_processing: NIQE approximation degraded by more than 0.5 -> 0.6 enhanced + 0.4 original
    {
        const int ns = c.select(dirty);
        if (ns > 0) {
            rc = mdimg_quality(cur, n, h, w, c.b.sel, ns, 1, c.b.quality, c.b.sub, c.b.sub_bytes, stream);
            if (rc) return rc;
            rc = c.fetch(tmpd.data(), c.b.quality, sizeof(double) * 2 * n);
            if (rc) return rc;
            for (int i = 0; i < n; ++i)
                if (dirty[i]) niqe1[i] = tmpd[(size_t)i * 2 + 1];
        }
        for (int i = 0; i < n; ++i) over[i] = (niqe1[i] - nb[i]) > 0.5;
        const int no = c.select(over);
        if (no > 0) {
            rc = mdimg_axpby(cur, in, cur, n, h, w, c.b.sel, no, 0.6, 0.4, 1, stream);
            if (rc) return rc;
            for (int i = 0; i < n; ++i)
                if (over[i]) dirty[i] = 1;
        }
    }
    {
        const int ns = c.select(dirty);
        if (ns > 0) { rc = metrics(cur, rows, c.b.sel, ns); if (rc) return rc; }
    }
    // slices on which the reference would have raised: returned unchanged
    {
        std::vector<char> bad(n, 0);
        for (int i = 0; i < n; ++i) bad[i] = err[i] != 0;
        const int ns = c.select(bad);
        if (ns > 0) {
            rc = mdimg_copy(in, cur, n, h, w, c.b.sel, ns, stream);
            if (rc) return rc;
            rc = metrics(cur, rows, c.b.sel, ns);
            if (rc) return rc;
        }
    }
    if (cur != out) cudaMemcpyAsync(out, cur, img_bytes, cudaMemcpyDeviceToDevice, st_);
    for (int i = 0; i < n; ++i)
        flags_out[i] = (halo[i] ? MDIMG_FLAG_HALO : 0) | (noise[i] ? MDIMG_FLAG_NOISE_GUARD : 0) |
                       (over[i] ? MDIMG_FLAG_OVER_PROCESSED : 0) | err[i];
    if (tv_iters_out)
        for (int i = 0; i < n; ++i) tv_iters_out[i] = tv_main ? tv_iters[i] : 0;
    cudaError_t e = cudaStreamSynchronize(st_);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "enhance: %s", cudaGetErrorString(e));
    return check_launch("enhance");
}

int mdimg_enhance_issues(const float* in, float* out, int n, int h, int w, int issues,
                         const mdimg_enhance_tables* tables, const double* sigma_before, int32_t* flags_out,
                         void* ws, size_t ws_bytes, void* stream) {
    if (n < 0 || h < 1 || w < 1 || !in || !out || !tables || !flags_out || in == out)
        return set_error(MDIMG_ERR_INVALID, "enhance_issues: bad arguments");
    if (n == 0) return MDIMG_OK;
    // ENHANCEMENT_PARAMS (pipeline/enhancement.py:32-42)
    Ctx c;
    c.n = n; c.h = h; c.w = w; c.in = in; c.t = tables; c.stream = stream;
    std::memset(&c.q, 0, sizeof(c.q));
    c.q.clahe_clip_limit = 0.015; c.q.clahe_tile_size = 16; c.q.unsharp_radius = 0.8; c.q.unsharp_amount = 0.5;
    c.q.denoise_hard = 0; c.q.post_denoise_strength = 0.3;
    Arena a(ws, ws_bytes);
    carve(a, n, h, w, c.q.clahe_tile_size, c.b);
    if (!a.ok()) return set_error(MDIMG_ERR_WORKSPACE, "enhance_issues: workspace too small (%zu > %zu)", a.off, ws_bytes);
    cudaStream_t st_ = (cudaStream_t)stream;
    const size_t img_bytes = sizeof(float) * (size_t)n * h * w;
    float* cur = out;
    float* tmp = c.b.tmp;
    cudaMemcpyAsync(cur, in, img_bytes, cudaMemcpyDeviceToDevice, st_);
    Pass p;
    int rc = MDIMG_OK;
    auto has = [&](int bit) { return (issues & bit) != 0; };
    if (has(MDIMG_ISSUE_NOISE)) {
        rc = apply_step(c, MDIMG_STEP_DENOISE, 0.0, cur, tmp, nullptr, 0, p, c.b.iters);
        if (rc) return rc;
    }
    if (has(MDIMG_ISSUE_LOW_CONTRAST) || has(MDIMG_ISSUE_CLIPPING_LOW) || has(MDIMG_ISSUE_CLIPPING_HIGH)) {
        rc = apply_step(c, MDIMG_STEP_CLAHE, 0.0, cur, tmp, nullptr, 0, p, c.b.iters);
        if (rc) return rc;
    }
    double g = 0.0;
    if (has(MDIMG_ISSUE_CLIPPING_LOW) && !has(MDIMG_ISSUE_CLIPPING_HIGH)) g = 0.95;          // gamma_brighten
    else if (has(MDIMG_ISSUE_CLIPPING_HIGH) && !has(MDIMG_ISSUE_CLIPPING_LOW)) g = 1.05;     // gamma_darken
    if (g != 0.0) {
        c.q.gamma = g;
        rc = apply_step(c, MDIMG_STEP_GAMMA, 0.0, cur, tmp, nullptr, 0, p, c.b.iters);
        if (rc) return rc;
    }
    if (has(MDIMG_ISSUE_BLUR)) {
        rc = apply_step(c, MDIMG_STEP_UNSHARP, c.q.unsharp_amount, cur, tmp, nullptr, 0, p, c.b.iters);
        if (rc) return rc;
        rc = apply_step(c, MDIMG_STEP_POST_DENOISE, 0.0, cur, tmp, nullptr, 0, p, c.b.iters);
        if (rc) return rc;
    }
    rc = mdimg_clip01(cur, cur, n, h, w, nullptr, 0, stream);
    if (rc) return rc;
    std::vector<int32_t> err(n, 0);
    rc = flush_checks(c, p, err);
    if (rc) return rc;

    // _check_noise_amplification + corrective light denoise (enhancement.py:55-63,221-225)
    std::vector<double> s0(n), s1(n);
    const double* sb = sigma_before;
    if (!sb) {
        rc = mdimg_estimate_sigma(in, n, h, w, nullptr, 0, c.b.quality, c.b.sub, c.b.sub_bytes, stream);
        if (rc) return rc;
        sb = c.b.quality;
    }
    rc = c.fetch(s0.data(), sb, sizeof(double) * n);
    if (rc) return rc;
    rc = mdimg_estimate_sigma(cur, n, h, w, nullptr, 0, c.b.sigma, c.b.sub, c.b.sub_bytes, stream);
    if (rc) return rc;
    rc = c.fetch(s1.data(), c.b.sigma, sizeof(double) * n);
    if (rc) return rc;
    std::vector<char> noise(n, 0), bad(n, 0);
    for (int i = 0; i < n; ++i) noise[i] = !(s0[i] < 1e-8) && (s1[i] > s0[i] * 1.3);
    int ns = c.select(noise);
    if (ns > 0) {
        rc = mdimg_light_denoise(cur, cur, n, h, w, c.b.sel, ns, 0.4, c.b.skipped, c.b.sub, c.b.sub_bytes, stream);
        if (rc) return rc;
        rc = mdimg_clip01(cur, cur, n, h, w, c.b.sel, ns, stream);
        if (rc) return rc;
    }
    for (int i = 0; i < n; ++i) bad[i] = err[i] != 0;
    ns = c.select(bad);
    if (ns > 0) { rc = mdimg_copy(in, cur, n, h, w, c.b.sel, ns, stream); if (rc) return rc; }
    if (cur != out) cudaMemcpyAsync(out, cur, img_bytes, cudaMemcpyDeviceToDevice, st_);
    for (int i = 0; i < n; ++i) flags_out[i] = (noise[i] ? MDIMG_FLAG_NOISE_GUARD : 0) | err[i];
    cudaError_t e = cudaStreamSynchronize(st_);
    if (e != cudaSuccess) return set_error(MDIMG_ERR_CUDA, "enhance_issues: %s", cudaGetErrorString(e));
    return check_launch("enhance_issues");
}

// ---- host scalar logic (python-float arithmetic of pipeline/metrics.py, statement by statement) ----
static inline double py_max(double a, double b) { return b > a ? b : a; }      // python max(a, b): a unless b > a
static inline double py_min(double a, double b) { return b < a ? b : a; }      // python min(a, b)

int mdimg_detect_issues(const double* m) {
    if (!m) return 0;
    int mask = 0;
    if (m[MC_SIGMA] > 0.08) mask |= MDIMG_ISSUE_NOISE;
    if (m[MC_LAP_VAR] < 0.001) mask |= MDIMG_ISSUE_BLUR;
    if (m[MC_STD] < 0.12) mask |= MDIMG_ISSUE_LOW_CONTRAST;
    if (m[MC_PCT_LOW] > 0.01) mask |= MDIMG_ISSUE_CLIPPING_LOW;
    if (m[MC_PCT_HIGH] > 0.01) mask |= MDIMG_ISSUE_CLIPPING_HIGH;
    return mask;
}

int mdimg_validation_scalars_of(const double* row, mdimg_validation_scalars* o) {
    if (!row || !o) return set_error(MDIMG_ERR_INVALID, "validation scalars: bad arguments");
    const double* mb = row;
    const double* ma = row + MC_COLS;
    const double eps = 1e-8;
    o->ssim = row[2 * MC_COLS];
    o->psnr = row[2 * MC_COLS + 1];
    o->niqe_before = mb[MC_NIQE];
    o->niqe_after = ma[MC_NIQE];
    o->niqe_improved = o->niqe_after <= o->niqe_before;
    o->contrast_gain = (ma[MC_STD] - mb[MC_STD]) / py_max(mb[MC_STD], eps);
    o->sharpness_gain = (ma[MC_LAP_VAR] - mb[MC_LAP_VAR]) / py_max(mb[MC_LAP_VAR], eps);
    const double noise_reduction = (mb[MC_SIGMA] - ma[MC_SIGMA]) / py_max(mb[MC_SIGMA], eps);
    o->quality_improvement = 0.35 * o->contrast_gain + 0.35 * o->sharpness_gain + 0.30 * noise_reduction;
    o->meets_ssim = o->ssim >= 0.70;
    o->meets_psnr = o->psnr >= 22.0;
    o->meets_improvement = o->quality_improvement >= 0.10;
    o->passes = (o->meets_ssim && o->meets_psnr) || (o->meets_ssim && o->meets_improvement) ||
                (o->meets_psnr && o->meets_improvement && o->niqe_improved);
    o->noise_change = -noise_reduction;
    o->entropy_change = ma[MC_ENTROPY] - mb[MC_ENTROPY];
    o->snr_change = ma[MC_SNR] - mb[MC_SNR];
    o->cnr_change = ma[MC_CNR] - mb[MC_CNR];
    o->edge_density_change = ma[MC_EDGE_DENSITY] - mb[MC_EDGE_DENSITY];
    o->histogram_spread_change = ma[MC_HIST_SPREAD] - mb[MC_HIST_SPREAD];
    o->edge_ratio = ma[MC_EDGE_RATIO];
    o->local_contrast_change = ma[MC_LOCAL_CONTRAST] - mb[MC_LOCAL_CONTRAST];
    o->gradient_strength_change = ma[MC_GRAD_STRENGTH] - mb[MC_GRAD_STRENGTH];
    o->gradient_entropy_change = ma[MC_GRAD_ENTROPY] - mb[MC_GRAD_ENTROPY];
    return MDIMG_OK;
}

int mdimg_objective_score(const mdimg_validation_scalars* v, double* score, double parts[11]) {
    if (!v || !score || !parts) return set_error(MDIMG_ERR_INVALID, "objective score: bad arguments");
    auto capped = [](double x, double cap) { return py_max(0.0, py_min(x, cap)); };
    parts[0] = v->contrast_gain;
    parts[1] = v->sharpness_gain;
    parts[2] = py_max(0.0, v->noise_change);
    parts[3] = py_max(0.0, v->niqe_after - v->niqe_before);
    parts[4] = py_max(0.0, v->edge_ratio - 1.0) * 5.0;
    parts[5] = py_max(0.0, std::fabs(v->entropy_change) - 0.5) * 2.0;
    parts[6] = capped(v->snr_change * 0.1, 0.5);
    parts[7] = capped(v->histogram_spread_change * 0.5, 0.3);
    parts[8] = capped(v->local_contrast_change * 0.3, 0.3);
    parts[9] = capped(v->gradient_strength_change * 0.2, 0.2);
    parts[10] = py_max(0.0, std::fabs(v->gradient_entropy_change) - 0.3) * 1.5;
    *score = 0.35 * parts[0] + 0.35 * parts[1] - 0.30 * parts[2] - 5.0 * parts[3] - 10.0 * (v->passes ? 0 : 1)
             - parts[4] - parts[5] + parts[6] + parts[7] + parts[8] + parts[9] - parts[10];
    return MDIMG_OK;
}

}  // extern "C"
