// C-ABI layer (include/ripcurrents_b200.h): context, buffer management, host<->device staging, call order.
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdio.h>
#include <new>

#include "rc_internal.h"

#define RC_VERSION 100

const char* const rc_kernel_names[K_COUNT] = {"pyr_h", "pyr_v", "polyexp", "update_matrices", "flow_iter_fused",
                                              "flow_iter_final", "polar_hist", "thresholds", "classify", "window_mean",
                                              "advect", "streakline", "misc"};

namespace {

int fail(rc_ctx* c, int code, const char* fmt, const char* detail = "")
{
    if (c) {
        char buf[512];
        snprintf(buf, sizeof buf, fmt, detail);
        c->err = buf;
    }
    return code;
}

#define CUDA_TRY(c, expr)                                                                     \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) return fail((c), RC_ERR_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

#define CHECK_LAUNCH(c)                                                                       \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) return fail((c), RC_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
    } while (0)

bool is_device_ptr(const void* p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure(rc_ctx* c, void** buf, size_t* cap, size_t need)
{
    if (*cap >= need && *buf) return RC_OK;
    if (*buf) { cudaStreamSynchronize(c->stream); cudaFree(*buf); *buf = nullptr; *cap = 0; }
    size_t sz = need < 256 ? 256 : need;
    if (cudaMalloc(buf, sz) != cudaSuccess) { cudaGetLastError(); return fail(c, RC_ERR_NOMEM, "cudaMalloc failed%s"); }
    *cap = sz;
    return RC_OK;
}

int dev_alloc(rc_ctx* c, void** p, size_t bytes)
{
    if (cudaMalloc(p, bytes < 256 ? 256 : bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(c, RC_ERR_NOMEM, "cudaMalloc failed%s");
    }
    c->allocs.push_back(*p);
    return RC_OK;
}

void free_farneback(rc_ctx* c)
{
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (void* p : c->allocs) cudaFree(p);
    c->allocs.clear();
    for (auto& L : c->layer) L = Layer();
    c->configured = false; c->nlayers = 0; c->frames_seen = 0; c->have_flow = false; c->flow_out = nullptr;
}

int round_half_even(double v) { return (int)nearbyint(v); }

int alloc_planes(rc_ctx* c, Planes& P, int w, int h)
{
    P.w = w; P.h = h; P.pitch = (w + 31) / 32 * 32;
    P.pstride = (size_t)P.pitch * h;
    return dev_alloc(c, (void**)&P.p, sizeof(float) * P.pstride * 5);
}

// Appendix A.3 kernels and the closed-form inverse of the moment matrix
void make_poly(PolyCoef& pc, int n, double sigma, bool strict)
{
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    std::vector<float> G(2 * n + 1);
    double s = 0;
    for (int x = -n; x <= n; x++) { G[x + n] = (float)exp(-x * x / (2 * sigma * sigma)); s += G[x + n]; }
    s = 1.0 / s;
    for (int x = -n; x <= n; x++) G[x + n] = (float)(G[x + n] * s);
    for (int x = 0; x <= n; x++) {
        pc.g[x] = G[x + n];
        pc.xg[x] = (float)(x * G[x + n]);
        pc.xxg[x] = (float)(x * x * G[x + n]);
    }
    double a = 0, b = 0, cc = 0, d = 0;
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            float gyx = G[y + n] * G[x + n];
            a += gyx; b += gyx * x * x; cc += gyx * x * x * x * x; d += gyx * x * x * y * y;
        }
    double det3 = (cc - d) * (a * (cc + d) - 2 * b * b);
    pc.ig11 = 1.0 / b;
    pc.ig03 = -b * (cc - d) / det3;
    pc.ig33 = (a * cc - b * b) / det3;
    pc.ig55 = 1.0 / d;
    pc.n = n;
    pc.n_eff = n;
    if (!strict) {
        // drop taps whose x^2-weighted contribution is below 1e-12 of the centre weight
        int k = n;
        while (k > 1 && (double)pc.xxg[k] < 1e-12 * (double)pc.g[0]) k--;
        pc.n_eff = k;
    }
}

void make_smooth(SmoothCoef& sc, double sigma, int ksize)
{
    sc.ksize = ksize;
    if (sigma <= 0 && ksize == 3) { sc.k[0] = 0.25f; sc.k[1] = 0.5f; sc.k[2] = 0.25f; return; }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> t(ksize);
    double s2 = -0.5 / (sigma * sigma), sum = 0;
    for (int i = 0; i < ksize; i++) { double x = i - (ksize - 1) * 0.5; t[i] = exp(s2 * x * x); sum += t[i]; }
    sum = 1.0 / sum;
    for (int i = 0; i < ksize; i++) sc.k[i] = (float)(t[i] * sum);
}

void make_gwin(GaussWin& g, int winsize)
{
    int m = winsize / 2;
    double sigma = m * 0.3, s = 1.0;
    g.m = m;
    g.k[0] = 1.f;
    for (int i = 1; i <= m; i++) { float t = (float)exp(-i * i / (2 * sigma * sigma)); g.k[i] = t; s += t * 2; }
    s = 1.0 / s;
    for (int i = 0; i <= m; i++) g.k[i] = (float)(g.k[i] * s);
}

int ensure_aggregate(rc_ctx* c)
{
    if (c->d_hist2d) return RC_OK;
    CUDA_TRY(c, cudaMalloc((void**)&c->d_hist2d, sizeof(unsigned long long) * RC_HIST_ROWS * RC_HIST_BINS));
    CUDA_TRY(c, cudaMemsetAsync(c->d_hist2d, 0, sizeof(unsigned long long) * RC_HIST_ROWS * RC_HIST_BINS, c->stream));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_thr, sizeof(float) * RC_THR_FLOATS));
    CUDA_TRY(c, cudaMemsetAsync(c->d_thr, 0, sizeof(float) * RC_THR_FLOATS, c->stream));
    return RC_OK;
}

int ensure_accumulator(rc_ctx* c, int w, int h)
{
    if (c->d_acc && c->acc_w == w && c->acc_h == h) return RC_OK;
    if (c->d_acc) { cudaStreamSynchronize(c->stream); cudaFree(c->d_acc); cudaFree(c->d_mask); cudaFree(c->d_cls); }
    c->d_acc = nullptr; c->d_mask = nullptr; c->d_cls = nullptr;
    const size_t n = (size_t)w * h;
    CUDA_TRY(c, cudaMalloc((void**)&c->d_acc, sizeof(float) * n));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_mask, n));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_cls, 2 * n));
    CUDA_TRY(c, cudaMemsetAsync(c->d_acc, 0, sizeof(float) * n, c->stream));
    c->acc_w = w; c->acc_h = h;
    return RC_OK;
}

// Brings a (possibly host, possibly strided) flow field to the device as a dense w*h*2 array when needed.
// Returns the device pointer + step to use.
int stage_flow_in(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float** d_flow, size_t* d_step)
{
    if (!flow) {
        if (!c->have_flow) return fail(c, RC_ERR_STATE, "no flow has been computed on this context%s");
        *d_flow = c->flow_out; *d_step = (size_t)c->prm.w * 8;
        return RC_OK;
    }
    if (w <= 0 || h <= 0 || flow_step < (size_t)w * 8) return fail(c, RC_ERR_INVALID, "bad flow geometry%s");
    if (is_device_ptr(flow)) { *d_flow = flow; *d_step = flow_step; return RC_OK; }
    int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, (size_t)w * h * 8);
    if (rc) return rc;
    CUDA_TRY(c, cudaMemcpy2DAsync(c->d_tmp, (size_t)w * 8, flow, flow_step, (size_t)w * 8, h, cudaMemcpyHostToDevice,
                                  c->stream));
    *d_flow = reinterpret_cast<const float*>(c->d_tmp); *d_step = (size_t)w * 8;
    return RC_OK;
}

int copy_out(rc_ctx* c, void* dst, size_t dst_step, const void* d_src, size_t src_step, size_t row_bytes, int rows,
             bool* host_written)
{
    const bool dev = is_device_ptr(dst);
    CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, d_src, src_step, row_bytes, rows,
                                  dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (!dev) *host_written = true;
    return RC_OK;
}

}  // namespace

static void prof_drain(rc_ctx* c)
{
    if (c->prof.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (auto& r : c->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            c->prof_ms[r.id] += ms; c->prof_bytes[r.id] += r.bytes; c->prof_n[r.id] += 1;
        } else cudaGetLastError();
        c->ev_pool.push_back(r.a); c->ev_pool.push_back(r.b);
    }
    c->prof.clear();
}


// =====================================================================================================
extern "C" {

int rc_version(void) { return RC_VERSION; }

const char* rc_error_string(int code)
{
    switch (code) {
    case RC_OK: return "ok";
    case RC_ERR_INVALID: return "invalid argument";
    case RC_ERR_CUDA: return "CUDA error";
    case RC_ERR_NOMEM: return "out of device memory";
    case RC_ERR_STATE: return "invalid call order";
    case RC_ERR_UNSUPPORTED: return "unsupported parameter";
    default: return "unknown error";
    }
}

const char* rc_last_error(const rc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int64_t rc_kernel_launches(const rc_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rc_create(rc_ctx** out, int device)
{
    if (!out) return RC_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return RC_ERR_CUDA; }
    if (device < 0 || device >= ndev) return RC_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return RC_ERR_CUDA; }
    rc_ctx* c = new (std::nothrow) rc_ctx();
    if (!c) return RC_ERR_NOMEM;
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError(); delete c; return RC_ERR_CUDA;
    }
    c->own_stream = true;
    *out = c;
    return RC_OK;
}

void rc_destroy(rc_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    free_farneback(c);
    if (c->stream) cudaStreamSynchronize(c->stream);
    void* bufs[] = {c->d_frame, c->d_tmp, c->d_tmp2, c->d_hist2d, c->d_thr, c->d_acc, c->d_mask, c->d_cls, c->d_ring,
                    c->d_avg};
    for (void* p : bufs) if (p) cudaFree(p);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    prof_drain(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rc_set_stream(rc_ctx* c, void* s)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->own_stream && c->stream) { cudaStreamDestroy(c->stream); c->own_stream = false; c->stream = nullptr; }
    if (!s) {
        CUDA_TRY(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)s;
    }
    return RC_OK;
}

int rc_synchronize(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// ---- per-kernel device timing ----------------------------------------------------------------------------
int rc_profile_enable(rc_ctx* c, int on)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    prof_drain(c);
    c->prof_on = on != 0;
    return RC_OK;
}

int rc_profile_reset(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    prof_drain(c);
    for (int i = 0; i < K_COUNT; i++) { c->prof_ms[i] = 0; c->prof_bytes[i] = 0; c->prof_n[i] = 0; }
    return RC_OK;
}

int rc_profile_count(void) { return K_COUNT; }

int rc_profile_get(rc_ctx* c, int idx, const char** name, double* total_ms, int64_t* launches, double* alg_bytes)
{
    if (!c || idx < 0 || idx >= K_COUNT) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    prof_drain(c);
    if (name) *name = rc_kernel_names[idx];
    if (total_ms) *total_ms = c->prof_ms[idx];
    if (launches) *launches = c->prof_n[idx];
    if (alg_bytes) *alg_bytes = c->prof_bytes[idx];
    return RC_OK;
}

// ---- A1 -----------------------------------------------------------------------------------------------
int rc_flow_configure(rc_ctx* c, int w, int h, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                      double poly_sigma, int flags)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (w < 2 || h < 2 || levels < 0) return fail(c, RC_ERR_INVALID, "bad geometry / levels%s");
    if (!(pyr_scale > 0.0 && pyr_scale < 1.0)) return fail(c, RC_ERR_INVALID, "pyr_scale must be in (0,1)%s");
    if (flags & 4) return fail(c, RC_ERR_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW is not supported%s");
    if (poly_n < 1 || poly_n > RC_MAX_POLY_N) return fail(c, RC_ERR_UNSUPPORTED, "poly_n must be in [1,32]%s");
    if (winsize < 1 || winsize / 2 > RC_MAX_WIN_HALF) return fail(c, RC_ERR_UNSUPPORTED, "winsize must be in [1,129]%s");
    if (iterations < 1) return fail(c, RC_ERR_UNSUPPORTED, "iterations must be >= 1%s");
    FarnebackParams p;
    p.w = w; p.h = h; p.pyr_scale = pyr_scale; p.levels = levels; p.winsize = winsize; p.iterations = iterations;
    p.poly_n = poly_n; p.poly_sigma = poly_sigma; p.flags = flags;
    if (c->configured && c->prm == p) { c->frames_seen = 0; c->have_flow = false; return RC_OK; }
    free_farneback(c);

    // Appendix A.1: layer selection
    int k; double scale = 1.0;
    for (k = 0; k < levels; k++) { scale *= pyr_scale; if (w * scale < 32 || h * scale < 32) break; }
    const int nl = k + 1;
    if (nl > RC_MAX_LAYERS) return fail(c, RC_ERR_UNSUPPORTED, "too many pyramid layers%s");
    scale = 1.0;
    for (k = 0; k < nl; k++) {
        Layer& L = c->layer[k];
        L.w = round_half_even(w * scale); L.h = round_half_even(h * scale);
        double sigma = (1.0 / scale - 1.0) * 0.5;
        int ks = round_half_even(sigma * 5.0) | 1; if (ks < 3) ks = 3;
        if (ks > RC_MAX_SMOOTH_TAPS) { free_farneback(c); return fail(c, RC_ERR_UNSUPPORTED, "pyramid too deep (smoothing kernel > 255 taps)%s"); }
        make_smooth(L.smooth, sigma, ks);
        const int pitch = (L.w + 31) / 32 * 32;
        int rc;
        if ((rc = dev_alloc(c, (void**)&L.I, sizeof(float) * (size_t)pitch * L.h)) ||
            (rc = dev_alloc(c, (void**)&L.htmp, sizeof(float) * 2 * (size_t)L.w * h)) ||
            (rc = alloc_planes(c, L.R[0], L.w, L.h)) || (rc = alloc_planes(c, L.R[1], L.w, L.h)) ||
            (rc = alloc_planes(c, L.M[0], L.w, L.h)) || (rc = alloc_planes(c, L.M[1], L.w, L.h)) ||
            (rc = dev_alloc(c, (void**)&L.flow, sizeof(float) * 2 * (size_t)L.w * L.h))) {
            free_farneback(c);
            return rc;
        }
        scale *= pyr_scale;
    }
    c->nlayers = nl;
    make_poly(c->poly, poly_n, poly_sigma, (flags & RC_FARNEBACK_STRICT) != 0);
    make_gwin(c->gwin, winsize);
    c->prm = p; c->configured = true; c->frames_seen = 0; c->have_flow = false; c->cur = 0;
    c->flow_out = c->layer[0].flow;
    return RC_OK;
}

static int flow_push_impl(rc_ctx* c, const uint8_t* frame, size_t step)
{
    const int w = c->prm.w, h = c->prm.h;
    if (!frame || step < (size_t)w) return fail(c, RC_ERR_INVALID, "bad frame pointer / step%s");
    const uint8_t* d_img = frame; size_t d_step = step;
    if (!is_device_ptr(frame)) {
        int rc = ensure(c, (void**)&c->d_frame, &c->d_frame_cap, (size_t)w * h);
        if (rc) return rc;
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d_frame, w, frame, step, w, h, cudaMemcpyHostToDevice, c->stream));
        d_img = c->d_frame; d_step = w;
    }
    c->cur ^= 1;
    const int cur = c->cur, prev = cur ^ 1;
    for (int k = 0; k < c->nlayers; k++) {
        Layer& L = c->layer[k];
        rc_launch_pyr_layer(c, d_img, d_step, w, h, L);
        rc_launch_polyexp(c, L.I, L.w, L.h, (L.w + 31) / 32 * 32, L.R[cur]);
    }
    CHECK_LAUNCH(c);
    c->frames_seen++;
    if (c->frames_seen < 2) return 0;
    const int T = c->prm.iterations;
    const float fscale = (float)(1.0 / c->prm.pyr_scale);
    for (int k = c->nlayers - 1; k >= 0; k--) {
        Layer& L = c->layer[k];
        if (k == c->nlayers - 1) rc_launch_update_matrices(c, L.R[prev], L.R[cur], L.M[0], 0, nullptr, 0, 0, 1.f);
        else {
            Layer& C = c->layer[k + 1];
            rc_launch_update_matrices(c, L.R[prev], L.R[cur], L.M[0], 1, C.flow, C.w, C.h, fscale);
        }
        int mi = 0;
        for (int it = 0; it < T; it++) {
            if (it < T - 1) { rc_launch_update_flow(c, L.M[mi], L.R[prev], L.R[cur], L.M[mi ^ 1], nullptr, nullptr); mi ^= 1; }
            else rc_launch_update_flow(c, L.M[mi], L.R[prev], L.R[cur], Planes(), L.flow, nullptr);
        }
    }
    CHECK_LAUNCH(c);
    c->have_flow = true;
    return 1;
}

int rc_flow_push(rc_ctx* c, const uint8_t* frame, size_t step, float* flow, size_t flow_step)
{
    if (!c) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    cudaSetDevice(c->device);
    int produced = flow_push_impl(c, frame, step);
    if (produced < 0) return produced;
    bool host = !is_device_ptr(frame);   // the source host buffer must be consumed before we return
    if (produced == 1 && flow) {
        if (flow_step < (size_t)c->prm.w * 8) return fail(c, RC_ERR_INVALID, "flow_step too small%s");
        int rc = copy_out(c, flow, flow_step, c->flow_out, (size_t)c->prm.w * 8, (size_t)c->prm.w * 8, c->prm.h, &host);
        if (rc) return rc;
    }
    if (host) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return produced;
}

int rc_farneback(rc_ctx* c, const uint8_t* prev, size_t prev_step, const uint8_t* next, size_t next_step, int w, int h,
                 float* flow, size_t flow_step, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                 double poly_sigma, int flags)
{
    if (!c) return RC_ERR_INVALID;
    int rc = rc_flow_configure(c, w, h, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags);
    if (rc) return rc;
    c->frames_seen = 0; c->have_flow = false;
    rc = rc_flow_push(c, prev, prev_step, nullptr, 0);
    if (rc < 0) return rc;
    rc = rc_flow_push(c, next, next_step, flow, flow_step);
    return rc < 0 ? rc : RC_OK;
}

int rc_flow_device(rc_ctx* c, float** dev_flow, int* w, int* h)
{
    if (!c || !dev_flow) return RC_ERR_INVALID;
    if (!c->have_flow) return fail(c, RC_ERR_STATE, "no flow has been computed on this context%s");
    *dev_flow = c->flow_out;
    if (w) *w = c->prm.w;
    if (h) *h = c->prm.h;
    return RC_OK;
}

// ---- A2 + A3 --------------------------------------------------------------------------------------------
int rc_hist_reset(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    CUDA_TRY(c, cudaMemsetAsync(c->d_hist2d, 0, sizeof(unsigned long long) * RC_HIST_ROWS * RC_HIST_BINS, c->stream));
    return RC_OK;
}

int rc_polar_hist(rc_ctx* c, const float* flow, size_t flow_step, int w, int h)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    const float* d_flow; size_t d_step;
    rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    rc_launch_polar_hist(c, d_flow, d_step, w, h, c->d_hist2d);
    CHECK_LAUNCH(c);
    if (flow && !is_device_ptr(flow)) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_hist_get(rc_ctx* c, int64_t* hist, int64_t* histsum, int64_t* hist2d, int64_t* histsum2d)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    int64_t tmp[RC_HIST_ROWS * RC_HIST_BINS];
    CUDA_TRY(c, cudaMemcpyAsync(tmp, c->d_hist2d, sizeof tmp, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (hist2d) memcpy(hist2d, tmp, sizeof tmp);
    int64_t total = 0;
    if (hist) memset(hist, 0, sizeof(int64_t) * RC_HIST_BINS);
    for (int a = 0; a < RC_HIST_ROWS; a++) {
        int64_t s = 0;
        for (int b = 0; b < RC_HIST_BINS; b++) { s += tmp[a * RC_HIST_BINS + b]; if (hist) hist[b] += tmp[a * RC_HIST_BINS + b]; }
        if (histsum2d) histsum2d[a] = s;
        total += s;
    }
    if (histsum) *histsum = total;
    return RC_OK;
}

int rc_hist_add(rc_ctx* c, const int64_t* hist2d)
{
    if (!c || !hist2d) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    int64_t tmp[RC_HIST_ROWS * RC_HIST_BINS];
    CUDA_TRY(c, cudaMemcpyAsync(tmp, c->d_hist2d, sizeof tmp, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < RC_HIST_ROWS * RC_HIST_BINS; i++) tmp[i] += hist2d[i];
    CUDA_TRY(c, cudaMemcpyAsync(c->d_hist2d, tmp, sizeof tmp, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_hist_device(rc_ctx* c, int64_t** dev_hist2d)
{
    if (!c || !dev_hist2d) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    *dev_hist2d = reinterpret_cast<int64_t*>(c->d_hist2d);
    return RC_OK;
}

int rc_cart_to_polar(rc_ctx* c, const float* flow, size_t n, float* mag, float* ang)
{
    if (!c || !flow || !mag || !ang) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (!n) return RC_OK;
    const bool hin = !is_device_ptr(flow), hout = !is_device_ptr(mag);
    const float* d_in = flow; float* d_mag = mag; float* d_ang = ang;
    if (hin) {
        int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, n * 8); if (rc) return rc;
        CUDA_TRY(c, cudaMemcpyAsync(c->d_tmp, flow, n * 8, cudaMemcpyHostToDevice, c->stream));
        d_in = reinterpret_cast<const float*>(c->d_tmp);
    }
    if (hout) {
        int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, n * 8); if (rc) return rc;
        d_mag = reinterpret_cast<float*>(c->d_tmp2); d_ang = d_mag + n;
    }
    rc_launch_cart_to_polar(c, d_in, n, d_mag, d_ang);
    CHECK_LAUNCH(c);
    if (hout) {
        CUDA_TRY(c, cudaMemcpyAsync(mag, d_mag, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(ang, d_ang, n * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (hin || hout) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// ---- A4 ----------------------------------------------------------------------------------------------------
int rc_thresholds(rc_ctx* c, float* UPPER, float* UPPER2d, float* prop)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    rc_launch_thresholds(c, c->d_hist2d, c->d_thr);
    CHECK_LAUNCH(c);
    if (UPPER || UPPER2d || prop) {
        float t[RC_THR_FLOATS];
        CUDA_TRY(c, cudaMemcpyAsync(t, c->d_thr, sizeof t, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (UPPER) *UPPER = t[0];
        if (UPPER2d) memcpy(UPPER2d, t + 1, sizeof(float) * RC_HIST_DIRECTIONS);
        if (prop) memcpy(prop, t + 37, sizeof(float) * RC_HIST_DIRECTIONS);
    }
    return RC_OK;
}

// ---- A5 ----------------------------------------------------------------------------------------------------
int rc_accumulator_reset(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (c->d_acc) CUDA_TRY(c, cudaMemsetAsync(c->d_acc, 0, sizeof(float) * (size_t)c->acc_w * c->acc_h, c->stream));
    return RC_OK;
}

static int classify_impl(rc_ctx* c, const float* d_flow, size_t d_step, int w, int h, float upper, int framecount,
                         uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass, bool with_window, bool* host_written)
{
    int rc = ensure_aggregate(c); if (rc) return rc;
    rc = ensure_accumulator(c, w, h); if (rc) return rc;
    const size_t n = (size_t)w * h;
    uint8_t* d_mask = outmask ? (is_device_ptr(outmask) ? outmask : c->d_mask) : nullptr;
    uint8_t* d_wave = waveclass ? (is_device_ptr(waveclass) ? waveclass : c->d_cls) : nullptr;
    uint8_t* d_water = waterclass ? (is_device_ptr(waterclass) ? waterclass : c->d_cls + n) : nullptr;
    float* slot = nullptr; float* avg = nullptr;
    if (with_window && c->win_W > 0 && c->win_w == w && c->win_h == h) {
        slot = c->d_ring + (size_t)c->win_i * n * 2; avg = c->d_avg;
        c->win_i = (c->win_i + 1) % c->win_W;
    }
    rc_launch_classify(c, d_flow, d_step, w, h, upper, c->d_thr, framecount, c->d_acc, d_mask, d_wave, d_water, slot, avg,
                       c->win_W);
    CHECK_LAUNCH(c);
    if (outmask && d_mask != outmask) { CUDA_TRY(c, cudaMemcpyAsync(outmask, d_mask, n, cudaMemcpyDeviceToHost, c->stream)); *host_written = true; }
    if (waveclass && d_wave != waveclass) { CUDA_TRY(c, cudaMemcpyAsync(waveclass, d_wave, n, cudaMemcpyDeviceToHost, c->stream)); *host_written = true; }
    if (waterclass && d_water != waterclass) { CUDA_TRY(c, cudaMemcpyAsync(waterclass, d_water, n, cudaMemcpyDeviceToHost, c->stream)); *host_written = true; }
    return RC_OK;
}

int rc_classify_accumulate(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, int framecount,
                           uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    bool host = flow && !is_device_ptr(flow);
    rc = classify_impl(c, d_flow, d_step, w, h, upper, framecount, outmask, waveclass, waterclass, false, &host);
    if (rc) return rc;
    if (host) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_accumulator_get(rc_ctx* c, float* acc_x)
{
    if (!c || !acc_x) return RC_ERR_INVALID;
    if (!c->d_acc) return fail(c, RC_ERR_STATE, "no accumulator yet%s");
    cudaSetDevice(c->device);
    const bool dev = is_device_ptr(acc_x);
    CUDA_TRY(c, cudaMemcpyAsync(acc_x, c->d_acc, sizeof(float) * (size_t)c->acc_w * c->acc_h,
                                dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (!dev) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_accumulator_device(rc_ctx* c, float** dev_acc_x, int* w, int* h)
{
    if (!c || !dev_acc_x) return RC_ERR_INVALID;
    if (!c->d_acc) return fail(c, RC_ERR_STATE, "no accumulator yet%s");
    *dev_acc_x = c->d_acc;
    if (w) *w = c->acc_w;
    if (h) *h = c->acc_h;
    return RC_OK;
}

// ---- A6 ----------------------------------------------------------------------------------------------------
int rc_window_configure(rc_ctx* c, int w, int h, int W)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (w < 1 || h < 1 || W < 0) return fail(c, RC_ERR_INVALID, "bad window geometry%s");
    if (c->d_ring) { cudaStreamSynchronize(c->stream); cudaFree(c->d_ring); cudaFree(c->d_avg); c->d_ring = c->d_avg = nullptr; }
    c->win_W = 0; c->win_i = 0;
    if (W == 0) return RC_OK;
    const size_t nb = sizeof(float) * 2 * (size_t)w * h;
    if (cudaMalloc((void**)&c->d_ring, nb * W) != cudaSuccess || cudaMalloc((void**)&c->d_avg, nb) != cudaSuccess) {
        cudaGetLastError();
        if (c->d_ring) cudaFree(c->d_ring);
        c->d_ring = c->d_avg = nullptr;
        return fail(c, RC_ERR_NOMEM, "cudaMalloc failed (window ring)%s");
    }
    CUDA_TRY(c, cudaMemsetAsync(c->d_ring, 0, nb * W, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_avg, 0, nb, c->stream));
    c->win_W = W; c->win_w = w; c->win_h = h;
    return RC_OK;
}

int rc_window_update(rc_ctx* c, const float* flow, size_t flow_step)
{
    if (!c) return RC_ERR_INVALID;
    if (c->win_W <= 0) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    cudaSetDevice(c->device);
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, c->win_w, c->win_h, &d_flow, &d_step); if (rc) return rc;
    if (!flow && (c->prm.w != c->win_w || c->prm.h != c->win_h)) return fail(c, RC_ERR_INVALID, "window / flow size mismatch%s");
    const size_t n = (size_t)c->win_w * c->win_h;
    rc_launch_window_update(c, d_flow, d_step, c->win_w, c->win_h, c->d_ring + (size_t)c->win_i * n * 2, c->d_avg, c->win_W);
    CHECK_LAUNCH(c);
    c->win_i = (c->win_i + 1) % c->win_W;
    if (flow && !is_device_ptr(flow)) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_window_get(rc_ctx* c, float* avg, size_t avg_step)
{
    if (!c || !avg) return RC_ERR_INVALID;
    if (c->win_W <= 0) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    cudaSetDevice(c->device);
    bool host = false;
    int rc = copy_out(c, avg, avg_step, c->d_avg, (size_t)c->win_w * 8, (size_t)c->win_w * 8, c->win_h, &host);
    if (rc) return rc;
    if (host) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_window_device(rc_ctx* c, float** dev_avg)
{
    if (!c || !dev_avg) return RC_ERR_INVALID;
    if (c->win_W <= 0) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    *dev_avg = c->d_avg;
    return RC_OK;
}

int rc_subtract_mean(rc_ctx* c, float* flow, size_t flow_step, int w, int h, double* mean_xy)
{
    if (!c || !flow || w < 1 || h < 1 || flow_step < (size_t)w * 8) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const bool dev = is_device_ptr(flow);
    float* d_flow = flow; size_t d_step = flow_step;
    if (!dev) {
        int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, (size_t)w * h * 8); if (rc) return rc;
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d_tmp, (size_t)w * 8, flow, flow_step, (size_t)w * 8, h, cudaMemcpyHostToDevice, c->stream));
        d_flow = reinterpret_cast<float*>(c->d_tmp); d_step = (size_t)w * 8;
    }
    int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 64); if (rc) return rc;
    double* d_sums = reinterpret_cast<double*>(c->d_tmp2);
    rc_launch_subtract_mean(c, d_flow, d_step, w, h, d_sums);
    CHECK_LAUNCH(c);
    if (!dev) CUDA_TRY(c, cudaMemcpy2DAsync(flow, flow_step, d_flow, d_step, (size_t)w * 8, h, cudaMemcpyDeviceToHost, c->stream));
    if (mean_xy || !dev) {
        double s[2];
        CUDA_TRY(c, cudaMemcpyAsync(s, d_sums, sizeof s, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        if (mean_xy) { mean_xy[0] = s[0] / ((double)w * h); mean_xy[1] = s[1] / ((double)w * h); }
    }
    return RC_OK;
}

// ---- A7 ----------------------------------------------------------------------------------------------------
int rc_advect(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* seeds, size_t n, float dt,
              int iterations, float upper, int variant, float* dist, const int32_t* home)
{
    if (!c || !seeds) return RC_ERR_INVALID;
    if (variant < RC_ADV_PATHLINE || variant > RC_ADV_GET_DELTA) return fail(c, RC_ERR_INVALID, "unknown advection variant%s");
    if (iterations < 0) return fail(c, RC_ERR_INVALID, "iterations < 0%s");
    cudaSetDevice(c->device);
    if (!n) return RC_OK;
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    const bool hs = !is_device_ptr(seeds);
    const bool hd = dist && !is_device_ptr(dist);
    const bool hh = home && !is_device_ptr(home);
    float* d_seeds = seeds; float* d_dist = dist; const int32_t* d_home = home;
    if (hs || hd || hh) {
        // layout of the scratch: seeds | dist | home
        size_t need = n * 8 + n * 4 + n * 8;
        rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, need); if (rc) return rc;
        char* base = reinterpret_cast<char*>(c->d_tmp2);
        if (hs) { d_seeds = reinterpret_cast<float*>(base); CUDA_TRY(c, cudaMemcpyAsync(d_seeds, seeds, n * 8, cudaMemcpyHostToDevice, c->stream)); }
        if (hd) { d_dist = reinterpret_cast<float*>(base + n * 8); CUDA_TRY(c, cudaMemcpyAsync(d_dist, dist, n * 4, cudaMemcpyHostToDevice, c->stream)); }
        if (hh) { int32_t* dh = reinterpret_cast<int32_t*>(base + n * 12); CUDA_TRY(c, cudaMemcpyAsync(dh, home, n * 8, cudaMemcpyHostToDevice, c->stream)); d_home = dh; }
    }
    rc_launch_advect(c, d_flow, d_step, w, h, d_seeds, n, dt, iterations, upper, variant, d_dist, d_home);
    CHECK_LAUNCH(c);
    if (hs) CUDA_TRY(c, cudaMemcpyAsync(seeds, d_seeds, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (hd) CUDA_TRY(c, cudaMemcpyAsync(dist, d_dist, n * 4, cudaMemcpyDeviceToHost, c->stream));
    if (hs || hd || hh || (flow && !is_device_ptr(flow))) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_streakline_step(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* emitters, int E,
                       float* vertices, int32_t* count, int cap, float dt)
{
    if (!c || !emitters || !vertices || !count || E < 0 || cap < 1) return RC_ERR_INVALID;
    if (E > 65535) return fail(c, RC_ERR_UNSUPPORTED, "at most 65535 emitters per call%s");
    cudaSetDevice(c->device);
    if (!E) return RC_OK;
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    const size_t vb = (size_t)E * cap * 8;
    const bool hv = !is_device_ptr(vertices), he = !is_device_ptr(emitters), hc = !is_device_ptr(count);
    if (hv != hc) return fail(c, RC_ERR_INVALID, "vertices and count must both be host or both be device pointers%s");
    // scratch layout in d_tmp2: [vout vb][vin vb (host case)][emitters E*8][count E*4]
    rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 2 * vb + (size_t)E * 12 + 64); if (rc) return rc;
    char* base = reinterpret_cast<char*>(c->d_tmp2);
    float* d_v = vertices; const float* d_e = emitters; int32_t* d_c = count;
    if (hv) {
        d_v = reinterpret_cast<float*>(base + vb); d_c = reinterpret_cast<int32_t*>(base + 2 * vb + (size_t)E * 8);
        CUDA_TRY(c, cudaMemcpyAsync(d_v, vertices, vb, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(d_c, count, (size_t)E * 4, cudaMemcpyHostToDevice, c->stream));
    }
    if (he) {
        float* de = reinterpret_cast<float*>(base + 2 * vb);
        CUDA_TRY(c, cudaMemcpyAsync(de, emitters, (size_t)E * 8, cudaMemcpyHostToDevice, c->stream));
        d_e = de;
    }
    rc_launch_streakline(c, d_flow, d_step, w, h, d_e, E, d_v, d_c, cap, dt);
    CHECK_LAUNCH(c);
    if (hv) {
        CUDA_TRY(c, cudaMemcpyAsync(vertices, d_v, vb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(count, d_c, (size_t)E * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (hv || he || (flow && !is_device_ptr(flow))) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// ---- fused per-frame step ------------------------------------------------------------------------------------
int rc_process_frame(rc_ctx* c, const uint8_t* frame, size_t step, int framecount, uint8_t* outmask,
                     rc_frame_result* result)
{
    if (!c) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    int produced = flow_push_impl(c, frame, step);
    if (produced < 0) return produced;
    bool host = !is_device_ptr(frame);
    const int w = c->prm.w, h = c->prm.h;
    if (produced) {
        rc_launch_polar_hist(c, c->flow_out, (size_t)w * 8, w, h, c->d_hist2d);
        rc_launch_thresholds(c, c->d_hist2d, c->d_thr);
        CHECK_LAUNCH(c);
        rc = classify_impl(c, c->flow_out, (size_t)w * 8, w, h, NAN, framecount, outmask, nullptr, nullptr, true, &host);
        if (rc) return rc;
    }
    if (result) {
        memset(result, 0, sizeof *result);
        result->produced = produced;
        if (produced) {
            float t[RC_THR_FLOATS];
            CUDA_TRY(c, cudaMemcpyAsync(t, c->d_thr, sizeof t, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(c, cudaStreamSynchronize(c->stream));
            result->UPPER = t[0];
            memcpy(result->UPPER2d, t + 1, sizeof(float) * RC_HIST_DIRECTIONS);
            memcpy(result->prop_above_upper, t + 37, sizeof(float) * RC_HIST_DIRECTIONS);
            memcpy(&result->histsum, t + 74, sizeof(int64_t));
            host = false;
        }
    }
    if (host) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return produced;
}

}  // extern "C"
