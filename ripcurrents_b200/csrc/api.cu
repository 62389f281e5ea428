// C-ABI layer (include/ripcurrents_b200.h): context, buffer management, host<->device staging, call order.
#include <math.h>
#include <float.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <new>

#include "rc_internal.h"

#define RC_VERSION 101

const char* const rc_kernel_names[K_COUNT] = {"reserved", "pyramid", "polyexp", "update_matrices", "flow_iter_fused",
                                              "flow_iter_final", "flow_layer_fused", "polar_hist", "thresholds",
                                              "classify", "window_mean", "advect", "streakline", "misc",
                                              "particle_fields", "diagnostics"};

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
        if (!(c)->launch_err.empty()) {                                                       \
            const std::string _m = (c)->launch_err; (c)->launch_err.clear();                  \
            return fail((c), RC_ERR_CUDA, "kernel launch failed: %s", _m.c_str());            \
        }                                                                                     \
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
        *p = nullptr;
        return fail(c, RC_ERR_NOMEM, "cudaMalloc failed%s");
    }
    c->allocs.push_back(*p);
    return RC_OK;
}

void sync_all(rc_ctx* c)
{
    if (c->s_in) cudaStreamSynchronize(c->s_in);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->s_out) cudaStreamSynchronize(c->s_out);
    rc_comm_fence(c, false);
}

void free_farneback(rc_ctx* c)
{
    sync_all(c);
    for (void* p : c->allocs) cudaFree(p);
    c->allocs.clear();
    for (auto& L : c->layer) L = Layer();
    for (int s = 0; s < 2; s++) {
        c->d_frames[s] = nullptr; c->d_masks[s] = nullptr; c->d_thr_batch[s] = nullptr;
        if (c->h_thr[s]) { cudaFreeHost(c->h_thr[s]); c->h_thr[s] = nullptr; }
        c->pending_results[s] = nullptr; c->pending_count[s] = 0;
    }
    c->flow_ring = nullptr; c->ring_slots = 0; c->d_avg = nullptr; c->d_hist_delta = nullptr;
    c->configured = false; c->nlayers = 0; c->frames_seen = 0; c->n_flows = 0; c->pairs_done = 0; c->win_start = 0; c->submitted = 0;
}

int round_half_even(double v) { return (int)nearbyint(v); }

// Appendix A.3 kernels and the closed-form inverse of the moment matrix
void make_poly(PolyCoef& pc, int n, double sigma, bool strict)
{
    memset(&pc, 0, sizeof pc);
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
        // fast mode: drop taps whose x^2-weighted weight is below 1e-7 of the centre weight (below fp32 rounding of the
        // sums they would join: the flow error against cv2 is the same to three digits as with 1e-9, which keeps one more
        // tap at sigma = 1.2 -- profiles/r02_polyexp_truncation_{8,7,6}taps.json; 1e-5 is visibly outside)
        static const double trunc = getenv("RC_POLY_TRUNC") ? atof(getenv("RC_POLY_TRUNC")) : 1e-7;
        int k = n;
        while (k > 1 && (double)pc.xxg[k] < trunc * (double)pc.g[0]) k--;
        pc.n_eff = k;
    }
}

void make_smooth(SmoothCoef& sc, double sigma, int ksize)
{
    memset(&sc, 0, sizeof sc);
    sc.ksize = ksize;
    if (sigma <= 0 && ksize == 3) { sc.k[0] = 0.25f; sc.k[1] = 0.5f; sc.k[2] = 0.25f; return; }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    std::vector<double> t(ksize);
    double s2 = -0.5 / (sigma * sigma), sum = 0;
    for (int i = 0; i < ksize; i++) { double x = i - (ksize - 1) * 0.5; t[i] = exp(s2 * x * x); sum += t[i]; }
    sum = 1.0 / sum;
    for (int i = 0; i < ksize; i++) sc.k[i] = (float)(t[i] * sum);
}

void make_win(WinCoef& g, int winsize, bool gaussian)
{
    memset(&g, 0, sizeof g);
    int m = winsize / 2;
    g.m = m; g.gaussian = gaussian ? 1 : 0;
    if (!gaussian) {
        for (int i = 0; i <= m; i++) g.k[i] = 1.f;
        g.post_scale_d = 1.0 / ((double)winsize * winsize);
        g.post_scale = (float)g.post_scale_d;
        return;
    }
    double sigma = m * 0.3, s = 1.0;
    g.k[0] = 1.f;
    for (int i = 1; i <= m; i++) { float t = (float)exp(-i * i / (2 * sigma * sigma)); g.k[i] = t; s += t * 2; }
    s = 1.0 / s;
    for (int i = 0; i <= m; i++) g.k[i] = (float)(g.k[i] * s);
    g.post_scale = 1.f; g.post_scale_d = 1.0;
}

int ensure_aggregate(rc_ctx* c)
{
    if (c->d_hist2d) return RC_OK;
    // [0, CELLS): the cumulative counters; [CELLS, 2*CELLS): scratch for the batched thresholds launch
    CUDA_TRY(c, cudaMalloc((void**)&c->d_hist2d, sizeof(unsigned long long) * RC_HIST_CELLS * 2));
    CUDA_TRY(c, cudaMemsetAsync(c->d_hist2d, 0, sizeof(unsigned long long) * RC_HIST_CELLS * 2, c->stream));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_thr, sizeof(float) * RC_THR_FLOATS));
    CUDA_TRY(c, cudaMemsetAsync(c->d_thr, 0, sizeof(float) * RC_THR_FLOATS, c->stream));
    return RC_OK;
}

int ensure_accumulator(rc_ctx* c, int w, int h)
{
    if (c->d_acc && c->acc_w == w && c->acc_h == h) return RC_OK;
    if (c->d_acc) { sync_all(c); cudaFree(c->d_acc); cudaFree(c->d_cls); }
    c->d_acc = nullptr; c->d_cls = nullptr;
    const size_t n = (size_t)w * h;
    CUDA_TRY(c, cudaMalloc((void**)&c->d_acc, sizeof(float) * (n + 4)));
    CUDA_TRY(c, cudaMalloc((void**)&c->d_cls, 3 * n));
    CUDA_TRY(c, cudaMemsetAsync(c->d_acc, 0, sizeof(float) * (n + 4), c->stream));
    c->acc_w = w; c->acc_h = h;
    return RC_OK;
}

// (Re)allocates the layer-0 flow ring: W + B slots so that the flow leaving a W-frame window is still resident
// while a batch of up to B new flows is written.
int ensure_flow_ring(rc_ctx* c)
{
    const int want = (c->win_W > 0 ? c->win_W : 0) + c->B;
    if (c->flow_ring && c->ring_slots == want) return RC_OK;
    const size_t per = sizeof(float) * 2 * (size_t)c->prm.w * c->prm.h;
    sync_all(c);
    int rc;
    // (an older ring, if any, stays in `allocs` until the context is reconfigured: window changes are rare)
    if ((rc = dev_alloc(c, (void**)&c->flow_ring, per * want + 64))) return rc;
    if (!c->d_avg && (rc = dev_alloc(c, (void**)&c->d_avg, per + 64))) return rc;
    CUDA_TRY(c, cudaMemsetAsync(c->flow_ring, 0, per * want, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_avg, 0, per, c->stream));
    c->ring_slots = want;
    c->pairs_done = 0; c->win_start = 0;
    return RC_OK;
}

float* ring_slot(rc_ctx* c, long long pair)
{
    return c->flow_ring + (size_t)(pair % c->ring_slots) * 2 * (size_t)c->prm.w * c->prm.h;
}

// Brings a (possibly host, possibly strided) flow field to the device when needed.
int stage_flow_in(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float** d_flow, size_t* d_step)
{
    if (!flow) {
        if (!c->pairs_done) return fail(c, RC_ERR_STATE, "no flow has been computed on this context%s");
        *d_flow = ring_slot(c, c->pairs_done - 1); *d_step = (size_t)c->prm.w * 8;
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

void prof_drain(rc_ctx* c)
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

// Expands + computes flows for `count` device-resident frames (rows of `step` bytes, frames `fstride` apart),
// in sub-batches of at most B.  Returns the number of flows produced (count, or count-1 when priming).
// With `aggregate` each sub-batch is followed, in stream order, by thresholds + classify (+ window); the per-frame
// threshold records / masks go to thr_out / masks_out (device, indexed by produced flow).
// aggregate: 0 = flows only, 1 = flows + per-frame counts + thresholds + classify (+ window), 2 = flows + per-frame counts
// (left in d_hist_delta; sharded streams).  dst_override (optional): where the layer-0 flow of produced pair j goes instead
// of the context's flow ring.
int run_frames(rc_ctx* c, const uint8_t* d_frames, size_t step, size_t fstride, int count, int aggregate,
               int framecount0, float* thr_out, uint8_t* masks_out, float* const* dst_override = nullptr)
{
    const int w = c->prm.w, h = c->prm.h, B = c->B, nslots = B + 1;
    const size_t n = (size_t)w * h;
    int produced = 0, done = 0;
    if (c->frames_seen == 0 && count > 0) {           // priming frame: expand only
        rc_launch_expand(c, d_frames, step, fstride, 1, c->r_base);
        c->frames_seen = 1; done = 1;
    }
    while (done < count) {
        const int nb = count - done < B ? count - done : B;
        const int first = (c->r_base + 1) % nslots;
        static const double overlap_max_px = getenv("RC_OVERLAP_MAXPX") ? atof(getenv("RC_OVERLAP_MAXPX")) : 17e6;
        rc_launch_expand(c, d_frames + (size_t)done * fstride, step, fstride, nb, first, (double)nb * n <= overlap_max_px);
        float* dst[RC_MAX_BATCH];
        for (int j = 0; j < nb; j++) dst[j] = dst_override ? dst_override[produced + j] : ring_slot(c, c->pairs_done + j);
        rc_launch_flows(c, nb, c->r_base, dst, aggregate ? c->d_hist_delta : nullptr);
        if (aggregate == 1) {
            float* thr = thr_out + (size_t)produced * RC_THR_FLOATS;
            rc_launch_thresholds_batch(c, c->d_hist2d, c->d_hist_delta, nb, thr, c->d_thr);
            ClassifyBatch cb;
            cb.nb = nb;
            for (int j = 0; j < nb; j++) {
                cb.flow[j] = dst[j];
                const long long p = c->pairs_done + j;
                cb.old[j] = (c->win_W > 0 && p - c->win_start >= c->win_W) ? ring_slot(c, p - c->win_W) : nullptr;
            }
            // framecount of a flow = the caller's loop counter of the frame that completed the pair
            const size_t mbytes = c->mask_format == RC_MASK_PACKED ? n / 8 : n;
            rc_launch_classify_batch(c, cb, w, h, thr, framecount0 + done, c->d_acc,
                                     masks_out ? masks_out + (size_t)produced * mbytes : nullptr,
                                     c->win_W > 0 ? c->d_avg : nullptr, c->win_W, c->mask_format == RC_MASK_PACKED ? 1 : 0);
        }
        c->r_base = (c->r_base + nb) % nslots;
        c->pairs_done += nb; c->frames_seen += nb;
        produced += nb; done += nb;
    }
    c->n_flows = produced;
    return produced;
}

void fill_results(const float* h_thr, int produced, int first_produced, int count, rc_frame_result* results)
{
    if (!results) return;
    memset(results, 0, sizeof(rc_frame_result) * count);
    for (int i = 0; i < count; i++) {
        const int j = i - first_produced;             // index among the produced flows
        if (j < 0 || j >= produced) continue;
        const float* t = h_thr + (size_t)j * RC_THR_FLOATS;
        results[i].produced = 1;
        results[i].UPPER = t[0];
        memcpy(results[i].UPPER2d, t + 1, sizeof(float) * RC_HIST_DIRECTIONS);
        memcpy(results[i].prop_above_upper, t + 37, sizeof(float) * RC_HIST_DIRECTIONS);
        memcpy(&results[i].histsum, t + 74, sizeof(int64_t));
    }
}

// Finishes the batch staged in `slot`: waits for its D2H copies and fills the caller's result records.
int finish_slot(rc_ctx* c, int slot)
{
    if (!c->pending_count[slot]) return RC_OK;
    CUDA_TRY(c, cudaEventSynchronize(c->ev_out[slot]));
    const int count = c->pending_count[slot];
    const int first = c->pending_first_produced[slot];
    fill_results(c->h_thr[slot], count - first, first, count, c->pending_results[slot]);
    c->pending_count[slot] = 0; c->pending_results[slot] = nullptr;
    return RC_OK;
}

}  // namespace

// ---- helpers shared with comm.cu ----------------------------------------------------------------------------------
int rc_fail(rc_ctx* c, int code, const char* fmt, const char* detail) { return fail(c, code, fmt, detail); }
bool rc_is_device_ptr(const void* p) { return is_device_ptr(p); }
int rc_ensure_aggregate(rc_ctx* c) { return ensure_aggregate(c); }
int rc_ensure_accumulator(rc_ctx* c, int w, int h) { return ensure_accumulator(c, w, h); }
float* rc_ring_slot(rc_ctx* c, long long pair) { return ring_slot(c, pair); }
void rc_fill_results(const float* h_thr, int produced, int first_produced, int count, rc_frame_result* results)
{
    fill_results(h_thr, produced, first_produced, count, results);
}
int rc_run_frames_hist(rc_ctx* c, const uint8_t* d_frames, size_t step, size_t fstride, int count, float* const* dst_override)
{
    return run_frames(c, d_frames, step, fstride, count, 2, 0, nullptr, nullptr, dst_override);
}

// =====================================================================================================
namespace {
// Host pointers are staged through one scratch arena (d_tmp); device pointers pass through.
struct Stager {
    rc_ctx* c; char* base = nullptr; size_t off = 0;
    struct Out { void* host; void* dev; size_t bytes; };
    std::vector<Out> outs;
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <class T> const T* in(const T* p, size_t bytes)
    {
        if (!p || is_device_ptr(p)) return p;
        T* d = reinterpret_cast<T*>(base + off); off += pad(bytes);
        cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream);
        return d;
    }
    template <class T> T* out(T* p, size_t bytes, bool also_in = false)
    {
        if (!p || is_device_ptr(p)) return p;
        T* d = reinterpret_cast<T*>(base + off); off += pad(bytes);
        if (also_in) cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream);
        outs.push_back(Out{p, d, bytes});
        return d;
    }
    int finish()
    {
        for (auto& o : outs) cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream);
        if (!outs.empty() && cudaStreamSynchronize(c->stream) != cudaSuccess)
            return fail(c, RC_ERR_CUDA, "particle fields: %s", cudaGetErrorString(cudaGetLastError()));
        return RC_OK;
    }
};

int fields_impl(rc_ctx* c, const float* field, const float* src0, const float* dist, size_t n, int flags, int want,
                float* mag_out, float* ratio_out, uint8_t* const gray[3], uint8_t* const bgr[3], double* maxes)
{
    cudaSetDevice(c->device);
    const size_t need = 36 * n + 16 * 256;     // upper bound: every array on the host
    int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, need); if (rc) return rc;
    Stager st{c, reinterpret_cast<char*>(c->d_tmp)};
    unsigned* d_enc = reinterpret_cast<unsigned*>(st.base); st.off = 256;
    double* d_max = nullptr;
    if (maxes) { d_max = reinterpret_cast<double*>(st.base + st.off); st.off += 256; st.outs.push_back({maxes, d_max, 24}); }
    field = st.in(field, n * 8); src0 = st.in(src0, n * 4); dist = st.in(dist, n * 4);
    mag_out = st.out(mag_out, n * 4); ratio_out = st.out(ratio_out, n * 4);
    uint8_t* g[3]; uint8_t* b[3];
    for (int k = 0; k < 3; k++) { g[k] = st.out(gray[k], n); b[k] = st.out(bgr[k], 3 * n); }
    rc_launch_fields(c, field, src0, dist, n, (flags & RC_FIELDS_DIV_ZERO_IS_ZERO) ? 1 : 0, want, mag_out, ratio_out, g, b,
                     d_enc, d_max);
    CHECK_LAUNCH(c);
    return st.finish();
}
}  // namespace

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
    rc_farneback_init_device(device);
    bool ok = cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_pyr, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_poly[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&c->ev_poly[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; ok && s < 2; s++)
        ok = cudaEventCreateWithFlags(&c->ev_in[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_compute[s], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&c->ev_out[s], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); delete c; return RC_ERR_CUDA; }
    c->own_stream = true;
    *out = c;
    return RC_OK;
}

void rc_destroy(rc_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    rc_comm_destroy(c);
    rc_comm_fence(c, true);
    free_farneback(c);
    void* bufs[] = {c->d_bgr[0], c->d_bgr[1], c->d_tmp, c->d_tmp2, c->d_diag, c->d_hist2d, c->d_thr, c->d_acc, c->d_cls, c->d_swin_ring, c->d_swin_avg};
    for (void* p : bufs) if (p) cudaFree(p);
    prof_drain(c);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    for (int s = 0; s < 2; s++) {
        if (c->ev_in[s]) cudaEventDestroy(c->ev_in[s]);
        if (c->ev_compute[s]) cudaEventDestroy(c->ev_compute[s]);
        if (c->ev_out[s]) cudaEventDestroy(c->ev_out[s]);
    }
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    if (c->s_aux) cudaStreamDestroy(c->s_aux);
    if (c->ev_pyr) cudaEventDestroy(c->ev_pyr);
    for (int i = 0; i < 2; i++) if (c->ev_poly[i]) cudaEventDestroy(c->ev_poly[i]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int rc_set_stream(rc_ctx* c, void* s)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    sync_all(c);
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
    cudaSetDevice(c->device);
    CUDA_TRY(c, cudaStreamSynchronize(c->s_in));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_out));
    if (rc_comm_fence(c, false)) return fail(c, RC_ERR_CUDA, "communication stream: %s", cudaGetErrorString(cudaGetLastError()));
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
int rc_flow_configure_batch(rc_ctx* c, int w, int h, double pyr_scale, int levels, int winsize, int iterations,
                            int poly_n, double poly_sigma, int flags, int max_batch)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    if (w < 2 || h < 2 || levels < 0) return fail(c, RC_ERR_INVALID, "bad geometry / levels%s");
    if (!(pyr_scale > 0.0 && pyr_scale < 1.0)) return fail(c, RC_ERR_INVALID, "pyr_scale must be in (0,1)%s");
    if (flags & 4) return fail(c, RC_ERR_UNSUPPORTED, "OPTFLOW_USE_INITIAL_FLOW is not supported%s");
    if (poly_n < 1 || poly_n > RC_MAX_POLY_N) return fail(c, RC_ERR_UNSUPPORTED, "poly_n must be in [1,32]%s");
    if (winsize < 1 || winsize / 2 > RC_MAX_WIN_HALF) return fail(c, RC_ERR_UNSUPPORTED, "winsize must be in [1,129]%s");
    if (iterations < 1) return fail(c, RC_ERR_UNSUPPORTED, "iterations must be >= 1%s");
    if (max_batch < 1 || max_batch > RC_MAX_BATCH) return fail(c, RC_ERR_INVALID, "max_batch must be in [1,64]%s");
    FarnebackParams p;
    p.w = w; p.h = h; p.pyr_scale = pyr_scale; p.levels = levels; p.winsize = winsize; p.iterations = iterations;
    p.poly_n = poly_n; p.poly_sigma = poly_sigma; p.flags = flags; p.max_batch = max_batch;
    if (c->configured && c->prm == p) {
        // restart of the stream (a new clip): the next frame primes, and the sliding window starts from zero buffers
        // like the reference's (main.cpp:1084-1092) -- flows of the previous clip never leave a mean they did not enter
        sync_all(c);
        c->frames_seen = 0; c->n_flows = 0;
        c->win_start = c->pairs_done;
        if (c->d_avg) CUDA_TRY(c, cudaMemsetAsync(c->d_avg, 0, sizeof(float) * 2 * (size_t)w * h, c->stream));
        return RC_OK;
    }
    free_farneback(c);
    c->prm = p; c->B = max_batch;
    const int B = max_batch;

    // Appendix A.1: layer selection
    int k; double scale = 1.0;
    for (k = 0; k < levels; k++) { scale *= pyr_scale; if (w * scale < 32 || h * scale < 32) break; }
    const int nl = k + 1;
    if (nl > RC_MAX_LAYERS) return fail(c, RC_ERR_UNSUPPORTED, "too many pyramid layers%s");
    c->strict = (flags & RC_FARNEBACK_STRICT) != 0;
    scale = 1.0;
    int rc = RC_OK;
    for (k = 0; k < nl && !rc; k++) {
        Layer& L = c->layer[k];
        L.w = round_half_even(w * scale); L.h = round_half_even(h * scale);
        L.pitch = (L.w + 31) / 32 * 32;
        L.plane = (size_t)L.pitch * L.h;
        double sigma = (1.0 / scale - 1.0) * 0.5;
        int ks = round_half_even(sigma * 5.0) | 1; if (ks < 3) ks = 3;
        if (ks > RC_MAX_SMOOTH_TAPS) { rc = fail(c, RC_ERR_UNSUPPORTED, "pyramid too deep (smoothing kernel > 255 taps)%s"); break; }
        make_smooth(L.smooth, sigma, ks);
        const bool fused = !c->strict && winsize / 2 == 1 && iterations <= 3;
        if ((rc = dev_alloc(c, (void**)&L.I, sizeof(float) * L.plane * B)) ||
            (rc = dev_alloc(c, (void**)&L.R, sizeof(float) * 5 * L.plane * (B + 1))) ||
            (!fused && (rc = dev_alloc(c, (void**)&L.M, sizeof(float) * 2 * 5 * L.plane * B))) ||
            (k > 0 && (rc = dev_alloc(c, (void**)&L.flow, sizeof(float) * 2 * (size_t)L.w * L.h * B))))
            break;
        scale *= pyr_scale;
    }
    const size_t n = (size_t)w * h;
    for (int s = 0; s < 2 && !rc; s++) {
        if ((rc = dev_alloc(c, (void**)&c->d_frames[s], n * B + 64)) || (rc = dev_alloc(c, (void**)&c->d_masks[s], n * B + 64)) ||
            (rc = dev_alloc(c, (void**)&c->d_thr_batch[s], sizeof(float) * RC_THR_FLOATS * B)))
            break;
        if (cudaMallocHost((void**)&c->h_thr[s], sizeof(float) * RC_THR_FLOATS * B) != cudaSuccess) {
            cudaGetLastError(); rc = fail(c, RC_ERR_NOMEM, "cudaMallocHost failed%s");
        }
    }
    if (!rc) rc = dev_alloc(c, (void**)&c->d_hist_delta, sizeof(unsigned int) * RC_HIST_CELLS * B);
    if (rc) { free_farneback(c); return rc; }
    c->nlayers = nl;
    make_poly(c->poly, poly_n, poly_sigma, c->strict);
    make_win(c->win, winsize, (flags & RC_FARNEBACK_GAUSSIAN) != 0);
    c->configured = true; c->frames_seen = 0; c->n_flows = 0; c->r_base = 0; c->pairs_done = 0; c->submitted = 0;
    if (c->win_W > 0 && (c->swin_w != w || c->swin_h != h)) c->win_W = 0;
    rc = ensure_flow_ring(c);
    if (rc) { free_farneback(c); return rc; }
    return RC_OK;
}

int rc_flow_configure(rc_ctx* c, int w, int h, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                      double poly_sigma, int flags)
{
    return rc_flow_configure_batch(c, w, h, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, 1);
}

int rc_flow_push_batch(rc_ctx* c, const uint8_t* frames, size_t step, size_t frame_stride, int count, float* flows,
                       size_t flow_step, size_t flow_stride)
{
    if (!c) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    if (count < 1) return fail(c, RC_ERR_INVALID, "count < 1%s");
    cudaSetDevice(c->device);
    const int w = c->prm.w, h = c->prm.h;
    if (!frames || step < (size_t)w || (count > 1 && frame_stride < step * (size_t)(h - 1) + w))
        return fail(c, RC_ERR_INVALID, "bad frame pointer / step / stride%s");
    const uint8_t* d = frames; size_t ds = step, dfs = frame_stride;
    bool host = false;
    if (!is_device_ptr(frames)) {
        if (count > c->B) return fail(c, RC_ERR_INVALID, "more host frames than max_batch%s");
        host = true;
        for (int j = 0; j < count; j++)
            CUDA_TRY(c, cudaMemcpy2DAsync(c->d_frames[0] + (size_t)j * w * h, w, frames + (size_t)j * frame_stride, step, w, h,
                                          cudaMemcpyHostToDevice, c->stream));
        d = c->d_frames[0]; ds = w; dfs = (size_t)w * h;
    }
    if (flows && count - (c->frames_seen == 0 ? 1 : 0) > c->ring_slots)
        return fail(c, RC_ERR_INVALID, "more flows requested than the flow ring holds%s");
    const long long first_pair = c->pairs_done;
    int produced = run_frames(c, d, ds, dfs, count, 0, 0, nullptr, nullptr);
    CHECK_LAUNCH(c);
    if (flows && produced > 0) {
        if (flow_step < (size_t)w * 8) return fail(c, RC_ERR_INVALID, "flow_step too small%s");
        const bool dev = is_device_ptr(flows);
        for (int j = 0; j < produced; j++)
            CUDA_TRY(c, cudaMemcpy2DAsync(reinterpret_cast<char*>(flows) + (size_t)j * flow_stride, flow_step,
                                          ring_slot(c, first_pair + j), (size_t)w * 8, (size_t)w * 8, h,
                                          dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
        if (!dev) host = true;
    }
    if (host) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return produced;
}

int rc_flow_push(rc_ctx* c, const uint8_t* frame, size_t step, float* flow, size_t flow_step)
{
    return rc_flow_push_batch(c, frame, step, 0, 1, flow, flow_step, 0);
}

int rc_farneback(rc_ctx* c, const uint8_t* prev, size_t prev_step, const uint8_t* next, size_t next_step, int w, int h,
                 float* flow, size_t flow_step, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                 double poly_sigma, int flags)
{
    if (!c) return RC_ERR_INVALID;
    int rc;
    const bool same = c->configured && c->prm.w == w && c->prm.h == h && c->prm.pyr_scale == pyr_scale &&
                      c->prm.levels == levels && c->prm.winsize == winsize && c->prm.iterations == iterations &&
                      c->prm.poly_n == poly_n && c->prm.poly_sigma == poly_sigma && c->prm.flags == flags;
    if (!same) {
        rc = rc_flow_configure_batch(c, w, h, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags, 1);
        if (rc) return rc;
    }
    c->frames_seen = 0;      // independent two-frame call: nothing cached is reused
    rc = rc_flow_push(c, prev, prev_step, nullptr, 0);
    if (rc < 0) return rc;
    rc = rc_flow_push(c, next, next_step, flow, flow_step);
    return rc < 0 ? rc : RC_OK;
}

int rc_flow_device(rc_ctx* c, float** dev_flow, int* w, int* h)
{
    if (!c || !dev_flow) return RC_ERR_INVALID;
    if (!c->pairs_done) return fail(c, RC_ERR_STATE, "no flow has been computed on this context%s");
    *dev_flow = ring_slot(c, c->pairs_done - 1);
    if (w) *w = c->prm.w;
    if (h) *h = c->prm.h;
    return RC_OK;
}

int rc_flow_device_at(rc_ctx* c, int back, float** dev_flow)
{
    if (!c || !dev_flow || back < 0) return RC_ERR_INVALID;
    if (back >= c->pairs_done || back >= c->ring_slots) return fail(c, RC_ERR_STATE, "that flow is no longer resident%s");
    *dev_flow = ring_slot(c, c->pairs_done - 1 - back);
    return RC_OK;
}

// ---- A2 + A3 --------------------------------------------------------------------------------------------
int rc_hist_reset(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    CUDA_TRY(c, cudaMemsetAsync(c->d_hist2d, 0, sizeof(unsigned long long) * RC_HIST_CELLS, c->stream));
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
    int64_t tmp[RC_HIST_CELLS];
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
    int64_t tmp[RC_HIST_CELLS];
    CUDA_TRY(c, cudaMemcpyAsync(tmp, c->d_hist2d, sizeof tmp, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < RC_HIST_CELLS; i++) tmp[i] += hist2d[i];
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
    rc_launch_thresholds_batch(c, c->d_hist2d, nullptr, 0, nullptr, c->d_thr);
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

int rc_classify_accumulate(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, int framecount,
                           uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    bool host = flow && !is_device_ptr(flow);
    rc = ensure_aggregate(c); if (rc) return rc;
    rc = ensure_accumulator(c, w, h); if (rc) return rc;
    const size_t n = (size_t)w * h;
    uint8_t* d_mask = outmask ? (is_device_ptr(outmask) ? outmask : c->d_cls) : nullptr;
    uint8_t* d_wave = waveclass ? (is_device_ptr(waveclass) ? waveclass : c->d_cls + n) : nullptr;
    uint8_t* d_water = waterclass ? (is_device_ptr(waterclass) ? waterclass : c->d_cls + 2 * n) : nullptr;
    rc_launch_classify(c, d_flow, d_step, w, h, upper, c->d_thr, framecount, c->d_acc, d_mask, d_wave, d_water);
    CHECK_LAUNCH(c);
    if (outmask && d_mask != outmask) { CUDA_TRY(c, cudaMemcpyAsync(outmask, d_mask, n, cudaMemcpyDeviceToHost, c->stream)); host = true; }
    if (waveclass && d_wave != waveclass) { CUDA_TRY(c, cudaMemcpyAsync(waveclass, d_wave, n, cudaMemcpyDeviceToHost, c->stream)); host = true; }
    if (waterclass && d_water != waterclass) { CUDA_TRY(c, cudaMemcpyAsync(waterclass, d_water, n, cudaMemcpyDeviceToHost, c->stream)); host = true; }
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
    if (!c->d_acc) {
        if (!c->configured) return fail(c, RC_ERR_STATE, "no accumulator yet%s");
        cudaSetDevice(c->device);
        int rc = ensure_accumulator(c, c->prm.w, c->prm.h); if (rc) return rc;
    }
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
    sync_all(c);
    if (c->d_swin_ring) { cudaFree(c->d_swin_ring); cudaFree(c->d_swin_avg); c->d_swin_ring = c->d_swin_avg = nullptr; }
    c->swin_W = 0; c->swin_i = 0; c->swin_w = w; c->swin_h = h;
    // pipeline window (rc_process_frame / rc_process_frames): lives in the flow ring
    c->win_W = (!c->configured || (c->prm.w == w && c->prm.h == h)) ? W : 0;
    if (c->configured) {
        int rc = ensure_flow_ring(c); if (rc) return rc;
        CUDA_TRY(c, cudaMemsetAsync(c->d_avg, 0, sizeof(float) * 2 * (size_t)c->prm.w * c->prm.h, c->stream));
        c->win_start = c->pairs_done;      // re-armed: the ring's older flows were never added to the new mean
    }
    if (W == 0) return RC_OK;
    // stand-alone window for caller-provided flows (rc_window_update)
    const size_t nb = sizeof(float) * 2 * (size_t)w * h;
    if (cudaMalloc((void**)&c->d_swin_ring, nb * W) != cudaSuccess || cudaMalloc((void**)&c->d_swin_avg, nb) != cudaSuccess) {
        cudaGetLastError();
        if (c->d_swin_ring) cudaFree(c->d_swin_ring);
        c->d_swin_ring = c->d_swin_avg = nullptr;
        return fail(c, RC_ERR_NOMEM, "cudaMalloc failed (window ring)%s");
    }
    CUDA_TRY(c, cudaMemsetAsync(c->d_swin_ring, 0, nb * W, c->stream));
    CUDA_TRY(c, cudaMemsetAsync(c->d_swin_avg, 0, nb, c->stream));
    c->swin_W = W;
    return RC_OK;
}

int rc_window_update(rc_ctx* c, const float* flow, size_t flow_step)
{
    if (!c) return RC_ERR_INVALID;
    if (c->swin_W <= 0) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    cudaSetDevice(c->device);
    const float* d_flow; size_t d_step;
    int rc = stage_flow_in(c, flow, flow_step, c->swin_w, c->swin_h, &d_flow, &d_step); if (rc) return rc;
    if (!flow && (c->prm.w != c->swin_w || c->prm.h != c->swin_h)) return fail(c, RC_ERR_INVALID, "window / flow size mismatch%s");
    const size_t n = (size_t)c->swin_w * c->swin_h;
    rc_launch_window_update(c, d_flow, d_step, c->swin_w, c->swin_h, c->d_swin_ring + (size_t)c->swin_i * n * 2,
                            c->d_swin_avg, c->swin_W);
    CHECK_LAUNCH(c);
    c->swin_i = (c->swin_i + 1) % c->swin_W;
    c->win_W = 0;            // the caller drives the window explicitly: rc_process_frame(s) leaves it alone
    if (flow && !is_device_ptr(flow)) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// The mean maintained by rc_process_frame(s) when the pipeline window is active, otherwise the stand-alone mean
// driven by rc_window_update.
static float* active_avg(rc_ctx* c, int* w, int* h)
{
    if (c->win_W > 0) { *w = c->prm.w; *h = c->prm.h; return c->d_avg; }
    if (c->swin_W > 0) { *w = c->swin_w; *h = c->swin_h; return c->d_swin_avg; }
    return nullptr;
}

int rc_window_get(rc_ctx* c, float* avg, size_t avg_step)
{
    if (!c || !avg) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int w, h;
    float* src = active_avg(c, &w, &h);
    if (!src) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    if (avg_step < (size_t)w * 8) return fail(c, RC_ERR_INVALID, "avg_step too small%s");
    const bool dev = is_device_ptr(avg);
    CUDA_TRY(c, cudaMemcpy2DAsync(avg, avg_step, src, (size_t)w * 8, (size_t)w * 8, h,
                                  dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    if (!dev) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_window_device(rc_ctx* c, float** dev_avg)
{
    if (!c || !dev_avg) return RC_ERR_INVALID;
    int w, h;
    float* src = active_avg(c, &w, &h);
    if (!src) return fail(c, RC_ERR_STATE, "rc_window_configure has not been called%s");
    *dev_avg = src;
    return RC_OK;
}

int rc_average_vector(rc_ctx* c, const float* old_slot, const float* flow, size_t flow_step, int w, int h, float* average,
                      float* new_slot, int frames, float dt, float upper)
{
    if (!c || !flow || !average || w < 1 || h < 1 || frames < 1 || flow_step < (size_t)w * 8) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const size_t nb = (size_t)w * h * 8;
    const bool dev = is_device_ptr(average);
    if ((old_slot && is_device_ptr(old_slot) != dev) || (new_slot && is_device_ptr(new_slot) != dev) || is_device_ptr(flow) != dev)
        return fail(c, RC_ERR_INVALID, "old_slot / flow / average / new_slot must all be host or all be device pointers%s");
    const float *d_old = old_slot, *d_flow = flow; float *d_avg = average, *d_new = new_slot; size_t d_step = flow_step;
    if (!dev) {
        int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 4 * nb); if (rc) return rc;
        char* base = reinterpret_cast<char*>(c->d_tmp2);
        float* df = reinterpret_cast<float*>(base); d_avg = reinterpret_cast<float*>(base + nb);
        CUDA_TRY(c, cudaMemcpy2DAsync(df, (size_t)w * 8, flow, flow_step, (size_t)w * 8, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(d_avg, average, nb, cudaMemcpyHostToDevice, c->stream));
        d_flow = df; d_step = (size_t)w * 8;
        if (old_slot) {
            float* dold = reinterpret_cast<float*>(base + 2 * nb);
            CUDA_TRY(c, cudaMemcpyAsync(dold, old_slot, nb, cudaMemcpyHostToDevice, c->stream));
            d_old = dold;
        }
        if (new_slot) d_new = reinterpret_cast<float*>(base + 3 * nb);
    }
    rc_launch_average_vector(c, d_flow, d_step, w, h, d_old, d_avg, d_new, frames, dt, upper);
    CHECK_LAUNCH(c);
    if (!dev) {
        CUDA_TRY(c, cudaMemcpyAsync(average, d_avg, nb, cudaMemcpyDeviceToHost, c->stream));
        if (new_slot) CUDA_TRY(c, cudaMemcpyAsync(new_slot, d_new, nb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
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
        size_t need = n * 8 + n * 4 + n * 8;        // seeds | dist | home
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

// ---- two-phase aggregation for a stream sharded by frame pair ---------------------------
int rc_batch_hist(rc_ctx* c, int nb, int64_t* deltas)
{
    if (!c || !deltas) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    if (nb < 1 || nb > c->B || nb > c->pairs_done) return fail(c, RC_ERR_INVALID, "nb must be in [1, min(max_batch, flows produced)]%s");
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    float* f[RC_MAX_BATCH];
    for (int j = 0; j < nb; j++) f[j] = ring_slot(c, c->pairs_done - nb + j);
    rc_launch_hist_of_flows(c, f, nb, c->prm.w, c->prm.h, c->d_hist_delta);
    const size_t n = (size_t)nb * RC_HIST_CELLS;
    const bool dev = is_device_ptr(deltas);
    long long* d_out = reinterpret_cast<long long*>(deltas);
    if (!dev) { rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, n * 8); if (rc) return rc; d_out = reinterpret_cast<long long*>(c->d_tmp2); }
    rc_launch_widen_counts(c, c->d_hist_delta, d_out, n);
    CHECK_LAUNCH(c);
    if (!dev) {
        CUDA_TRY(c, cudaMemcpyAsync(deltas, d_out, n * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return RC_OK;
}

int rc_aggregate_last(rc_ctx* c, int nb, int framecount0, rc_frame_result* results)
{
    if (!c) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    if (nb < 1 || nb > c->B || nb > c->pairs_done) return fail(c, RC_ERR_INVALID, "nb must be in [1, min(max_batch, flows produced)]%s");
    cudaSetDevice(c->device);
    const int w = c->prm.w, h = c->prm.h;
    int rc = ensure_aggregate(c); if (rc) return rc;
    rc = ensure_accumulator(c, w, h); if (rc) return rc;
    // uses the per-frame counts left in d_hist_delta by rc_batch_hist(nb)
    rc_launch_thresholds_batch(c, c->d_hist2d, c->d_hist_delta, nb, c->d_thr_batch[0], c->d_thr);
    ClassifyBatch cb;
    cb.nb = nb;
    for (int j = 0; j < nb; j++) { cb.flow[j] = ring_slot(c, c->pairs_done - nb + j); cb.old[j] = nullptr; }
    rc_launch_classify_batch(c, cb, w, h, c->d_thr_batch[0], framecount0, c->d_acc, nullptr, nullptr, 0);
    CHECK_LAUNCH(c);
    if (results) {
        CUDA_TRY(c, cudaMemcpyAsync(c->h_thr[0], c->d_thr_batch[0], sizeof(float) * RC_THR_FLOATS * nb, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        fill_results(c->h_thr[0], nb, 0, nb, results);
    }
    return RC_OK;
}

int rc_accumulator_mask(rc_ctx* c, int framecount, uint8_t* outmask)
{
    if (!c || !outmask) return RC_ERR_INVALID;
    if (!c->d_acc) return fail(c, RC_ERR_STATE, "no accumulator yet%s");
    cudaSetDevice(c->device);
    const size_t n = (size_t)c->acc_w * c->acc_h;
    const bool dev = is_device_ptr(outmask);
    uint8_t* d = dev ? outmask : c->d_cls;
    rc_launch_acc_mask(c, c->d_acc, n, framecount, d);
    CHECK_LAUNCH(c);
    if (!dev) {
        CUDA_TRY(c, cudaMemcpyAsync(outmask, d, n, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return RC_OK;
}

// ---- reference intermediate formats (compat.cu) ------------------------------------------------------------------
// stages `bytes` of a host image into (*scratch), or passes a device pointer through
static int stage3(rc_ctx* c, const void* p, size_t step, size_t row_bytes, int h, void** scratch, size_t* cap, void** d,
                  size_t* d_step)
{
    if (is_device_ptr(p)) { *d = const_cast<void*>(p); *d_step = step; return RC_OK; }
    int rc = ensure(c, scratch, cap, row_bytes * h); if (rc) return rc;
    CUDA_TRY(c, cudaMemcpy2DAsync(*scratch, row_bytes, p, step, row_bytes, h, cudaMemcpyHostToDevice, c->stream));
    *d = *scratch; *d_step = row_bytes;
    return RC_OK;
}

int rc_hist_from_polar(rc_ctx* c, const float* polar3, size_t step, int w, int h)
{
    if (!c || !polar3 || w < 1 || h < 1 || step < (size_t)w * 12) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_aggregate(c); if (rc) return rc;
    void* d; size_t ds;
    rc = stage3(c, polar3, step, (size_t)w * 12, h, &c->d_tmp, &c->d_tmp_cap, &d, &ds); if (rc) return rc;
    rc_launch_hist_polar(c, reinterpret_cast<const float*>(d), ds, w, h, c->d_hist2d);
    CHECK_LAUNCH(c);
    if (!is_device_ptr(polar3)) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_create_flow(rc_ctx* c, float* cur, size_t cstep, float* wc, size_t wstep, float* acc2, size_t astep, int w, int h,
                   float UPPER, float MID, float LOWER, const float* UPPER2d)
{
    if (!c || !cur || !wc || !acc2 || !UPPER2d || w < 1 || h < 1) return RC_ERR_INVALID;
    if (cstep < (size_t)w * 12 || wstep < (size_t)w * 12 || astep < (size_t)w * 12) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const bool dev = is_device_ptr(cur);
    if (dev != is_device_ptr(wc) || dev != is_device_ptr(acc2))
        return fail(c, RC_ERR_INVALID, "current / waterclass / accumulator2 must all be host or all be device%s");
    const size_t rb = (size_t)w * 12, img = rb * h;
    int rc = ensure_aggregate(c); if (rc) return rc;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_thr + 1, UPPER2d, sizeof(float) * RC_HIST_DIRECTIONS,
                                is_device_ptr(UPPER2d) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    float *dc = cur, *dw = wc, *da = acc2; size_t sc = cstep, sw = wstep, sa = astep;
    if (!dev) {
        rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 3 * img); if (rc) return rc;
        char* base = reinterpret_cast<char*>(c->d_tmp2);
        dc = reinterpret_cast<float*>(base); dw = reinterpret_cast<float*>(base + img); da = reinterpret_cast<float*>(base + 2 * img);
        CUDA_TRY(c, cudaMemcpy2DAsync(dc, rb, cur, cstep, rb, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(dw, rb, wc, wstep, rb, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(da, rb, acc2, astep, rb, h, cudaMemcpyHostToDevice, c->stream));
        sc = sw = sa = rb;
    }
    rc_launch_create_flow(c, dc, sc, dw, sw, da, sa, w, h, UPPER, MID, LOWER, c->d_thr + 1);
    CHECK_LAUNCH(c);
    if (!dev) {
        CUDA_TRY(c, cudaMemcpy2DAsync(cur, cstep, dc, rb, rb, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(wc, wstep, dw, rb, rb, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(acc2, astep, da, rb, rb, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return RC_OK;
}

int rc_create_accumulationbuffer(rc_ctx* c, float* acc, size_t astep, const float* acc2, size_t a2step, float* out,
                                 size_t ostep, uint8_t* mask, size_t mstep, int w, int h, int framecount)
{
    if (!c || !acc || !acc2 || !out || !mask || w < 1 || h < 1) return RC_ERR_INVALID;
    if (astep < (size_t)w * 12 || a2step < (size_t)w * 12 || ostep < (size_t)w * 12 || mstep < (size_t)w) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const bool dev = is_device_ptr(acc);
    if (dev != is_device_ptr(acc2) || dev != is_device_ptr(out) || dev != is_device_ptr(mask))
        return fail(c, RC_ERR_INVALID, "accumulator / accumulator2 / out / outmask must all be host or all be device%s");
    const size_t rb = (size_t)w * 12, img = rb * h;
    float *da = acc, *do_ = out; const float* da2 = acc2; uint8_t* dm = mask;
    size_t sa = astep, sa2 = a2step, so = ostep, sm = mstep;
    if (!dev) {
        int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 3 * img + (size_t)w * h + 64); if (rc) return rc;
        char* base = reinterpret_cast<char*>(c->d_tmp2);
        da = reinterpret_cast<float*>(base); float* d2 = reinterpret_cast<float*>(base + img);
        do_ = reinterpret_cast<float*>(base + 2 * img); dm = reinterpret_cast<uint8_t*>(base + 3 * img);
        CUDA_TRY(c, cudaMemcpy2DAsync(da, rb, acc, astep, rb, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(d2, rb, acc2, a2step, rb, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(do_, rb, out, ostep, rb, h, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(dm, w, mask, mstep, w, h, cudaMemcpyHostToDevice, c->stream));
        da2 = d2; sa = sa2 = so = rb; sm = w;
    }
    rc_launch_accumulate(c, da, sa, da2, sa2, do_, so, dm, sm, w, h, framecount);
    CHECK_LAUNCH(c);
    if (!dev) {
        CUDA_TRY(c, cudaMemcpy2DAsync(acc, astep, da, rb, rb, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(out, ostep, do_, rb, rb, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaMemcpy2DAsync(mask, mstep, dm, w, w, h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return RC_OK;
}

int rc_mask_edges(rc_ctx* c, const uint8_t* masks, size_t mask_step, size_t mask_stride, int w, int h, int count,
                  uint8_t* edges, size_t edges_step, size_t edges_stride)
{
    if (!c || !masks || !edges || w < 1 || h < 1 || count < 1 || mask_step < (size_t)w || edges_step < (size_t)w)
        return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const size_t n = (size_t)w * h;
    const bool hin = !is_device_ptr(masks), hout = !is_device_ptr(edges);
    const uint8_t* d_in = masks; size_t istep = mask_step, istride = mask_stride;
    uint8_t* d_out = edges; size_t ostep = edges_step, ostride = edges_stride;
    if (hin) {
        int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, n * count); if (rc) return rc;
        for (int j = 0; j < count; j++)
            CUDA_TRY(c, cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(c->d_tmp) + (size_t)j * n, w, masks + (size_t)j * mask_stride,
                                          mask_step, w, h, cudaMemcpyHostToDevice, c->stream));
        d_in = reinterpret_cast<const uint8_t*>(c->d_tmp); istep = w; istride = n;
    }
    if (hout) {
        int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, n * count); if (rc) return rc;
        d_out = reinterpret_cast<uint8_t*>(c->d_tmp2); ostep = w; ostride = n;
    } else if (d_out == d_in) {
        return fail(c, RC_ERR_INVALID, "in-place edges need host pointers%s");
    }
    rc_launch_edges(c, d_in, istep, istride, w, h, d_out, ostep, ostride, count);
    CHECK_LAUNCH(c);
    if (hout)
        for (int j = 0; j < count; j++)
            CUDA_TRY(c, cudaMemcpy2DAsync(edges + (size_t)j * edges_stride, edges_step, d_out + (size_t)j * n, w, w, h,
                                          cudaMemcpyDeviceToHost, c->stream));
    if (hin || hout) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// ---- derived particle fields ---------------------------------------------------------------
int rc_particle_fields(rc_ctx* c, const float* field, const float* dist, int w, int h, int flags, float* streamfield,
                       uint8_t* disp_bgr, uint8_t* motion_bgr, uint8_t* ratio_bgr, float* density, double* maxes)
{
    if (!c || !field || w < 1 || h < 1) return RC_ERR_INVALID;
    if (!dist && (motion_bgr || ratio_bgr)) return fail(c, RC_ERR_INVALID, "motion/ratio outputs need the distance field%s");
    const size_t n = (size_t)w * h;
    uint8_t* gray[3] = {nullptr, nullptr, nullptr};
    uint8_t* bgr[3] = {disp_bgr, motion_bgr, ratio_bgr};
    const int want = 1 | (dist ? 6 : 0);
    double m[3];
    int rc = fields_impl(c, field, nullptr, dist, n, flags, want, streamfield, nullptr, gray, bgr, maxes ? m : nullptr);
    if (rc) return rc;
    if (maxes) { maxes[0] = m[0]; maxes[1] = dist ? m[1] : 0.0; maxes[2] = dist ? m[2] : 0.0; }
    if (density) return rc_streamline_positions(c, field, w, h, density, flags);
    return RC_OK;
}

int rc_normalize_jet(rc_ctx* c, const float* src, size_t n, uint8_t* gray, uint8_t* bgr, double* maxval)
{
    if (!c || !src || n < 1) return RC_ERR_INVALID;
    uint8_t* g[3] = {gray, nullptr, nullptr}; uint8_t* b[3] = {bgr, nullptr, nullptr};
    double m[3];
    int rc = fields_impl(c, nullptr, src, nullptr, n, 0, 1, nullptr, nullptr, g, b, maxval ? m : nullptr);
    if (!rc && maxval) *maxval = m[0];
    return rc;
}

int rc_ratio_jet(rc_ctx* c, const float* a, const float* b, size_t n, int flags, float* ratio, uint8_t* gray, uint8_t* bgr,
                 double* maxval)
{
    if (!c || !a || !b || n < 1) return RC_ERR_INVALID;
    uint8_t* g[3] = {nullptr, nullptr, gray}; uint8_t* o[3] = {nullptr, nullptr, bgr};
    double m[3];
    int rc = fields_impl(c, nullptr, a, b, n, flags, 4, nullptr, ratio, g, o, maxval ? m : nullptr);
    if (!rc && maxval) *maxval = m[2];
    return rc;
}

int rc_field_magnitude(rc_ctx* c, const float* field, size_t n, float* mag)
{
    if (!c || !field || !mag || n < 1) return RC_ERR_INVALID;
    uint8_t* z[3] = {nullptr, nullptr, nullptr};
    return fields_impl(c, field, nullptr, nullptr, n, 0, 0, mag, nullptr, z, z, nullptr);
}

int rc_streamline_positions(rc_ctx* c, const float* field, int w, int h, float* density, int flags)
{
    if (!c || !field || !density || w < 1 || h < 1) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const size_t n = (size_t)w * h;
    int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, 20 * n + 512); if (rc) return rc;
    Stager st{c, reinterpret_cast<char*>(c->d_tmp2)};
    const bool keep = (flags & RC_FIELDS_KEEP_DENSITY) != 0;
    field = st.in(field, n * 8);
    density = st.out(density, n * 12, keep);
    rc_launch_positions(c, field, w, h, density, keep ? 0 : 1);
    CHECK_LAUNCH(c);
    return st.finish();
}

// ---- flow-derived diagnostics -------------------------------------------------------------
static int ensure_diag(rc_ctx* c)
{
    if (c->d_diag) return RC_OK;
    const size_t bytes = 64 + 148 * 4 * sizeof(double);
    if (cudaMalloc((void**)&c->d_diag, bytes) != cudaSuccess) { cudaGetLastError(); return fail(c, RC_ERR_NOMEM, "cudaMalloc failed%s"); }
    CUDA_TRY(c, cudaMemsetAsync(c->d_diag, 0, bytes, c->stream));
    return RC_OK;
}

int rc_subtract_mean_magnitude(rc_ctx* c, float* flow, size_t flow_step, int w, int h, int flags, float* meanval)
{
    if (!c || !flow || w < 1 || h < 1 || flow_step < (size_t)w * 8) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_diag(c); if (rc) return rc;
    const bool dev = is_device_ptr(flow);
    float* d_flow = flow; size_t d_step = flow_step;
    if (!dev) {
        rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, (size_t)w * h * 8); if (rc) return rc;
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d_tmp, (size_t)w * 8, flow, flow_step, (size_t)w * 8, h, cudaMemcpyHostToDevice, c->stream));
        d_flow = reinterpret_cast<float*>(c->d_tmp); d_step = (size_t)w * 8;
    }
    rc_launch_sub_mean_magnitude(c, d_flow, d_step, w, h, (flags & RC_DIAG_SEQUENTIAL_SUM) ? 1 : 0,
                                 reinterpret_cast<double*>(reinterpret_cast<char*>(c->d_diag) + 64), c->d_diag + 4);
    CHECK_LAUNCH(c);
    if (!dev) CUDA_TRY(c, cudaMemcpy2DAsync(flow, flow_step, d_flow, d_step, (size_t)w * 8, h, cudaMemcpyDeviceToHost, c->stream));
    if (meanval) CUDA_TRY(c, cudaMemcpyAsync(meanval, c->d_diag + 4, 4, cudaMemcpyDeviceToHost, c->stream));
    if (meanval || !dev) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

// shared body of the two colourisations: which = 0 vectorToColor, 1 shearRateToColor
static int color_impl(rc_ctx* c, int which, const float* flow, size_t flow_step, int w, int h, uint8_t* img, size_t img_step,
                      float* max_io, int flags)
{
    if (!c || !img) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    int rc = ensure_diag(c); if (rc) return rc;
    const float* d_flow; size_t d_step;
    rc = stage_flow_in(c, flow, flow_step, w, h, &d_flow, &d_step); if (rc) return rc;
    if (!flow) { w = c->prm.w; h = c->prm.h; }
    if (img_step < (size_t)w * 3) return fail(c, RC_ERR_INVALID, "bad image step%s");
    const bool himg = !is_device_ptr(img);
    uint8_t* d_img = img; size_t d_istep = img_step;
    if (himg) {
        rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, (size_t)w * h * 3); if (rc) return rc;
        d_img = reinterpret_cast<uint8_t*>(c->d_tmp2); d_istep = (size_t)w * 3;
        if (which == 1)       // in/out: the border keeps the caller's pixels
            CUDA_TRY(c, cudaMemcpy2DAsync(d_img, d_istep, img, img_step, (size_t)w * 3, h, cudaMemcpyHostToDevice, c->stream));
    }
    float* prev = c->d_diag + 2 * which;
    unsigned* next = reinterpret_cast<unsigned*>(c->d_diag + 2 * which + 1);
    if (max_io) CUDA_TRY(c, cudaMemcpyAsync(prev, max_io, 4, cudaMemcpyHostToDevice, c->stream));
    const int fma = (flags & RC_DIAG_HSV_NOFMA) ? 0 : 1;
    if (which == 0) rc_launch_vector_color(c, d_flow, d_step, w, h, d_img, d_istep, prev, next, fma);
    else rc_launch_shear_color(c, d_flow, d_step, w, h, d_img, d_istep, prev, next, fma);
    CHECK_LAUNCH(c);
    CUDA_TRY(c, cudaMemcpyAsync(prev, next, 4, cudaMemcpyDeviceToDevice, c->stream));     // `static max = max_new;`
    if (himg) CUDA_TRY(c, cudaMemcpy2DAsync(img, img_step, d_img, d_istep, (size_t)w * 3, h, cudaMemcpyDeviceToHost, c->stream));
    if (max_io) CUDA_TRY(c, cudaMemcpyAsync(max_io, prev, 4, cudaMemcpyDeviceToHost, c->stream));
    if (himg || max_io) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_vector_to_color(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, uint8_t* bgr, size_t bgr_step,
                       float* max_displacement, int flags)
{
    return color_impl(c, 0, flow, flow_step, w, h, bgr, bgr_step, max_displacement, flags);
}

int rc_shear_rate_to_color(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, uint8_t* img, size_t img_step,
                           float* max_frobenius, int flags)
{
    return color_impl(c, 1, flow, flow_step, w, h, img, img_step, max_frobenius, flags);
}

// ---- fused per-frame / per-batch step -------------------------------------------------------------------------
static int submit_impl(rc_ctx* c, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                       uint8_t* outmasks, size_t mask_stride, rc_frame_result* results, int src_w, int src_h,
                       int ingest_flags)
{
    const bool is_bgr = src_w > 0;
    if (!c) return RC_ERR_INVALID;
    if (!c->configured) return fail(c, RC_ERR_STATE, "rc_flow_configure has not been called%s");
    if (count < 1 || count > c->B) return fail(c, RC_ERR_INVALID, "count must be in [1, max_batch]%s");
    cudaSetDevice(c->device);
    const int w = c->prm.w, h = c->prm.h;
    const size_t n = (size_t)w * h;
    if (!is_bgr && (!frames || step < (size_t)w || (count > 1 && frame_stride < step * (size_t)(h - 1) + w)))
        return fail(c, RC_ERR_INVALID, "bad frame pointer / step / stride%s");
    if (is_bgr && (!frames || src_h < 1 || step < (size_t)src_w * 3 ||
                   (count > 1 && frame_stride < step * (size_t)(src_h - 1) + (size_t)src_w * 3)))
        return fail(c, RC_ERR_INVALID, "bad BGR frame pointer / step / stride%s");
    const bool packed = c->mask_format == RC_MASK_PACKED;
    if (packed && outmasks && (n % 8 != 0 || w % 8 != 0))
        return fail(c, RC_ERR_UNSUPPORTED, "RC_MASK_PACKED needs an image width that is a multiple of 8%s");
    const size_t mb = packed ? n / 8 : n;          // bytes of one outmask
    if (outmasks && count > 1 && mask_stride < mb) return fail(c, RC_ERR_INVALID, "mask_stride too small%s");
    int rc = ensure_aggregate(c); if (rc) return rc;
    rc = ensure_accumulator(c, w, h); if (rc) return rc;
    const int slot = (int)(c->submitted & 1);
    rc = finish_slot(c, slot); if (rc) return rc;          // the batch that used this slot two submits ago

    const bool dev_in = is_device_ptr(frames);
    const uint8_t* d = frames; size_t ds = step, dfs = frame_stride;
    if (is_bgr) {
        const uint8_t* d_src = frames; size_t sstep = step, sfs = frame_stride;
        if (!dev_in) {
            const size_t row = (size_t)src_w * 3, per = row * src_h;
            rc = ensure(c, &c->d_bgr[slot], &c->d_bgr_cap[slot], per * count); if (rc) return rc;
            CUDA_TRY(c, cudaStreamWaitEvent(c->s_in, c->ev_compute[slot], 0));
            if (step == row && (count == 1 || frame_stride == per))
                CUDA_TRY(c, cudaMemcpyAsync(c->d_bgr[slot], frames, per * count, cudaMemcpyHostToDevice, c->s_in));
            else
                for (int j = 0; j < count; j++)
                    CUDA_TRY(c, cudaMemcpy2DAsync(reinterpret_cast<uint8_t*>(c->d_bgr[slot]) + (size_t)j * per, row,
                                                  frames + (size_t)j * frame_stride, step, row, src_h, cudaMemcpyHostToDevice, c->s_in));
            CUDA_TRY(c, cudaEventRecord(c->ev_in[slot], c->s_in));
            CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_in[slot], 0));
            d_src = reinterpret_cast<const uint8_t*>(c->d_bgr[slot]); sstep = row; sfs = per;
        }
        if ((ingest_flags & RC_INGEST_AREA) && (src_w < w || src_h < h))
            return fail(c, RC_ERR_UNSUPPORTED, "RC_INGEST_AREA is built for downscaling only%s");
        rc_launch_ingest_bgr(c, d_src, sstep, sfs, src_w, src_h, c->d_frames[slot], w, n, w, h, count,
                             (ingest_flags & RC_INGEST_GRAY14) ? 1 : 0, (ingest_flags & RC_INGEST_AREA) ? 1 : 0);
        d = c->d_frames[slot]; ds = w; dfs = n;
    } else if (!dev_in) {
        // H2D on the copy-in stream; it may only overwrite the staging slot once the kernels that read it are done
        CUDA_TRY(c, cudaStreamWaitEvent(c->s_in, c->ev_compute[slot], 0));
        if (step == (size_t)w && (count == 1 || frame_stride == n))
            CUDA_TRY(c, cudaMemcpyAsync(c->d_frames[slot], frames, n * count, cudaMemcpyHostToDevice, c->s_in));
        else
            for (int j = 0; j < count; j++)
                CUDA_TRY(c, cudaMemcpy2DAsync(c->d_frames[slot] + (size_t)j * n, w, frames + (size_t)j * frame_stride, step,
                                              w, h, cudaMemcpyHostToDevice, c->s_in));
        CUDA_TRY(c, cudaEventRecord(c->ev_in[slot], c->s_in));
        CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_in[slot], 0));
        d = c->d_frames[slot]; ds = w; dfs = n;
    }
    const bool dev_masks = outmasks && is_device_ptr(outmasks);
    // masks / thresholds of this batch go to staging slot `slot`; its previous D2H must have drained
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_out[slot], 0));
    const int first_produced = c->frames_seen == 0 ? 1 : 0;
    const bool direct = dev_masks && (count == 1 || mask_stride == mb);
    uint8_t* d_masks = outmasks ? (direct ? outmasks + (size_t)first_produced * mask_stride : c->d_masks[slot]) : nullptr;
    int produced = run_frames(c, d, ds, dfs, count, 1, framecount0, c->d_thr_batch[slot], d_masks);
    CHECK_LAUNCH(c);
    CUDA_TRY(c, cudaEventRecord(c->ev_compute[slot], c->stream));
    bool host_out = false;
    if ((outmasks || results) && produced > 0) {
        CUDA_TRY(c, cudaStreamWaitEvent(c->s_out, c->ev_compute[slot], 0));
        if (outmasks && !direct) {
            const cudaMemcpyKind kind = dev_masks ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
            if (count == 1 || mask_stride == mb)
                CUDA_TRY(c, cudaMemcpyAsync(outmasks + (size_t)first_produced * mask_stride, d_masks, mb * produced, kind, c->s_out));
            else
                for (int j = 0; j < produced; j++)
                    CUDA_TRY(c, cudaMemcpyAsync(outmasks + (size_t)(first_produced + j) * mask_stride, d_masks + (size_t)j * mb, mb,
                                                kind, c->s_out));
            if (!dev_masks) host_out = true;
        }
        if (results) {
            CUDA_TRY(c, cudaMemcpyAsync(c->h_thr[slot], c->d_thr_batch[slot], sizeof(float) * RC_THR_FLOATS * produced,
                                        cudaMemcpyDeviceToHost, c->s_out));
            host_out = true;
        }
    }
    CUDA_TRY(c, cudaEventRecord(c->ev_out[slot], c->s_out));
    if (results || host_out) {
        c->pending_results[slot] = results; c->pending_count[slot] = count; c->pending_first_produced[slot] = first_produced;
    }
    c->submitted++;
    return produced;
}

int rc_submit_frames(rc_ctx* c, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                     uint8_t* outmasks, size_t mask_stride, rc_frame_result* results)
{
    return submit_impl(c, frames, step, frame_stride, count, framecount0, outmasks, mask_stride, results, 0, 0, 0);
}

int rc_submit_frames_bgr(rc_ctx* c, const uint8_t* bgr_frames, size_t step, size_t frame_stride, int src_w, int src_h,
                         int count, int framecount0, int ingest_flags, uint8_t* outmasks, size_t mask_stride,
                         rc_frame_result* results)
{
    if (src_w < 1 || src_h < 1) return RC_ERR_INVALID;
    return submit_impl(c, bgr_frames, step, frame_stride, count, framecount0, outmasks, mask_stride, results, src_w, src_h,
                       ingest_flags);
}

int rc_ingest_bgr(rc_ctx* c, const uint8_t* bgr, size_t step, int src_w, int src_h, uint8_t* gray, size_t gray_step,
                  int dst_w, int dst_h, int flags)
{
    if (!c || !bgr || !gray || src_w < 1 || src_h < 1 || dst_w < 1 || dst_h < 1) return RC_ERR_INVALID;
    if (step < (size_t)src_w * 3 || gray_step < (size_t)dst_w) return fail(c, RC_ERR_INVALID, "bad step%s");
    cudaSetDevice(c->device);
    const bool hin = !is_device_ptr(bgr), hout = !is_device_ptr(gray);
    const uint8_t* d_in = bgr; size_t d_step = step;
    uint8_t* d_out = gray; size_t o_step = gray_step;
    if (hin) {
        const size_t row = (size_t)src_w * 3;
        int rc = ensure(c, &c->d_tmp, &c->d_tmp_cap, row * src_h); if (rc) return rc;
        CUDA_TRY(c, cudaMemcpy2DAsync(c->d_tmp, row, bgr, step, row, src_h, cudaMemcpyHostToDevice, c->stream));
        d_in = reinterpret_cast<const uint8_t*>(c->d_tmp); d_step = row;
    }
    if (hout) {
        int rc = ensure(c, &c->d_tmp2, &c->d_tmp2_cap, (size_t)dst_w * dst_h); if (rc) return rc;
        d_out = reinterpret_cast<uint8_t*>(c->d_tmp2); o_step = dst_w;
    }
    if ((flags & RC_INGEST_AREA) && (src_w < dst_w || src_h < dst_h))
        return fail(c, RC_ERR_UNSUPPORTED, "RC_INGEST_AREA is built for downscaling only%s");
    rc_launch_ingest_bgr(c, d_in, d_step, 0, src_w, src_h, d_out, o_step, 0, dst_w, dst_h, 1, (flags & RC_INGEST_GRAY14) ? 1 : 0,
                         (flags & RC_INGEST_AREA) ? 1 : 0);
    CHECK_LAUNCH(c);
    if (hout) CUDA_TRY(c, cudaMemcpy2DAsync(gray, gray_step, d_out, o_step, dst_w, dst_h, cudaMemcpyDeviceToHost, c->stream));
    if (hin || hout) CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return RC_OK;
}

int rc_set_mask_format(rc_ctx* c, int format)
{
    if (!c || (format != RC_MASK_U8 && format != RC_MASK_PACKED)) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    sync_all(c);
    c->mask_format = format;
    return RC_OK;
}

int rc_wait(rc_ctx* c)
{
    if (!c) return RC_ERR_INVALID;
    cudaSetDevice(c->device);
    const int s0 = (int)(c->submitted & 1);      // oldest first
    int rc = finish_slot(c, s0); if (rc) return rc;
    rc = finish_slot(c, s0 ^ 1); if (rc) return rc;
    CUDA_TRY(c, cudaStreamSynchronize(c->s_in));
    CUDA_TRY(c, cudaStreamSynchronize(c->s_out));
    // an overlapped accumulator all-reduce (rc_allreduce_accumulators, out of place) joins the context's stream here
    if (c->ev_ar_done) CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->ev_ar_done, 0));
    return RC_OK;
}

int rc_process_frames(rc_ctx* c, const uint8_t* frames, size_t step, size_t frame_stride, int count, int framecount0,
                      uint8_t* outmasks, size_t mask_stride, rc_frame_result* results)
{
    int produced = rc_submit_frames(c, frames, step, frame_stride, count, framecount0, outmasks, mask_stride, results);
    if (produced < 0) return produced;
    const bool host_touch = !is_device_ptr(frames) || results || (outmasks && !is_device_ptr(outmasks));
    if (host_touch) { int rc = rc_wait(c); if (rc) return rc; }
    return produced;
}

int rc_process_frame(rc_ctx* c, const uint8_t* frame, size_t step, int framecount, uint8_t* outmask,
                     rc_frame_result* result)
{
    return rc_process_frames(c, frame, step, 0, 1, framecount, outmask, 0, result);
}

}  // extern "C"
