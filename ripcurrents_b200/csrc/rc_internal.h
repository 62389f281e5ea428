// Internal declarations shared by the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/ripcurrents_b200.h"

#define RC_MAX_LAYERS 12
#define RC_MAX_POLY_N 32
#define RC_MAX_SMOOTH_TAPS 255
#define RC_MAX_WIN_HALF 64
#define RC_MAX_BATCH 64
#define RC_THR_FLOATS 80       // per-frame thresholds record: [0] UPPER, [1..36] UPPER2d, [37..72] prop, [74..75] int64 histsum
#define RC_HIST_CELLS (RC_HIST_ROWS * RC_HIST_BINS)

// kernel classes for rc_profile_* (per-launch CUDA-event timing) -- order matches rc_kernel_names[]
enum RcKernelId { K_RESERVED = 0, K_PYR_V, K_POLYEXP, K_UPDATE_MATRICES, K_FLOW_ITER_FUSED, K_FLOW_ITER_FINAL,
                  K_FLOW_LAYER, K_POLAR_HIST, K_THRESHOLDS, K_CLASSIFY, K_WINDOW, K_ADVECT, K_STREAKLINE, K_MISC,
                  K_FIELDS, K_DIAG, K_COUNT };
extern const char* const rc_kernel_names[K_COUNT];

struct PolyCoef {          // polynomial-expansion kernels; entries beyond n are zero
    float g[RC_MAX_POLY_N + 1], xg[RC_MAX_POLY_N + 1], xxg[RC_MAX_POLY_N + 1];
    double ig11, ig03, ig33, ig55;
    int n;        // poly_n
    int n_eff;    // taps evaluated by the fast kernel (== n in strict mode)
};

struct SmoothCoef {        // presmooth kernel of one pyramid layer (A.2)
    float k[RC_MAX_SMOOTH_TAPS];
    int ksize;
};

struct WinCoef {           // updateFlow window: box (all ones, post-scale 1/winsize^2) or Gaussian (A.7, post-scale 1)
    float k[RC_MAX_WIN_HALF + 1];
    int m;
    float post_scale;      // fp32 post scale used by the fast kernels
    double post_scale_d;   // fp64 post scale used by the strict kernels
    int gaussian;
};

struct FarnebackParams {
    int w = 0, h = 0;
    double pyr_scale = 0.5;
    int levels = 0, winsize = 0, iterations = 0, poly_n = 0;
    double poly_sigma = 0;
    int flags = 0;
    int max_batch = 1;
    bool operator==(const FarnebackParams& o) const {
        return w == o.w && h == o.h && pyr_scale == o.pyr_scale && levels == o.levels && winsize == o.winsize &&
               iterations == o.iterations && poly_n == o.poly_n && poly_sigma == o.poly_sigma && flags == o.flags &&
               max_batch == o.max_batch;
    }
};

// One pyramid layer.  All arrays carry a leading batch / ring dimension.
//   I     [B]      presmoothed layer image of each new frame of the batch (row pitch `pitch`)
//   R     [B+1]    ring of polynomial expansions, "4+1" layout: float4 plane + float plane (plane = pitch*h pixels)
//   M     [B][2]   G/h matrices ping-pong, 5 planes each (only the unfused path uses them)
//   flow  [B]      dense w*h*2 (layers k >= 1; layer 0 writes into the context's flow ring)
struct Layer {
    int w = 0, h = 0, pitch = 0;
    size_t plane = 0;
    SmoothCoef smooth;
    float* I = nullptr;
    float* R = nullptr;
    float* M = nullptr;
    float* flow = nullptr;
    __host__ __device__ float* Rslot(int s) const { return R + (size_t)s * 5 * plane; }
};

struct ProfRec { int id; cudaEvent_t a, b; double bytes; };

struct rc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    std::string launch_err;        // first failed kernel launch since the last CHECK_LAUNCH (kernel class + CUDA error)
    int64_t launches = 0;
    int mask_format = 0;           // RC_MASK_U8 / RC_MASK_PACKED (rc_set_mask_format)

    bool prof_on = false;
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    double prof_ms[K_COUNT] = {0};
    double prof_bytes[K_COUNT] = {0};
    int64_t prof_n[K_COUNT] = {0};

    // Farneback state
    bool configured = false;
    FarnebackParams prm;
    int nlayers = 0;
    Layer layer[RC_MAX_LAYERS];
    PolyCoef poly;
    WinCoef win;
    bool strict = false;
    int B = 1;                     // max batch
    int r_base = 0;                // ring slot (in R) of the most recent expanded frame
    long long frames_seen = 0;
    int n_flows = 0;               // flows produced by the last push (0..B)
    std::vector<void*> allocs;

    // small-batch overlap: expansions of layers 1/0 on a second stream (rc_launch_expand / rc_launch_flows)
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_pyr = nullptr, ev_poly[2] = {nullptr, nullptr};
    bool expand_overlapped = false;

    // flow ring: layer-0 flows of the most recent frames; slot of frame-pair index p is p % ring_slots
    float* flow_ring = nullptr;
    int ring_slots = 0;
    long long pairs_done = 0;      // total flows produced since configure
    int win_W = 0;                 // sliding-window length (0 = off)
    long long win_start = 0;       // pair index at which the window was (re)armed: flows before it never entered the mean
    float* d_avg = nullptr;        // window mean (w*h*2)

    // staging for host frames (double-buffered) and outputs
    uint8_t* d_frames[2] = {nullptr, nullptr};
    void* d_bgr[2] = {nullptr, nullptr};           // staging for BGR camera frames (rc_submit_frames_bgr), lazily sized
    size_t d_bgr_cap[2] = {0, 0};
    uint8_t* d_masks[2] = {nullptr, nullptr};
    float* d_thr_batch[2] = {nullptr, nullptr};    // [B][RC_THR_FLOATS] per staging slot
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr};     // H2D of slot finished
    cudaEvent_t ev_compute[2] = {nullptr, nullptr};// kernels that read/wrote slot finished
    cudaEvent_t ev_out[2] = {nullptr, nullptr};    // D2H of slot finished
    long long submitted = 0;                       // batches submitted through the async path
    int slot_frames[2] = {0, 0};                   // frames staged in each slot (for rc_wait)
    float* h_thr[2] = {nullptr, nullptr};          // pinned host mirror of d_thr_batch
    rc_frame_result* pending_results[2] = {nullptr, nullptr};
    int pending_count[2] = {0, 0};
    int pending_first_produced[2] = {0, 0};

    // generic scratch for host-pointer arguments
    void* d_tmp = nullptr;  size_t d_tmp_cap = 0;
    void* d_tmp2 = nullptr; size_t d_tmp2_cap = 0;
    // diagnostics state (diag.cu): [0] vectorToColor max of the previous call, [1] its new max (bits), [2]/[3] the same
    // for shearRateToColor, [4] meanval; fp64 partial sums from byte 64 on.  The reference keeps these in function statics.
    float* d_diag = nullptr;

    // aggregation state
    unsigned long long* d_hist2d = nullptr;   // cumulative counters, RC_HIST_CELLS
    unsigned int* d_hist_delta = nullptr;     // [B][RC_HIST_CELLS] per-frame counts of the current batch
    float* d_thr = nullptr;                   // thresholds of the most recent frame (RC_THR_FLOATS)
    float* d_acc = nullptr;                   // accumulator.x
    int acc_w = 0, acc_h = 0;
    uint8_t* d_cls = nullptr;                 // waveclass / waterclass scratch (2 planes) for rc_classify_accumulate

    // multi-GPU (comm.cu): NCCL communicator + state of a stream sharded by frame pair
    void* nccl = nullptr;                     // ncclComm_t
    bool own_comm = false;
    int rank = 0, nranks = 1;
    unsigned int* d_gather = nullptr;         // [nranks][B][RC_HIST_CELLS] all-gathered per-frame counts
    unsigned long long* d_hist_global = nullptr;   // cumulative counts of the whole stream through the previous super-block
    float* d_acc_global = nullptr;            // all-reduced accumulator (reporting copy)
    int shard_owner = 0, shard_W = 0, shard_slots = 0;
    long long shard_pairs = 0;                // pairs of the whole stream processed so far
    float* d_shard_ring = nullptr;            // owner rank: flows of the stream in order, W + nranks*B slots
    float* d_shard_avg = nullptr;             // owner rank: the sliding-window mean
    float* d_shard_flows = nullptr;           // this rank's flows of the current / previous super-block: [2][B] full frames
    cudaStream_t s_comm = nullptr;            // collectives that overlap the compute stream (band exchange + window update,
                                              // out-of-place accumulator all-reduce); created on first use
    cudaEvent_t ev_ar_snap = nullptr, ev_ar_done = nullptr;
    cudaEvent_t ev_flows[2] = {nullptr, nullptr}, ev_comm[2] = {nullptr, nullptr};
    long long shard_steps = 0;
    bool shard_configured = false;

    // stand-alone window (rc_window_* with caller-provided flows)
    int swin_W = 0, swin_w = 0, swin_h = 0, swin_i = 0;
    float* d_swin_ring = nullptr;
    float* d_swin_avg = nullptr;
};

// RAII bracket around one kernel launch: counts it and, when profiling is on, times it with two CUDA events on
// the launching stream.  `bytes` = ALGORITHMIC bytes of this launch ("Kernels and their rooflines").
struct KScope {
    rc_ctx* c; int id; double bytes; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(rc_ctx* c)
    {
        cudaEvent_t e;
        if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    KScope(rc_ctx* c_, int id_, double bytes_, int nlaunch = 1) : c(c_), id(id_), bytes(bytes_)
    {
        c->launches += nlaunch;
        if (c->prof_on) { a = get(c); b = get(c); cudaEventRecord(a, c->stream); }
    }
    ~KScope()
    {
        // launch status: a bad configuration (grid, shared memory, registers) surfaces here, named by kernel class,
        // instead of at some later runtime call
        const cudaError_t e = cudaPeekAtLastError();
        if (e != cudaSuccess && c->launch_err.empty())
            c->launch_err = std::string(rc_kernel_names[id]) + ": " + cudaGetErrorString(e);
        if (a) { cudaEventRecord(b, c->stream); c->prof.push_back(ProfRec{id, a, b, bytes}); }
    }
};

// ---- farneback.cu ----------------------------------------------------------------------------------
void rc_farneback_init_device(int device);     // once per device: opt-in shared-memory sizes of the kernels
// Expands `nb` new frames (device, dense u8, frame stride `fstride`) into R ring slots first_slot.. (mod B+1).
void rc_launch_expand(rc_ctx* c, const uint8_t* d_frames, size_t step, size_t fstride, int nb, int first_slot,
                      bool overlap = false);
// Flows for `nb` consecutive pairs: pair j = (ring slot (prev_slot + j) % (B+1), next slot); layer-0 flow of pair j
// goes to flow_dst[j]; hist_delta (may be null) receives per-pair direction/speed counts [nb][RC_HIST_CELLS].
void rc_launch_flows(rc_ctx* c, int nb, int prev_slot, float* const* flow_dst_host, unsigned int* hist_delta);

void rc_launch_hist_of_flows(rc_ctx* c, float* const* flows, int nb, int w, int h, unsigned int* delta);

// ---- aggregate.cu ----------------------------------------------------------------------------------
void rc_launch_polar_hist(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, unsigned long long* hist2d);
void rc_launch_cart_to_polar(rc_ctx* c, const float* flow, size_t n, float* mag, float* ang);
// cumulative += delta[j] for j < nb, thresholds after each frame -> thr_batch[j]; last also copied to thr_last
void rc_launch_thresholds_batch(rc_ctx* c, unsigned long long* hist2d, const unsigned int* delta, int nb,
                                float* thr_batch, float* thr_last);
void rc_launch_classify(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, const float* thr,
                        int framecount, float* acc, uint8_t* mask, uint8_t* waveclass, uint8_t* waterclass);
// batched classify + accumulate + mask (+ sliding-window mean) over nb consecutive flows held in the flow ring
struct ClassifyBatch {
    const float* flow[RC_MAX_BATCH];      // flow of frame j
    const float* old[RC_MAX_BATCH];       // ring slot leaving the window at frame j (null while the window fills)
    int nb;
};
void rc_launch_classify_batch(rc_ctx* c, const ClassifyBatch& cb, int w, int h, const float* thr_batch, int framecount0,
                              float* acc, uint8_t* masks, float* avg, int W, int packed = 0);
void rc_launch_widen_counts(rc_ctx* c, const unsigned int* in, long long* out, size_t n);
void rc_launch_acc_mask(rc_ctx* c, const float* acc, size_t n, int framecount, uint8_t* mask);
void rc_launch_window_update(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* slot, float* avg,
                             int W);
void rc_launch_subtract_mean(rc_ctx* c, float* flow, size_t flow_step, int w, int h, double* d_sums);

// window mean alone over nb consecutive flows (the owner rank of a sharded stream)
void rc_launch_window_batch(rc_ctx* c, const ClassifyBatch& cb, int w, int h, float* avg, int W);
// sharded stream: start counters of this rank (global + frames of lower ranks) and the next global counters
void rc_launch_shard_prefix(rc_ctx* c, const unsigned int* gathered, int rank, int nranks, int B, unsigned long long* hist_global,
                            unsigned long long* hist_start);

// ---- comm.cu -----------------------------------------------------------------------------------------
int rc_comm_fence(rc_ctx* c, bool destroy);    // wait for the communication stream (and optionally destroy it)

// ---- api.cu helpers used by comm.cu ------------------------------------------------------------------
int rc_fail(rc_ctx* c, int code, const char* fmt, const char* detail);
bool rc_is_device_ptr(const void* p);
int rc_ensure_aggregate(rc_ctx* c);
int rc_ensure_accumulator(rc_ctx* c, int w, int h);
float* rc_ring_slot(rc_ctx* c, long long pair);
void rc_fill_results(const float* h_thr, int produced, int first_produced, int count, rc_frame_result* results);
int rc_run_frames_hist(rc_ctx* c, const uint8_t* d_frames, size_t step, size_t fstride, int count, float* const* dst_override);

// ---- compat.cu -------------------------------------------------------------------------------------
void rc_launch_hist_polar(rc_ctx* c, const float* polar, size_t step, int w, int h, unsigned long long* hist2d);
void rc_launch_create_flow(rc_ctx* c, float* cur, size_t cstep, float* wc, size_t wstep, float* acc2, size_t astep, int w,
                           int h, float UPPER, float MID, float LOWER, const float* d_upper2d);
void rc_launch_accumulate(rc_ctx* c, float* acc, size_t astep, const float* acc2, size_t a2step, float* out, size_t ostep,
                          unsigned char* mask, size_t mstep, int w, int h, int framecount);

void rc_launch_ingest_bgr(rc_ctx* c, const uint8_t* bgr, size_t step, size_t fstride, int sw, int sh, uint8_t* gray,
                          size_t gstep, size_t gstride, int dw, int dh, int nb, int legacy14, int area = 0);

void rc_launch_edges(rc_ctx* c, const uint8_t* mask, size_t step, size_t stride, int w, int h, uint8_t* out, size_t ostep,
                     size_t ostride, int nb);

// ---- fields.cu -------------------------------------------------------------------------------------
void rc_launch_fields(rc_ctx* c, const float* field, const float* src0, const float* dist, size_t n, int div0_zero, int want,
                      float* mag_out, float* ratio_out, uint8_t* const gray[3], uint8_t* const bgr[3], unsigned* d_maxenc,
                      double* d_max);
void rc_launch_positions(rc_ctx* c, const float* field, int w, int h, float* density, int zero_first);

// ---- diag.cu ---------------------------------------------------------------------------------------
void rc_launch_sub_mean_magnitude(rc_ctx* c, float* flow, size_t step, int w, int h, int sequential, double* d_partial,
                                  float* d_meanval);
void rc_launch_vector_color(rc_ctx* c, const float* flow, size_t step, int w, int h, uint8_t* bgr, size_t bstep,
                            const float* d_prev_max, unsigned* d_new_max, int fma);
void rc_launch_shear_color(rc_ctx* c, const float* flow, size_t step, int w, int h, uint8_t* img, size_t istep,
                           const float* d_prev_max, unsigned* d_new_max, int fma);

// ---- advect.cu -------------------------------------------------------------------------------------
void rc_launch_advect(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* seeds, size_t n, float dt,
                      int iterations, float upper, int variant, float* dist, const int32_t* home);
void rc_launch_average_vector(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* old_slot,
                              float* average, float* new_slot, int frames, float dt, float upper);
void rc_launch_streakline(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* emitters, int E,
                          float* vertices, int32_t* count, int cap, float dt);
