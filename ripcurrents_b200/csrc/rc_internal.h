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

// Planar fp32 image set: `planes` planes of h rows, row pitch `pitch` floats, plane stride `pstride` floats.
struct Planes {
    float* p = nullptr;
    int w = 0, h = 0, pitch = 0;
    size_t pstride = 0;
    __host__ __device__ float* plane(int c) const { return p + (size_t)c * pstride; }
};

struct PolyCoef {          // polynomial-expansion kernels
    float g[RC_MAX_POLY_N + 1], xg[RC_MAX_POLY_N + 1], xxg[RC_MAX_POLY_N + 1];
    double ig11, ig03, ig33, ig55;
    int n;        // poly_n
    int n_eff;    // taps actually evaluated (== n in strict mode)
};

struct SmoothCoef {        // presmooth kernel of one pyramid layer (A.2)
    float k[RC_MAX_SMOOTH_TAPS];
    int ksize;
};

struct GaussWin {          // updateFlow Gaussian window (A.7)
    float k[RC_MAX_WIN_HALF + 1];
    int m;
};

struct FarnebackParams {
    int w = 0, h = 0;
    double pyr_scale = 0.5;
    int levels = 0, winsize = 0, iterations = 0, poly_n = 0;
    double poly_sigma = 0;
    int flags = 0;
    bool operator==(const FarnebackParams& o) const {
        return w == o.w && h == o.h && pyr_scale == o.pyr_scale && levels == o.levels && winsize == o.winsize &&
               iterations == o.iterations && poly_n == o.poly_n && poly_sigma == o.poly_sigma && flags == o.flags;
    }
};

struct Layer {
    int w = 0, h = 0;
    SmoothCoef smooth;
    float* I = nullptr;            // w*h (pitch) fp32 presmoothed layer image of the frame being expanded
    float* htmp = nullptr;         // pass-1 scratch of the pyramid kernel
    Planes R[2];                   // polynomial expansion of the two cached frames (ping-pong)
    Planes M[2];                   // G/h matrices (ping-pong across iterations)
    float* flow = nullptr;         // w*h*2 fp32 (dense) flow of this layer
};

// kernel classes for rc_profile_* (per-launch CUDA-event timing) -- order matches rc_kernel_names[]
enum RcKernelId { K_PYR_H = 0, K_PYR_V, K_POLYEXP, K_UPDATE_MATRICES, K_FLOW_ITER_FUSED, K_FLOW_ITER_FINAL,
                  K_POLAR_HIST, K_THRESHOLDS, K_CLASSIFY, K_WINDOW, K_ADVECT, K_STREAKLINE, K_MISC, K_COUNT };
extern const char* const rc_kernel_names[K_COUNT];

struct ProfRec { int id; cudaEvent_t a, b; double bytes; };

struct rc_ctx {
    int device = 0;
    bool prof_on = false;
    std::vector<ProfRec> prof;          // pending records (events not yet read)
    std::vector<cudaEvent_t> ev_pool;   // recycled events
    double prof_ms[K_COUNT] = {0};
    double prof_bytes[K_COUNT] = {0};
    int64_t prof_n[K_COUNT] = {0};
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    int64_t launches = 0;

    // Farneback state
    bool configured = false;
    FarnebackParams prm;
    int nlayers = 0;
    Layer layer[RC_MAX_LAYERS];
    PolyCoef poly;
    GaussWin gwin;
    uint8_t* d_frame = nullptr;    // staging for a host frame (w*h u8, dense)
    size_t d_frame_cap = 0;
    int cur = 0;                   // R[cur] holds the most recent frame
    int frames_seen = 0;
    bool have_flow = false;
    float* flow_out = nullptr;     // == layer[0].flow
    std::vector<void*> allocs;

    // pinned staging
    void* h_pin = nullptr;
    size_t h_pin_cap = 0;
    void* d_tmp = nullptr;         // generic device scratch for host-pointer arguments
    size_t d_tmp_cap = 0;
    void* d_tmp2 = nullptr;
    size_t d_tmp2_cap = 0;

    // aggregation state
    unsigned long long* d_hist2d = nullptr;   // RC_HIST_ROWS*RC_HIST_BINS
    float* d_thr = nullptr;                   // [0]=UPPER, [1..36]=UPPER2d, [37..72]=prop, then int64 histsum
    float* d_acc = nullptr;                   // accumulator.x, acc_w*acc_h
    int acc_w = 0, acc_h = 0;
    uint8_t* d_mask = nullptr;                // outmask scratch
    uint8_t* d_cls = nullptr;                 // waveclass / waterclass scratch (2 planes)

    // sliding window
    int win_W = 0, win_w = 0, win_h = 0, win_i = 0;
    float* d_ring = nullptr;                  // W slots of w*h*2
    float* d_avg = nullptr;

    // advection scratch handled through d_tmp
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
        if (a) { cudaEventRecord(b, c->stream); c->prof.push_back(ProfRec{id, a, b, bytes}); }
    }
};

// ---- kernel launchers (farneback.cu) -------------------------------------------------------------
void rc_launch_pyr_layer(rc_ctx* c, const uint8_t* d_img, size_t step, int W, int H, Layer& L);
void rc_launch_polyexp(rc_ctx* c, const float* I, int w, int h, int pitch, const Planes& R);
// flow_mode: 0 zero flow, 1 upsample `coarse` (cw x ch) and scale, 2 read `flow` (w x h)
void rc_launch_update_matrices(rc_ctx* c, const Planes& R0, const Planes& R1, const Planes& M, int flow_mode,
                               const float* flow, int cw, int ch, float flow_scale);
// one updateFlow iteration: blur M_in, solve; if M_out.p: fused updateMatrices into M_out, else write flow_out.
// hist2d != nullptr additionally bins the produced flow (A2+A3 fused into the final iteration of layer 0).
void rc_launch_update_flow(rc_ctx* c, const Planes& M_in, const Planes& R0, const Planes& R1, const Planes& M_out,
                           float* flow_out, unsigned long long* hist2d);

// ---- aggregate.cu ----------------------------------------------------------------------------------
void rc_launch_polar_hist(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, unsigned long long* hist2d);
void rc_launch_cart_to_polar(rc_ctx* c, const float* flow, size_t n, float* mag, float* ang);
void rc_launch_thresholds(rc_ctx* c, const unsigned long long* hist2d, float* thr);
void rc_launch_classify(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, const float* thr,
                        int framecount, float* acc, uint8_t* mask, uint8_t* waveclass, uint8_t* waterclass,
                        float* ring_old, float* avg, int W);
void rc_launch_window_update(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* slot, float* avg,
                             int W);
void rc_launch_subtract_mean(rc_ctx* c, float* flow, size_t flow_step, int w, int h, double* d_sums);

// ---- advect.cu -------------------------------------------------------------------------------------
void rc_launch_advect(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* seeds, size_t n, float dt,
                      int iterations, float upper, int variant, float* dist, const int32_t* home);
void rc_launch_streakline(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* emitters, int E,
                          float* vertices, int32_t* count, int cap, float dt);

// thresholds buffer layout (floats): [0] UPPER, [1..36] UPPER2d, [37..72] prop_above_upper
#define RC_THR_FLOATS 80
