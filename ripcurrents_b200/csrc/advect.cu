// Seed-parallel Euler / bilinear-gather particle advection (pathlines, streamlines, per-pixel particle field,
// streaklines).  Reference: RipCurrents_main/pathlines.cpp:9-46, ripcurrents.cpp:611-698,
// ripcurrents_module.cpp:486-679, Streakline.cpp:11-48.  One thread per seed; the four flow taps of a step are
// plain (texture-free) 8-byte global loads that hit L1/L2 -- the flow field of a 1080p frame is 16.6 MB, far
// inside the 126 MB L2.  Compiled with -fmad=false: every cv::Point_<float> operator rounds separately.
#include "rc_internal.h"

namespace {

struct FlowView {
    const char* p;
    size_t step;
    int w, h;
    __device__ __forceinline__ float2 at(int x, int y) const
    {
        return __ldg(reinterpret_cast<const float2*>(p + (size_t)y * step) + x);
    }
};

// returns false when the particle has left the interior (reference: silent stop)
__device__ __forceinline__ bool gather(const FlowView& F, float x, float y, float& dx, float& dy)
{
    int xi = (int)floorf(x), yi = (int)floorf(y);
    float xr = x - (float)xi, yr = y - (float)yi;
    if (xi < 1 || yi < 1 || xi + 2 > F.w || yi + 2 > F.h) return false;
    float2 p00 = F.at(xi, yi), p01 = F.at(xi + 1, yi), p10 = F.at(xi, yi + 1), p11 = F.at(xi + 1, yi + 1);
    float ax = 1.f - xr, ay = 1.f - yr;
    float t0x = (p00.x * ax) * ay, t0y = (p00.y * ax) * ay;
    float t1x = (p01.x * xr) * ay, t1y = (p01.y * xr) * ay;
    float t2x = (p10.x * ax) * yr, t2y = (p10.y * ax) * yr;
    float t3x = (p11.x * xr) * yr, t3y = (p11.y * xr) * yr;
    dx = ((t0x + t1x) + t2x) + t3x;
    dy = ((t0y + t1y) + t2y) + t3y;
    return true;
}

__global__ void __launch_bounds__(256)
advect_kernel(FlowView F, float* __restrict__ seeds, size_t n, float dt, int iterations, float upper, int variant,
              float* __restrict__ dist, const int32_t* __restrict__ home)
{
    size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float2 p = reinterpret_cast<float2*>(seeds)[s];
    int xo = 0, yo = 0;
    if (variant == RC_ADV_FIELD || variant == RC_ADV_GET_DELTA) {
        if (home) { xo = home[2 * s]; yo = home[2 * s + 1]; }
        else { xo = (int)(s % (size_t)F.w); yo = (int)(s / (size_t)F.w); }
    }
    int nit = iterations;
    if (variant == RC_ADV_FIXED100) nit = 100;
    if (variant == RC_ADV_GET_DELTA) nit = 1;
    const float fit = (float)iterations;
    float d = dist ? dist[s] : 0.f;
    for (int it = 0; it < nit; it++) {
        float dx, dy;
        if (!gather(F, p.x + (float)xo, p.y + (float)yo, dx, dy)) break;
        float r = __fsqrt_rn(dx * dx + dy * dy);
        bool stop = false;
        switch (variant) {
        case RC_ADV_PATHLINE:
            p.x = p.x + __fdiv_rn(dx * dt, fit); p.y = p.y + __fdiv_rn(dy * dt, fit); break;
        case RC_ADV_LEGACY:
            if (r > upper) { stop = true; break; }
            p.x = p.x + __fdiv_rn(dx * dt, fit); p.y = p.y + __fdiv_rn(dy * dt, fit); break;
        case RC_ADV_MODULE:
        case RC_ADV_GET_DELTA:
            if (r > upper) { stop = true; break; }
            p.x = p.x + dx * dt; p.y = p.y + dy * dt; break;
        case RC_ADV_CUT5:
            if (r > 5.f) { stop = true; break; }
            p.x = p.x + dx * dt; p.y = p.y + dy * dt; break;
        case RC_ADV_FIXED100:
            p.x = p.x + (float)((double)dx * 0.1); p.y = p.y + (float)((double)dy * 0.1); break;
        case RC_ADV_FIELD:
            if (r > upper) { stop = true; break; }
            p.x = p.x + __fdiv_rn(dx * dt, fit); p.y = p.y + __fdiv_rn(dy * dt, fit);
            d = d + r;
            break;
        default: stop = true;
        }
        if (stop) break;
    }
    reinterpret_cast<float2*>(seeds)[s] = p;
    if (dist && variant == RC_ADV_FIELD) dist[s] = d;
}

// One Streakline frame: thread (e, i) moves vertex i of emitter e and writes it to slot i+1 of the output
// (the reference inserts the generation point at the FRONT, Streakline.cpp:46-48); slot 0 receives the emitter.
__global__ void __launch_bounds__(256)
streakline_kernel(FlowView F, const float* __restrict__ emitters, int E, const float* __restrict__ vin,
                  float* __restrict__ vout, int32_t* __restrict__ count, int cap, float dt, double lim_x, double lim_y)
{
    const int e = blockIdx.y;
    const int c = count[e];
    const float2* in = reinterpret_cast<const float2*>(vin) + (size_t)e * cap;
    float2* out = reinterpret_cast<float2*>(vout) + (size_t)e * cap;
    const bool grow = c < cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < c; i += gridDim.x * blockDim.x) {
        float2 v = in[i];
        float dx, dy;
        if (gather(F, v.x, v.y, dx, dy)) {
            float nx = v.x + dx * dt, ny = v.y + dy * dt;
            if (!((double)fabsf(v.x - nx) > lim_x || (double)fabsf(v.y - ny) > lim_y)) { v.x = nx; v.y = ny; }
        }
        out[grow ? i + 1 : i] = v;
    }
    if (grow && blockIdx.x == 0 && threadIdx.x == 0) out[0] = reinterpret_cast<const float2*>(emitters)[e];
}

// averageVector's window update (ripcurrents_module.cpp:392-400) in one pass: per pixel
//   average -= old / frames;  new = get_delta(pixel = 0, home (x, y), flow, dt, UPPER);  average += new / frames
// (`Mat / s` scales by (float)(1.0 / s), oracle/aggregate_oracle.c); new_slot (optional) receives the new field.
__global__ void __launch_bounds__(256)
average_vector_kernel(FlowView F, const float* __restrict__ old_slot, float* __restrict__ average,
                      float* __restrict__ new_slot, float inv, float dt, float upper)
{
    const size_t n = (size_t)F.w * F.h;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int x = (int)(i % (size_t)F.w), y = (int)(i / (size_t)F.w);
    float2 p = make_float2(0.f, 0.f);
    float dx, dy;
    if (gather(F, (float)x, (float)y, dx, dy) && !(__fsqrt_rn(dx * dx + dy * dy) > upper)) { p.x = p.x + dx * dt; p.y = p.y + dy * dt; }
    float2 a = reinterpret_cast<float2*>(average)[i];
    const float2 o = old_slot ? reinterpret_cast<const float2*>(old_slot)[i] : make_float2(0.f, 0.f);
    a.x = (a.x - o.x * inv) + p.x * inv;
    a.y = (a.y - o.y * inv) + p.y * inv;
    reinterpret_cast<float2*>(average)[i] = a;
    if (new_slot) reinterpret_cast<float2*>(new_slot)[i] = p;
}

__global__ void streakline_count_kernel(int32_t* count, int E, int cap)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E && count[e] < cap) count[e] += 1;
}

}  // namespace

void rc_launch_advect(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* seeds, size_t n, float dt,
                      int iterations, float upper, int variant, float* dist, const int32_t* home)
{
    if (!n) return;
    FlowView F{reinterpret_cast<const char*>(flow), flow_step, w, h};
    KScope ks(c, K_ADVECT, (16.0 + (dist ? 8.0 : 0.0)) * n + 8.0 * w * h);
    advect_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(F, seeds, n, dt, iterations, upper, variant, dist,
                                                                     home);
}

void rc_launch_average_vector(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* old_slot,
                              float* average, float* new_slot, int frames, float dt, float upper)
{
    FlowView F{reinterpret_cast<const char*>(flow), flow_step, w, h};
    const size_t n = (size_t)w * h;
    KScope ks(c, K_WINDOW, (8.0 + (old_slot ? 8.0 : 0.0) + 16.0 + (new_slot ? 8.0 : 0.0)) * n);
    average_vector_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(F, old_slot, average, new_slot,
                                                                             (float)(1.0 / (double)frames), dt, upper);
}

// vertices is updated out of place into `vout` by the caller-provided scratch; the launcher copies back.
void rc_launch_streakline(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, const float* emitters, int E,
                          float* vertices, int32_t* count, int cap, float dt)
{
    if (E <= 0) return;
    FlowView F{reinterpret_cast<const char*>(flow), flow_step, w, h};
    const size_t bytes = (size_t)E * cap * 2 * sizeof(float);
    float* vout = reinterpret_cast<float*>(c->d_tmp2);   // sized by the API layer
    int gx = (cap + 255) / 256; if (gx > 64) gx = 64;
    dim3 g(gx, E);
    KScope ks(c, K_STREAKLINE, 16.0 * E * cap + 8.0 * w * h, 2);
    streakline_kernel<<<g, 256, 0, c->stream>>>(F, emitters, E, vertices, vout, count, cap, dt, w * 0.1, h * 0.1);
    streakline_count_kernel<<<(E + 127) / 128, 128, 0, c->stream>>>(count, E, cap);
    cudaMemcpyAsync(vertices, vout, bytes, cudaMemcpyDeviceToDevice, c->stream);
}
