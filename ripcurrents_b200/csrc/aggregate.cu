// Aggregation that follows the flow in the reference's frame loop (RipCurrents_main/ripcurrents.cpp:305-439,
// main.cpp:1143-1153): polar conversion, cumulative speed/direction histograms, tail thresholds,
// classification + accumulation + mask, sliding-window flow mean.  Compiled with -fmad=false; the fused
// multiply-adds cv::cartToPolar performs are written explicitly as fmaf().
#include <float.h>
#include "rc_internal.h"

namespace {

// cv::cartToPolar(x, y, mag, angle, angleInDegrees=true), default path: bit-exact restatement
// (pinned against cv2 4.13.0 in tests/test_oracle_aggregate.py).
__device__ __forceinline__ void cart_to_polar(float x, float y, float& mag, float& ang)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float c = __fdiv_rn(mn, mx + (float)DBL_EPSILON);
    float c2 = c * c;
    float a = fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
    if (ax < ay) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    ang = a;
    mag = __fsqrt_rn(fmaf(x, x, y * y));
}

__device__ __forceinline__ int hist_key(float dx, float dy)
{
    float mag, ang;
    cart_to_polar(dx, dy, mag, ang);
    int bin = (int)(mag * (float)RC_HIST_RESOLUTION);
    int dir = (int)__fdiv_rn(ang * (float)RC_HIST_DIRECTIONS, 360.f);
    if (bin < RC_HIST_BINS && bin >= 0) return dir * RC_HIST_BINS + bin;
    return -1;
}

// Warp-aggregated shared-memory histogram: lanes with the same (direction, bin) key elect one leader which adds
// the group's population count, so a frame whose motion is nearly uniform (every pixel in one bin) costs one
// shared atomic per warp instead of 32 serialised ones.  One global atomic per touched bin per CTA at the end.
__global__ void __launch_bounds__(256)
polar_hist_kernel(const float* __restrict__ flow, size_t flow_step, int w, int h, unsigned long long* __restrict__ hist2d)
{
    __shared__ unsigned int sh[RC_HIST_ROWS * RC_HIST_BINS];
    for (int i = threadIdx.x; i < RC_HIST_ROWS * RC_HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t n = (size_t)w * h;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = (n + stride - 1) / stride * stride;
    const bool dense = flow_step == (size_t)w * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        int key = -1;
        if (i < n) {
            float2 f;
            if (dense) f = reinterpret_cast<const float2*>(flow)[i];
            else {
                size_t y = i / w, x = i - y * w;
                f = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flow) + y * flow_step + x * 8);
            }
            key = hist_key(f.x, f.y);
        }
        unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[key], __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RC_HIST_ROWS * RC_HIST_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist2d[i], (unsigned long long)sh[i]);
}

__global__ void cart_to_polar_kernel(const float* __restrict__ flow, size_t n, float* __restrict__ mag,
                                     float* __restrict__ ang)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 f = reinterpret_cast<const float2*>(flow)[i];
    float m, a;
    cart_to_polar(f.x, f.y, m, a);
    mag[i] = m; ang[i] = a;
}

// Thresholds (ripcurrents.cpp:333-366).  One warp: lane a < 36 owns direction a; lane 0 also does the global one.
__global__ void thresholds_kernel(const unsigned long long* __restrict__ hist2d, float* __restrict__ thr)
{
    __shared__ long long hist[RC_HIST_BINS];
    __shared__ long long s_threshsum;
    __shared__ int s_target;
    const int t = threadIdx.x;
    if (t < RC_HIST_BINS) {
        long long s = 0;
        for (int a = 0; a < RC_HIST_ROWS; a++) s += (long long)hist2d[a * RC_HIST_BINS + t];
        hist[t] = s;
    }
    __syncthreads();
    if (t == 0) {
        long long histsum = 0;
        for (int b = 0; b < RC_HIST_BINS; b++) histsum += hist[b];
        long long threshsum = 0;
        int bin = RC_HIST_BINS - 1;
        while ((double)threshsum < ((double)histsum * .05)) { threshsum += hist[bin]; bin--; }
        thr[0] = __fdiv_rn((float)bin, (float)RC_HIST_RESOLUTION);
        s_threshsum = threshsum; s_target = bin;
        reinterpret_cast<long long*>(thr + 74)[0] = histsum;   // 8-byte aligned slot after the 73 floats
    }
    __syncthreads();
    if (t < RC_HIST_DIRECTIONS) {
        const unsigned long long* row = hist2d + t * RC_HIST_BINS;
        long long sum = 0;
        for (int b = 0; b < RC_HIST_BINS; b++) sum += (long long)row[b];
        long long t2 = 0, t3 = 0;
        int b = RC_HIST_BINS - 1;
        while ((double)t2 < ((double)sum * .05)) { t2 += (long long)row[b]; b--; }
        float u = __fdiv_rn((float)b, (float)RC_HIST_RESOLUTION);
        if ((double)u < 0.01) u = (float)0.01;
        thr[1 + t] = u;
        b = RC_HIST_BINS - 1;
        while (b > s_target) { t3 += (long long)row[b]; b--; }
        thr[37 + t] = __fdiv_rn((float)t3, (float)s_threshsum);
    }
}

// classify (mag > UPPER -> accumulator2.x = 1), accumulate when framecount > 30, mask/out classes; optionally the
// sliding-window update of main.cpp:1143-1153 on the same read of the flow.
__global__ void __launch_bounds__(256)
classify_kernel(const float* __restrict__ flow, size_t flow_step, int w, int h, float upper_arg,
                const float* __restrict__ thr, int framecount, float* __restrict__ acc, uint8_t* __restrict__ mask,
                uint8_t* __restrict__ waveclass, uint8_t* __restrict__ waterclass, float* __restrict__ ring_slot,
                float* __restrict__ avg, float inv_w)
{
    const size_t n = (size_t)w * h;
    const float upper = isnan(upper_arg) ? thr[0] : upper_arg;
    const double lo = .1 * framecount, hi = .2 * framecount;
    const bool dense = flow_step == (size_t)w * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 f;
        if (dense) f = reinterpret_cast<const float2*>(flow)[i];
        else {
            size_t y = i / w, x = i - y * w;
            f = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flow) + y * flow_step + x * 8);
        }
        float m, a;
        cart_to_polar(f.x, f.y, m, a);
        float acc2 = m > upper ? 1.f : 0.f;
        float av = acc[i];
        if (framecount > 30) { av = acc2 + av; acc[i] = av; }
        int val = (int)av;
        bool wave = (double)val > lo;
        if (mask) mask[i] = wave ? 0 : 255;
        if (waveclass) waveclass[i] = wave ? (((double)val < hi) ? 1 : 2) : 0;
        if (waterclass) waterclass[i] = m > upper ? 3 : (m > 0.5f ? 2 : (m > 0.2f ? 1 : 0));
        if (avg) {
            float2* rs = reinterpret_cast<float2*>(ring_slot) + i;
            float2* ap = reinterpret_cast<float2*>(avg) + i;
            float2 o = *rs, v = *ap;
            v.x = (v.x - o.x * inv_w) + f.x * inv_w;
            v.y = (v.y - o.y * inv_w) + f.y * inv_w;
            *rs = f; *ap = v;
        }
    }
}

__global__ void window_update_kernel(const float* __restrict__ flow, size_t flow_step, int w, int h,
                                     float* __restrict__ slot, float* __restrict__ avg, float inv_w)
{
    const size_t n = (size_t)w * h;
    const bool dense = flow_step == (size_t)w * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float2 f;
        if (dense) f = reinterpret_cast<const float2*>(flow)[i];
        else {
            size_t y = i / w, x = i - y * w;
            f = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flow) + y * flow_step + x * 8);
        }
        float2* rs = reinterpret_cast<float2*>(slot) + i;
        float2* ap = reinterpret_cast<float2*>(avg) + i;
        float2 o = *rs, v = *ap;
        v.x = (v.x - o.x * inv_w) + f.x * inv_w;
        v.y = (v.y - o.y * inv_w) + f.y * inv_w;
        *rs = f; *ap = v;
    }
}

// subtructAverage: cv::mean (fp64 sums), then per-pixel subtraction in fp64 rounded to fp32
__global__ void sum_flow_kernel(const float* __restrict__ flow, size_t flow_step, int w, int h, double* __restrict__ sums)
{
    const size_t n = (size_t)w * h;
    double sx = 0, sy = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / w, x = i - y * w;
        float2 f = *reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flow) + y * flow_step + x * 8);
        sx += f.x; sy += f.y;
    }
    for (int o = 16; o; o >>= 1) {
        sx += __shfl_down_sync(0xffffffffu, sx, o);
        sy += __shfl_down_sync(0xffffffffu, sy, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], sx); atomicAdd(&sums[1], sy); }
}

__global__ void sub_mean_kernel(float* __restrict__ flow, size_t flow_step, int w, int h, const double* __restrict__ sums)
{
    const size_t n = (size_t)w * h;
    const double mx = sums[0] / (double)n, my = sums[1] / (double)n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t y = i / w, x = i - y * w;
        float2* p = reinterpret_cast<float2*>(reinterpret_cast<char*>(flow) + y * flow_step + x * 8);
        float2 f = *p;
        f.x = (float)((double)f.x - mx); f.y = (float)((double)f.y - my);
        *p = f;
    }
}

int grid_for(size_t n, int block, int per_sm)
{
    size_t g = (n + block - 1) / block;
    size_t cap = (size_t)148 * per_sm;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

void rc_launch_polar_hist(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, unsigned long long* hist2d)
{
    const size_t n = (size_t)w * h;
    KScope ks(c, K_POLAR_HIST, 8.0 * n);
    polar_hist_kernel<<<grid_for(n, 256, 4), 256, 0, c->stream>>>(flow, flow_step, w, h, hist2d);
}

void rc_launch_cart_to_polar(rc_ctx* c, const float* flow, size_t n, float* mag, float* ang)
{
    if (!n) return;
    KScope ks(c, K_MISC, 16.0 * n);
    cart_to_polar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(flow, n, mag, ang);
}

void rc_launch_thresholds(rc_ctx* c, const unsigned long long* hist2d, float* thr)
{
    KScope ks(c, K_THRESHOLDS, 8.0 * RC_HIST_ROWS * RC_HIST_BINS);
    thresholds_kernel<<<1, 64, 0, c->stream>>>(hist2d, thr);
}

void rc_launch_classify(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, const float* thr,
                        int framecount, float* acc, uint8_t* mask, uint8_t* waveclass, uint8_t* waterclass,
                        float* ring_slot, float* avg, int W)
{
    const size_t n = (size_t)w * h;
    const float inv = W > 0 ? (float)(1.0 / (double)W) : 0.f;
    KScope ks(c, K_CLASSIFY, (8.0 + 8.0 + (mask ? 1.0 : 0.0) + (avg ? 32.0 : 0.0)) * n);
    classify_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(flow, flow_step, w, h, upper, thr, framecount, acc, mask,
                                                               waveclass, waterclass, ring_slot, avg, inv);
}

void rc_launch_window_update(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float* slot, float* avg, int W)
{
    const size_t n = (size_t)w * h;
    KScope ks(c, K_WINDOW, 40.0 * n);
    window_update_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(flow, flow_step, w, h, slot, avg,
                                                                    (float)(1.0 / (double)W));
}

void rc_launch_subtract_mean(rc_ctx* c, float* flow, size_t flow_step, int w, int h, double* d_sums)
{
    const size_t n = (size_t)w * h;
    cudaMemsetAsync(d_sums, 0, 2 * sizeof(double), c->stream);
    KScope ks(c, K_MISC, 24.0 * n, 2);
    sum_flow_kernel<<<grid_for(n, 256, 4), 256, 0, c->stream>>>(flow, flow_step, w, h, d_sums);
    sub_mean_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(flow, flow_step, w, h, d_sums);
}
