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

// Thresholds (ripcurrents.cpp:333-366) after each frame of a batch.  CTA j owns frame j: it forms the cumulative
// counts through its frame (base + delta[0..j]) and runs the tail scans; thread a < 36 owns direction a, thread 0
// also does the global threshold.  The last CTA's sums become the new cumulative counters (histnext), so the batch
// is one parallel launch instead of nb dependent ones.
__global__ void __launch_bounds__(256)
thresholds_batch_kernel(const unsigned long long* hist2d, unsigned long long* histnext,       // may alias when nb == 1
                        const unsigned int* __restrict__ delta, int nb, float* __restrict__ thr_batch,
                        float* __restrict__ thr_last)
{
    __shared__ long long cum[RC_HIST_CELLS];
    __shared__ long long hist[RC_HIST_BINS];
    __shared__ long long s_threshsum;
    __shared__ int s_target;
    const int t = threadIdx.x, j = blockIdx.x;
    const int nd = nb > 0 ? j + 1 : 0;
    for (int i = t; i < RC_HIST_CELLS; i += blockDim.x) {
        long long s = (long long)hist2d[i];
        for (int d = 0; d < nd; d++) s += (long long)delta[(size_t)d * RC_HIST_CELLS + i];
        cum[i] = s;
    }
    __syncthreads();
    const bool is_last = nb <= 0 || j == nb - 1;
    float* thr = thr_batch ? thr_batch + (size_t)j * RC_THR_FLOATS : thr_last;
    if (t < RC_HIST_BINS) {
        long long s = 0;
        for (int a = 0; a < RC_HIST_ROWS; a++) s += cum[a * RC_HIST_BINS + t];
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
        const long long* row = cum + t * RC_HIST_BINS;
        long long sum = 0;
        for (int b = 0; b < RC_HIST_BINS; b++) sum += row[b];
        long long t2 = 0, t3 = 0;
        int b = RC_HIST_BINS - 1;
        while ((double)t2 < ((double)sum * .05)) { t2 += row[b]; b--; }
        float u = __fdiv_rn((float)b, (float)RC_HIST_RESOLUTION);
        if ((double)u < 0.01) u = (float)0.01;
        thr[1 + t] = u;
        b = RC_HIST_BINS - 1;
        while (b > s_target) { t3 += row[b]; b--; }
        thr[37 + t] = __fdiv_rn((float)t3, (float)s_threshsum);
    }
    __syncthreads();
    if (is_last) {
        if (thr_batch && thr_last) for (int i = t; i < RC_THR_FLOATS; i += blockDim.x) thr_last[i] = thr[i];
        if (nb > 0) for (int i = t; i < RC_HIST_CELLS; i += blockDim.x) histnext[i] = (unsigned long long)cum[i];
    }
}

// classify (mag > UPPER -> accumulator2.x = 1), accumulate when framecount > 30, mask/out classes: single frame
__global__ void __launch_bounds__(256)
classify_kernel(const float* __restrict__ flow, size_t flow_step, int w, int h, float upper_arg,
                const float* __restrict__ thr, int framecount, float* __restrict__ acc, uint8_t* __restrict__ mask,
                uint8_t* __restrict__ waveclass, uint8_t* __restrict__ waterclass)
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
    }
}

// Batched classify + accumulate + mask + sliding-window mean: each thread owns 4 adjacent pixels and walks the
// nb frames of the batch in order, keeping accumulator.x and the window mean in registers -- they are read and
// written ONCE per batch instead of once per frame.  Per frame and pixel: 8 B flow + 8 B leaving ring slot read,
// 1 B mask written.  (ripcurrents.cpp:376-439 + main.cpp:1143-1153, in the reference's order.)
__global__ void __launch_bounds__(256)
classify_batch_kernel(ClassifyBatch cb, size_t n4, size_t n, const float* __restrict__ thr_batch, int framecount0,
                      float* __restrict__ acc, uint8_t* __restrict__ masks, float* __restrict__ avg, float inv_w, int vec, int packed)
{
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i4 >= n4) return;
    const size_t i = i4 * 4;
    // vec: the host checked that w*h % 4 == 0 and that every flow / ring / mask / accumulator pointer is 16-byte (masks:
    // 4-byte) aligned -- odd image sizes put every other ring slot and mask row off alignment and take the scalar path
    const bool full = vec && i + 3 < n;
    float av[4]; float2 mean[4];
    if (full) {
        float4 t = *reinterpret_cast<const float4*>(acc + i);
        av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        if (avg) {
            float4 m0 = *reinterpret_cast<const float4*>(avg + 2 * i), m1 = *reinterpret_cast<const float4*>(avg + 2 * i + 4);
            mean[0] = make_float2(m0.x, m0.y); mean[1] = make_float2(m0.z, m0.w);
            mean[2] = make_float2(m1.x, m1.y); mean[3] = make_float2(m1.z, m1.w);
        }
    } else {
        for (int k = 0; k < 4; k++) {
            av[k] = i + k < n ? acc[i + k] : 0.f;
            mean[k] = (avg && i + k < n) ? reinterpret_cast<const float2*>(avg)[i + k] : make_float2(0.f, 0.f);
        }
    }
    for (int j = 0; j < cb.nb; j++) {
        const float upper = thr_batch[(size_t)j * RC_THR_FLOATS];
        const int fc = framecount0 + j;
        const double lo = .1 * fc;
        float2 f[4], o[4];
        if (full) {
            float4 a = __ldcs(reinterpret_cast<const float4*>(cb.flow[j] + 2 * i));
            float4 b = __ldcs(reinterpret_cast<const float4*>(cb.flow[j] + 2 * i + 4));
            f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w); f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
            if (avg && cb.old[j]) {
                float4 c = __ldcs(reinterpret_cast<const float4*>(cb.old[j] + 2 * i));
                float4 d = __ldcs(reinterpret_cast<const float4*>(cb.old[j] + 2 * i + 4));
                o[0] = make_float2(c.x, c.y); o[1] = make_float2(c.z, c.w); o[2] = make_float2(d.x, d.y); o[3] = make_float2(d.z, d.w);
            } else { o[0] = o[1] = o[2] = o[3] = make_float2(0.f, 0.f); }
        } else {
            for (int k = 0; k < 4; k++) {
                f[k] = i + k < n ? reinterpret_cast<const float2*>(cb.flow[j])[i + k] : make_float2(0.f, 0.f);
                o[k] = (avg && cb.old[j] && i + k < n) ? reinterpret_cast<const float2*>(cb.old[j])[i + k] : make_float2(0.f, 0.f);
            }
        }
        unsigned char mk[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float m, a;
            cart_to_polar(f[k].x, f[k].y, m, a);
            if (fc > 30) av[k] = (m > upper ? 1.f : 0.f) + av[k];
            mk[k] = ((double)(int)av[k] > lo) ? 0 : 255;
            if (avg) {
                mean[k].x = (mean[k].x - o[k].x * inv_w) + f[k].x * inv_w;
                mean[k].y = (mean[k].y - o[k].y * inv_w) + f[k].y * inv_w;
            }
        }
        if (masks && packed) {
            // RC_MASK_PACKED: 1 bit per pixel (1 = calm / 255, 0 = wave), pixel p in bit p & 7 of byte p >> 3; the launcher
            // guarantees n % 8 == 0, so lanes (2k, 2k+1) hold the two nibbles of one byte and are active together
            const unsigned nib = (mk[0] ? 1u : 0u) | (mk[1] ? 2u : 0u) | (mk[2] ? 4u : 0u) | (mk[3] ? 8u : 0u);
            const unsigned other = __shfl_xor_sync(__activemask(), nib, 1);
            if (!(i4 & 1)) masks[(size_t)j * (n >> 3) + (i4 >> 1)] = (uint8_t)(nib | (other << 4));
        } else if (masks) {
            uint8_t* mrow = masks + (size_t)j * n;
            if (full) *reinterpret_cast<uchar4*>(mrow + i) = make_uchar4(mk[0], mk[1], mk[2], mk[3]);
            else for (int k = 0; k < 4; k++) if (i + k < n) mrow[i + k] = mk[k];
        }
    }
    if (full) {
        *reinterpret_cast<float4*>(acc + i) = make_float4(av[0], av[1], av[2], av[3]);
        if (avg) {
            *reinterpret_cast<float4*>(avg + 2 * i) = make_float4(mean[0].x, mean[0].y, mean[1].x, mean[1].y);
            *reinterpret_cast<float4*>(avg + 2 * i + 4) = make_float4(mean[2].x, mean[2].y, mean[3].x, mean[3].y);
        }
    } else {
        for (int k = 0; k < 4; k++)
            if (i + k < n) { acc[i + k] = av[k]; if (avg) reinterpret_cast<float2*>(avg)[i + k] = mean[k]; }
    }
}

// Sliding-window mean alone (main.cpp:1143-1153) over nb consecutive flows: the mean stays in registers across the batch.
__global__ void __launch_bounds__(256)
window_batch_kernel(ClassifyBatch cb, size_t n2, float* __restrict__ avg, float inv_w)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // index of a float2 PAIR (two pixels, 16 bytes)
    if (i >= n2) return;
    float4 m = reinterpret_cast<const float4*>(avg)[i];
    for (int j = 0; j < cb.nb; j++) {
        const float4 f = __ldcs(reinterpret_cast<const float4*>(cb.flow[j]) + i);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cb.old[j]) o = __ldcs(reinterpret_cast<const float4*>(cb.old[j]) + i);
        m.x = (m.x - o.x * inv_w) + f.x * inv_w; m.y = (m.y - o.y * inv_w) + f.y * inv_w;
        m.z = (m.z - o.z * inv_w) + f.z * inv_w; m.w = (m.w - o.w * inv_w) + f.w * inv_w;
    }
    reinterpret_cast<float4*>(avg)[i] = m;
}

__global__ void __launch_bounds__(256)
window_batch_kernel_px(ClassifyBatch cb, size_t n, float* __restrict__ avg, float inv_w)     // odd image sizes: one pixel per thread
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float2 m = reinterpret_cast<const float2*>(avg)[i];
    for (int j = 0; j < cb.nb; j++) {
        const float2 f = reinterpret_cast<const float2*>(cb.flow[j])[i];
        float2 o = make_float2(0.f, 0.f);
        if (cb.old[j]) o = reinterpret_cast<const float2*>(cb.old[j])[i];
        m.x = (m.x - o.x * inv_w) + f.x * inv_w; m.y = (m.y - o.y * inv_w) + f.y * inv_w;
    }
    reinterpret_cast<float2*>(avg)[i] = m;
}

// Sharded stream: gathered = [nranks][B][CELLS] per-frame counts of this super-block (zero rows for frames
// a rank does not have).  hist_start = global + all frames of lower ranks (what this rank's thresholds start from);
// hist_global += all frames of all ranks (the stream's counters after this super-block).
__global__ void shard_prefix_kernel(const unsigned int* __restrict__ gathered, int rank, int nranks, int B,
                                    unsigned long long* __restrict__ hist_global, unsigned long long* __restrict__ hist_start)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= RC_HIST_CELLS) return;
    unsigned long long lower = 0, all = 0;
    for (int r = 0; r < nranks; r++) {
        unsigned long long s = 0;
        for (int j = 0; j < B; j++) s += gathered[((size_t)r * B + j) * RC_HIST_CELLS + i];
        if (r < rank) lower += s;
        all += s;
    }
    const unsigned long long g = hist_global[i];
    hist_start[i] = g + lower;
    hist_global[i] = g + all;
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

__global__ void widen_counts_kernel(const unsigned int* __restrict__ in, long long* __restrict__ out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (long long)in[i];
}

// outmask of ripcurrents.cpp:424-439 from the accumulator alone (reporting point of a sharded stream)
__global__ void acc_mask_kernel(const float* __restrict__ acc, size_t n, double lo, uint8_t* __restrict__ mask)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mask[i] = ((double)(int)acc[i] > lo) ? 0 : 255;
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

void rc_launch_thresholds_batch(rc_ctx* c, unsigned long long* hist2d, const unsigned int* delta, int nb,
                                float* thr_batch, float* thr_last)
{
    KScope ks(c, K_THRESHOLDS, (8.0 + 4.0 * nb) * RC_HIST_CELLS);
    // CTAs read the old counters and the last one writes the new ones: double-buffered to keep the launch race-free.  A
    // single frame is a single CTA, which has every counter in shared memory before it writes any: in place, no copy
    // (frame-by-frame use: the extra copy was ~7 us of a 195 us step)
    unsigned long long* next = nb > 1 ? hist2d + RC_HIST_CELLS : hist2d;
    thresholds_batch_kernel<<<nb > 0 ? nb : 1, 256, 0, c->stream>>>(hist2d, next, delta, nb, thr_batch, thr_last);
    if (nb > 1) cudaMemcpyAsync(hist2d, next, sizeof(unsigned long long) * RC_HIST_CELLS, cudaMemcpyDeviceToDevice, c->stream);
}

void rc_launch_classify(rc_ctx* c, const float* flow, size_t flow_step, int w, int h, float upper, const float* thr,
                        int framecount, float* acc, uint8_t* mask, uint8_t* waveclass, uint8_t* waterclass)
{
    const size_t n = (size_t)w * h;
    KScope ks(c, K_CLASSIFY, (8.0 + 8.0 + (mask ? 1.0 : 0.0)) * n);
    classify_kernel<<<grid_for(n, 256, 8), 256, 0, c->stream>>>(flow, flow_step, w, h, upper, thr, framecount, acc, mask,
                                                               waveclass, waterclass);
}

void rc_launch_classify_batch(rc_ctx* c, const ClassifyBatch& cb, int w, int h, const float* thr_batch, int framecount0,
                              float* acc, uint8_t* masks, float* avg, int W, int packed)
{
    const size_t n = (size_t)w * h, n4 = (n + 3) / 4;
    const float inv = W > 0 ? (float)(1.0 / (double)W) : 0.f;
    int nold = 0;
    for (int j = 0; j < cb.nb; j++) nold += (avg && cb.old[j]) ? 1 : 0;
    auto al = [](const void* p, size_t a) { return (reinterpret_cast<size_t>(p) & (a - 1)) == 0; };
    bool vec = n % 4 == 0 && al(acc, 16) && al(avg, 16) && al(masks, 4);
    for (int j = 0; j < cb.nb && vec; j++) vec = al(cb.flow[j], 16) && al(cb.old[j], 16);
    if (packed) vec = vec && true;      // packed masks are byte stores: no alignment requirement on `masks`
    KScope ks(c, K_CLASSIFY, ((8.0 + (masks ? (packed ? 0.125 : 1.0) : 0.0)) * cb.nb + 8.0 * nold + 8.0 + (avg ? 16.0 : 0.0)) * n);
    classify_batch_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, c->stream>>>(cb, n4, n, thr_batch, framecount0, acc, masks,
                                                                              avg, inv, vec ? 1 : 0, packed);
}

void rc_launch_window_batch(rc_ctx* c, const ClassifyBatch& cb, int w, int h, float* avg, int W)
{
    const size_t n = (size_t)w * h;
    int nold = 0;
    auto al = [](const void* p) { return (reinterpret_cast<size_t>(p) & 15) == 0; };
    bool vec = n % 2 == 0 && al(avg);
    for (int j = 0; j < cb.nb; j++) { nold += cb.old[j] ? 1 : 0; vec = vec && al(cb.flow[j]) && al(cb.old[j]); }
    KScope ks(c, K_WINDOW, (8.0 * cb.nb + 8.0 * nold + 16.0) * n);
    const float inv = (float)(1.0 / (double)W);
    if (vec) window_batch_kernel<<<(unsigned)((n / 2 + 255) / 256), 256, 0, c->stream>>>(cb, n / 2, avg, inv);
    else window_batch_kernel_px<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(cb, n, avg, inv);
}

void rc_launch_shard_prefix(rc_ctx* c, const unsigned int* gathered, int rank, int nranks, int B, unsigned long long* hist_global,
                            unsigned long long* hist_start)
{
    KScope ks(c, K_THRESHOLDS, 4.0 * nranks * B * RC_HIST_CELLS);
    shard_prefix_kernel<<<(RC_HIST_CELLS + 127) / 128, 128, 0, c->stream>>>(gathered, rank, nranks, B, hist_global, hist_start);
}

void rc_launch_widen_counts(rc_ctx* c, const unsigned int* in, long long* out, size_t n)
{
    KScope ks(c, K_MISC, 12.0 * n);
    widen_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(in, out, n);
}

void rc_launch_acc_mask(rc_ctx* c, const float* acc, size_t n, int framecount, uint8_t* mask)
{
    KScope ks(c, K_MISC, 5.0 * n);
    acc_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(acc, n, .1 * framecount, mask);
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
