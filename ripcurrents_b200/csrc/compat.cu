// Kernels behind the header-compatible C++ entry points (ripcurrents_b200/cpp) that take the reference's own
// intermediate formats: the merged polar image CV_32FC3 (angle, mag, mag) of ripcurrents.cpp:305-309 and the
// CV_32FC3 accumulator / class images of ripcurrents.cpp:371-439.  The device-resident pipeline (rc_process_frames)
// never materialises these; they exist so that create_histogram / create_flow / create_accumulationbuffer can be
// swapped in one at a time.  Compiled with -fmad=false.
#include <math.h>
#include "rc_internal.h"

namespace {

__global__ void __launch_bounds__(256)
hist_polar_kernel(const float* __restrict__ polar, size_t step, int w, int h, unsigned long long* __restrict__ hist2d)
{
    __shared__ unsigned int sh[RC_HIST_CELLS];
    for (int i = threadIdx.x; i < RC_HIST_CELLS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t n = (size_t)w * h, stride = (size_t)gridDim.x * blockDim.x;
    const size_t nround = (n + stride - 1) / stride * stride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += stride) {
        int key = -1;
        if (i < n) {
            const size_t y = i / w, x = i - y * w;
            const float* p = reinterpret_cast<const float*>(reinterpret_cast<const char*>(polar) + y * step) + 3 * x;
            const int bin = (int)(p[1] * (float)RC_HIST_RESOLUTION);                       // ripcurrents.cpp:323
            const int dir = (int)__fdiv_rn(p[0] * (float)RC_HIST_DIRECTIONS, 360.f);        // ripcurrents.cpp:324
            if (bin < RC_HIST_BINS && bin >= 0 && dir >= 0 && dir < RC_HIST_ROWS) key = dir * RC_HIST_BINS + bin;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[key], __popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RC_HIST_CELLS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist2d[i], (unsigned long long)sh[i]);
}

// create_flow (ripcurrents_module.cpp:153-182): classes, accumulator2.x += 1, display rescale of `current`
__global__ void create_flow_kernel(float* __restrict__ cur, size_t cstep, float* __restrict__ wc, size_t wstep,
                                   float* __restrict__ acc2, size_t astep, int w, int h, float UPPER, float MID,
                                   float LOWER, const float* __restrict__ upper2d)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float* p = reinterpret_cast<float*>(reinterpret_cast<char*>(cur) + (size_t)y * cstep) + 3 * x;
    float* c = reinterpret_cast<float*>(reinterpret_cast<char*>(wc) + (size_t)y * wstep) + 3 * x;
    float* a = reinterpret_cast<float*>(reinterpret_cast<char*>(acc2) + (size_t)y * astep) + 3 * x;
    int angle = (int)__fdiv_rn(p[0] * (float)RC_HIST_DIRECTIONS, 360.f);
    const float val = p[2];
    if (val > UPPER) { c[0] = .5f; a[0] = a[0] + 1.f; }
    else if (val > MID) c[2] = 1.f;
    else if (val > LOWER) c[2] = .5f;
    else c[1] = .5f;
    angle = angle < 0 ? 0 : (angle >= RC_HIST_DIRECTIONS ? RC_HIST_DIRECTIONS - 1 : angle);   // reference: UB for 36
    const float z = __fdiv_rn(val, upper2d[angle]);
    p[2] = z;
    p[1] = z > 1.f ? 1.f : .7f;
}

// create_accumulationbuffer (ripcurrents_module.cpp:189-212)
__global__ void accumulate_kernel(float* __restrict__ acc, size_t astep, const float* __restrict__ acc2, size_t a2step,
                                  float* __restrict__ out, size_t ostep, unsigned char* __restrict__ mask, size_t mstep,
                                  int w, int h, int framecount)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float* a = reinterpret_cast<float*>(reinterpret_cast<char*>(acc) + (size_t)y * astep) + 3 * x;
    const float* a2 = reinterpret_cast<const float*>(reinterpret_cast<const char*>(acc2) + (size_t)y * a2step) + 3 * x;
    float* o = reinterpret_cast<float*>(reinterpret_cast<char*>(out) + (size_t)y * ostep) + 3 * x;
    if (framecount > 30) { a[0] = a2[0] + a[0]; a[1] = a2[1] + a[1]; a[2] = a2[2] + a[2]; }
    const int val = (int)a[0];
    if ((double)val > .1 * framecount) {
        if ((double)val < .2 * framecount) o[2] = 1.f; else o[0] = 1.f;
    } else {
        o[1] = .5f;
        mask[(size_t)y * mstep + x] = 255;
    }
}

// Frame ingest: cv::resize(INTER_LINEAR) of the 8-bit BGR camera frame to the working
// size followed by cv::cvtColor(COLOR_BGR2GRAY) -- ripcurrents.cpp:209-210, main.cpp:258-259 -- in OpenCV's fixed
// point (oracle/ingest_oracle.c; bit-exact against cv2 4.13.0): one thread per destination pixel.
__device__ __forceinline__ void ingest_coef(int d, int sn, double scale, bool clamp_f, int& s0, int& a0, int& a1)
{
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_f) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    s0 = s;
    a0 = __float2int_rn((1.f - f) * 2048.f);
    a1 = __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256)
ingest_bgr_kernel(const uint8_t* __restrict__ bgr, size_t step, size_t fstride, int sw, int sh, uint8_t* __restrict__ gray,
                  size_t gstep, size_t gstride, int dw, int dh, double scale_x, double scale_y, int legacy14)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    bgr += (size_t)blockIdx.z * fstride;
    gray += (size_t)blockIdx.z * gstride;
    int sx, a0, a1, sy, b0, b1;
    ingest_coef(x, sw, scale_x, true, sx, a0, a1);
    ingest_coef(y, sh, scale_y, false, sy, b0, b1);
    const int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
    const int sy1 = sy + 1 < 0 ? 0 : (sy + 1 < sh ? sy + 1 : sh - 1);
    sy = sy < 0 ? 0 : (sy < sh ? sy : sh - 1);
    const uint8_t* r0 = bgr + (size_t)sy * step;
    const uint8_t* r1 = bgr + (size_t)sy1 * step;
    int px[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int h0 = r0[3 * sx + c] * a0 + r0[3 * sx1 + c] * a1;
        const int h1 = r1[3 * sx + c] * a0 + r1[3 * sx1 + c] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        px[c] = v > 255 ? 255 : v;
    }
    const int g = legacy14 ? (px[0] * 1868 + px[1] * 9617 + px[2] * 4899 + (1 << 13)) >> 14
                           : (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
    gray[(size_t)y * gstep + x] = (uint8_t)g;
}

// INTER_AREA ingest (the PRIMING frame of every loop: ripcurrents.cpp:186-187, main.cpp:223,570,...): cv::resize(INTER_AREA)
// + cvtColor, downscaling only, bit-exact against cv2 4.13.0 (oracle/ingest_oracle.c states the two OpenCV code paths).
// One thread per destination pixel; the (source index, fp32 weight) entries of its column and row are generated on the fly.
__device__ __forceinline__ int area_range(int d, int sn, double scale, int& s1, int& s2, float& a_lead, float& a_full, float& a_trail)
{
    const double f1 = d * scale, f2 = f1 + scale, cell = fmin(scale, (double)sn - f1);
    s1 = (int)ceil(f1); s2 = (int)floor(f2);
    if (s2 > sn - 1) s2 = sn - 1;
    if (s1 > s2) s1 = s2;
    a_lead = (s1 - f1 > 1e-3) ? (float)((s1 - f1) / cell) : -1.f;          // < 0: no leading partial cell
    a_full = (float)(1.0 / cell);
    a_trail = (f2 - s2 > 1e-3) ? (float)(fmin(fmin(f2 - s2, 1.0), cell) / cell) : -1.f;
    return 0;
}

__global__ void __launch_bounds__(256)
ingest_bgr_area_kernel(const uint8_t* __restrict__ bgr, size_t step, size_t fstride, int sw, int sh, uint8_t* __restrict__ gray,
                       size_t gstep, size_t gstride, int dw, int dh, double scale_x, double scale_y, int ix, int iy, int legacy14)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    bgr += (size_t)blockIdx.z * fstride;
    gray += (size_t)blockIdx.z * gstride;
    int px[3];
    if (ix > 0) {                                   // integer ratios: OpenCV's "area fast" path
        const float sc = 1.f / (float)(ix * iy);
        int sum[3] = {0, 0, 0};
        for (int j = 0; j < iy; j++) {
            const uint8_t* row = bgr + (size_t)(y * iy + j) * step + 3 * (x * ix);
            for (int i = 0; i < ix; i++) { sum[0] += row[3 * i]; sum[1] += row[3 * i + 1]; sum[2] += row[3 * i + 2]; }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            int v = (ix == 2 && iy == 2) ? (sum[c] + 2) >> 2 : __float2int_rn((float)sum[c] * sc);
            px[c] = v > 255 ? 255 : v;
        }
    } else {
        int xs1, xs2, ys1, ys2; float xl, xf, xt, yl, yf, yt;
        area_range(x, sw, scale_x, xs1, xs2, xl, xf, xt);
        area_range(y, sh, scale_y, ys1, ys2, yl, yf, yt);
        float sum[3] = {0.f, 0.f, 0.f};
        bool first = true;
        auto do_row = [&](int sy, float beta) {
            const uint8_t* row = bgr + (size_t)sy * step;
            float buf[3] = {0.f, 0.f, 0.f};
            if (xl >= 0.f) { const uint8_t* p = row + 3 * (xs1 - 1); buf[0] = buf[0] + (float)p[0] * xl; buf[1] = buf[1] + (float)p[1] * xl; buf[2] = buf[2] + (float)p[2] * xl; }
            for (int sx = xs1; sx < xs2; sx++) { const uint8_t* p = row + 3 * sx; buf[0] = buf[0] + (float)p[0] * xf; buf[1] = buf[1] + (float)p[1] * xf; buf[2] = buf[2] + (float)p[2] * xf; }
            if (xt >= 0.f) { const uint8_t* p = row + 3 * xs2; buf[0] = buf[0] + (float)p[0] * xt; buf[1] = buf[1] + (float)p[1] * xt; buf[2] = buf[2] + (float)p[2] * xt; }
#pragma unroll
            for (int c = 0; c < 3; c++) sum[c] = first ? beta * buf[c] : sum[c] + beta * buf[c];
            first = false;
        };
        if (yl >= 0.f) do_row(ys1 - 1, yl);
        for (int sy = ys1; sy < ys2; sy++) do_row(sy, yf);
        if (yt >= 0.f) do_row(ys2, yt);
#pragma unroll
        for (int c = 0; c < 3; c++) { int v = __float2int_rn(sum[c]); px[c] = v < 0 ? 0 : (v > 255 ? 255 : v); }
    }
    const int g = legacy14 ? (px[0] * 1868 + px[1] * 9617 + px[2] * 4899 + (1 << 13)) >> 14
                           : (px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15;
    gray[(size_t)y * gstep + x] = (uint8_t)g;
}

// Mask clean-up: create_edges, ripcurrents_module.cpp:216-220 -- dilate with OpenCV's
// 5x5 ellipse, then morphological gradient (dilate - erode) with the same element; taps outside the image are ignored.
// One 32x32 tile per CTA, all stages in shared memory (halo 4), any number of masks per launch (blockIdx.z).
// The ellipse is a 5-wide x 3-tall rectangle plus the centre column of rows +-2, so each max / min is a 5-wide row pass
// followed by three rows of it and two centre taps: 8 compares per pixel and stage instead of 21.
__device__ __forceinline__ int max5(const uint8_t* p) { return max(max(max((int)p[0], (int)p[1]), max((int)p[2], (int)p[3])), (int)p[4]); }
__device__ __forceinline__ int min5(const uint8_t* p) { return min(min(min((int)p[0], (int)p[1]), min((int)p[2], (int)p[3])), (int)p[4]); }

__global__ void __launch_bounds__(256)
edges_kernel(const uint8_t* __restrict__ mask, size_t step, size_t stride, int w, int h, uint8_t* __restrict__ out,
             size_t ostep, size_t ostride)
{
    __shared__ uint8_t sIn[40][40];                    // mask on (y0-4.., x0-4..), 0 outside the image
    __shared__ uint8_t sH[38][36];                     // 5-wide row maximum of sIn: rows y0-3.., columns x0-2..
    __shared__ uint8_t sDx[36][36], sDn[36][36];       // dilated mask on (y0-2.., x0-2..); outside the image 0 / 255
    __shared__ uint8_t sHx[34][32], sHn[34][32];       // 5-wide row maximum of sDx / minimum of sDn: rows y0-1.., columns x0..
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32, tid = threadIdx.x;
    mask += (size_t)blockIdx.z * stride;
    out += (size_t)blockIdx.z * ostride;
    for (int i = tid; i < 40 * 40; i += 256) {
        const int cy = i / 40, cx = i - cy * 40, x = x0 - 4 + cx, y = y0 - 4 + cy;
        sIn[cy][cx] = (x >= 0 && y >= 0 && x < w && y < h) ? mask[(size_t)y * step + x] : 0;
    }
    __syncthreads();
    for (int i = tid; i < 38 * 36; i += 256) {
        const int r = i / 36, c = i - r * 36;
        sH[r][c] = (uint8_t)max5(&sIn[r + 1][c]);
    }
    __syncthreads();
    for (int i = tid; i < 36 * 36; i += 256) {
        const int cy = i / 36, cx = i - cy * 36, x = x0 - 2 + cx, y = y0 - 2 + cy;
        const bool inside = x >= 0 && y >= 0 && x < w && y < h;
        const int v = max(max(max((int)sH[cy][cx], (int)sH[cy + 1][cx]), max((int)sH[cy + 2][cx], (int)sIn[cy][cx + 2])),
                          (int)sIn[cy + 4][cx + 2]);
        sDx[cy][cx] = inside ? (uint8_t)v : 0;
        sDn[cy][cx] = inside ? (uint8_t)v : 255;
    }
    __syncthreads();
    for (int i = tid; i < 34 * 32; i += 256) {
        const int r = i >> 5, c = i & 31;
        sHx[r][c] = (uint8_t)max5(&sDx[r + 1][c]);
        sHn[r][c] = (uint8_t)min5(&sDn[r + 1][c]);
    }
    __syncthreads();
    for (int i = tid; i < 32 * 32; i += 256) {
        const int cy = i >> 5, cx = i & 31, x = x0 + cx, y = y0 + cy;
        if (x >= w || y >= h) continue;
        const int dd = max(max(max((int)sHx[cy][cx], (int)sHx[cy + 1][cx]), max((int)sHx[cy + 2][cx], (int)sDx[cy][cx + 2])),
                           (int)sDx[cy + 4][cx + 2]);
        const int de = min(min(min((int)sHn[cy][cx], (int)sHn[cy + 1][cx]), min((int)sHn[cy + 2][cx], (int)sDn[cy][cx + 2])),
                           (int)sDn[cy + 4][cx + 2]);
        out[(size_t)y * ostep + x] = (uint8_t)(dd - de);
    }
}

}  // namespace

void rc_launch_edges(rc_ctx* c, const uint8_t* mask, size_t step, size_t stride, int w, int h, uint8_t* out, size_t ostep,
                     size_t ostride, int nb)
{
    dim3 g((w + 31) / 32, (h + 31) / 32, nb);
    KScope ks(c, K_MISC, 2.0 * w * h * nb);
    edges_kernel<<<g, 256, 0, c->stream>>>(mask, step, stride, w, h, out, ostep, ostride);
}

void rc_launch_ingest_bgr(rc_ctx* c, const uint8_t* bgr, size_t step, size_t fstride, int sw, int sh, uint8_t* gray,
                          size_t gstep, size_t gstride, int dw, int dh, int nb, int legacy14, int area)
{
    dim3 g((dw + 31) / 32, (dh + 7) / 8, nb);
    KScope ks(c, K_MISC, (3.0 * sw * sh + (double)dw * dh) * nb);
    if (area) {
        const double fx = (double)sw / (double)dw, fy = (double)sh / (double)dh;
        const int ix = (int)nearbyint(fx), iy = (int)nearbyint(fy);
        const bool fast = fabs(fx - ix) < 2.220446049250313e-16 && fabs(fy - iy) < 2.220446049250313e-16;
        ingest_bgr_area_kernel<<<g, 256, 0, c->stream>>>(bgr, step, fstride, sw, sh, gray, gstep, gstride, dw, dh, fx, fy,
                                                        fast ? ix : 0, fast ? iy : 0, legacy14);
        return;
    }
    ingest_bgr_kernel<<<g, 256, 0, c->stream>>>(bgr, step, fstride, sw, sh, gray, gstep, gstride, dw, dh,
                                               (double)sw / (double)dw, (double)sh / (double)dh, legacy14);
}

void rc_launch_hist_polar(rc_ctx* c, const float* polar, size_t step, int w, int h, unsigned long long* hist2d)
{
    const size_t n = (size_t)w * h;
    size_t g = (n + 255) / 256; if (g > 148 * 4) g = 148 * 4;
    KScope ks(c, K_POLAR_HIST, 12.0 * n);
    hist_polar_kernel<<<(unsigned)g, 256, 0, c->stream>>>(polar, step, w, h, hist2d);
}

void rc_launch_create_flow(rc_ctx* c, float* cur, size_t cstep, float* wc, size_t wstep, float* acc2, size_t astep, int w,
                           int h, float UPPER, float MID, float LOWER, const float* d_upper2d)
{
    dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
    KScope ks(c, K_MISC, 48.0 * w * h);
    create_flow_kernel<<<g, b, 0, c->stream>>>(cur, cstep, wc, wstep, acc2, astep, w, h, UPPER, MID, LOWER, d_upper2d);
}

void rc_launch_accumulate(rc_ctx* c, float* acc, size_t astep, const float* acc2, size_t a2step, float* out, size_t ostep,
                          unsigned char* mask, size_t mstep, int w, int h, int framecount)
{
    dim3 b(32, 8), g((w + 31) / 32, (h + 7) / 8);
    KScope ks(c, K_MISC, 49.0 * w * h);
    accumulate_kernel<<<g, b, 0, c->stream>>>(acc, astep, acc2, a2step, out, ostep, mask, mstep, w, h, framecount);
}
