// Flow-derived diagnostics: subtructMeanMagnitude, vectorToColor, shearRateToColor
// (ripcurrents_module.cpp:900-1138).  One thread per pixel, one pass each: the reference normalises with the PREVIOUS
// frame's maximum (a function-local static), so the new maximum is reduced in the same pass (warp redux + one atomicMax
// per warp on the float bits, valid for non-negative values; NaNs never win, as with the reference's `>` tests).
// Arithmetic follows oracle/diag_oracle.c operation by operation (compiled with -fmad=false; the two fused
// multiply-adds of cv2's HSV->BGR are explicit): bit-exact, including atan2f (atan2f_ref.h) and the x86 float->uchar
// store semantics.
#include "rc_internal.h"
#include "atan2f_ref.h"

namespace {

// `uchar = float` as GCC/x86-64 compiles it: cvttss2si, low byte
__device__ __forceinline__ unsigned to_uchar(float f)
{
    const int r = (f >= -2147483648.f && f < 2147483648.f) ? __float2int_rz(f) : (int)0x80000000;
    return (unsigned)r & 0xffu;
}

// cvtColor(COLOR_HSV2BGR), 8-bit, cv2's block path (oracle/diag_oracle.c); returns b | g << 8 | r << 16
__device__ __forceinline__ unsigned hsv2bgr(unsigned hb, unsigned sb, unsigned vb, bool fma)
{
    float h = __fmul_rn((float)hb, 6.f / 180.f);
    const float s = __fmul_rn((float)sb, 1.f / 255.f), v = __fmul_rn((float)vb, 1.f / 255.f);
    while (h >= 6.f) h = __fsub_rn(h, 6.f);
    int sec = (int)floorf(h);
    float f = __fsub_rn(h, (float)sec);
    if ((unsigned)sec >= 6u) { sec = 0; f = 0.f; }
    const float omf = __fsub_rn(1.f, f);
    const float q2 = fma ? __fmaf_rn(-s, f, 1.f) : __fsub_rn(1.f, __fmul_rn(s, f));
    const float q3 = fma ? __fmaf_rn(-s, omf, 1.f) : __fsub_rn(1.f, __fmul_rn(s, omf));
    const float t0 = v, t1 = __fmul_rn(v, __fsub_rn(1.f, s)), t2 = __fmul_rn(v, q2), t3 = __fmul_rn(v, q3);
    // sector table {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} as selects
    const float b = sec == 0 || sec == 1 ? t1 : sec == 2 ? t3 : sec == 5 ? t2 : t0;
    const float g = sec == 0 ? t3 : sec == 1 || sec == 2 ? t0 : sec == 3 ? t2 : t1;
    const float r = sec == 0 || sec == 5 ? t0 : sec == 1 ? t2 : sec == 2 || sec == 3 ? t1 : t3;
    auto q = [](float t) { return (unsigned)min(max(__float2int_rz(__fmul_rn(t, 255.f)), 0), 255); };
    return q(b) | q(g) << 8 | q(r) << 16;
}

__device__ __forceinline__ float magnitude(float x, float y)
{
    return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
}

__device__ __forceinline__ const float2* flow_row(const float* flow, size_t step, int y)
{
    return reinterpret_cast<const float2*>(reinterpret_cast<const char*>(flow) + (size_t)y * step);
}

__device__ __forceinline__ void warp_max_nonneg(float m, unsigned* dst)
{
    const unsigned r = __reduce_max_sync(0xffffffffu, __float_as_uint(m));      // m >= +0, never NaN
    if ((threadIdx.x & 31) == 0 && r) atomicMax(dst, r);
}

// ---- subtructMeanMagnitude ---------------------------------------------------------------------------------------
// default: magnitudes summed in fp64, fixed two-stage order (deterministic); the reference's loop accumulates 2 M
// values sequentially in fp32, which no parallel order reproduces -- `sequential` does exactly that with one thread
// (slow, for verification).
__global__ void __launch_bounds__(256)
mag_sum_partial_kernel(const float* __restrict__ flow, size_t step, int w, int h, double* __restrict__ partial)
{
    __shared__ double sh[256];
    double s = 0;
    const size_t n = (size_t)w * h;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
        const float2 f = flow_row(flow, step, y)[x];
        s += (double)magnitude(f.x, f.y);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void mag_sum_final_kernel(const double* __restrict__ partial, int nparts, int n, float* __restrict__ meanval)
{
    __shared__ double sh[256];
    double s = 0;
    for (int i = threadIdx.x; i < nparts; i += 256) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *meanval = (float)(sh[0] / (double)n);
}

__global__ void __launch_bounds__(256)
mag_sum_sequential_kernel(const float* __restrict__ flow, size_t step, int w, int h, float* __restrict__ meanval)
{
    constexpr int CH = 4096;
    __shared__ float sh[CH];
    const size_t n = (size_t)w * h;
    float acc = 0.f;
    for (size_t base = 0; base < n; base += CH) {
        for (int k = threadIdx.x; k < CH && base + k < n; k += 256) {
            const size_t i = base + k;
            const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
            const float2 f = flow_row(flow, step, y)[x];
            sh[k] = magnitude(f.x, f.y);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const int m = (int)min((size_t)CH, n - base);
            for (int k = 0; k < m; k++) acc = __fadd_rn(acc, sh[k]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *meanval = __fdiv_rn(acc, (float)(int)n);
}

__global__ void __launch_bounds__(256)
sub_mean_mag_kernel(float* __restrict__ flow, size_t step, int w, int h, const float* __restrict__ meanval)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    float2* row = reinterpret_cast<float2*>(reinterpret_cast<char*>(flow) + (size_t)y * step);
    const float2 f = row[x];
    const float mag = magnitude(f.x, f.y), mv = *meanval;
    float ux = 0.f, uy = 0.f;
    if (mag != 0.f) { ux = __fdiv_rn(f.x, mag); uy = __fdiv_rn(f.y, mag); }
    const float d = __fsub_rn(mag, mv);
    row[x] = make_float2(__fmul_rn(ux, d), __fmul_rn(uy, d));
}

// ---- vectorToColor -----------------------------------------------------------------------------------------------
__device__ __forceinline__ float theta_deg(float y, float x)
{
    float theta = (float)(((double)rc_atan2f_ref(y, x) * 180.0) / 3.14159265358979323846);
    if (theta < 0) theta = __fadd_rn(theta, 360.f);
    return theta;
}

// One pixel of vectorToColor: packed b | g << 8 | r << 16; *mag receives |v| (0 for NaN, which never raises the maximum)
__device__ __forceinline__ unsigned vector_px(float2 f, float prev_max, bool fma, float* mag)
{
    const float m = magnitude(f.x, f.y);
    const unsigned hue = to_uchar(__fdiv_rn(theta_deg(f.y, f.x), 2.f));
    const unsigned val = to_uchar(__fdiv_rn(__fmul_rn(m, 255.f), prev_max));
    *mag = m > 0.f ? m : 0.f;
    return hsv2bgr(hue, 255u, val, fma);
}

// PX = 4: four adjacent pixels per thread, two 16-byte flow loads and three 32-bit stores (w % 4 == 0, 4-byte aligned
// image rows); PX = 1: any geometry.
template <int PX>
__global__ void __launch_bounds__(256)
vector_color_kernel(const float* __restrict__ flow, size_t step, int w, int h, uint8_t* __restrict__ bgr, size_t bstep,
                    const float* __restrict__ prev_max, unsigned* __restrict__ new_max, int fma)
{
    const float pm = *prev_max;
    const int x = (blockIdx.x * 256 + threadIdx.x) * PX, y = blockIdx.y;
    float best = 0.f;
    if (x < w) {
        uint8_t* o = bgr + (size_t)y * bstep + 3 * (size_t)x;
        if (PX == 4) {
            const float4* fp = reinterpret_cast<const float4*>(flow_row(flow, step, y) + x);
            const float4 a = fp[0], b = fp[1];
            float m0, m1, m2, m3;
            const unsigned p0 = vector_px(make_float2(a.x, a.y), pm, fma != 0, &m0), p1 = vector_px(make_float2(a.z, a.w), pm, fma != 0, &m1);
            const unsigned p2 = vector_px(make_float2(b.x, b.y), pm, fma != 0, &m2), p3 = vector_px(make_float2(b.z, b.w), pm, fma != 0, &m3);
            unsigned* o4 = reinterpret_cast<unsigned*>(o);
            o4[0] = p0 | p1 << 24; o4[1] = p1 >> 8 | p2 << 16; o4[2] = p2 >> 16 | p3 << 8;
            best = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
            const unsigned p = vector_px(flow_row(flow, step, y)[x], pm, fma != 0, &best);
            o[0] = (uint8_t)p; o[1] = (uint8_t)(p >> 8); o[2] = (uint8_t)(p >> 16);
        }
    }
    warp_max_nonneg(best, new_max);
}

// ---- shearRateToColor --------------------------------------------------------------------------------------------
// One pixel: interior pixels get (hue, 255, 255) from the Frobenius norm of the Jacobian, border pixels keep the caller's
// bytes `old`; both are then converted HSV -> BGR, as the reference's full-image cvtColor does.
__device__ __forceinline__ unsigned shear_px(const float* __restrict__ flow, size_t step, int w, int h, int x, int y,
                                             unsigned old, float prev_max, bool fma, float* frob_out)
{
    constexpr int OFF = 10;
    unsigned hh = old & 0xffu, ss = (old >> 8) & 0xffu, vv = (old >> 16) & 0xffu;
    float frob = 0.f;
    if (x >= OFF && x < w - OFF && y >= OFF && y < h - OFF) {
        const float2 above = flow_row(flow, step, y - OFF)[x], below = flow_row(flow, step, y + OFF)[x];
        const float2 left = flow_row(flow, step, y)[x - OFF], right = flow_row(flow, step, y)[x + OFF];
        const float j00 = __fsub_rn(right.x, left.x), j01 = __fsub_rn(above.x, below.x);
        const float j10 = __fsub_rn(right.y, left.y), j11 = __fsub_rn(above.y, below.y);
        float fr = __fadd_rn(__fmul_rn(j00, j00), __fmul_rn(j01, j01));
        fr = __fadd_rn(fr, __fmul_rn(j10, j10));
        fr = __fadd_rn(fr, __fmul_rn(j11, j11));
        frob = __fsqrt_rn(fr);
        hh = to_uchar(__fsub_rn(128.f, __fdiv_rn(__fmul_rn(frob, 128.f), prev_max)));
        ss = 255u; vv = 255u;
        if (!(frob > 0.f)) frob = 0.f;
    }
    *frob_out = frob;
    return hsv2bgr(hh, ss, vv, fma);
}

template <int PX>
__global__ void __launch_bounds__(256)
shear_color_kernel(const float* __restrict__ flow, size_t step, int w, int h, uint8_t* __restrict__ img, size_t istep,
                   const float* __restrict__ prev_max, unsigned* __restrict__ new_max, int fma)
{
    constexpr int OFF = 10;
    const float pm = *prev_max;
    const int x = (blockIdx.x * 256 + threadIdx.x) * PX, y = blockIdx.y;
    float best = 0.f;
    if (x < w) {
        uint8_t* o = img + (size_t)y * istep + 3 * (size_t)x;
        if (PX == 4) {
            unsigned* o4 = reinterpret_cast<unsigned*>(o);
            unsigned old[4] = {0, 0, 0, 0};
            if (!(x >= OFF && x + 3 < w - OFF && y >= OFF && y < h - OFF)) {      // some of the four are border pixels
                const unsigned w0 = o4[0], w1 = o4[1], w2 = o4[2];
                old[0] = w0 & 0xffffffu; old[1] = (w0 >> 24) | (w1 & 0xffffu) << 8;
                old[2] = (w1 >> 16) | (w2 & 0xffu) << 16; old[3] = w2 >> 8;
            }
            unsigned p[4]; float m[4];
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = shear_px(flow, step, w, h, x + i, y, old[i], pm, fma != 0, &m[i]);
            o4[0] = p[0] | p[1] << 24; o4[1] = p[1] >> 8 | p[2] << 16; o4[2] = p[2] >> 16 | p[3] << 8;
            best = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
        } else {
            const unsigned old = (unsigned)o[0] | (unsigned)o[1] << 8 | (unsigned)o[2] << 16;
            const unsigned p = shear_px(flow, step, w, h, x, y, old, pm, fma != 0, &best);
            o[0] = (uint8_t)p; o[1] = (uint8_t)(p >> 8); o[2] = (uint8_t)(p >> 16);
        }
    }
    warp_max_nonneg(best, new_max);
}

}  // namespace

void rc_launch_sub_mean_magnitude(rc_ctx* c, float* flow, size_t step, int w, int h, int sequential, double* d_partial,
                                  float* d_meanval)
{
    const double n = (double)w * h;
    {
        KScope ks(c, K_DIAG, 8.0 * n, 2);
        if (sequential) {
            mag_sum_sequential_kernel<<<1, 256, 0, c->stream>>>(flow, step, w, h, d_meanval);
        } else {
            const int parts = 148 * 4;
            mag_sum_partial_kernel<<<parts, 256, 0, c->stream>>>(flow, step, w, h, d_partial);
            mag_sum_final_kernel<<<1, 256, 0, c->stream>>>(d_partial, parts, w * h, d_meanval);
        }
    }
    KScope ks(c, K_DIAG, 16.0 * n);
    sub_mean_mag_kernel<<<dim3((w + 255) / 256, h), 256, 0, c->stream>>>(flow, step, w, h, d_meanval);
}

void rc_launch_vector_color(rc_ctx* c, const float* flow, size_t step, int w, int h, uint8_t* bgr, size_t bstep,
                            const float* d_prev_max, unsigned* d_new_max, int fma)
{
    cudaMemsetAsync(d_new_max, 0, 4, c->stream);
    KScope ks(c, K_DIAG, 11.0 * w * h);
    const bool v4 = w % 4 == 0 && bstep % 4 == 0 && step % 16 == 0 && (reinterpret_cast<uintptr_t>(bgr) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(flow) & 15) == 0;
    if (v4) vector_color_kernel<4><<<dim3((w / 4 + 255) / 256, h), 256, 0, c->stream>>>(flow, step, w, h, bgr, bstep, d_prev_max, d_new_max, fma);
    else vector_color_kernel<1><<<dim3((w + 255) / 256, h), 256, 0, c->stream>>>(flow, step, w, h, bgr, bstep, d_prev_max, d_new_max, fma);
}

void rc_launch_shear_color(rc_ctx* c, const float* flow, size_t step, int w, int h, uint8_t* img, size_t istep,
                           const float* d_prev_max, unsigned* d_new_max, int fma)
{
    cudaMemsetAsync(d_new_max, 0, 4, c->stream);
    KScope ks(c, K_DIAG, 11.0 * w * h);
    const bool v4 = w % 4 == 0 && istep % 4 == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0;
    if (v4) shear_color_kernel<4><<<dim3((w / 4 + 255) / 256, h), 256, 0, c->stream>>>(flow, step, w, h, img, istep, d_prev_max, d_new_max, fma);
    else shear_color_kernel<1><<<dim3((w + 255) / 256, h), 256, 0, c->stream>>>(flow, step, w, h, img, istep, d_prev_max, d_new_max, fma);
}
