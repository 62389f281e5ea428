// Derived fields of the per-pixel particle state: what ripcurrents.cpp:231-279 and
// ripcurrents_module.cpp:13-59 compute from streamlines_mat (displacement, CV_32FC2) and streamlines_distance
// (path length, CV_32FC1) with split/magnitude/minMaxLoc/convertTo/applyColorMap/divide and the position scatter.
//
// The reference makes ~12 passes over image-sized Mats (split, magnitude, three minMaxLoc, divide, three convertTo,
// three applyColorMap).  Here: one reduction pass (three maxima, one read of 12 B/px) and one output pass that
// recomputes the fields from the same 12 B/px and writes the three colour images; the maxima never leave the device.
// Arithmetic follows oracle/fields_oracle.c operation by operation (compiled with -fmad=false): bit-exact.
#include "rc_internal.h"

namespace {

// order-preserving map float -> unsigned (0 is reserved for "no element yet"); NaNs are never encoded
__device__ __forceinline__ unsigned enc_f(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned e)
{
    if (e == 0) return __int_as_float(0x7fc00000);                       // no non-NaN element: NaN, as cv2 does
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// applyColorMap(COLORMAP_JET) for one level, packed b | g << 8 | r << 16 (closed form: see the oracle header)
__device__ __forceinline__ unsigned jet_bgr(int i)
{
    unsigned out = 0;
#pragma unroll
    for (int c = 1; c <= 3; c++) {
        const int base = 382 - abs(4 * i - 255 * c);
        int v = min(max(base + (base & 1), 0), 255);
        if (c == 1 && i == 159) v = 1;
        out |= (unsigned)v << (8 * (c - 1));
    }
    return out;
}

// convertTo(CV_8UC1, alpha): cvRound(src * (float)alpha) saturated; cvtss2si semantics for NaN / out of range
__device__ __forceinline__ int convert_u8(float v, float alpha)
{
    const float t = __fmul_rn(v, alpha);
    if (!(t >= -2147483648.f && t < 2147483648.f)) return 0;
    return min(max(__float2int_rn(t), 0), 255);
}

struct FieldsArgs {
    const float* field;      // displacement (x, y) per pixel, or null
    const float* src0;       // channel 0 when field is null
    const float* dist;       // channel 1 (path length / divisor), or null
    size_t n;
    int div0_zero;           // OpenCV 3.x divide: zero divisor -> 0
    int want;                // bit k: channel k (0 displacement length, 1 path length, 2 ratio) has an output
    float* mag_out;          // optional copies of channel 0 and channel 2
    float* ratio_out;
    unsigned* maxenc;        // [3]
    uint8_t* gray[3];
    uint8_t* bgr[3];
};

__device__ __forceinline__ void channels(const FieldsArgs& a, size_t i, float v[3])
{
    if (a.field) {
        const float2 f = reinterpret_cast<const float2*>(a.field)[i];
        v[0] = __fsqrt_rn(__fadd_rn(__fmul_rn(f.x, f.x), __fmul_rn(f.y, f.y)));           // magnitude(x, y)
    } else {
        v[0] = a.src0 ? a.src0[i] : 0.f;
    }
    v[1] = a.dist ? a.dist[i] : 0.f;
    v[2] = (a.div0_zero && v[1] == 0.f) ? 0.f : __fdiv_rn(v[0], v[1]);                    // divide(a, b)
}

// 4 consecutive pixels with 16-byte loads (all pointers 16-byte aligned, i a multiple of 4, i + 4 <= n)
__device__ __forceinline__ void channels4(const FieldsArgs& a, size_t i, float v[4][3])
{
    float c0[4], c1[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.field) {
        const float4 p = reinterpret_cast<const float4*>(a.field)[i / 2], q = reinterpret_cast<const float4*>(a.field)[i / 2 + 1];
        c0[0] = __fsqrt_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)));
        c0[1] = __fsqrt_rn(__fadd_rn(__fmul_rn(p.z, p.z), __fmul_rn(p.w, p.w)));
        c0[2] = __fsqrt_rn(__fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y)));
        c0[3] = __fsqrt_rn(__fadd_rn(__fmul_rn(q.z, q.z), __fmul_rn(q.w, q.w)));
    } else if (a.src0) {
        const float4 p = reinterpret_cast<const float4*>(a.src0)[i / 4];
        c0[0] = p.x; c0[1] = p.y; c0[2] = p.z; c0[3] = p.w;
    } else {
        c0[0] = c0[1] = c0[2] = c0[3] = 0.f;
    }
    if (a.dist) {
        const float4 p = reinterpret_cast<const float4*>(a.dist)[i / 4];
        c1[0] = p.x; c1[1] = p.y; c1[2] = p.z; c1[3] = p.w;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k][0] = c0[k]; v[k][1] = c1[k];
        v[k][2] = (a.div0_zero && c1[k] == 0.f) ? 0.f : __fdiv_rn(c0[k], c1[k]);
    }
}

// pass 1: maxima of the wanted channels (minMaxLoc), optional copies of the magnitude and the ratio
template <bool VEC>
__global__ void __launch_bounds__(256)
fields_max_kernel(FieldsArgs a)
{
    unsigned m[3] = {0, 0, 0};
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    auto take = [&](const float v[3]) {
#pragma unroll
        for (int c = 0; c < 3; c++)
            if (v[c] == v[c]) m[c] = max(m[c], enc_f(v[c]));
    };
    size_t done = 0;
    if (VEC) {
        const size_t n4 = a.n / 4;
        for (size_t q = tid; q < n4; q += nthr) {
            float v[4][3];
            channels4(a, q * 4, v);
#pragma unroll
            for (int k = 0; k < 4; k++) take(v[k]);
            if (a.mag_out) reinterpret_cast<float4*>(a.mag_out)[q] = make_float4(v[0][0], v[1][0], v[2][0], v[3][0]);
            if (a.ratio_out) reinterpret_cast<float4*>(a.ratio_out)[q] = make_float4(v[0][2], v[1][2], v[2][2], v[3][2]);
        }
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < a.n; i += nthr) {
        float v[3];
        channels(a, i, v);
        take(v);
        if (a.mag_out) a.mag_out[i] = v[0];
        if (a.ratio_out) a.ratio_out[i] = v[2];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        if (!((a.want >> c) & 1)) continue;
        const unsigned r = __reduce_max_sync(0xffffffffu, m[c]);
        if ((threadIdx.x & 31) == 0 && r) atomicMax(&a.maxenc[c], r);
    }
}

// pass 2: convertTo(CV_8UC1, 255 / max) + applyColorMap(JET) of every wanted channel
template <bool VEC>
__global__ void __launch_bounds__(256)
fields_color_kernel(FieldsArgs a)
{
    float alpha[3];
#pragma unroll
    for (int c = 0; c < 3; c++) alpha[c] = (float)(255.0 / (double)dec_f(a.maxenc[c]));
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (VEC) {
        const size_t n4 = a.n / 4;
        for (size_t q = tid; q < n4; q += nthr) {
            float v[4][3];
            channels4(a, q * 4, v);
#pragma unroll
            for (int c = 0; c < 3; c++) {
                if (!((a.want >> c) & 1)) continue;
                int g[4];
#pragma unroll
                for (int k = 0; k < 4; k++) g[k] = convert_u8(v[k][c], alpha[c]);
                if (a.gray[c]) reinterpret_cast<unsigned*>(a.gray[c])[q] = g[0] | g[1] << 8 | g[2] << 16 | g[3] << 24;
                if (a.bgr[c]) {
                    const unsigned p0 = jet_bgr(g[0]), p1 = jet_bgr(g[1]), p2 = jet_bgr(g[2]), p3 = jet_bgr(g[3]);
                    unsigned* o = reinterpret_cast<unsigned*>(a.bgr[c]) + q * 3;
                    o[0] = p0 | p1 << 24; o[1] = p1 >> 8 | p2 << 16; o[2] = p2 >> 16 | p3 << 8;
                }
            }
        }
        done = n4 * 4;
    }
    for (size_t i = done + tid; i < a.n; i += nthr) {
        float v[3];
        channels(a, i, v);
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (!((a.want >> c) & 1)) continue;
            const int g = convert_u8(v[c], alpha[c]);
            if (a.gray[c]) a.gray[c][i] = (uint8_t)g;
            if (a.bgr[c]) {
                const unsigned p = jet_bgr(g);
                uint8_t* o = a.bgr[c] + 3 * i;
                o[0] = (uint8_t)p; o[1] = (uint8_t)(p >> 8); o[2] = (uint8_t)(p >> 16);
            }
        }
    }
}

__global__ void decode_max_kernel(const unsigned* __restrict__ enc, double* __restrict__ out)
{
    if (threadIdx.x < 3) out[threadIdx.x] = (double)dec_f(enc[threadIdx.x]);
}

// streamline_positions (ripcurrents_module.cpp:44-59): every particle marks the pixel it ended up on.  All writers of a
// pixel store the same value, so the scatter needs no atomics.
__global__ void __launch_bounds__(256)
positions_kernel(const float* __restrict__ field, int w, int h, float* __restrict__ density)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const float2 p = reinterpret_cast<const float2*>(field)[(size_t)y * w + x];
    const float fx = floorf(__fadd_rn(p.x, (float)x)), fy = floorf(__fadd_rn(p.y, (float)y));
    if (!(fx >= 1.f && fy >= 1.f && fx <= (float)(w - 2) && fy <= (float)(h - 2))) return;   // also NaN / inf
    float* d = density + 3 * ((size_t)(int)fy * w + (int)fx);
    d[0] = 1.f; d[1] = 1.f; d[2] = 1.f;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// All pointers are device pointers; d_maxenc holds 3 unsigned, d_max (optional) receives the 3 maxima as doubles.
void rc_launch_fields(rc_ctx* c, const float* field, const float* src0, const float* dist, size_t n, int div0_zero, int want,
                      float* mag_out, float* ratio_out, uint8_t* const gray[3], uint8_t* const bgr[3], unsigned* d_maxenc,
                      double* d_max)
{
    FieldsArgs a;
    a.field = field; a.src0 = src0; a.dist = dist; a.n = n; a.div0_zero = div0_zero; a.want = want;
    a.mag_out = mag_out; a.ratio_out = ratio_out; a.maxenc = d_maxenc;
    bool vec = al16(field) && al16(src0) && al16(dist) && al16(mag_out) && al16(ratio_out);
    bool color = false;
    for (int k = 0; k < 3; k++) {
        a.gray[k] = gray[k]; a.bgr[k] = bgr[k];
        vec = vec && (reinterpret_cast<uintptr_t>(gray[k]) & 3) == 0 && (reinterpret_cast<uintptr_t>(bgr[k]) & 3) == 0;
        color = color || gray[k] || bgr[k];
    }
    const int chans = (want & 1) + ((want >> 1) & 1) + ((want >> 2) & 1);
    const double in_bytes = (field ? 8.0 : src0 ? 4.0 : 0.0) + (dist ? 4.0 : 0.0);
    double out_bytes = 0;
    for (int k = 0; k < 3; k++) out_bytes += (gray[k] ? 1.0 : 0.0) + (bgr[k] ? 3.0 : 0.0);
    const int grid = 148 * 8;
    cudaMemsetAsync(d_maxenc, 0, 3 * sizeof(unsigned), c->stream);
    {
        KScope ks(c, K_FIELDS, n * (in_bytes + (mag_out ? 4.0 : 0.0) + (ratio_out ? 4.0 : 0.0)));
        if (vec) fields_max_kernel<true><<<grid, 256, 0, c->stream>>>(a);
        else fields_max_kernel<false><<<grid, 256, 0, c->stream>>>(a);
    }
    if (color && chans) {
        KScope ks(c, K_FIELDS, n * (in_bytes + out_bytes));
        if (vec) fields_color_kernel<true><<<grid, 256, 0, c->stream>>>(a);
        else fields_color_kernel<false><<<grid, 256, 0, c->stream>>>(a);
    }
    if (d_max) {
        KScope ks(c, K_FIELDS, 0);
        decode_max_kernel<<<1, 32, 0, c->stream>>>(d_maxenc, d_max);
    }
}

void rc_launch_positions(rc_ctx* c, const float* field, int w, int h, float* density, int zero_first)
{
    if (zero_first) cudaMemsetAsync(density, 0, (size_t)w * h * 12, c->stream);
    KScope ks(c, K_FIELDS, (size_t)w * h * (8.0 + 12.0 + (zero_first ? 12.0 : 0.0)));
    positions_kernel<<<dim3((w + 255) / 256, h), 256, 0, c->stream>>>(field, w, h, density);
}
