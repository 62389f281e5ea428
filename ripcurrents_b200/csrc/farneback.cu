// Farneback dense optical flow for sm_100a -- kernels and launchers.
//
// Functional specification: SURVEY.md Appendix A (the algorithm cv::calcOpticalFlowFarneback runs for the
// reference's calls at RipCurrents_main/ripcurrents.cpp:215 and main.cpp:264,609,742,961,1119,1481).
//
// Two arithmetic modes (DESIGN.md "Arithmetic"):
//   strict : every tap, fp64 accumulators where OpenCV has them, products and sums rounded separately
//            (written with __fmul_rn/__fadd_rn so the compiler cannot contract them);
//   fast   : (default) fp32 with FMA, Gaussian tails of the expansion kernel below 1e-9 of the centre dropped,
//            2x2 solve in fp32 with error-free determinants.  Measured as close to cv2 as the strict mode is.
//
// HBM layout: the polynomial expansion R of a frame is stored "4+1": one float4 plane {dI/dy, dI/dx, Iyy, Ixx} and
// one float plane {Ixy} (row pitch rounded up to 32 pixels), so that the data-dependent bilinear gather of
// updateMatrices costs 4 x (16 B + 4 B) loads instead of 20 scalar ones and every access of a warp is a run of
// full 32-byte sectors; the structure matrices M of the unfused path are 5 fp32 planes; flow is interleaved float2
// (CV_32FC2, the output format).  Every array has a leading batch dimension: one
// launch processes all frame pairs of a batch (blockIdx.z), which is what fills 148 SMs on the small pyramid layers.
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include "rc_internal.h"
#include <cuda.h>
#include <cuda_pipeline.h>
#include <mutex>

namespace {

template <bool S> __device__ __forceinline__ float fmul(float a, float b) { return S ? __fmul_rn(a, b) : a * b; }
template <bool S> __device__ __forceinline__ float fadd(float a, float b) { return S ? __fadd_rn(a, b) : a + b; }
template <bool S> __device__ __forceinline__ float fsub(float a, float b) { return S ? __fsub_rn(a, b) : a - b; }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cv::resize(INTER_LINEAR) source index + weight for destination index d (Appendix A.2)
__device__ __forceinline__ void resize_coef(int d, int src, double scale, int& s0, float& f)
{
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src - 1) { s = src - 1; fx = 0.f; }
    s0 = s; f = fx;
}

// ---------------------------------------------------------------------------------------------------
// Pyramid layer k of one frame (Appendix A.2): u8 -> fp32, Gaussian blur of the FULL-RESOLUTION image
// (REFLECT_101, rows then columns, fp32, taps accumulated in order with separately rounded products), bilinear
// resize to (dw, dh).  One kernel per layer: a CTA owns a TXD x TYD destination tile, stages the u8 source region
// it depends on in shared memory (reflected at the image border), blurs horizontally only at the (up to) two source
// columns each destination column samples, then vertically only at the two sampled rows, and combines.
// The staged region stays u8 (converted tap by tap in the horizontal pass): interior tiles copy it with 4-byte cp.async
// -- every word of the CTA in flight at once, no registers, no conversion pass -- and it is a quarter of the fp32 size, so
// deep layers (scale 1/8, 1/16: 17- and 39-tap windows) get 4x larger destination tiles and a third less halo.
// HBM traffic: the u8 frame once per layer (L2-resident after the first layer) + 4 B per destination pixel.
// ---------------------------------------------------------------------------------------------------
struct PyrArgs {
    const uint8_t* img; size_t step, fstride; int W, H;
    int dw, dh; double sxs, sys; int two;
    float* out; int pitch; size_t ostride;
    int TXD, TYD, SRW, SRH, SRWB;      // SRWB: row pitch of the staged region in bytes (multiple of 4, odd word count)
};

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(256)
pyr_tile_kernel(PyrArgs a, SmoothCoef sc)
{
    extern __shared__ __align__(16) unsigned char psm[];
    const int TXD = a.TXD, TYD = a.TYD, SRW = a.SRW, SRH = a.SRH, SRWB = a.SRWB;
    const int SRHP = SRH | 1;                                          // odd pitch of the transposed H-blur buffer
    const int lx = 31 - __clz(TXD);                                    // TXD is a power of two
    float2* sT = reinterpret_cast<float2*>(psm);                       // [TXD][SRHP] horizontally blurred pairs
    uint8_t* sS = reinterpret_cast<uint8_t*>(sT + (size_t)TXD * SRHP); // [SRH][SRWB] source region, u8
    int* cS = reinterpret_cast<int*>(sS + (size_t)SRH * SRWB);         // [TXD] sx, [TYD] sy
    float* cF = reinterpret_cast<float*>(cS + TXD + TYD);              // [TXD] fx, [TYD] fy
    int* gX = reinterpret_cast<int*>(cF + TXD + TYD);                  // [SRW] reflected source column of region col
    const int tid = threadIdx.x, r = sc.ksize / 2, ks = sc.ksize;
    const int lane = tid & 31, wrp = tid >> 5;
    const int X0 = blockIdx.x * TXD, Y0 = blockIdx.y * TYD;
    const uint8_t* img = a.img + (size_t)blockIdx.z * a.fstride;
    float* out = a.out + (size_t)blockIdx.z * a.ostride;

    for (int i = tid; i < TXD + TYD; i += 256) {
        int s; float f = 0.f;
        if (i < TXD) {
            const int X = min(X0 + i, a.dw - 1);
            if (a.two) resize_coef(X, a.W, a.sxs, s, f); else s = X;
        } else {
            const int Y = min(Y0 + i - TXD, a.dh - 1);
            if (a.two) resize_coef(Y, a.H, a.sys, s, f); else s = Y;
        }
        cS[i] = s; cF[i] = f;
    }
    __syncthreads();
    const int ox = cS[0] - r, oy = cS[TXD] - r;
    for (int i = tid; i < SRW; i += 256) gX[i] = reflect101(ox + i, a.W);
    __syncthreads();
    // stage the source region: warp = row.  4-byte aligned images copy whole 32-bit words asynchronously (region column 0
    // sits `sh` bytes into the row).  Tiles on the image border do the same for the part of every row that lies inside the
    // image (rows outside it copy their REFLECT_101 mirror row) and then fill the few columns outside the image from their
    // mirror columns, which are inside the staged region; unaligned images gather bytes through the reflected column table.
    const bool aligned = (a.step & 3) == 0 && (reinterpret_cast<size_t>(img) & 3) == 0 && (a.W & 3) == 0;
    const int sh = aligned ? (ox & 3) : 0;
    if (aligned) {
        const int a0 = ox & ~3;                                          // image column of the row's first word (may be < 0)
        const int w_lo = a0 < 0 ? (-a0) >> 2 : 0;                        // first / one-past-last word inside the image
        const int w_hi = min((SRW + sh + 3) >> 2, (a.W - a0) >> 2);
        for (int ry = wrp; ry < SRH; ry += 8) {
            const unsigned int* grow = reinterpret_cast<const unsigned int*>(img + (size_t)reflect101(oy + ry, a.H) * a.step) + (a0 >> 2);
            unsigned int* srow = reinterpret_cast<unsigned int*>(sS + ry * SRWB);
            for (int j = w_lo + lane; j < w_hi; j += 32) cp_async4(srow + j, grow + j);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (ox < 0 || ox + SRW > a.W) {
            __syncthreads();
            const int nl = ox < 0 ? -ox : 0, nr = ox + SRW > a.W ? ox + SRW - a.W : 0;      // columns outside, left / right
            for (int ry = wrp; ry < SRH; ry += 8) {
                uint8_t* srow = sS + ry * SRWB + sh;
                for (int i = lane; i < nl + nr; i += 32) {
                    const int cidx = i < nl ? i : SRW - nr + (i - nl);
                    srow[cidx] = srow[gX[cidx] - ox];
                }
            }
        }
    } else {
        for (int ry = wrp; ry < SRH; ry += 8) {
            const uint8_t* grow = img + (size_t)reflect101(oy + ry, a.H) * a.step;
            uint8_t* srow = sS + ry * SRWB;
            for (int rx = lane; rx < SRW; rx += 32) srow[rx] = grow[gX[rx]];
        }
    }
    __syncthreads();
    // horizontal blur at the two sampled columns of every destination column, every staged row.
    // warp = destination column, lane = staged row: the odd row pitch makes the strided reads conflict-free, and the
    // two sampled columns (adjacent except at the right border) share their taps.
    for (int tx = wrp; tx < TXD; tx += 8) {
        const int sx = cS[tx];
        const int d1 = a.two ? (sx + 1 < a.W ? sx + 1 : a.W - 1) - sx : 0;
        const int c0 = sx - r - ox;
        for (int ry = lane; ry < SRH; ry += 32) {
            const uint8_t* row = sS + ry * SRWB + sh + c0;
            float s0 = __fmul_rn(sc.k[0], (float)row[0]), s1 = 0.f;
            if (d1 == 1) {
                for (int i = 1; i < ks; i++) {
                    const float cur = (float)row[i];
                    s0 = __fadd_rn(s0, __fmul_rn(sc.k[i], cur));
                    s1 = i == 1 ? __fmul_rn(sc.k[0], cur) : __fadd_rn(s1, __fmul_rn(sc.k[i - 1], cur));
                }
                s1 = __fadd_rn(s1, __fmul_rn(sc.k[ks - 1], (float)row[ks]));
            } else {
                for (int i = 1; i < ks; i++) s0 = __fadd_rn(s0, __fmul_rn(sc.k[i], (float)row[i]));
                s1 = s0;        // identity resize, or the clamped last column (d1 == 0): both samples coincide
            }
            sT[tx * SRHP + ry] = make_float2(s0, s1);
        }
    }
    __syncthreads();
    // vertical blur at the two sampled rows + bilinear combination
    for (int idx = tid; idx < (TYD << lx); idx += 256) {
        const int ty = idx >> lx, tx = idx & (TXD - 1);
        const int X = X0 + tx, Y = Y0 + ty;
        if (X >= a.dw || Y >= a.dh) continue;
        const int sy = cS[TXD + ty];
        const float2* colp = sT + tx * SRHP + (sy - r - oy);
        float v;
        if (a.two) {
            const int d1 = (sy + 1 < a.H ? sy + 1 : a.H - 1) - sy;
            const float fx = cF[tx], fy = cF[TXD + ty];
            float2 p = colp[0];
            float b00 = __fmul_rn(sc.k[0], p.x), b01 = __fmul_rn(sc.k[0], p.y), b10, b11;
            if (d1 == 1) {
                b10 = 0.f; b11 = 0.f;
                for (int jj = 1; jj < ks; jj++) {
                    p = colp[jj];
                    b00 = __fadd_rn(b00, __fmul_rn(sc.k[jj], p.x)); b01 = __fadd_rn(b01, __fmul_rn(sc.k[jj], p.y));
                    if (jj == 1) { b10 = __fmul_rn(sc.k[0], p.x); b11 = __fmul_rn(sc.k[0], p.y); }
                    else { b10 = __fadd_rn(b10, __fmul_rn(sc.k[jj - 1], p.x)); b11 = __fadd_rn(b11, __fmul_rn(sc.k[jj - 1], p.y)); }
                }
                p = colp[ks];
                b10 = __fadd_rn(b10, __fmul_rn(sc.k[ks - 1], p.x)); b11 = __fadd_rn(b11, __fmul_rn(sc.k[ks - 1], p.y));
            } else {
                for (int jj = 1; jj < ks; jj++) {
                    p = colp[jj];
                    b00 = __fadd_rn(b00, __fmul_rn(sc.k[jj], p.x)); b01 = __fadd_rn(b01, __fmul_rn(sc.k[jj], p.y));
                }
                b10 = b00; b11 = b01;
            }
            const float top = __fadd_rn(__fmul_rn(b00, 1.f - fx), __fmul_rn(b01, fx));
            const float bot = __fadd_rn(__fmul_rn(b10, 1.f - fx), __fmul_rn(b11, fx));
            v = __fadd_rn(__fmul_rn(top, 1.f - fy), __fmul_rn(bot, fy));
        } else {
            float s = __fmul_rn(sc.k[0], colp[0].x);
            for (int jj = 1; jj < ks; jj++) s = __fadd_rn(s, __fmul_rn(sc.k[jj], colp[jj].x));
            v = s;
        }
        out[(size_t)Y * a.pitch + X] = v;
    }
}

// Exact 2:1 layer with a 3-tap presmooth (pyr_scale 0.5, layer 1): the bilinear resize samples source columns
// 2X, 2X+1 with weight 1/2 each (same for rows), so a destination pixel is a fixed 4x4 source window.  Each thread
// produces two adjacent destination pixels from four rows of (one aligned 4-byte load + two edge bytes); arithmetic
// order identical to pyr_tile_kernel (taps in order, products and sums rounded separately).
__global__ void __launch_bounds__(256)
pyr_half_kernel(const uint8_t* __restrict__ img, size_t step, size_t fstride, int W, int H, int dw, int dh, float k0,
                float k1, float k2, float* __restrict__ out, int pitch, size_t ostride)
{
    const int X = (blockIdx.x * 64 + (threadIdx.x & 63)) * 2;       // dw is even (W % 4 == 0)
    const int Y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (X >= dw || Y >= dh) return;
    img += (size_t)blockIdx.z * fstride;
    out += (size_t)blockIdx.z * ostride;
    const int x = 2 * X;                                            // source columns x-1 .. x+4
    const int xl = x > 0 ? x - 1 : 1, xr = x + 4 < W ? x + 4 : W - 2;
    float hb[4][4];                                                 // [source row 2Y-1+rr][source column x+i] after the H blur
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
        int yy = 2 * Y - 1 + rr;
        yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);
        const uint8_t* row = img + (size_t)yy * step;
        const uchar4 m = *reinterpret_cast<const uchar4*>(row + x);
        const float v[6] = {(float)row[xl], (float)m.x, (float)m.y, (float)m.z, (float)m.w, (float)row[xr]};
#pragma unroll
        for (int i = 0; i < 4; i++)
            hb[rr][i] = __fadd_rn(__fadd_rn(__fmul_rn(k0, v[i]), __fmul_rn(k1, v[i + 1])), __fmul_rn(k2, v[i + 2]));
    }
    float r[2];
#pragma unroll
    for (int d = 0; d < 2; d++) {
        float b[2][2];                                              // [sampled row][sampled column]
#pragma unroll
        for (int sr = 0; sr < 2; sr++)
#pragma unroll
            for (int sc = 0; sc < 2; sc++) {
                const int c = 2 * d + sc;
                b[sr][sc] = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[sr][c]), __fmul_rn(k1, hb[sr + 1][c])), __fmul_rn(k2, hb[sr + 2][c]));
            }
        const float top = __fadd_rn(__fmul_rn(b[0][0], 0.5f), __fmul_rn(b[0][1], 0.5f));
        const float bot = __fadd_rn(__fmul_rn(b[1][0], 0.5f), __fmul_rn(b[1][1], 0.5f));
        r[d] = __fadd_rn(__fmul_rn(top, 0.5f), __fmul_rn(bot, 0.5f));
    }
    *reinterpret_cast<float2*>(out + (size_t)Y * pitch + X) = make_float2(r[0], r[1]);
}

// Exact 4:1 layer with the 9-tap presmooth (pyr_scale 0.5, layer 2: sigma 1.5): the resize samples source columns
// 4X+1, 4X+2 (rows 4Y+1, 4Y+2) with weight 1/2, so a destination pixel is a fixed 10x10 source window.  One thread per
// destination pixel; interior pixels read three aligned 32-bit words per row, border pixels gather bytes with
// REFLECT_101.  Arithmetic order identical to pyr_tile_kernel.
__global__ void __launch_bounds__(256)
pyr_quarter_kernel(const uint8_t* __restrict__ img, size_t step, size_t fstride, int W, int H, int dw, int dh, SmoothCoef sc,
                   float* __restrict__ out, int pitch, size_t ostride)
{
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= dw || Y >= dh) return;
    img += (size_t)blockIdx.z * fstride;
    out += (size_t)blockIdx.z * ostride;
    const int cx = 4 * X - 3, cy = 4 * Y - 3;                      // window origin: columns cx..cx+9, rows cy..cy+9
    const bool inner = cx >= 1 && cx + 11 <= W && cy >= 0 && cy + 9 < H;
    float k[9];
#pragma unroll
    for (int i = 0; i < 9; i++) k[i] = sc.k[i];
    float h0[10], h1[10];                                          // H-blurred values at columns 4X+1 / 4X+2, per window row
#pragma unroll
    for (int rr = 0; rr < 10; rr++) {
        float v[10];
        if (inner) {
            const unsigned int* p = reinterpret_cast<const unsigned int*>(img + (size_t)(cy + rr) * step + (cx - 1));
            const unsigned int w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);    // bytes cx-1 .. cx+10
            v[0] = (float)((w0 >> 8) & 0xff); v[1] = (float)((w0 >> 16) & 0xff); v[2] = (float)(w0 >> 24);
            v[3] = (float)(w1 & 0xff); v[4] = (float)((w1 >> 8) & 0xff); v[5] = (float)((w1 >> 16) & 0xff); v[6] = (float)(w1 >> 24);
            v[7] = (float)(w2 & 0xff); v[8] = (float)((w2 >> 8) & 0xff); v[9] = (float)((w2 >> 16) & 0xff);
        } else {
            const uint8_t* row = img + (size_t)reflect101(cy + rr, H) * step;
#pragma unroll
            for (int i = 0; i < 10; i++) v[i] = (float)row[reflect101(cx + i, W)];
        }
        float s0 = __fmul_rn(k[0], v[0]), s1 = __fmul_rn(k[0], v[1]);
#pragma unroll
        for (int i = 1; i < 9; i++) { s0 = __fadd_rn(s0, __fmul_rn(k[i], v[i])); s1 = __fadd_rn(s1, __fmul_rn(k[i], v[i + 1])); }
        h0[rr] = s0; h1[rr] = s1;
    }
    float b00 = __fmul_rn(k[0], h0[0]), b01 = __fmul_rn(k[0], h1[0]), b10 = __fmul_rn(k[0], h0[1]), b11 = __fmul_rn(k[0], h1[1]);
#pragma unroll
    for (int j = 1; j < 9; j++) {
        b00 = __fadd_rn(b00, __fmul_rn(k[j], h0[j])); b01 = __fadd_rn(b01, __fmul_rn(k[j], h1[j]));
        b10 = __fadd_rn(b10, __fmul_rn(k[j], h0[j + 1])); b11 = __fadd_rn(b11, __fmul_rn(k[j], h1[j + 1]));
    }
    const float top = __fadd_rn(__fmul_rn(b00, 0.5f), __fmul_rn(b01, 0.5f));
    const float bot = __fadd_rn(__fmul_rn(b10, 0.5f), __fmul_rn(b11, 0.5f));
    out[(size_t)Y * pitch + X] = __fadd_rn(__fmul_rn(top, 0.5f), __fmul_rn(bot, 0.5f));
}

// Layer 0 (no resize; OpenCV's 3-tap [1/4 1/2 1/4] presmooth): each thread produces 4 adjacent pixels from three
// 4-byte row loads + two edge bytes per row.  Same tap order and separately rounded products as the tile kernel.
__global__ void __launch_bounds__(256)
pyr0_kernel(const uint8_t* __restrict__ img, size_t step, size_t fstride, int W, int H, float k0, float k1, float k2,
            float* __restrict__ out, int pitch, size_t ostride)
{
    const int x = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4;       // W % 4 == 0 is guaranteed by the launcher
    const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    img += (size_t)blockIdx.z * fstride;
    out += (size_t)blockIdx.z * ostride;
    const int xl = x > 0 ? x - 1 : 1, xr = x + 4 < W ? x + 4 : W - 2;   // REFLECT_101
    float hb[3][4];
#pragma unroll
    for (int rr = 0; rr < 3; rr++) {
        int yy = y + rr - 1;
        yy = yy < 0 ? 1 : (yy >= H ? H - 2 : yy);
        const uint8_t* row = img + (size_t)yy * step;
        const uchar4 m = *reinterpret_cast<const uchar4*>(row + x);
        const float v[6] = {(float)row[xl], (float)m.x, (float)m.y, (float)m.z, (float)m.w, (float)row[xr]};
#pragma unroll
        for (int i = 0; i < 4; i++)
            hb[rr][i] = __fadd_rn(__fadd_rn(__fmul_rn(k0, v[i]), __fmul_rn(k1, v[i + 1])), __fmul_rn(k2, v[i + 2]));
    }
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[0][0]), __fmul_rn(k1, hb[1][0])), __fmul_rn(k2, hb[2][0]));
    o.y = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[0][1]), __fmul_rn(k1, hb[1][1])), __fmul_rn(k2, hb[2][1]));
    o.z = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[0][2]), __fmul_rn(k1, hb[1][2])), __fmul_rn(k2, hb[2][2]));
    o.w = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[0][3]), __fmul_rn(k1, hb[1][3])), __fmul_rn(k2, hb[2][3]));
    *reinterpret_cast<float4*>(out + (size_t)y * pitch + x) = o;
}

// ---------------------------------------------------------------------------------------------------
// Layers 0, 1 and 2 of a pyr_scale = 0.5 pyramid in ONE pass over the frame (the reference default: three layers with
// 3-, 3- and 9-tap presmooth).  A CTA stages a 128x32 pixel tile (+3 halo, REFLECT_101) as fp32 in shared memory once
// and produces the 128x32, 64x16 and 32x8 pixels of the three layer images from it, so the u8 frame is read once
// instead of three times and one launch replaces three.  Per-pixel arithmetic (tap order, separately rounded products
// and sums) is that of pyr0_kernel / pyr_half_kernel / pyr_quarter_kernel: the results are bit-identical.
// ---------------------------------------------------------------------------------------------------
struct Pyr3Args {
    const uint8_t* img; size_t step, fstride; int W, H;
    float* out[3]; int pitch[3]; size_t ostride[3];
    float k3a[3], k3b[3];       // presmooth taps of layer 0 and layer 1
    SmoothCoef k9;              // 9 taps of layer 2
};

// 12 consecutive floats from a 16-byte aligned shared-memory address (three conflict-free 128-bit loads)
__device__ __forceinline__ void lds12(const float* p, float w[12])
{
    *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(p);
    *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(p + 4);
    *reinterpret_cast<float4*>(w + 8) = *reinterpret_cast<const float4*>(p + 8);
}

__global__ void __launch_bounds__(256)
pyr3_kernel(Pyr3Args a)
{
    constexpr int TW = 128, TH = 32, SP = 136, SR = TH + 6;       // staged: columns X0-4 .. X0+131, rows Y0-3 .. Y0+34
    __shared__ __align__(16) float sS[SR * SP];
    __shared__ float2 sHq[SR * 32];                                  // layer 2: H-blurred pair per (staged row, output column)
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH, W = a.W, H = a.H;
    const uint8_t* img = a.img + (size_t)blockIdx.z * a.fstride;
    // ---- stage (sS column c <-> image column X0 - 4 + c)
    const bool interior = X0 >= 4 && X0 + 132 <= W && Y0 >= 3 && Y0 + TH + 3 <= H;
    if (interior) {
        for (int r = wrp; r < SR; r += 8) {
            const unsigned int* grow = reinterpret_cast<const unsigned int*>(img + (size_t)(Y0 - 3 + r) * a.step + (X0 - 4));
            float* srow = sS + r * SP;
            for (int q = lane; q < SP / 4; q += 32) {
                const unsigned int v = __ldg(grow + q);
                *reinterpret_cast<float4*>(srow + 4 * q) =
                    make_float4((float)(v & 0xffu), (float)((v >> 8) & 0xffu), (float)((v >> 16) & 0xffu), (float)(v >> 24));
            }
        }
    } else {
        for (int r = wrp; r < SR; r += 8) {
            const uint8_t* grow = img + (size_t)reflect101(Y0 - 3 + r, H) * a.step;
            float* srow = sS + r * SP;
            for (int cidx = lane; cidx < SP; cidx += 32) srow[cidx] = (float)grow[reflect101(X0 - 4 + cidx, W)];
        }
    }
    __syncthreads();
    // ---- layer 2, horizontal pass: columns 4X+1 and 4X+2 of every staged row (window columns 4X-3 .. 4X+6)
    {
        float k[9];
#pragma unroll
        for (int i = 0; i < 9; i++) k[i] = a.k9.k[i];
        for (int it = tid; it < SR * 32; it += 256) {
            const int r = it >> 5, X = it & 31;
            float w12[12];
            lds12(sS + r * SP + 4 * X, w12);
            const float* v = w12 + 1;                                // image column X0 + 4X - 3
            float s0 = __fmul_rn(k[0], v[0]), s1 = __fmul_rn(k[0], v[1]);
#pragma unroll
            for (int i = 1; i < 9; i++) { s0 = __fadd_rn(s0, __fmul_rn(k[i], v[i])); s1 = __fadd_rn(s1, __fmul_rn(k[i], v[i + 1])); }
            sHq[it] = make_float2(s0, s1);
        }
    }
    // ---- layer 0: item = 4 adjacent pixels x 4 rows: the six horizontally blurred rows it needs are formed once (a 4 x 1
    // item recomputes three per output row); per-pixel arithmetic and order unchanged
    {
        const float k0 = a.k3a[0], k1 = a.k3a[1], k2 = a.k3a[2];
        float* out = a.out[0] + (size_t)blockIdx.z * a.ostride[0];
        for (int it = tid; it < (TH / 4) * (TW / 4); it += 256) {
            const int y = (it >> 5) * 4, x4 = (it & 31) * 4;
            if (X0 + x4 >= W || Y0 + y >= H) continue;
            float hb[6][4];
#pragma unroll
            for (int rr = 0; rr < 6; rr++) {
                float w12[12];
                lds12(sS + (y + 2 + rr) * SP + x4, w12);
                const float* v = w12 + 3;                            // image row Y0 + y - 1 + rr, column X0 + x4 - 1
#pragma unroll
                for (int i = 0; i < 4; i++)
                    hb[rr][i] = __fadd_rn(__fadd_rn(__fmul_rn(k0, v[i]), __fmul_rn(k1, v[i + 1])), __fmul_rn(k2, v[i + 2]));
            }
#pragma unroll
            for (int oy = 0; oy < 4; oy++) {
                if (Y0 + y + oy >= H) break;
                float4 o;
                o.x = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[oy][0]), __fmul_rn(k1, hb[oy + 1][0])), __fmul_rn(k2, hb[oy + 2][0]));
                o.y = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[oy][1]), __fmul_rn(k1, hb[oy + 1][1])), __fmul_rn(k2, hb[oy + 2][1]));
                o.z = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[oy][2]), __fmul_rn(k1, hb[oy + 1][2])), __fmul_rn(k2, hb[oy + 2][2]));
                o.w = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[oy][3]), __fmul_rn(k1, hb[oy + 1][3])), __fmul_rn(k2, hb[oy + 2][3]));
                *reinterpret_cast<float4*>(out + (size_t)(Y0 + y + oy) * a.pitch[0] + X0 + x4) = o;
            }
        }
    }
    // ---- layer 1: item = 2 adjacent pixels x 2 rows (source window 6 rows x 6 columns; the two rows share two of their
    // four horizontally blurred source rows)
    {
        const float k0 = a.k3b[0], k1 = a.k3b[1], k2 = a.k3b[2];
        float* out = a.out[1] + (size_t)blockIdx.z * a.ostride[1];
        const int dw = W / 2, dh = H / 2;
        for (int it = tid; it < (TH / 4) * (TW / 4); it += 256) {
            const int y = (it >> 5) * 2, x2 = (it & 31) * 2;          // layer-1 pixels (X0/2 + x2 + {0,1}, Y0/2 + y + {0,1})
            const int X = X0 / 2 + x2, Y = Y0 / 2 + y;
            if (X >= dw || Y >= dh) continue;
            float hb[6][4];
#pragma unroll
            for (int rr = 0; rr < 6; rr++) {
                float w12[12];
                lds12(sS + (2 * y + 2 + rr) * SP + 2 * x2, w12);
                const float* v = w12 + 3;                            // image row 2Y - 1 + rr, column 2X - 1
#pragma unroll
                for (int i = 0; i < 4; i++)
                    hb[rr][i] = __fadd_rn(__fadd_rn(__fmul_rn(k0, v[i]), __fmul_rn(k1, v[i + 1])), __fmul_rn(k2, v[i + 2]));
            }
#pragma unroll
            for (int oy = 0; oy < 2; oy++) {
                if (Y + oy >= dh) break;
                float r[2];
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    float b[2][2];
#pragma unroll
                    for (int sr = 0; sr < 2; sr++)
#pragma unroll
                        for (int sc = 0; sc < 2; sc++) {
                            const int c = 2 * d + sc, r0 = 2 * oy + sr;
                            b[sr][sc] = __fadd_rn(__fadd_rn(__fmul_rn(k0, hb[r0][c]), __fmul_rn(k1, hb[r0 + 1][c])), __fmul_rn(k2, hb[r0 + 2][c]));
                        }
                    const float top = __fadd_rn(__fmul_rn(b[0][0], 0.5f), __fmul_rn(b[0][1], 0.5f));
                    const float bot = __fadd_rn(__fmul_rn(b[1][0], 0.5f), __fmul_rn(b[1][1], 0.5f));
                    r[d] = __fadd_rn(__fmul_rn(top, 0.5f), __fmul_rn(bot, 0.5f));
                }
                *reinterpret_cast<float2*>(out + (size_t)(Y + oy) * a.pitch[1] + X) = make_float2(r[0], r[1]);
            }
        }
    }
    __syncthreads();
    // ---- layer 2, vertical pass: one pixel per thread (rows 4Y-3 .. 4Y+6 = staged rows 4y .. 4y+9)
    {
        const int y = tid >> 5, X = X0 / 4 + lane, Y = Y0 / 4 + y;
        if (X < W / 4 && Y < H / 4) {
            float k[9];
#pragma unroll
            for (int i = 0; i < 9; i++) k[i] = a.k9.k[i];
            const float2* hq = sHq + (4 * y) * 32 + lane;
            float2 p = hq[0], q = hq[32];
            float b00 = __fmul_rn(k[0], p.x), b01 = __fmul_rn(k[0], p.y), b10 = __fmul_rn(k[0], q.x), b11 = __fmul_rn(k[0], q.y);
#pragma unroll
            for (int jj = 1; jj < 9; jj++) {
                p = q; q = hq[(jj + 1) * 32];
                b00 = __fadd_rn(b00, __fmul_rn(k[jj], p.x)); b01 = __fadd_rn(b01, __fmul_rn(k[jj], p.y));
                b10 = __fadd_rn(b10, __fmul_rn(k[jj], q.x)); b11 = __fadd_rn(b11, __fmul_rn(k[jj], q.y));
            }
            const float top = __fadd_rn(__fmul_rn(b00, 0.5f), __fmul_rn(b01, 0.5f));
            const float bot = __fadd_rn(__fmul_rn(b10, 0.5f), __fmul_rn(b11, 0.5f));
            float* out = a.out[2] + (size_t)blockIdx.z * a.ostride[2];
            out[(size_t)Y * a.pitch[2] + X] = __fadd_rn(__fmul_rn(top, 0.5f), __fmul_rn(bot, 0.5f));
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Polynomial expansion (Appendix A.3), STRICT tile kernel: vertical pass fp32 -> shared memory,
// horizontal pass with fp64 accumulators, replicate borders, every tap, no contraction.
// ---------------------------------------------------------------------------------------------------
template <int TX, int TY>
__global__ void polyexp_strict_kernel(const float* __restrict__ I, size_t istride, int w, int h, int pitch,
                                      float* __restrict__ R, size_t plane, int first_slot, int nslots, PolyCoef pc)
{
    extern __shared__ float sm[];
    const int n = pc.n;
    const int SW = TX + 2 * n;
    float* sI = sm;
    float* sr = sm + (TY + 2 * n) * SW;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x, nt = blockDim.x;
    I += (size_t)blockIdx.z * istride;
    R += (size_t)((first_slot + blockIdx.z) % nslots) * 5 * plane;

    for (int idx = tid; idx < (TY + 2 * n) * SW; idx += nt) {
        int r = idx / SW, c = idx - r * SW;
        int gy = clampi(y0 + r - n, 0, h - 1), gx = clampi(x0 + c - n, 0, w - 1);
        sI[idx] = I[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int idx = tid; idx < TY * SW; idx += nt) {
        int r = idx / SW, c = idx - r * SW;
        const float* col = sI + (r + n) * SW + c;
        float r0 = __fmul_rn(col[0], pc.g[0]), r1 = 0.f, r2 = 0.f;
        for (int k = 1; k <= n; k++) {
            float up = col[-k * SW], dn = col[k * SW];
            float p = __fadd_rn(up, dn);
            r0 = __fadd_rn(r0, __fmul_rn(pc.g[k], p));
            r1 = __fadd_rn(r1, __fmul_rn(pc.xg[k], __fsub_rn(dn, up)));
            r2 = __fadd_rn(r2, __fmul_rn(pc.xxg[k], p));
        }
        sr[idx] = r0; sr[TY * SW + idx] = r1; sr[2 * TY * SW + idx] = r2;
    }
    __syncthreads();
    for (int idx = tid; idx < TY * TX; idx += nt) {
        int r = idx / TX, c = idx - r * TX;
        int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        const float* q0 = sr + r * SW + c + n;
        const float* q1 = q0 + TY * SW;
        const float* q2 = q1 + TY * SW;
        double b1 = __fmul_rn(q0[0], pc.g[0]), b2 = 0, b3 = __fmul_rn(q1[0], pc.g[0]), b4 = 0,
               b5 = __fmul_rn(q2[0], pc.g[0]), b6 = 0;
        for (int k = 1; k <= n; k++) {
            double tg = __fadd_rn(q0[k], q0[-k]);
            b1 = __dadd_rn(b1, __dmul_rn(tg, (double)pc.g[k]));
            b4 = __dadd_rn(b4, __dmul_rn(tg, (double)pc.xxg[k]));
            b2 = __dadd_rn(b2, __dmul_rn((double)__fsub_rn(q0[k], q0[-k]), (double)pc.xg[k]));
            b3 = __dadd_rn(b3, __dmul_rn((double)__fadd_rn(q1[k], q1[-k]), (double)pc.g[k]));
            b6 = __dadd_rn(b6, __dmul_rn((double)__fsub_rn(q1[k], q1[-k]), (double)pc.xg[k]));
            b5 = __dadd_rn(b5, __dmul_rn((double)__fadd_rn(q2[k], q2[-k]), (double)pc.g[k]));
        }
        size_t o = (size_t)y * pitch + x;
        reinterpret_cast<float4*>(R)[o] =
            make_float4((float)__dmul_rn(b3, pc.ig11), (float)__dmul_rn(b2, pc.ig11),
                        (float)__dadd_rn(__dmul_rn(b1, pc.ig03), __dmul_rn(b5, pc.ig33)),
                        (float)__dadd_rn(__dmul_rn(b1, pc.ig03), __dmul_rn(b4, pc.ig33)));
        R[4 * plane + o] = (float)__dmul_rn(b6, pc.ig55);
    }
}

// ---------------------------------------------------------------------------------------------------
// Polynomial expansion, FAST kernel.  NP = evaluated taps per side (multiple of 4; weights beyond poly_n are 0).
// CTA tile: TX = 128 - 2*NP outputs wide, 32 rows.  Shared row width is exactly 128 floats.
//  phase V: thread = (column, 16-row segment): its 16 + 2*NP input values come straight from global memory into
//           registers (coalesced 128 B per warp and row, all loads independent), the three vertical sums of each
//           of its rows are formed with FMAs and stored to shared memory.
//  phase H: thread = (row, 4 adjacent outputs): its (4 + 2*NP)-wide window of each vertical sum is read with
//           conflict-free 16-byte shared loads; six symmetric/antisymmetric FMA chains; 16-byte plane stores.
// ---------------------------------------------------------------------------------------------------
#ifndef RC_POLY_SW
#define RC_POLY_SW 144
#endif
struct PolyCoefF {
    float g[RC_MAX_POLY_N + 1], xg[RC_MAX_POLY_N + 1], xxg[RC_MAX_POLY_N + 1];
    float ig11, ig03, ig33, ig55;
};

// TY = 32 rows per CTA (256 threads) or 16 (128 threads: half the shared memory, twice the resident CTAs; RC_POLYEXP_ROWS)
// NC = taps per side that are evaluated (<= NP, the tile radius, which the 16-byte shared-memory windows need even)
template <int NP, int TY, int NC = NP>
__global__ void __launch_bounds__(TY * 8)
polyexp_fast_kernel(const float* __restrict__ I, size_t istride, int w, int h, int pitch, float* __restrict__ R,
                    size_t plane, int first_slot, int nslots, PolyCoefF pc)
{
    // 128 columns are used; with 16-row tiles the row pitch of 144 floats (the 32-row form is at the 48 KB static limit) shifts consecutive rows by 16 banks, so a quarter-warp of phase H
    // whose eight 16-byte windows straddle two rows (28 groups per row) still touches 32 distinct banks
    constexpr int TX = 128 - 2 * NP, SW = TY == 16 ? RC_POLY_SW : 128, VB = 16, WIN = VB + 2 * NP, NTHR = TY * 8;
    __shared__ __align__(16) float sr[3][TY][SW];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    I += (size_t)blockIdx.z * istride;
    R += (size_t)((first_slot + blockIdx.z) % nslots) * 5 * plane;

    {   // ---- phase V
        const int col = tid & 127, seg = tid >> 7;
        const int gx = clampi(x0 - NP + col, 0, w - 1);
        const int ybase = y0 + seg * VB - NP;
        float win[WIN];
#pragma unroll
        for (int j = NP - NC; j < WIN - (NP - NC); j++) win[j] = __ldg(I + (size_t)clampi(ybase + j, 0, h - 1) * pitch + gx);
#pragma unroll
        for (int i = 0; i < VB; i++) {
            float r0 = win[i + NP] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 1; k <= NC; k++) {
                float up = win[i + NP - k], dn = win[i + NP + k];
                float p = up + dn;
                r0 = fmaf(pc.g[k], p, r0);
                r1 = fmaf(pc.xg[k], dn - up, r1);
                r2 = fmaf(pc.xxg[k], p, r2);
            }
            sr[0][seg * VB + i][col] = r0; sr[1][seg * VB + i][col] = r1; sr[2][seg * VB + i][col] = r2;
        }
    }
    __syncthreads();
    {   // ---- phase H
        constexpr int GPR = TX / 4;            // 4-wide groups per row
        constexpr int NW = 4 + 2 * NP;         // window width
        for (int it = tid; it < TY * GPR; it += NTHR) {
            const int row = it / GPR, xg4 = it - row * GPR;
            const int x = x0 + 4 * xg4, y = y0 + row;
            if (x >= w || y >= h) continue;
            float wv[NW];
            float b1[4], b2[4], b4[4], t1[4];
            // r0 window: b1 (g), b2 (xg, antisymmetric), b4 (xxg)
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[0][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a1 = wv[NP + i] * pc.g[0], a2 = 0.f, a4 = 0.f;
#pragma unroll
                for (int k = 1; k <= NC; k++) {
                    float p = wv[NP + i + k], m = wv[NP + i - k];
                    float tg = p + m;
                    a1 = fmaf(tg, pc.g[k], a1);
                    a4 = fmaf(tg, pc.xxg[k], a4);
                    a2 = fmaf(p - m, pc.xg[k], a2);
                }
                b1[i] = a1; b2[i] = a2; b4[i] = a4;
            }
            float o1[4], o3[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                o1[i] = b2[i] * pc.ig11;
                t1[i] = b1[i] * pc.ig03;
                o3[i] = fmaf(b4[i], pc.ig33, t1[i]);
            }
            // r2 window: b5 (g)
            float o2[4];
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[2][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a5 = wv[NP + i] * pc.g[0];
#pragma unroll
                for (int k = 1; k <= NC; k++) a5 = fmaf(wv[NP + i + k] + wv[NP + i - k], pc.g[k], a5);
                o2[i] = fmaf(a5, pc.ig33, t1[i]);
            }
            // r1 window: b3 (g), b6 (xg, antisymmetric)
            float o0[4], o4[4];
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[1][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a3 = wv[NP + i] * pc.g[0], a6 = 0.f;
#pragma unroll
                for (int k = 1; k <= NC; k++) {
                    float p = wv[NP + i + k], m = wv[NP + i - k];
                    a3 = fmaf(p + m, pc.g[k], a3);
                    a6 = fmaf(p - m, pc.xg[k], a6);
                }
                o0[i] = a3 * pc.ig11; o4[i] = a6 * pc.ig55;
            }
            const size_t o = (size_t)y * pitch + x;
            float4* A = reinterpret_cast<float4*>(R) + o;
            if (x + 3 < w) {
#pragma unroll
                for (int i = 0; i < 4; i++) A[i] = make_float4(o0[i], o1[i], o2[i], o3[i]);
                *reinterpret_cast<float4*>(R + 4 * plane + o) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (x + i < w) { A[i] = make_float4(o0[i], o1[i], o2[i], o3[i]); R[4 * plane + o + i] = o4[i]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Polynomial expansion, FAST kernel, packed-fp32 variant (RC_POLYEXP=packed; sm_100 FFMA2 / FADD2).  Same tiling, same
// shared-memory traffic and the same symmetric / antisymmetric tap pairing as polyexp_fast_kernel; the difference is that
// every thread's four outputs (vertical phase: four rows of its column; horizontal phase: four adjacent pixels) are
// computed as register-aligned PAIRS: taps at an even distance k pair the outputs (0,1)(2,3), taps at an odd distance pair
// (-1,0)(1,2)(3,4) -- the window values of such a pair are two consecutive, even-aligned registers, so the symmetric sum
// w[c+k] + w[c-k], the antisymmetric difference and the multiply-accumulates are one packed instruction for two outputs.
// The two partial sums of every output are added at the end.  ~150 instead of ~190 thread instructions per pixel.
// ---------------------------------------------------------------------------------------------------
struct PolyCoefF2 {
    float2 g[17], xg[17], xxg[17];      // every weight duplicated into both halves: a 64-bit constant operand per tap
    float ig11, ig03, ig33, ig55;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

template <int NP>
__global__ void __launch_bounds__(256, 4)
polyexp_packed_kernel(const float* __restrict__ I, size_t istride, int w, int h, int pitch, float* __restrict__ R,
                      size_t plane, int first_slot, int nslots, PolyCoefF2 pc)
{
    static_assert(NP % 2 == 0 && NP <= 16, "even tap radius");
    constexpr int TX = 128 - 2 * NP, TY = 32, SW = 128, VB = 16, WIN = VB + 2 * NP;
    __shared__ __align__(16) float sr[3][TY][SW];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    I += (size_t)blockIdx.z * istride;
    R += (size_t)((first_slot + blockIdx.z) % nslots) * 5 * plane;

    {   // ---- phase V: thread = (column, 16-row segment), four output rows at a time
        const int col = tid & 127, seg = tid >> 7;
        const int gx = clampi(x0 - NP + col, 0, w - 1);
        const int ybase = y0 + seg * VB - NP;
        float win[WIN];
#pragma unroll
        for (int j = 0; j < WIN; j++) win[j] = __ldg(I + (size_t)clampi(ybase + j, 0, h - 1) * pitch + gx);
#pragma unroll
        for (int i0 = 0; i0 < VB; i0 += 4) {
            const int c = i0 + NP;                                   // centre of output row i0 (even)
            float2 e0[2], o0[3], e1[2], o1[3], e2[2], o2[3];         // r0 (g), r1 (xg, antisymmetric), r2 (xxg)
#pragma unroll
            for (int p = 0; p < 2; p++) {
                e0[p] = __fmul2_rn(f2(win[c + 2 * p], win[c + 2 * p + 1]), pc.g[0]);
                e1[p] = e2[p] = f2(0.f, 0.f);
            }
#pragma unroll
            for (int p = 0; p < 3; p++) o0[p] = o1[p] = o2[p] = f2(0.f, 0.f);
#pragma unroll
            for (int k = 1; k <= NP; k++) {
                if ((k & 1) == 0) {
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const float2 dn = f2(win[c + 2 * p + k], win[c + 2 * p + 1 + k]), up = f2(win[c + 2 * p - k], win[c + 2 * p + 1 - k]);
                        const float2 sm = __fadd2_rn(dn, up);
                        e0[p] = __ffma2_rn(sm, pc.g[k], e0[p]);
                        e2[p] = __ffma2_rn(sm, pc.xxg[k], e2[p]);
                        e1[p] = __ffma2_rn(sub2(dn, up), pc.xg[k], e1[p]);
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < 3; p++) {
                        const float2 dn = f2(win[c + 2 * p - 1 + k], win[c + 2 * p + k]), up = f2(win[c + 2 * p - 1 - k], win[c + 2 * p - k]);
                        const float2 sm = __fadd2_rn(dn, up);
                        o0[p] = __ffma2_rn(sm, pc.g[k], o0[p]);
                        o2[p] = __ffma2_rn(sm, pc.xxg[k], o2[p]);
                        o1[p] = __ffma2_rn(sub2(dn, up), pc.xg[k], o1[p]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const int r = seg * VB + i0 + 2 * p;
                sr[0][r][col] = e0[p].x + o0[p].y; sr[0][r + 1][col] = e0[p].y + o0[p + 1].x;
                sr[1][r][col] = e1[p].x + o1[p].y; sr[1][r + 1][col] = e1[p].y + o1[p + 1].x;
                sr[2][r][col] = e2[p].x + o2[p].y; sr[2][r + 1][col] = e2[p].y + o2[p + 1].x;
            }
        }
    }
    __syncthreads();
    {   // ---- phase H: thread = (row, 4 adjacent outputs)
        constexpr int GPR = TX / 4;
        constexpr int NW = 4 + 2 * NP;
        for (int it = tid; it < TY * GPR; it += 256) {
            const int row = it / GPR, xg4 = it - row * GPR;
            const int x = x0 + 4 * xg4, y = y0 + row;
            if (x >= w || y >= h) continue;
            float wv[NW];
            float b1[4], b2[4], b4[4], b5[4], b3[4], b6[4];
            // symmetric (weights ws) + optional second symmetric (ws2) + optional antisymmetric (wa) sums of one window
            auto load = [&](int plane_idx) {
#pragma unroll
                for (int j = 0; j < NW / 4; j++)
                    *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[plane_idx][row][4 * xg4 + 4 * j]);
            };
            // ---- r0 window: b1 (g), b4 (xxg), b2 (xg, antisymmetric)
            load(0);
            {
                float2 e1[2], o1[3], e4[2], o4[3], e2[2], o2[3];
#pragma unroll
                for (int p = 0; p < 2; p++) { e1[p] = __fmul2_rn(f2(wv[NP + 2 * p], wv[NP + 2 * p + 1]), pc.g[0]); e4[p] = e2[p] = f2(0.f, 0.f); }
#pragma unroll
                for (int p = 0; p < 3; p++) o1[p] = o4[p] = o2[p] = f2(0.f, 0.f);
#pragma unroll
                for (int k = 1; k <= NP; k++) {
                    if ((k & 1) == 0) {
#pragma unroll
                        for (int p = 0; p < 2; p++) {
                            const float2 P = f2(wv[NP + 2 * p + k], wv[NP + 2 * p + 1 + k]), M = f2(wv[NP + 2 * p - k], wv[NP + 2 * p + 1 - k]);
                            const float2 tg = __fadd2_rn(P, M);
                            e1[p] = __ffma2_rn(tg, pc.g[k], e1[p]); e4[p] = __ffma2_rn(tg, pc.xxg[k], e4[p]);
                            e2[p] = __ffma2_rn(sub2(P, M), pc.xg[k], e2[p]);
                        }
                    } else {
#pragma unroll
                        for (int p = 0; p < 3; p++) {
                            const float2 P = f2(wv[NP + 2 * p - 1 + k], wv[NP + 2 * p + k]), M = f2(wv[NP + 2 * p - 1 - k], wv[NP + 2 * p - k]);
                            const float2 tg = __fadd2_rn(P, M);
                            o1[p] = __ffma2_rn(tg, pc.g[k], o1[p]); o4[p] = __ffma2_rn(tg, pc.xxg[k], o4[p]);
                            o2[p] = __ffma2_rn(sub2(P, M), pc.xg[k], o2[p]);
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    b1[2 * p] = e1[p].x + o1[p].y; b1[2 * p + 1] = e1[p].y + o1[p + 1].x;
                    b4[2 * p] = e4[p].x + o4[p].y; b4[2 * p + 1] = e4[p].y + o4[p + 1].x;
                    b2[2 * p] = e2[p].x + o2[p].y; b2[2 * p + 1] = e2[p].y + o2[p + 1].x;
                }
            }
            // ---- r2 window: b5 (g)
            load(2);
            {
                float2 e5[2], o5[3];
#pragma unroll
                for (int p = 0; p < 2; p++) e5[p] = __fmul2_rn(f2(wv[NP + 2 * p], wv[NP + 2 * p + 1]), pc.g[0]);
#pragma unroll
                for (int p = 0; p < 3; p++) o5[p] = f2(0.f, 0.f);
#pragma unroll
                for (int k = 1; k <= NP; k++) {
                    if ((k & 1) == 0) {
#pragma unroll
                        for (int p = 0; p < 2; p++)
                            e5[p] = __ffma2_rn(__fadd2_rn(f2(wv[NP + 2 * p + k], wv[NP + 2 * p + 1 + k]), f2(wv[NP + 2 * p - k], wv[NP + 2 * p + 1 - k])), pc.g[k], e5[p]);
                    } else {
#pragma unroll
                        for (int p = 0; p < 3; p++)
                            o5[p] = __ffma2_rn(__fadd2_rn(f2(wv[NP + 2 * p - 1 + k], wv[NP + 2 * p + k]), f2(wv[NP + 2 * p - 1 - k], wv[NP + 2 * p - k])), pc.g[k], o5[p]);
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; p++) { b5[2 * p] = e5[p].x + o5[p].y; b5[2 * p + 1] = e5[p].y + o5[p + 1].x; }
            }
            // ---- r1 window: b3 (g), b6 (xg, antisymmetric)
            load(1);
            {
                float2 e3[2], o3[3], e6[2], o6[3];
#pragma unroll
                for (int p = 0; p < 2; p++) { e3[p] = __fmul2_rn(f2(wv[NP + 2 * p], wv[NP + 2 * p + 1]), pc.g[0]); e6[p] = f2(0.f, 0.f); }
#pragma unroll
                for (int p = 0; p < 3; p++) o3[p] = o6[p] = f2(0.f, 0.f);
#pragma unroll
                for (int k = 1; k <= NP; k++) {
                    if ((k & 1) == 0) {
#pragma unroll
                        for (int p = 0; p < 2; p++) {
                            const float2 P = f2(wv[NP + 2 * p + k], wv[NP + 2 * p + 1 + k]), M = f2(wv[NP + 2 * p - k], wv[NP + 2 * p + 1 - k]);
                            e3[p] = __ffma2_rn(__fadd2_rn(P, M), pc.g[k], e3[p]);
                            e6[p] = __ffma2_rn(sub2(P, M), pc.xg[k], e6[p]);
                        }
                    } else {
#pragma unroll
                        for (int p = 0; p < 3; p++) {
                            const float2 P = f2(wv[NP + 2 * p - 1 + k], wv[NP + 2 * p + k]), M = f2(wv[NP + 2 * p - 1 - k], wv[NP + 2 * p - k]);
                            o3[p] = __ffma2_rn(__fadd2_rn(P, M), pc.g[k], o3[p]);
                            o6[p] = __ffma2_rn(sub2(P, M), pc.xg[k], o6[p]);
                        }
                    }
                }
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    b3[2 * p] = e3[p].x + o3[p].y; b3[2 * p + 1] = e3[p].y + o3[p + 1].x;
                    b6[2 * p] = e6[p].x + o6[p].y; b6[2 * p + 1] = e6[p].y + o6[p + 1].x;
                }
            }
            const size_t o = (size_t)y * pitch + x;
            float4* A = reinterpret_cast<float4*>(R) + o;
            float o4v[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float t1 = b1[i] * pc.ig03;
                const float4 v = make_float4(b3[i] * pc.ig11, b2[i] * pc.ig11, fmaf(b5[i], pc.ig33, t1), fmaf(b4[i], pc.ig33, t1));
                o4v[i] = b6[i] * pc.ig55;
                if (x + i < w) A[i] = v;
            }
            if (x + 3 < w) *reinterpret_cast<float4*>(R + 4 * plane + o) = make_float4(o4v[0], o4v[1], o4v[2], o4v[3]);
            else
#pragma unroll
                for (int i = 0; i < 4; i++) if (x + i < w) R[4 * plane + o + i] = o4v[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Polynomial expansion, FAST kernel, TMA variant (RC_POLYEXP=tma).  Identical arithmetic and identical bits; the only
// difference is how the (32 + 2 NP) x 128 input tile reaches the SM: ONE cp.async.bulk.tensor.3d (TMA) issued by one thread
// into shared memory behind an mbarrier, instead of 32 coalesced global loads per thread into registers.  The tensor map
// describes the layer images [frame][row][column]; coordinates outside the image are zero-filled by the copy engine and
// never read: the replicate border is resolved when the vertical phase indexes the tile (clamped row / column -> a position
// inside the tile, because every tile intersects the image).  The tile aliases the memory of the vertical sums, so the
// occupancy is that of the direct kernel (4 CTAs per SM); measured in profiles/r02_polyexp_variants.txt.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int NP>
__global__ void __launch_bounds__(256, 4)
polyexp_tma_kernel(const __grid_constant__ CUtensorMap tmap, int w, int h, int pitch, float* __restrict__ R, size_t plane,
                   int first_slot, int nslots, PolyCoefF pc)
{
    constexpr int TX = 128 - 2 * NP, TY = 32, SW = 128, VB = 16, WIN = VB + 2 * NP, TR = TY + 2 * NP;
    constexpr unsigned TILE_BYTES = TR * SW * sizeof(float);
    extern __shared__ __align__(128) unsigned char tsm[];
    // the input tile ALIASES the first 24.5 KB of the vertical sums: every thread has its column segment in registers before
    // the first sum is stored (one extra barrier), so the TMA variant needs no more shared memory than the direct one
    float* tile = reinterpret_cast<float*>(tsm);                                        // [TR][SW]
    float (*sr)[TY][SW] = reinterpret_cast<float (*)[TY][SW]>(tsm);                      // [3][TY][SW]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(tsm + 3 * TY * SW * sizeof(float));
    static_assert(TILE_BYTES <= 3 * TY * SW * sizeof(float), "tile fits under the sums");
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x;
    R += (size_t)((first_slot + blockIdx.z) % nslots) * 5 * plane;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(TILE_BYTES) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(tile)), "l"(&tmap), "r"(x0 - NP), "r"(y0 - NP), "r"((int)blockIdx.z), "r"(smem_u32(bar))
                     : "memory");
    }
    {
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    {   // ---- phase V from the shared tile (replicate border = clamped tile coordinates)
        const int col = tid & 127, seg = tid >> 7;
        const int cj = clampi(x0 - NP + col, 0, w - 1) - (x0 - NP);
        const int ybase = y0 + seg * VB - NP;
        float win[WIN];
#pragma unroll
        for (int j = 0; j < WIN; j++) win[j] = tile[(clampi(ybase + j, 0, h - 1) - (y0 - NP)) * SW + cj];
        __syncthreads();                      // the tile is dead from here on: its memory receives the vertical sums
#pragma unroll
        for (int i = 0; i < VB; i++) {
            float r0 = win[i + NP] * pc.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 1; k <= NP; k++) {
                float up = win[i + NP - k], dn = win[i + NP + k];
                float p = up + dn;
                r0 = fmaf(pc.g[k], p, r0);
                r1 = fmaf(pc.xg[k], dn - up, r1);
                r2 = fmaf(pc.xxg[k], p, r2);
            }
            sr[0][seg * VB + i][col] = r0; sr[1][seg * VB + i][col] = r1; sr[2][seg * VB + i][col] = r2;
        }
    }
    __syncthreads();
    {   // ---- phase H (as polyexp_fast_kernel)
        constexpr int GPR = TX / 4;
        constexpr int NW = 4 + 2 * NP;
        for (int it = tid; it < TY * GPR; it += 256) {
            const int row = it / GPR, xg4 = it - row * GPR;
            const int x = x0 + 4 * xg4, y = y0 + row;
            if (x >= w || y >= h) continue;
            float wv[NW];
            float b1[4], b2[4], b4[4], t1[4];
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[0][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a1 = wv[NP + i] * pc.g[0], a2 = 0.f, a4 = 0.f;
#pragma unroll
                for (int k = 1; k <= NP; k++) {
                    float p = wv[NP + i + k], m = wv[NP + i - k];
                    float tg = p + m;
                    a1 = fmaf(tg, pc.g[k], a1);
                    a4 = fmaf(tg, pc.xxg[k], a4);
                    a2 = fmaf(p - m, pc.xg[k], a2);
                }
                b1[i] = a1; b2[i] = a2; b4[i] = a4;
            }
            float o1[4], o3[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                o1[i] = b2[i] * pc.ig11;
                t1[i] = b1[i] * pc.ig03;
                o3[i] = fmaf(b4[i], pc.ig33, t1[i]);
            }
            float o2[4];
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[2][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a5 = wv[NP + i] * pc.g[0];
#pragma unroll
                for (int k = 1; k <= NP; k++) a5 = fmaf(wv[NP + i + k] + wv[NP + i - k], pc.g[k], a5);
                o2[i] = fmaf(a5, pc.ig33, t1[i]);
            }
            float o0[4], o4[4];
#pragma unroll
            for (int j = 0; j < NW / 4; j++)
                *reinterpret_cast<float4*>(wv + 4 * j) = *reinterpret_cast<const float4*>(&sr[1][row][4 * xg4 + 4 * j]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float a3 = wv[NP + i] * pc.g[0], a6 = 0.f;
#pragma unroll
                for (int k = 1; k <= NP; k++) {
                    float p = wv[NP + i + k], m = wv[NP + i - k];
                    a3 = fmaf(p + m, pc.g[k], a3);
                    a6 = fmaf(p - m, pc.xg[k], a6);
                }
                o0[i] = a3 * pc.ig11; o4[i] = a6 * pc.ig55;
            }
            const size_t o = (size_t)y * pitch + x;
            float4* A = reinterpret_cast<float4*>(R) + o;
            if (x + 3 < w) {
#pragma unroll
                for (int i = 0; i < 4; i++) A[i] = make_float4(o0[i], o1[i], o2[i], o3[i]);
                *reinterpret_cast<float4*>(R + 4 * plane + o) = make_float4(o4[0], o4[1], o4[2], o4[3]);
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (x + i < w) { A[i] = make_float4(o0[i], o1[i], o2[i], o3[i]); R[4 * plane + o + i] = o4[i]; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// updateMatrices for one pixel (Appendix A.5).  R0/R1 = 5 planes each.  Result -> m[5].
// ---------------------------------------------------------------------------------------------------
// R view of one frame: A = {c0..c3} per pixel, B = c4
struct RView { const float4* A; const float* B; };
__device__ __forceinline__ RView rview(const float* slot, size_t plane)
{
    return RView{reinterpret_cast<const float4*>(slot), slot + 4 * plane};
}

template <bool S>
__device__ __forceinline__ void update_matrices_core(int x, int y, float dx, float dy, int w, int h, const RView& R0,
                                                     const RView& R1, int pitch, float m[5])
{
    const int p = y * pitch + x;
    float fx = fadd<S>((float)x, dx), fy = fadd<S>((float)y, dy);
    const float flx = floorf(fx), fly = floorf(fy);
    const int x1 = (int)flx, y1 = (int)fly;
    fx = fsub<S>(fx, flx); fy = fsub<S>(fy, fly);
    const float4 r0 = __ldg(R0.A + p);
    const float r0_4 = __ldg(R0.B + p);
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const int q = y1 * pitch + x1;
        const float4 t00 = __ldg(R1.A + q), t01 = __ldg(R1.A + q + 1), t10 = __ldg(R1.A + q + pitch),
                     t11 = __ldg(R1.A + q + pitch + 1);
        const float u00 = __ldg(R1.B + q), u01 = __ldg(R1.B + q + 1), u10 = __ldg(R1.B + q + pitch),
                    u11 = __ldg(R1.B + q + pitch + 1);
        const float gx = 1.f - fx, gy = 1.f - fy;
        const float a00 = fmul<S>(gx, gy), a01 = fmul<S>(fx, gy), a10 = fmul<S>(gx, fy), a11 = fmul<S>(fx, fy);
#define RC_BILERP(c00, c01, c10, c11) \
    fadd<S>(fadd<S>(fadd<S>(fmul<S>(a00, c00), fmul<S>(a01, c01)), fmul<S>(a10, c10)), fmul<S>(a11, c11))
        r2 = RC_BILERP(t00.x, t01.x, t10.x, t11.x);
        r3 = RC_BILERP(t00.y, t01.y, t10.y, t11.y);
        r4 = fmul<S>(fadd<S>(r0.z, RC_BILERP(t00.z, t01.z, t10.z, t11.z)), 0.5f);
        r5 = fmul<S>(fadd<S>(r0.w, RC_BILERP(t00.w, t01.w, t10.w, t11.w)), 0.5f);
        r6 = fmul<S>(fadd<S>(r0_4, RC_BILERP(u00, u01, u10, u11)), 0.25f);
#undef RC_BILERP
    } else {
        r2 = r3 = 0.f;
        r4 = r0.z; r5 = r0.w; r6 = fmul<S>(r0_4, 0.5f);
    }
    r2 = fmul<S>(fsub<S>(r0.x, r2), 0.5f);
    r3 = fmul<S>(fsub<S>(r0.y, r3), 0.5f);
    r2 = fadd<S>(r2, fadd<S>(fmul<S>(r4, dy), fmul<S>(r6, dx)));
    r3 = fadd<S>(r3, fadd<S>(fmul<S>(r6, dy), fmul<S>(r5, dx)));
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        // border = {0.14, 0.14, 0.4472, 0.4472, 0.4472} indexed by the distance to each edge
        auto bw = [](int d) { return d >= 5 ? 1.f : (d < 2 ? 0.14f : 0.4472f); };
        float scale = __fmul_rn(__fmul_rn(__fmul_rn(bw(x), bw(w - x - 1)), bw(y)), bw(h - y - 1));
        r2 = __fmul_rn(r2, scale); r3 = __fmul_rn(r3, scale); r4 = __fmul_rn(r4, scale);
        r5 = __fmul_rn(r5, scale); r6 = __fmul_rn(r6, scale);
    }
    m[0] = fadd<S>(fmul<S>(r4, r4), fmul<S>(r6, r6));
    m[1] = fmul<S>(fadd<S>(r4, r5), r6);
    m[2] = fadd<S>(fmul<S>(r5, r5), fmul<S>(r6, r6));
    m[3] = fadd<S>(fmul<S>(r4, r2), fmul<S>(r6, r3));
    m[4] = fadd<S>(fmul<S>(r6, r2), fmul<S>(r5, r3));
}

// Flow initialisation of Appendix A.4 at pixel (x, y) of a w x h layer from the coarser layer's flow.
template <bool S>
__device__ __forceinline__ float2 upsample_flow(const float* __restrict__ coarse, int cw, int ch, int x, int y,
                                                double sxs, double sys, float fscale)
{
    int sx, sy; float fx, fy;
    resize_coef(x, cw, sxs, sx, fx);
    resize_coef(y, ch, sys, sy, fy);
    int sx1 = sx + 1 < cw ? sx + 1 : cw - 1, sy1 = sy + 1 < ch ? sy + 1 : ch - 1;
    const float2* cf = reinterpret_cast<const float2*>(coarse);
    float2 a = __ldg(cf + (size_t)sy * cw + sx), b = __ldg(cf + (size_t)sy * cw + sx1);
    float2 c = __ldg(cf + (size_t)sy1 * cw + sx), d = __ldg(cf + (size_t)sy1 * cw + sx1);
    float tx = fadd<S>(fmul<S>(a.x, 1.f - fx), fmul<S>(b.x, fx)), ty = fadd<S>(fmul<S>(a.y, 1.f - fx), fmul<S>(b.y, fx));
    float bx = fadd<S>(fmul<S>(c.x, 1.f - fx), fmul<S>(d.x, fx)), by = fadd<S>(fmul<S>(c.y, 1.f - fx), fmul<S>(d.y, fx));
    float2 r;
    r.x = fmul<S>(fadd<S>(fmul<S>(tx, 1.f - fy), fmul<S>(bx, fy)), fscale);
    r.y = fmul<S>(fadd<S>(fmul<S>(ty, 1.f - fy), fmul<S>(by, fy)), fscale);
    return r;
}

// the same with the per-axis resize coefficients already known
__device__ __forceinline__ float2 upsample_flow_tab(const float2* __restrict__ cf, int cw, int ch, int sx, float fx,
                                                    int sy, float fy, float fscale)
{
    const int sx1 = sx + 1 < cw ? sx + 1 : cw - 1, sy1 = sy + 1 < ch ? sy + 1 : ch - 1;
    const float2 a = __ldg(cf + sy * cw + sx), b = __ldg(cf + sy * cw + sx1);
    const float2 c = __ldg(cf + sy1 * cw + sx), d = __ldg(cf + sy1 * cw + sx1);
    const float gx = 1.f - fx, gy = 1.f - fy;
    const float tx = a.x * gx + b.x * fx, ty = a.y * gx + b.y * fx;
    const float bx = c.x * gx + d.x * fx, by = c.y * gx + d.y * fx;
    return make_float2((tx * gy + bx * fy) * fscale, (ty * gy + by * fy) * fscale);
}

// arguments shared by the per-layer flow kernels
struct FlowArgs {
    const float* R; size_t plane; int pitch; int w, h; int nslots, prev_slot;
    const float* coarse; int cw, ch; size_t coarse_stride; double sxs, sys; float fscale;   // coarse == null: zero flow
    float* M; size_t m_stride;             // unfused path: [B][2][5 planes]
    float* flow; size_t flow_stride;       // this layer's flow (layers >= 1), or null when flow_dst is used
    float* flow_dst[RC_MAX_BATCH];         // layer 0: per-pair destination (flow ring slots)
    unsigned int* hist_delta;              // layer 0: per-pair histogram counts or null
    WinCoef win;
    __device__ __forceinline__ const float* R0(int j) const { return R + (size_t)((prev_slot + j) % nslots) * 5 * plane; }
    __device__ __forceinline__ const float* R1(int j) const { return R + (size_t)((prev_slot + j + 1) % nslots) * 5 * plane; }
    __device__ __forceinline__ float* out(int j) const { return flow ? flow + (size_t)j * flow_stride : flow_dst[j]; }
};

// ---- unfused path (strict mode, or windows too large for the fused kernel) -----------------------------------
template <bool S>
__global__ void update_matrices_kernel(FlowArgs a, int mi)
{
    const int w = a.w, h = a.h, j = blockIdx.z;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float2 f = make_float2(0.f, 0.f);
    if (a.coarse) f = upsample_flow<S>(a.coarse + (size_t)j * a.coarse_stride, a.cw, a.ch, x, y, a.sxs, a.sys, a.fscale);
    float m[5];
    update_matrices_core<S>(x, y, f.x, f.y, w, h, rview(a.R0(j), a.plane), rview(a.R1(j), a.plane), a.pitch, m);
    float* M = a.M + (size_t)j * a.m_stride + (size_t)mi * 5 * a.plane;
    const size_t o = (size_t)y * a.pitch + x;
#pragma unroll
    for (int c = 0; c < 5; c++) M[c * a.plane + o] = m[c];
}

__device__ __forceinline__ float2 solve_strict(double g11, double g12, double g22, double h1, double h2)
{
    double det = __dadd_rn(__dsub_rn(__dmul_rn(g11, g22), __dmul_rn(g12, g12)), 1e-3);
    double idet = __ddiv_rn(1.0, det);
    return make_float2((float)__dmul_rn(__dsub_rn(__dmul_rn(g11, h2), __dmul_rn(g12, h1)), idet),
                       (float)__dmul_rn(__dsub_rn(__dmul_rn(g22, h1), __dmul_rn(g12, h2)), idet));
}

// a*b - c*d with one rounding error (Kahan): exact product error of c*d recovered with an FMA
__device__ __forceinline__ float det2(float a, float b, float c, float d)
{
    float w = c * d;
    float e = fmaf(c, d, -w);
    float f = fmaf(a, b, -w);
    return f - e;
}

__device__ __forceinline__ float2 solve_fast(float g11, float g12, float g22, float h1, float h2)
{
    float idet = 1.f / (det2(g11, g22, g12, g12) + 1e-3f);
    return make_float2(det2(g11, h2, g12, h1) * idet, det2(g22, h1, g12, h2) * idet);
}

// The same solve on UNSCALED window sums S = s / ps (box windows): numerators and determinant both carry 1/ps^2, so only
// the regulariser changes: eps = 1e-3 / ps^2.  Saves the five post-scale multiplies; the reciprocal is the hardware
// approximation (1 ulp) -- both far below the 1e-3 px tolerance of the fast path (strict mode has its own solve).
__device__ __forceinline__ float2 solve_fast_unscaled(float g11, float g12, float g22, float h1, float h2, float eps)
{
    float idet;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(idet) : "f"(det2(g11, g22, g12, g12) + eps));
    return make_float2(det2(g11, h2, g12, h1) * idet, det2(g22, h1, g12, h2) * idet);
}

// one updateFlow iteration, per-pixel reference form: blur(M[mi]) -> solve -> (fused updateMatrices -> M[mi^1] | flow)
template <bool FUSE>
__global__ void update_flow_strict_kernel(FlowArgs a, int mi)
{
    const int w = a.w, h = a.h, j = blockIdx.z, m = a.win.m;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float* Min = a.M + (size_t)j * a.m_stride + (size_t)mi * 5 * a.plane;
    float2 f;
    if (!a.win.gaussian) {
        double s[5] = {0, 0, 0, 0, 0};
        for (int i = -m; i <= m; i++) {
            int xx = clampi(x + i, 0, w - 1);
            double v[5] = {0, 0, 0, 0, 0};
            for (int jj = -m; jj <= m; jj++) {
                size_t o = (size_t)clampi(y + jj, 0, h - 1) * a.pitch + xx;
#pragma unroll
                for (int c = 0; c < 5; c++) v[c] = __dadd_rn(v[c], (double)__ldg(Min + c * a.plane + o));
            }
#pragma unroll
            for (int c = 0; c < 5; c++) s[c] = __dadd_rn(s[c], v[c]);
        }
        const double sc = a.win.post_scale_d;
        f = solve_strict(__dmul_rn(s[0], sc), __dmul_rn(s[1], sc), __dmul_rn(s[2], sc), __dmul_rn(s[3], sc),
                         __dmul_rn(s[4], sc));
    } else {
        auto vcol = [&](int xx, float* v) {
            size_t o = (size_t)y * a.pitch + xx;
#pragma unroll
            for (int c = 0; c < 5; c++) v[c] = __fmul_rn(__ldg(Min + c * a.plane + o), a.win.k[0]);
            for (int jj = 1; jj <= m; jj++) {
                size_t ou = (size_t)clampi(y - jj, 0, h - 1) * a.pitch + xx;
                size_t od = (size_t)clampi(y + jj, 0, h - 1) * a.pitch + xx;
#pragma unroll
                for (int c = 0; c < 5; c++)
                    v[c] = __fadd_rn(v[c], __fmul_rn(__fadd_rn(__ldg(Min + c * a.plane + od), __ldg(Min + c * a.plane + ou)),
                                                     a.win.k[jj]));
            }
        };
        float hs[5], v0[5], va[5], vb[5];
        vcol(x, v0);
#pragma unroll
        for (int c = 0; c < 5; c++) hs[c] = __fmul_rn(v0[c], a.win.k[0]);
        for (int i = 1; i <= m; i++) {
            vcol(clampi(x - i, 0, w - 1), va);
            vcol(clampi(x + i, 0, w - 1), vb);
#pragma unroll
            for (int c = 0; c < 5; c++) hs[c] = __fadd_rn(hs[c], __fmul_rn(a.win.k[i], __fadd_rn(va[c], vb[c])));
        }
        f = solve_strict(hs[0], hs[1], hs[2], hs[3], hs[4]);
    }
    if (FUSE) {
        float mm[5];
        update_matrices_core<true>(x, y, f.x, f.y, w, h, rview(a.R0(j), a.plane), rview(a.R1(j), a.plane), a.pitch, mm);
        float* Mo = a.M + (size_t)j * a.m_stride + (size_t)(mi ^ 1) * 5 * a.plane;
        const size_t o = (size_t)y * a.pitch + x;
#pragma unroll
        for (int c = 0; c < 5; c++) Mo[c * a.plane + o] = mm[c];
    } else {
        reinterpret_cast<float2*>(a.out(j))[(size_t)y * w + x] = f;
    }
}

// ---------------------------------------------------------------------------------------------------
// FAST path, general window (winsize 4..33, box or Gaussian): one updateFlow iteration per launch.
// CTA = 64x16 output pixels.  The five channels of M go through shared memory ONE AT A TIME (tile + halo staged with
// replicate clamping, separable blur: vertical pass -> shared, horizontal pass -> registers), so a CTA needs only
// (16+2m)(64+2m)+16(64+2m) floats of shared memory and the five blurred sums of each pixel stay in registers for
// the fp32 solve; then either the fused updateMatrices (-> M of the next iteration) or the flow store.
// ---------------------------------------------------------------------------------------------------
template <bool FUSE>
__global__ void __launch_bounds__(256)
flow_iter_tiled_kernel(FlowArgs a, int mi)
{
    constexpr int TX = 64, TY = 16;
    extern __shared__ float fsm[];
    const int m = a.win.m, WP = TX + 2 * m, HP = TY + 2 * m;
    float* sIn = fsm;                   // [HP][WP]
    float* sV = fsm + HP * WP;          // [TY][WP]
    const int w = a.w, h = a.h, j = blockIdx.x, tid = threadIdx.x;       // grid = (pairs, tiles_x, tiles_y)
    const int x0 = blockIdx.y * TX, y0 = blockIdx.z * TY;
    const float* Min = a.M + (size_t)j * a.m_stride + (size_t)mi * 5 * a.plane;
    const int lane = tid & 31, wrp = tid >> 5;
    const int tx = tid & 63, tyb = tid >> 6;          // this thread's pixels: (tx, tyb + 4*i), i = 0..3
    float s[4][5];
#pragma unroll
    for (int c = 0; c < 5; c++) {
        const float* P = Min + (size_t)c * a.plane;
        for (int ry = wrp; ry < HP; ry += 8) {
            const float* grow = P + (size_t)clampi(y0 - m + ry, 0, h - 1) * a.pitch;
            for (int rx = lane; rx < WP; rx += 32) sIn[ry * WP + rx] = __ldg(grow + clampi(x0 - m + rx, 0, w - 1));
        }
        __syncthreads();
        for (int idx = tid; idx < TY * WP; idx += 256) {
            const int ty = idx / WP, cx = idx - ty * WP;
            const float* col = sIn + (ty + m) * WP + cx;
            float v = col[0] * a.win.k[0];
            for (int i = 1; i <= m; i++) v = fmaf(col[i * WP] + col[-i * WP], a.win.k[i], v);
            sV[idx] = v;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float* row = sV + (tyb + 4 * i) * WP + tx + m;
            float v = row[0] * a.win.k[0];
            for (int k = 1; k <= m; k++) v = fmaf(row[k] + row[-k], a.win.k[k], v);
            s[i][c] = v * a.win.post_scale;
        }
        __syncthreads();
    }
    const RView R0 = rview(a.R0(j), a.plane), R1 = rview(a.R1(j), a.plane);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int x = x0 + tx, y = y0 + tyb + 4 * i;
        if (x >= w || y >= h) continue;
        const float2 f = solve_fast(s[i][0], s[i][1], s[i][2], s[i][3], s[i][4]);
        if (FUSE) {
            float mm[5];
            update_matrices_core<false>(x, y, f.x, f.y, w, h, R0, R1, a.pitch, mm);
            float* Mo = a.M + (size_t)j * a.m_stride + (size_t)(mi ^ 1) * 5 * a.plane;
            const size_t o = (size_t)y * a.pitch + x;
#pragma unroll
            for (int c = 0; c < 5; c++) Mo[c * a.plane + o] = mm[c];
        } else {
            reinterpret_cast<float2*>(a.out(j))[(size_t)y * w + x] = f;
        }
    }
}

__device__ __forceinline__ int hist_key_fast(float dx, float dy);   // aggregate.cu twin, defined below

// ---------------------------------------------------------------------------------------------------
// FAST path, 3x3 window (winsize 2 or 3): ONE kernel per pyramid layer.  A 32x32 output tile carries a halo of
// NT pixels; the structure matrices M never leave shared memory:
//   stage 0      M0 = updateMatrices(initial flow) on (32+2NT)^2 cells (out-of-image cells = clamped pixel: replicate)
//   iteration i  flow_i = solve(blur3x3(M)) on the region shrunk by 1; if not last: M = updateMatrices(flow_i)
//   last         flow -> HBM (+ the frame's direction/speed histogram on layer 0)
// HBM traffic per pixel: R0 (20 B) + R1 gather (~20 B) + coarse flow (2 B) + flow out (8 B) instead of
// 62 + 80 (NT-1) + 28 B for the per-iteration schedule.
// ---------------------------------------------------------------------------------------------------

template <int NT, bool BOX>
__global__ void __launch_bounds__(256, 3)
flow_layer_kernel(FlowArgs a)
{
    constexpr int T = 32, HALO = NT, RS = T + 2 * HALO, RP = RS + 1;
    constexpr int FS = T + 2 * (HALO - 1);                 // side of the largest intermediate flow region
    __shared__ float sM[5][RS][RP];
    __shared__ float2 sF[NT > 1 ? FS * FS : 1];
    __shared__ unsigned int sH[RC_HIST_CELLS];
    __shared__ unsigned short sKeys[256];      // (direction, bin) cells this CTA touched -> only those are flushed
    __shared__ int sNKeys;
    __shared__ int sSX[RS], sSY[RS];
    __shared__ float sFX[RS], sFY[RS];
    // grid = (pairs, tiles_x, tiles_y): the pair index varies fastest, so the CTAs that read frame t+1's expansion as
    // R1 (pair t) and as R0 (pair t+1) run back to back and the second read hits L2
    const int w = a.w, h = a.h, j = blockIdx.x, tid = threadIdx.x, pitch = a.pitch;
    const int x0 = blockIdx.y * T, y0 = blockIdx.z * T;
    const RView R0 = rview(a.R0(j), a.plane), R1 = rview(a.R1(j), a.plane);
    const bool do_hist = a.hist_delta != nullptr;
    if (do_hist) {
        for (int i = tid; i < RC_HIST_CELLS; i += 256) sH[i] = 0;
        if (tid == 0) sNKeys = 0;
    }
    const float2* coarse = a.coarse ? reinterpret_cast<const float2*>(a.coarse + (size_t)j * a.coarse_stride) : nullptr;
    if (coarse) {
        // resize coefficients of the (clamped) columns / rows of this tile's halo region, once per CTA
        if (tid < RS) resize_coef(clampi(x0 - HALO + tid, 0, w - 1), a.cw, a.sxs, sSX[tid], sFX[tid]);
        else if (tid < 2 * RS) resize_coef(clampi(y0 - HALO + tid - RS, 0, h - 1), a.ch, a.sys, sSY[tid - RS], sFY[tid - RS]);
        __syncthreads();
    }

    // ---- stage 0: M0 on the whole halo region
    for (int idx = tid; idx < RS * RS; idx += 256) {
        const int cy = idx / RS, cx = idx - cy * RS;
        const int x = clampi(x0 - HALO + cx, 0, w - 1), y = clampi(y0 - HALO + cy, 0, h - 1);
        float2 f = make_float2(0.f, 0.f);
        if (coarse) f = upsample_flow_tab(coarse, a.cw, a.ch, sSX[cx], sFX[cx], sSY[cy], sFY[cy], a.fscale);
        float m[5];
        update_matrices_core<false>(x, y, f.x, f.y, w, h, R0, R1, pitch, m);
#pragma unroll
        for (int c = 0; c < 5; c++) sM[c][cy][cx] = m[c];
    }
    __syncthreads();

    // same arithmetic as flow_strip_kernel (unscaled sums, regulariser eps / ps^2): the two forms give identical bits, so
    // results do not depend on which one a launch size selects
    const float k0 = a.win.k[0], k1 = a.win.k[1];
    const float eps_unscaled = 1e-3f / (a.win.post_scale * a.win.post_scale);
#pragma unroll
    for (int it = 0; it < NT; it++) {
        constexpr int dummy = 0; (void)dummy;
        const int hs = HALO - (it + 1);            // halo of this iteration's output region
        const int side = T + 2 * hs, off = HALO - hs;
        const bool last = it == NT - 1;
        // (a) blur + solve: thread = (column, run of L rows); the 3x3 window slides down in registers
        const int runs = 256 / side, L = (side + runs - 1) / runs;
        const int col = tid % side, run = tid / side;
        const int r_begin = run * L, r_end = min(side, r_begin + L);
        const bool active = run < runs && r_begin < side;
        float win[3][3][5];                         // [row][col][channel]
        if (active) {
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int cc = 0; cc < 3; cc++)
#pragma unroll
                    for (int c = 0; c < 5; c++) win[rr + 1][cc][c] = sM[c][off + r_begin - 1 + rr][off + col - 1 + cc];
        }
#pragma unroll
        for (int i = 0; i < L; i++) {
            const int r = r_begin + i;
            const bool ok = active && r < r_end;
            float2 f = make_float2(0.f, 0.f);
            if (ok) {
#pragma unroll
                for (int cc = 0; cc < 3; cc++)
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        win[0][cc][c] = win[1][cc][c]; win[1][cc][c] = win[2][cc][c];
                        win[2][cc][c] = sM[c][off + r + 1][off + col - 1 + cc];
                    }
                float s[5];
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    if (BOX) {
                        const float v0 = win[1][0][c] + (win[0][0][c] + win[2][0][c]);
                        const float v1 = win[1][1][c] + (win[0][1][c] + win[2][1][c]);
                        const float v2 = win[1][2][c] + (win[0][2][c] + win[2][2][c]);
                        s[c] = v1 + (v0 + v2);
                    } else {
                        const float v0 = fmaf(win[0][0][c] + win[2][0][c], k1, win[1][0][c] * k0);
                        const float v1 = fmaf(win[0][1][c] + win[2][1][c], k1, win[1][1][c] * k0);
                        const float v2 = fmaf(win[0][2][c] + win[2][2][c], k1, win[1][2][c] * k0);
                        s[c] = fmaf(v0 + v2, k1, v1 * k0);
                    }
                }
                f = solve_fast_unscaled(s[0], s[1], s[2], s[3], s[4], eps_unscaled);
            }
            if (last) {
                const int x = x0 + col, y = y0 + r;
                const bool in = ok && x < w && y < h;
                if (in) reinterpret_cast<float2*>(a.out(j))[y * w + x] = f;
                if (do_hist) {
                    const int key = in ? hist_key_fast(f.x, f.y) : -1;
                    const unsigned peers = __match_any_sync(0xffffffffu, key);
                    if (key >= 0 && (int)(__ffs(peers) - 1) == (tid & 31)) {
                        if (atomicAdd(&sH[key], __popc(peers)) == 0) {
                            const int slot = atomicAdd(&sNKeys, 1);
                            if (slot < 256) sKeys[slot] = (unsigned short)key;
                        }
                    }
                }
            } else if (ok) {
                sF[r * side + col] = f;
            }
        }
        if (!last) {
            __syncthreads();
            // (b) M = updateMatrices(flow_i) on the same region; out-of-image cells take the clamped pixel
            for (int idx = tid; idx < side * side; idx += 256) {
                const int cy = idx / side, cx = idx - cy * side;
                const int x = clampi(x0 - hs + cx, 0, w - 1), y = clampi(y0 - hs + cy, 0, h - 1);
                const float2 f = sF[(y - (y0 - hs)) * side + (x - (x0 - hs))];
                float m[5];
                update_matrices_core<false>(x, y, f.x, f.y, w, h, R0, R1, pitch, m);
#pragma unroll
                for (int c = 0; c < 5; c++) sM[c][off + cy][off + cx] = m[c];
            }
            __syncthreads();
        }
    }
    if (do_hist) {
        __syncthreads();
        unsigned int* dst = a.hist_delta + (size_t)j * RC_HIST_CELLS;
        const int nk = sNKeys;
        if (nk <= 256) {
            if (tid < nk) { const int k = sKeys[tid]; atomicAdd(&dst[k], sH[k]); }
        } else {
            for (int i = tid; i < RC_HIST_CELLS; i += 256)
                if (sH[i]) atomicAdd(&dst[i], sH[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// FAST path, windows of half-width M = 2, 5, 10 (winsize 5 of the Android fork, 10 of main.cpp:1119,1481, 20/21 of
// main.cpp:609,961 and of the 4K configuration): one updateFlow iteration per launch, MARCHING formulation.
// A CTA owns a strip of TX = 64 columns and walks down a segment of rows in steps of RB = 16:
//   stage   RB rows of the five M planes (TX + 2M columns, replicate-clamped) into shared memory  -- every load of
//           the step is issued before any is consumed;
//   h-blur  thread = 4 adjacent pixels of one (row, channel): 16-byte shared loads, result into a ring of RB + 2M
//           horizontally blurred rows;
//   v-blur  thread = (column, 4 rows): its 4 + 2M ring values per channel are read once into registers, the five sums
//           of its four pixels stay in registers for the fp32 solve, then the fused updateMatrices (-> M of the next
//           iteration) or the flow store (+ the frame's direction/speed histogram on layer 0).
// Compared with square tiles nothing is recomputed vertically (2M warm-up rows per segment instead of 2M per 16 rows),
// all 256 threads work in every phase, and there are two barriers per 16 rows.  Box windows use running sums.
// ---------------------------------------------------------------------------------------------------
// FIRST: the first iteration of a layer.  M does not exist yet: the staged rows are COMPUTED (flow initialisation from
// the coarser layer + updateMatrices, at the replicate-clamped pixel) instead of copied, which removes the separate
// updateMatrices launch and its 20 B/px write + 20 B/px read.
// NTHR = 256 (16 rows per step) or 128 (8 rows per step: half the shared memory per CTA, twice the resident CTAs -- more
// independent phase machines per SM to cover each other's updateMatrices gathers and barriers; RC_MARCH_THREADS)
// Staging of M rows (iterations after the first): interior steps -- every staged column and row inside the image -- are ONE
// cp.async.bulk.tensor.3d (TMA) per step, issued by one thread behind an mbarrier: box = WPA columns x RB rows x 5 planes of
// the layer's M tensor [pair][buffer][plane][row][column].  Steps that touch the image border (replicate clamping, which the
// copy engine's zero fill cannot express) and contexts without a tensor map use per-element cp.async.  ncu before this:
// a fifth of the final iteration's instructions (and 78 % of its issue slots were busy) were staging address arithmetic.
template <int M, bool FUSE, bool BOX, bool FIRST, int NTHR>
__global__ void __launch_bounds__(NTHR, NTHR == 256 ? (FUSE ? 3 : 2) : (FUSE ? 6 : 4))
flow_march_kernel(FlowArgs a, int mi, int SEG, const __grid_constant__ CUtensorMap tmM, int use_tma)
{
    // D: the staged rows start at column x0 - M - D, the first 16-byte aligned column at or before x0 - M (the TMA copy
    // needs that alignment; the per-element path then copies two floats per cp.async); the blurs read from column D on
    constexpr int TX = 64, RB = NTHR / 16, NWARP = NTHR / 32, D = ((M + 3) & ~3) - M, WP = TX + 2 * M, WPA = (WP + D + 3) & ~3,
                  RING = RB + 2 * M;
    extern __shared__ __align__(16) float msm_raw[];
    // the TMA destination must be 128-byte aligned; static shared memory precedes the dynamic part, so align by hand
    float* msm = msm_raw + ((128u - (smem_u32(msm_raw) & 127u)) & 127u) / 4;
    float* sRaw = msm;                         // [5][RB][WPA]
    float* sRing = msm + RB * 5 * WPA;         // [RING][5][TX]
    __shared__ __align__(8) unsigned long long tbar;
    __shared__ int sColX[FIRST ? WP : 1], sColS[FIRST ? WP : 1];       // first iteration: clamped column, coarse column,
    __shared__ float sColF[FIRST ? WP : 1];                            // coarse interpolation weight of every staged column
    __shared__ unsigned int sH[FUSE ? 1 : RC_HIST_CELLS];
    __shared__ unsigned short sKeys[FUSE ? 1 : 256];
    __shared__ int sNKeys;
    const int w = a.w, h = a.h, j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    const int x0 = blockIdx.y * TX, y0 = blockIdx.z * SEG;
    const float* Min = a.M + (size_t)j * a.m_stride + (size_t)mi * 5 * a.plane;
    const bool do_hist = !FUSE && a.hist_delta != nullptr;
    if (do_hist) {
        for (int i = tid; i < RC_HIST_CELLS; i += NTHR) sH[i] = 0;
        if (tid == 0) sNKeys = 0;
    }
    float kk[M + 1];
#pragma unroll
    for (int i = 0; i <= M; i++) kk[i] = a.win.k[i];
    // window sums stay unscaled: the solve only needs the regulariser rescaled (solve_fast_unscaled)
    const float eps_unscaled = 1e-3f / (a.win.post_scale * a.win.post_scale);
    const RView R0 = rview(a.R0(j), a.plane), R1 = rview(a.R1(j), a.plane);
    const int nout = min(SEG, h - y0), total = nout + 2 * M;          // staged row s <-> image row y0 - M + s (clamped)
    const int col = tid & 63, rbk = tid >> 6;                          // v-blur role
    int ns = 0, no = 0;
    // stage: warp per (row, channel); asynchronous copies (global -> shared without registers), so the rows of step b+1
    // are in flight while step b is blurred vertically and solved
    const bool interior_x = x0 - M - D >= 0 && x0 - M + WP + 1 <= w;
    const bool tma_x = !FIRST && use_tma && interior_x;
    const unsigned long long tm_addr = reinterpret_cast<unsigned long long>(&tmM);     // param space (__grid_constant__)
    bool staged_tma = false;
    unsigned tphase = 0;
    if (!FIRST && use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&tbar)));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const float2* coarse = FIRST && a.coarse ? reinterpret_cast<const float2*>(a.coarse + (size_t)j * a.coarse_stride) : nullptr;
    if (FIRST) {
        for (int i = tid; i < WP; i += NTHR) {
            const int xc = clampi(x0 - M + i, 0, w - 1);
            int cs = 0; float cf = 0.f;
            if (coarse) resize_coef(xc, a.cw, a.sxs, cs, cf);
            sColX[i] = xc; sColS[i] = cs; sColF[i] = cf;
        }
        __syncthreads();
    }
    auto stage = [&](int first, int cnt) {
        if constexpr (FIRST) {
            // item = one staged (row, column): the cnt * WP items of the step are dealt to the threads in order, so every
            // round but the last has all lanes busy (a warp-per-row split leaves WP mod 32 lanes in its last round)
            for (int idx = tid; idx < cnt * WP; idx += NTHR) {
                const int r = idx / WP, rx = idx - r * WP;
                const int y = clampi(y0 - M + first + r, 0, h - 1);
                const int x = sColX[rx];
                float2 fi = make_float2(0.f, 0.f);
                if (coarse) {
                    int csy; float cfy;
                    resize_coef(y, a.ch, a.sys, csy, cfy);
                    fi = upsample_flow_tab(coarse, a.cw, a.ch, sColS[rx], sColF[rx], csy, cfy, a.fscale);
                }
                float mm[5];
                update_matrices_core<false>(x, y, fi.x, fi.y, w, h, R0, R1, a.pitch, mm);
#pragma unroll
                for (int c = 0; c < 5; c++) sRaw[(c * RB + r) * WPA + D + rx] = mm[c];
            }
        } else {
            const int ytop = y0 - M + first;
            staged_tma = tma_x && ytop >= 0 && ytop + cnt <= h;                // CTA-uniform
            if (staged_tma) {
                if (tid == 0) {
                    constexpr unsigned BYTES = 5 * RB * WPA * sizeof(float);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar)), "r"(BYTES) : "memory");
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(smem_u32(sRaw)), "l"(tm_addr), "r"(x0 - M - D), "r"(ytop), "r"((j * 2 + mi) * 5), "r"(smem_u32(&tbar))
                                 : "memory");
                }
                return;
            }
            for (int rc = wrp; rc < cnt * 5; rc += NWARP) {
                const int r = rc / 5, c = rc - 5 * r;
                const float* grow = Min + (size_t)c * a.plane + (size_t)clampi(y0 - M + first + r, 0, h - 1) * a.pitch;
                float* dst = sRaw + (c * RB + r) * WPA;
                if (interior_x) {                          // no column clamping, 8-byte aligned: half the copies
#pragma unroll
                    for (int q = 0; q < ((WP + D + 1) / 2 + 31) / 32; q++) {
                        const int rx = lane + 32 * q;
                        if (rx < (WP + D + 1) / 2) __pipeline_memcpy_async(dst + 2 * rx, grow + (x0 - M - D) + 2 * rx, 8);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < (WP + 31) / 32; q++) {
                        const int rx = lane + 32 * q;
                        if (rx < WP) __pipeline_memcpy_async(dst + D + rx, grow + clampi(x0 - M + rx, 0, w - 1), 4);
                    }
                }
            }
            __pipeline_commit();
        }
    };
    if (!FIRST) stage(0, min(RB, total));
    while (ns < total) {
        const int cnt = min(RB, total - ns);
        if (FIRST) stage(ns, cnt);
        else if (staged_tma) {
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&tbar)), "r"(tphase) : "memory");
            tphase ^= 1;
        } else __pipeline_wait_prior(0);
        __syncthreads();
        // ---- horizontal blur: item = (channel, row, group of 4 pixels); staged row rc = c * RB + r
        for (int it = tid; it < 5 * RB * (TX / 4); it += NTHR) {
            const int xg = it & 15, rc = it >> 4;
            const int c = rc / RB, r = rc - c * RB;
            if (r >= cnt) continue;
            float wbuf[(4 + 2 * M + D + 3) / 4 * 4];
#pragma unroll
            for (int q = 0; q < (4 + 2 * M + D + 3) / 4; q++)
                *reinterpret_cast<float4*>(wbuf + 4 * q) = *reinterpret_cast<const float4*>(sRaw + rc * WPA + 4 * xg + 4 * q);
            const float* win = wbuf + D;
            float o[4];
            if (BOX) {
                float run = win[0];
#pragma unroll
                for (int i = 1; i <= 2 * M; i++) run += win[i];
                o[0] = run;
#pragma unroll
                for (int i = 1; i < 4; i++) { run += win[i + 2 * M] - win[i - 1]; o[i] = run; }
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float v = win[i + M] * kk[0];
#pragma unroll
                    for (int k = 1; k <= M; k++) v = fmaf(win[i + M + k] + win[i + M - k], kk[k], v);
                    o[i] = v;
                }
            }
            *reinterpret_cast<float4*>(sRing + (((ns + r) % RING) * 5 + c) * TX + 4 * xg) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncthreads();
        ns += cnt;
        if (!FIRST && ns < total) stage(ns, min(RB, total - ns));
        // ---- vertical blur + solve for the output rows whose window is complete: no .. min(ns - 2M, nout) - 1
        const int lim = min(ns - 2 * M, nout);
        const int o0 = no + 4 * rbk;                                    // this thread's rows o0 .. o0 + 3 (local)
        if (o0 < lim) {
            float s[4][5];
            int slot0 = o0 % RING;
#pragma unroll
            for (int c = 0; c < 5; c++) {
                float win[4 + 2 * M];
                int sl = slot0;
#pragma unroll
                for (int i = 0; i < 4 + 2 * M; i++) {
                    win[i] = sRing[(sl * 5 + c) * TX + col];
                    sl = sl + 1 == RING ? 0 : sl + 1;
                }
                if (BOX) {
                    float run = win[0];
#pragma unroll
                    for (int i = 1; i <= 2 * M; i++) run += win[i];
                    s[0][c] = run;
#pragma unroll
                    for (int i = 1; i < 4; i++) { run += win[i + 2 * M] - win[i - 1]; s[i][c] = run; }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        float v = win[i + M] * kk[0];
#pragma unroll
                        for (int k = 1; k <= M; k++) v = fmaf(win[i + M + k] + win[i + M - k], kk[k], v);
                        s[i][c] = v;
                    }
                }
            }
            const int x = x0 + col;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int y = y0 + o0 + i;
                const bool in = x < w && o0 + i < lim;
                float2 f = make_float2(0.f, 0.f);
                if (in) f = solve_fast_unscaled(s[i][0], s[i][1], s[i][2], s[i][3], s[i][4], eps_unscaled);
                if (FUSE) {
                    if (in) {
                        float mm[5];
                        update_matrices_core<false>(x, y, f.x, f.y, w, h, R0, R1, a.pitch, mm);
                        float* Mo = a.M + (size_t)j * a.m_stride + (size_t)(mi ^ 1) * 5 * a.plane;
                        const size_t o = (size_t)y * a.pitch + x;
#pragma unroll
                        for (int c = 0; c < 5; c++) Mo[c * a.plane + o] = mm[c];
                    }
                } else {
                    if (in) reinterpret_cast<float2*>(a.out(j))[(size_t)y * w + x] = f;
                    if (do_hist) {
                        const int key = in ? hist_key_fast(f.x, f.y) : -1;
                        const unsigned peers = __match_any_sync(0xffffffffu, key);
                        if (key >= 0 && (int)(__ffs(peers) - 1) == lane) {
                            if (atomicAdd(&sH[key], __popc(peers)) == 0) {
                                const int slot = atomicAdd(&sNKeys, 1);
                                if (slot < 256) sKeys[slot] = (unsigned short)key;
                            }
                        }
                    }
                }
            }
        }
        no = max(lim, 0);
    }
    if (do_hist) {
        __syncthreads();
        unsigned int* dst = a.hist_delta + (size_t)j * RC_HIST_CELLS;
        const int nk = sNKeys;
        if (nk <= 256) {
            for (int t = tid; t < nk; t += NTHR) { const int k = sKeys[t]; atomicAdd(&dst[k], sH[k]); }
        } else {
            for (int i = tid; i < RC_HIST_CELLS; i += NTHR)
                if (sH[i]) atomicAdd(&dst[i], sH[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// FAST path, 3x3 window, register/shuffle formulation of the same fused layer: a WARP owns a 32-column strip
// (32 - 2*NT useful columns) and marches down a segment of rows.  Lane = column.  The last two rows of M of every
// iteration level live in registers; the vertical 3-sum is formed from them, the horizontal one with two shuffles
// per channel -- no shared memory for M, no block barriers, every warp independent.  Replicate borders: level-0 rows
// are evaluated at clamped pixels, deeper levels duplicate their first/last real row and copy the flow of the
// clamped column before updateMatrices, which reproduces M at the clamped pixel exactly.
// ---------------------------------------------------------------------------------------------------
// The two previous rows of M of every level live in per-warp shared memory slots (row r in slot r & 1): 64 registers,
// 32 warps per SM.  (Variants that kept them in registers and prefetched the gathers were measured and dropped: fewer
// instructions but half the occupancy)
template <int NT, bool BOX>
__global__ void __launch_bounds__(256, 4)
flow_strip_kernel(FlowArgs a, int SEG)
{
    constexpr int UW = 32 - 2 * NT;
    __shared__ unsigned int sH[RC_HIST_CELLS];
    __shared__ float sWin[8][NT][2][5][32];
    const int w = a.w, h = a.h, j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, pitch = a.pitch;
    const bool do_hist = a.hist_delta != nullptr;
    if (do_hist) {
        for (int i = tid; i < RC_HIST_CELLS; i += 256) sH[i] = 0;
        __syncthreads();
    }
    const int sx0 = (blockIdx.y * 8 + wrp) * UW;             // grid = (pairs, strip groups, segments), pair fastest
    const int y0 = blockIdx.z * SEG;
    if (sx0 < w) {
        const int xv = sx0 - NT + lane;
        const int x = clampi(xv, 0, w - 1);
        const bool col_out = lane >= NT && lane < 32 - NT && xv < w;
        const bool need_fix = (sx0 - NT < 0) || (sx0 - NT + 31 >= w);
        const int src_lane = x - (sx0 - NT);                 // lane that owns the clamped column
        const RView R0 = rview(a.R0(j), a.plane), R1 = rview(a.R1(j), a.plane);
        const float2* coarse = a.coarse ? reinterpret_cast<const float2*>(a.coarse + (size_t)j * a.coarse_stride) : nullptr;
        int csx = 0; float cfx = 0.f;
        if (coarse) resize_coef(x, a.cw, a.sxs, csx, cfx);
        const float k0 = a.win.k[0], k1 = a.win.k[1];
        const float eps_unscaled = 1e-3f / (a.win.post_scale * a.win.post_scale);     // post_scale is 1 for Gaussian windows
        float2* outp = reinterpret_cast<float2*>(a.out(j));

        int nfed[NT];
#pragma unroll
        for (int it = 0; it < NT; it++) nfed[it] = 0;
        const int vbeg = max(y0 - NT, 0);
        const int vend = min(y0 + SEG - 1 + NT, h + NT - 1);
        auto init_flow = [&](int r) {
            float2 fi = make_float2(0.f, 0.f);
            if (coarse) {
                int csy; float cfy;
                resize_coef(r, a.ch, a.sys, csy, cfy);
                fi = upsample_flow_tab(coarse, a.cw, a.ch, csx, cfx, csy, cfy, a.fscale);
            }
            return fi;
        };
        // emits row yo of the final flow (+ its histogram contribution)
        auto emit = [&](int yo, float2 f) {
            if (col_out) outp[yo * w + xv] = f;
            if (do_hist) {
                const int key = col_out ? hist_key_fast(f.x, f.y) : -1;
                const unsigned peers = __match_any_sync(0xffffffffu, key);
                if (key >= 0 && (int)(__ffs(peers) - 1) == lane) atomicAdd(&sH[key], __popc(peers));
            }
        };
        // blur of level `it` at row r - 1 from the two stored rows and the new one, then the 2x2 solve
        auto blur_solve = [&](int it, int r, const float (&C)[5]) {
            float sv[5];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const float ra = sWin[wrp][it][r & 1][c][lane];          // row r-2
                const float rb = sWin[wrp][it][(r - 1) & 1][c][lane];    // row r-1
                const float vs = BOX ? rb + (ra + C[c]) : fmaf(ra + C[c], k1, rb * k0);
                const float lft = __shfl_up_sync(0xffffffffu, vs, 1), rgt = __shfl_down_sync(0xffffffffu, vs, 1);
                sv[c] = BOX ? vs + (lft + rgt) : fmaf(lft + rgt, k1, vs * k0);
            }
            return solve_fast_unscaled(sv[0], sv[1], sv[2], sv[3], sv[4], eps_unscaled);
        };
        // Rows [vs0, vs1] are the steady state: every level has its two previous rows, every row index is inside the
        // image and every iteration emits one output row, so the loop body needs none of the warm-up / bottom-edge tests.
        const int vs0 = vbeg + 2 * NT + 2, vs1 = min(vend, h - 1);
#pragma unroll 1
        for (int v = vbeg; v <= vend; v++) {
            if (v >= vs0 && v <= vs1) {
                float2 f;
#pragma unroll
                for (int it = 0; it < NT; it++) {
                    const int r = v - it;
                    float C[5];
                    if (it == 0) {
                        const float2 fi = init_flow(r);
                        update_matrices_core<false>(x, r, fi.x, fi.y, w, h, R0, R1, pitch, C);
                    } else {
                        if (need_fix) { f.x = __shfl_sync(0xffffffffu, f.x, src_lane); f.y = __shfl_sync(0xffffffffu, f.y, src_lane); }
                        update_matrices_core<false>(x, r, f.x, f.y, w, h, R0, R1, pitch, C);
                    }
                    f = blur_solve(it, r, C);
#pragma unroll
                    for (int c = 0; c < 5; c++) sWin[wrp][it][r & 1][c][lane] = C[c];
                    nfed[it]++;
                }
                emit(v - NT, f);
                continue;
            }
            bool produced = false;
            float2 f = make_float2(0.f, 0.f);
#pragma unroll
            for (int it = 0; it < NT; it++) {
                const int r = v - it;
                float C[5];
                bool have = false;
                if (r >= 0 && r < h) {
                    if (it == 0) {
                        const float2 fi = init_flow(r);
                        update_matrices_core<false>(x, r, fi.x, fi.y, w, h, R0, R1, pitch, C);
                        have = true;
                    } else if (produced) {
                        if (need_fix) { f.x = __shfl_sync(0xffffffffu, f.x, src_lane); f.y = __shfl_sync(0xffffffffu, f.y, src_lane); }
                        update_matrices_core<false>(x, r, f.x, f.y, w, h, R0, R1, pitch, C);
                        have = true;
                    }
                } else if (r == h && nfed[it] > 0) {
#pragma unroll
                    for (int c = 0; c < 5; c++) C[c] = sWin[wrp][it][(r - 1) & 1][c][lane];
                    have = true;
                }
                produced = false;
                if (!have) continue;
                if (nfed[it] == 0) {
#pragma unroll
                    for (int c = 0; c < 5; c++) { sWin[wrp][it][0][c][lane] = C[c]; sWin[wrp][it][1][c][lane] = C[c]; }
                    nfed[it] = (r == 0) ? 2 : 1;
                    continue;
                }
                if (nfed[it] >= 2) {
                    f = blur_solve(it, r, C);
                    produced = true;
                }
#pragma unroll
                for (int c = 0; c < 5; c++) sWin[wrp][it][r & 1][c][lane] = C[c];
                nfed[it]++;
            }
            const int yo = v - NT;
            if (produced && yo >= y0 && yo < y0 + SEG && yo < h) emit(yo, f);      // warp-uniform
        }
    }
    if (do_hist) {
        __syncthreads();
        unsigned int* dst = a.hist_delta + (size_t)j * RC_HIST_CELLS;
        for (int i = tid; i < RC_HIST_CELLS; i += 256)
            if (sH[i]) atomicAdd(&dst[i], sH[i]);
    }
}

// direction/speed key of one flow vector: bit-exact twin of aggregate.cu's hist_key (cv::cartToPolar restated,
// SURVEY.md section 8(c)); intrinsics keep it independent of this file's FMA contraction.
__device__ __forceinline__ int hist_key_fast(float x, float y)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float c = __fdiv_rn(mn, __fadd_rn(mx, 2.220446049250313e-16f));
    float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1), c);
    if (ax < ay) a = __fsub_rn(90.f, a);
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    float mag = __fsqrt_rn(fmaf(x, x, __fmul_rn(y, y)));
    int bin = (int)__fmul_rn(mag, (float)RC_HIST_RESOLUTION);
    // (int)(a * 36 / 360) with the division as multiply + two FMAs (Markstein): the truncated result was checked against
    // IEEE division for every float in [0, 12960] (it differs only for denormal quotients, which truncate to 0 either way)
    const float t = __fmul_rn(a, (float)RC_HIST_DIRECTIONS), rc360 = 1.0f / 360.0f;
    const float q0 = __fmul_rn(t, rc360);
    int dir = (int)__fmaf_rn(__fmaf_rn(-q0, 360.f, t), rc360, q0);
    if (bin < RC_HIST_BINS && bin >= 0) return dir * RC_HIST_BINS + bin;
    return -1;
}

// histogram of a batch of dense flows (used when the final iteration ran in the unfused path)
__global__ void __launch_bounds__(256)
hist_batch_kernel(FlowArgs a, int n)
{
    __shared__ unsigned int sH[RC_HIST_CELLS];
    const int j = blockIdx.y;
    for (int i = threadIdx.x; i < RC_HIST_CELLS; i += 256) sH[i] = 0;
    __syncthreads();
    const float2* f = reinterpret_cast<const float2*>(a.out(j));
    const int stride = gridDim.x * 256;
    const int nround = (n + stride - 1) / stride * stride;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < nround; i += stride) {
        int key = -1;
        if (i < n) { float2 v = f[i]; key = hist_key_fast(v.x, v.y); }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sH[key], __popc(peers));
    }
    __syncthreads();
    unsigned int* dst = a.hist_delta + (size_t)j * RC_HIST_CELLS;
    for (int i = threadIdx.x; i < RC_HIST_CELLS; i += 256)
        if (sH[i]) atomicAdd(&dst[i], sH[i]);
}

}  // namespace

// ===================================================================================================
// launchers
// ===================================================================================================
// Opt-in dynamic shared memory sizes: set ONCE per device (cudaFuncSetAttribute is per device), from rc_create, under a
// std::once_flag -- contexts on several GPUs may be created and driven from different threads.
void rc_farneback_init_device(int device)
{
    static std::once_flag once[64];
    std::call_once(once[device & 63], [] {
        const int strict_max = (int)(sizeof(float) * ((size_t)(32 + 2 * RC_MAX_POLY_N) * (64 + 2 * RC_MAX_POLY_N) +
                                                      3 * (size_t)32 * (64 + 2 * RC_MAX_POLY_N)));
        cudaFuncSetAttribute(polyexp_strict_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, strict_max);
        cudaFuncSetAttribute(pyr_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(polyexp_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32 * 128 * 4 + 16);
        const int tiled_max = (int)(sizeof(float) * ((size_t)(16 + 32) * (64 + 32) + 16 * (size_t)(64 + 32)));
        cudaFuncSetAttribute(flow_iter_tiled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tiled_max);
        cudaFuncSetAttribute(flow_iter_tiled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tiled_max);
        const int big = (int)(128 + sizeof(float) * (16 * 5 * 88 + 36 * 5 * 64));
#define RC_CFG1(MM, FU, BX) \
    cudaFuncSetAttribute(flow_march_kernel<MM, FU, BX, false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big); \
    cudaFuncSetAttribute(flow_march_kernel<MM, FU, BX, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big); \
    cudaFuncSetAttribute(flow_march_kernel<MM, FU, BX, false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big); \
    cudaFuncSetAttribute(flow_march_kernel<MM, FU, BX, true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big)
#define RC_CFG(MM) RC_CFG1(MM, true, true); RC_CFG1(MM, true, false); RC_CFG1(MM, false, true); RC_CFG1(MM, false, false)
        RC_CFG(2); RC_CFG(5); RC_CFG(10);
#undef RC_CFG
#undef RC_CFG1
        cudaGetLastError();
    });
}

// cuTensorMapEncodeTiled through the runtime's driver entry point, so that the library keeps linking cudart only
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(fn);
    }();
    return encode;
}

static void launch_polyexp(rc_ctx* c, Layer& L, int nb, int first_slot)
{
    const int nslots = c->B + 1;
    const size_t istride = (size_t)L.pitch * L.h;
    KScope ks(c, K_POLYEXP, 24.0 * L.w * L.h * nb);
    int np = c->strict ? 0 : (c->poly.n_eff + 3) / 4 * 4;
    const bool seven = !c->strict && c->poly.n_eff == 7;                      // tile radius 8, seven taps evaluated
    if (!c->strict && c->poly.n_eff > 4 && c->poly.n_eff <= 6) np = 6;      // taps per side: 4, 6, 8, 12 or 16
    if (np > 16) np = 0;
    static const bool packed = getenv("RC_POLYEXP") && !strcmp(getenv("RC_POLYEXP"), "packed");
    static const bool use_tma = getenv("RC_POLYEXP") && !strcmp(getenv("RC_POLYEXP"), "tma");
    if (np == 8 && use_tma) {
        // tensor map of this layer's images [nb][h][w] (row pitch L.pitch); driver entry point resolved at run time so that
        // the library keeps linking cudart only
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = [] {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult q;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess) fn = nullptr;
            return reinterpret_cast<EncodeFn>(fn);
        }();
        CUtensorMap tm;
        const cuuint64_t gdim[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)c->B};
        const cuuint64_t gstr[2] = {(cuuint64_t)L.pitch * 4, (cuuint64_t)L.pitch * L.h * 4};
        const cuuint32_t box[3] = {128, 32 + 2 * 8, 1}, estr[3] = {1, 1, 1};
        if (encode && encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L.I, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
            PolyCoefF pf;
            for (int i = 0; i <= RC_MAX_POLY_N; i++) {      // taps beyond n_eff are dropped in every fast variant
                const bool on = i <= c->poly.n_eff;
                pf.g[i] = on ? c->poly.g[i] : 0.f; pf.xg[i] = on ? c->poly.xg[i] : 0.f; pf.xxg[i] = on ? c->poly.xxg[i] : 0.f;
            }
            pf.ig11 = (float)c->poly.ig11; pf.ig03 = (float)c->poly.ig03; pf.ig33 = (float)c->poly.ig33; pf.ig55 = (float)c->poly.ig55;
            const int TX = 128 - 2 * 8;
            dim3 g((L.w + TX - 1) / TX, (L.h + 31) / 32, nb);
            const size_t smem = (size_t)3 * 32 * 128 * 4 + 16;
            polyexp_tma_kernel<8><<<g, 256, smem, c->stream>>>(tm, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, pf);
            return;
        }
    }
    if (np && packed && np % 4 == 0) {
        PolyCoefF2 p2;
        for (int i = 0; i <= 16; i++) {
            const float g = i <= c->poly.n_eff ? c->poly.g[i] : 0.f, xg = i <= c->poly.n_eff ? c->poly.xg[i] : 0.f;
            const float xxg = i <= c->poly.n_eff ? c->poly.xxg[i] : 0.f;
            p2.g[i] = make_float2(g, g); p2.xg[i] = make_float2(xg, xg); p2.xxg[i] = make_float2(xxg, xxg);
        }
        p2.ig11 = (float)c->poly.ig11; p2.ig03 = (float)c->poly.ig03; p2.ig33 = (float)c->poly.ig33; p2.ig55 = (float)c->poly.ig55;
        const int TX = 128 - 2 * np;
        dim3 g((L.w + TX - 1) / TX, (L.h + 31) / 32, nb);
        switch (np) {
        case 4: polyexp_packed_kernel<4><<<g, 256, 0, c->stream>>>(L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, p2); break;
        case 8: polyexp_packed_kernel<8><<<g, 256, 0, c->stream>>>(L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, p2); break;
        case 12: polyexp_packed_kernel<12><<<g, 256, 0, c->stream>>>(L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, p2); break;
        default: polyexp_packed_kernel<16><<<g, 256, 0, c->stream>>>(L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, p2); break;
        }
        return;
    }
    if (np) {
        PolyCoefF pf;
        for (int i = 0; i <= RC_MAX_POLY_N; i++) {
            const bool on = i <= c->poly.n_eff;
            pf.g[i] = on ? c->poly.g[i] : 0.f; pf.xg[i] = on ? c->poly.xg[i] : 0.f; pf.xxg[i] = on ? c->poly.xxg[i] : 0.f;
        }
        pf.ig11 = (float)c->poly.ig11; pf.ig03 = (float)c->poly.ig03; pf.ig33 = (float)c->poly.ig33; pf.ig55 = (float)c->poly.ig55;
        const int TX = 128 - 2 * np;
        // 16-row tiles (128-thread CTAs, 24 KB of shared memory, eight resident CTAs) overlap the load-latency phase of one
        // CTA with the arithmetic of the others better than 32-row tiles: 422 vs 446 us per launch (RC_POLYEXP_ROWS=32: old form)
        static const int rows = getenv("RC_POLYEXP_ROWS") ? atoi(getenv("RC_POLYEXP_ROWS")) : 16;
#define RC_PX(NPV) \
    do { if (rows == 16) polyexp_fast_kernel<NPV, 16><<<dim3((L.w + TX - 1) / TX, (L.h + 15) / 16, nb), 128, 0, c->stream>>>( \
                 L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, pf); \
         else polyexp_fast_kernel<NPV, 32><<<dim3((L.w + TX - 1) / TX, (L.h + 31) / 32, nb), 256, 0, c->stream>>>( \
                 L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, pf); } while (0)
        if (seven && rows == 16) {
            polyexp_fast_kernel<8, 16, 7><<<dim3((L.w + TX - 1) / TX, (L.h + 15) / 16, nb), 128, 0, c->stream>>>(
                L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot, nslots, pf);
            return;
        }
        switch (np) {
        case 4: RC_PX(4); break;
        case 6: RC_PX(6); break;
        case 8: RC_PX(8); break;
        case 12: RC_PX(12); break;
        default: RC_PX(16); break;
        }
#undef RC_PX
        return;
    }
    constexpr int TX = 64, TY = 32;
    const int n = c->poly.n;
    const size_t smem = sizeof(float) * ((size_t)(TY + 2 * n) * (TX + 2 * n) + 3 * (size_t)TY * (TX + 2 * n));
    dim3 g((L.w + TX - 1) / TX, (L.h + TY - 1) / TY, nb);
    polyexp_strict_kernel<TX, TY><<<g, 256, smem, c->stream>>>(L.I, istride, L.w, L.h, L.pitch, L.R, L.plane, first_slot,
                                                              nslots, c->poly);
}

void rc_launch_expand(rc_ctx* c, const uint8_t* d_frames, size_t step, size_t fstride, int nb, int first_slot, bool overlap)
{
    const int W = c->prm.w, H = c->prm.h;
    int k_first = 0;
    c->expand_overlapped = false;
    // layers 0-2 of the default pyramid in one pass over the frame (RC_PYR=separate: one kernel per layer)
    static const bool pyr_fused = !(getenv("RC_PYR") && !strcmp(getenv("RC_PYR"), "separate"));
    if (pyr_fused && c->nlayers >= 3 && W % 4 == 0 && H % 4 == 0 && W >= 16 && H >= 16 && step % 4 == 0 && fstride % 4 == 0 &&
        (reinterpret_cast<size_t>(d_frames) & 3) == 0 && c->layer[0].w == W && c->layer[0].h == H &&
        c->layer[1].w * 2 == W && c->layer[1].h * 2 == H && c->layer[2].w * 4 == W && c->layer[2].h * 4 == H &&
        c->layer[0].smooth.ksize == 3 && c->layer[1].smooth.ksize == 3 && c->layer[2].smooth.ksize == 9) {
        Pyr3Args p;
        p.img = d_frames; p.step = step; p.fstride = fstride; p.W = W; p.H = H;
        double bytes = (double)W * H;
        for (int k = 0; k < 3; k++) {
            Layer& L = c->layer[k];
            p.out[k] = L.I; p.pitch[k] = L.pitch; p.ostride[k] = (size_t)L.pitch * L.h;
            bytes += 4.0 * L.w * L.h;
        }
        for (int i = 0; i < 3; i++) { p.k3a[i] = c->layer[0].smooth.k[i]; p.k3b[i] = c->layer[1].smooth.k[i]; }
        p.k9 = c->layer[2].smooth;
        {
            KScope ks(c, K_PYR_V, bytes * nb);
            pyr3_kernel<<<dim3((W + 127) / 128, (H + 31) / 32, nb), 256, 0, c->stream>>>(p);
        }
        if (overlap && c->nlayers == 3 && c->s_aux) {
            // small batches (frame-by-frame use): the 480x270 and 960x540 layers of one frame do not fill 148 SMs, so the
            // expansions of layers 1 and 0 run on a second stream while the main stream expands layer 2 and starts the
            // coarse-to-fine flow; rc_launch_flows waits for ev_poly[k] before the flow of layer k
            cudaEventRecord(c->ev_pyr, c->stream);
            cudaStreamWaitEvent(c->s_aux, c->ev_pyr, 0);
            cudaStream_t keep = c->stream;
            c->stream = c->s_aux;
            for (int k = 1; k >= 0; k--) {
                launch_polyexp(c, c->layer[k], nb, first_slot);
                cudaEventRecord(c->ev_poly[k], c->s_aux);
            }
            c->stream = keep;
            launch_polyexp(c, c->layer[2], nb, first_slot);
            c->expand_overlapped = true;
        } else {
            for (int k = 0; k < 3; k++) launch_polyexp(c, c->layer[k], nb, first_slot);
        }
        k_first = 3;
    }
    for (int k = k_first; k < c->nlayers; k++) {
        Layer& L = c->layer[k];
        PyrArgs a;
        a.img = d_frames; a.step = step; a.fstride = fstride; a.W = W; a.H = H; a.dw = L.w; a.dh = L.h;
        a.two = !(L.w == W && L.h == H);
        a.sxs = 1.0 / ((double)L.w / (double)W); a.sys = 1.0 / ((double)L.h / (double)H);
        a.out = L.I; a.pitch = L.pitch; a.ostride = (size_t)L.pitch * L.h;
        const int r = L.smooth.ksize / 2;
        if (!a.two && L.smooth.ksize == 3 && W % 4 == 0 && W >= 8 && H >= 2 && step % 4 == 0 && fstride % 4 == 0 &&
            (reinterpret_cast<size_t>(d_frames) & 3) == 0) {
            dim3 g((W / 4 + 63) / 64, (H + 3) / 4, nb);
            {
                KScope ks(c, K_PYR_V, ((double)W * H + 4.0 * L.w * L.h) * nb);
                pyr0_kernel<<<g, 256, 0, c->stream>>>(d_frames, step, fstride, W, H, L.smooth.k[0], L.smooth.k[1],
                                                     L.smooth.k[2], L.I, L.pitch, (size_t)L.pitch * L.h);
            }
            launch_polyexp(c, L, nb, first_slot);
            continue;
        }
        if (a.two && L.smooth.ksize == 3 && W == 2 * L.w && H == 2 * L.h && W % 4 == 0 && W >= 8 && H >= 4 && step % 4 == 0 &&
            fstride % 4 == 0 && (reinterpret_cast<size_t>(d_frames) & 3) == 0) {
            dim3 g((L.w / 2 + 63) / 64, (L.h + 3) / 4, nb);
            {
                KScope ks(c, K_PYR_V, 4.0 * L.w * L.h * nb);
                pyr_half_kernel<<<g, 256, 0, c->stream>>>(d_frames, step, fstride, W, H, L.w, L.h, L.smooth.k[0],
                                                         L.smooth.k[1], L.smooth.k[2], L.I, L.pitch, (size_t)L.pitch * L.h);
            }
            launch_polyexp(c, L, nb, first_slot);
            continue;
        }
        if (a.two && L.smooth.ksize == 9 && W == 4 * L.w && H == 4 * L.h && W >= 16 && H >= 16 && step % 4 == 0 &&
            fstride % 4 == 0 && (reinterpret_cast<size_t>(d_frames) & 3) == 0) {
            dim3 g((L.w + 31) / 32, (L.h + 7) / 8, nb);
            {
                KScope ks(c, K_PYR_V, 4.0 * L.w * L.h * nb);
                pyr_quarter_kernel<<<g, 256, 0, c->stream>>>(d_frames, step, fstride, W, H, L.w, L.h, L.smooth, L.I, L.pitch,
                                                            (size_t)L.pitch * L.h);
            }
            launch_polyexp(c, L, nb, first_slot);
            continue;
        }
        static const int cand[6][2] = {{64, 16}, {32, 16}, {32, 8}, {16, 8}, {8, 8}, {8, 4}};
        size_t smem = 0;
        for (int i = 0; i < 6; i++) {
            a.TXD = cand[i][0]; a.TYD = cand[i][1];
            a.SRW = (int)(a.sxs * (a.TXD - 1)) + 2 * r + 4; a.SRH = (int)(a.sys * (a.TYD - 1)) + 2 * r + 4;
            // u8 rows, up to 3 bytes of alignment shift, an odd number of 32-bit words per row: the row-strided byte reads
            // of the horizontal pass (lane = staged row) spread over the banks
            a.SRWB = 4 * (((a.SRW + 3 + 3) >> 2) | 1);
            smem = sizeof(float2) * (size_t)(a.SRH | 1) * a.TXD + (size_t)a.SRH * a.SRWB + 8 * (size_t)(a.TXD + a.TYD) +
                   4 * (size_t)a.SRW;
            if (smem <= 72 * 1024) break;        // three or more resident CTAs
        }
        {
            dim3 g((L.w + a.TXD - 1) / a.TXD, (L.h + a.TYD - 1) / a.TYD, nb);
            KScope ks(c, K_PYR_V, ((k == 0 ? (double)W * H : 0.0) + 4.0 * L.w * L.h) * nb);
            pyr_tile_kernel<<<g, 256, smem, c->stream>>>(a, L.smooth);
        }
        launch_polyexp(c, L, nb, first_slot);
    }
}

// per-frame direction/speed counts of nb dense w*h flows (device pointers) -> delta[nb][RC_HIST_CELLS]
void rc_launch_hist_of_flows(rc_ctx* c, float* const* flows, int nb, int w, int h, unsigned int* delta)
{
    FlowArgs a;
    memset(&a, 0, sizeof a);
    a.flow = nullptr;
    for (int j = 0; j < nb; j++) a.flow_dst[j] = flows[j];
    a.hist_delta = delta;
    const int n = w * h;
    int gx = (n + 255) / 256; if (gx > 148 * 4) gx = 148 * 4;
    cudaMemsetAsync(delta, 0, sizeof(unsigned int) * RC_HIST_CELLS * nb, c->stream);
    KScope ks(c, K_POLAR_HIST, 8.0 * n * nb);
    hist_batch_kernel<<<dim3(gx, nb), 256, 0, c->stream>>>(a, n);
}

void rc_launch_flows(rc_ctx* c, int nb, int prev_slot, float* const* flow_dst_host, unsigned int* hist_delta)
{
    const int T = c->prm.iterations;
    const bool fused_ok = !c->strict && c->win.m == 1 && T <= 3;
    if (hist_delta) cudaMemsetAsync(hist_delta, 0, sizeof(unsigned int) * RC_HIST_CELLS * nb, c->stream);
    for (int k = c->nlayers - 1; k >= 0; k--) {
        Layer& L = c->layer[k];
        if (c->expand_overlapped && k < 2) cudaStreamWaitEvent(c->stream, c->ev_poly[k], 0);
        FlowArgs a;
        a.R = L.R; a.plane = L.plane; a.pitch = L.pitch; a.w = L.w; a.h = L.h; a.nslots = c->B + 1; a.prev_slot = prev_slot;
        if (k == c->nlayers - 1) { a.coarse = nullptr; a.cw = a.ch = 0; a.coarse_stride = 0; a.sxs = a.sys = 1.0; a.fscale = 1.f; }
        else {
            Layer& C = c->layer[k + 1];
            a.coarse = C.flow; a.cw = C.w; a.ch = C.h; a.coarse_stride = (size_t)C.w * C.h * 2;
            a.sxs = 1.0 / ((double)L.w / (double)C.w); a.sys = 1.0 / ((double)L.h / (double)C.h);
            a.fscale = (float)(1.0 / c->prm.pyr_scale);
        }
        a.M = L.M; a.m_stride = 2 * 5 * L.plane;
        if (k == 0) { a.flow = nullptr; a.flow_stride = 0; for (int j = 0; j < nb; j++) a.flow_dst[j] = flow_dst_host[j]; }
        else { a.flow = L.flow; a.flow_stride = (size_t)L.w * L.h * 2; }
        a.hist_delta = (k == 0) ? hist_delta : nullptr;
        a.win = c->win;
        const double npx = (double)L.w * L.h * nb;
        if (fused_ok) {
            dim3 g(nb, (L.w + 31) / 32, (L.h + 31) / 32);
            KScope ks(c, K_FLOW_LAYER, (a.coarse ? 50.0 : 48.0) * npx);
            // Formulation of the fused layer (both pass the same parity tests): warp strips with M rows in
            // per-warp shared memory on big launches, 32x32 shared-memory tiles otherwise.  RC_FLOW_KERNEL=tile|strip forces
            // one of them, RC_STRIP_MINPX moves the switch-over point (pixel*pairs per launch).
            static const int force = [] {
                const char* e = getenv("RC_FLOW_KERNEL");
                return !e ? 0 : (!strcmp(e, "tile") ? 1 : (!strcmp(e, "strip") ? 2 : 0));
            }();
            static const double strip_min_px = getenv("RC_STRIP_MINPX") ? atof(getenv("RC_STRIP_MINPX")) : 16e6;
            static const int seg_env = getenv("RC_STRIP_SEG") ? atoi(getenv("RC_STRIP_SEG")) : 0;
            if (force == 2 || (force == 0 && npx >= strip_min_px)) {
                // row segments, evened out over the layer height (2T halo rows are recomputed per segment): ~72 rows when
                // that still leaves at least four waves of 148 x 4 CTAs, ~48 rows otherwise (measured: 1 % at B = 64)
                const int UW = 32 - 2 * T;
                const int groups = ((L.w + UW - 1) / UW + 7) / 8;
                int nseg = (L.h + 71) / 72;
                if ((long long)nb * groups * nseg < 148 * 4 * 4) nseg = (L.h + 47) / 48;
                const int SEG = seg_env > 0 ? seg_env : (L.h + nseg - 1) / nseg;
                dim3 gs(nb, groups, (L.h + SEG - 1) / SEG);
                if (!c->win.gaussian) {
                    if (T == 1) flow_strip_kernel<1, true><<<gs, 256, 0, c->stream>>>(a, SEG);
                    else if (T == 2) flow_strip_kernel<2, true><<<gs, 256, 0, c->stream>>>(a, SEG);
                    else flow_strip_kernel<3, true><<<gs, 256, 0, c->stream>>>(a, SEG);
                } else {
                    if (T == 1) flow_strip_kernel<1, false><<<gs, 256, 0, c->stream>>>(a, SEG);
                    else if (T == 2) flow_strip_kernel<2, false><<<gs, 256, 0, c->stream>>>(a, SEG);
                    else flow_strip_kernel<3, false><<<gs, 256, 0, c->stream>>>(a, SEG);
                }
                continue;
            }
            if (!c->win.gaussian) {
                if (T == 1) flow_layer_kernel<1, true><<<g, 256, 0, c->stream>>>(a);
                else if (T == 2) flow_layer_kernel<2, true><<<g, 256, 0, c->stream>>>(a);
                else flow_layer_kernel<3, true><<<g, 256, 0, c->stream>>>(a);
            } else {
                if (T == 1) flow_layer_kernel<1, false><<<g, 256, 0, c->stream>>>(a);
                else if (T == 2) flow_layer_kernel<2, false><<<g, 256, 0, c->stream>>>(a);
                else flow_layer_kernel<3, false><<<g, 256, 0, c->stream>>>(a);
            }
            continue;
        }
        dim3 b(32, 8), g((L.w + 31) / 32, (L.h + 7) / 8, nb);
        const bool tiled = !c->strict && c->win.m >= 1 && c->win.m <= 16;
        const int m = c->win.m;
        const bool march = tiled && (m == 2 || m == 5 || m == 10);
        // RC_MARCH_FIRST=0 / =10: M of the first iteration from the stand-alone updateMatrices kernel (no halo recomputation,
        // 20 B/px written and read back through the TMA staging) instead of inside the first marching launch -- for every
        // half-width / only for half-width 10
        static const int first_mode = getenv("RC_MARCH_FIRST") ? atoi(getenv("RC_MARCH_FIRST")) : 1;
        const bool first_fused = march && !(first_mode == 0 || (first_mode == 10 && m == 10));
        if (!first_fused) {      // otherwise the marching kernel computes M inside its first iteration
            KScope ks(c, K_UPDATE_MATRICES, (a.coarse ? 62.0 : 60.0) * npx);
            if (c->strict) update_matrices_kernel<true><<<g, b, 0, c->stream>>>(a, 0);
            else update_matrices_kernel<false><<<g, b, 0, c->stream>>>(a, 0);
        }
        const size_t tsm = sizeof(float) * ((size_t)(16 + 2 * m) * (64 + 2 * m) + 16 * (size_t)(64 + 2 * m));
        dim3 gt(nb, (L.w + 63) / 64, (L.h + 15) / 16);
        const bool box = !c->win.gaussian;
        const bool spec = march;      // marching kernel; other half-widths use the generic square-tile kernel
        static const int march_seg_env = getenv("RC_MARCH_SEG") ? atoi(getenv("RC_MARCH_SEG")) : 0;
        // row segments: every segment re-stages 2m halo rows, so they are as tall as the launch allows -- ~360 rows at
        // half-width 10, ~216 below (measured: +3.7 % on Gaussian winsize 20, +2 % on the 4K configuration) -- unless that
        // leaves fewer than ~8 CTAs per resident slot, in which case the 128-row segments keep the GPU full
        const int seg_target = m >= 8 ? 360 : 216;
        int mnseg = (L.h + seg_target - 1) / seg_target;
        if ((long long)nb * ((L.w + 63) / 64) * mnseg < 148LL * 3 * 8) mnseg = (L.h + 127) / 128;
        const int MSEG = march_seg_env > 0 ? march_seg_env : ((L.h + mnseg - 1) / mnseg + 3) & ~3;
        dim3 gm(nb, (L.w + 63) / 64, (L.h + MSEG - 1) / MSEG);
        const size_t msm = 128 + sizeof(float) * (16 * 5 * (size_t)((64 + 2 * m + (((m + 3) & ~3) - m) + 3) & ~3) + (size_t)(16 + 2 * m) * 5 * 64);
        const size_t msm128 = 128 + sizeof(float) * (8 * 5 * (size_t)((64 + 2 * m + (((m + 3) & ~3) - m) + 3) & ~3) + (size_t)(8 + 2 * m) * 5 * 64);
        // 128-thread CTAs (8 rows per step, 35 KB of shared memory, six resident CTAs) for half-widths up to 5: +2.5 % on
        // Gaussian winsize 10; at half-width 10 the ring alone is 36 KB, only four such CTAs fit and the 256-thread form wins
        static const int march_threads_env = getenv("RC_MARCH_THREADS") ? atoi(getenv("RC_MARCH_THREADS")) : 0;
        const int march_threads = march_threads_env ? march_threads_env : (m <= 5 ? 128 : 256);
        // tensor map of this layer's M buffers seen as [B * 2 * 5 planes][h][pitch] (zero fill outside the image: such steps
        // do not use it); box = staged columns x rows per step x 5 planes.  RC_MARCH_TMA=0 keeps the per-element copies.
        CUtensorMap tmM;
        memset(&tmM, 0, sizeof tmM);
        int use_tma = 0;
        static const bool tma_off = getenv("RC_MARCH_TMA") && atoi(getenv("RC_MARCH_TMA")) == 0;
        if (march && T > 1 && !tma_off) {
            const int rb = march_threads / 16, wpa = (64 + 2 * m + (((m + 3) & ~3) - m) + 3) & ~3;
            const cuuint64_t gdim[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)10 * c->B};
            const cuuint64_t gstr[2] = {(cuuint64_t)L.pitch * 4, (cuuint64_t)L.plane * 4};
            const cuuint32_t bx[3] = {(cuuint32_t)wpa, (cuuint32_t)rb, 5}, es[3] = {1, 1, 1};
            TensorMapEncodeFn enc = tensor_map_encoder();
            if (enc && enc(&tmM, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, L.M, gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                use_tma = 1;
            if (getenv("RC_MARCH_TMA_DEBUG")) fprintf(stderr, "march tma: layer %dx%d pitch %d m %d rb %d wpa %d B %d -> %d\n", L.w, L.h, L.pitch, m, rb, wpa, c->B, use_tma);
        }
        bool hist_fused = false;
        int mi = 0;
        for (int it = 0; it < T; it++) {
            const double first_extra = first_fused && it == 0 ? (a.coarse ? 22.0 : 20.0) : 0.0;   // R0 + R1 (+ coarse flow) instead of M
            if (it < T - 1) {
                KScope ks(c, K_FLOW_ITER_FUSED, (80.0 + first_extra) * npx);
                if (spec) {
#define RC_LAUNCH_MT(MM, FU, NT, SM) \
    do { if (it == 0 && first_fused) { \
             if (box) flow_march_kernel<MM, FU, true, true, NT><<<gm, NT, SM, c->stream>>>(a, mi, MSEG, tmM, 0); \
             else flow_march_kernel<MM, FU, false, true, NT><<<gm, NT, SM, c->stream>>>(a, mi, MSEG, tmM, 0); \
         } else if (box) flow_march_kernel<MM, FU, true, false, NT><<<gm, NT, SM, c->stream>>>(a, mi, MSEG, tmM, use_tma); \
         else flow_march_kernel<MM, FU, false, false, NT><<<gm, NT, SM, c->stream>>>(a, mi, MSEG, tmM, use_tma); } while (0)
#define RC_LAUNCH_M(MM, FU) \
    do { if (march_threads == 128) RC_LAUNCH_MT(MM, FU, 128, msm128); else RC_LAUNCH_MT(MM, FU, 256, msm); } while (0)
                    if (m == 2) RC_LAUNCH_M(2, true); else if (m == 5) RC_LAUNCH_M(5, true); else RC_LAUNCH_M(10, true);
                } else if (tiled) flow_iter_tiled_kernel<true><<<gt, 256, tsm, c->stream>>>(a, mi);
                else update_flow_strict_kernel<true><<<g, b, 0, c->stream>>>(a, mi);
                mi ^= 1;
            } else {
                KScope ks(c, K_FLOW_ITER_FINAL, (28.0 + first_extra) * npx);
                if (spec) {
                    if (m == 2) RC_LAUNCH_M(2, false); else if (m == 5) RC_LAUNCH_M(5, false); else RC_LAUNCH_M(10, false);
                    hist_fused = true;
#undef RC_LAUNCH_M
#undef RC_LAUNCH_MT
                } else if (tiled) flow_iter_tiled_kernel<false><<<gt, 256, tsm, c->stream>>>(a, mi);
                else update_flow_strict_kernel<false><<<g, b, 0, c->stream>>>(a, mi);
            }
        }
        if (k == 0 && hist_delta && !hist_fused) {
            const int n = L.w * L.h;
            int gx = (n + 255) / 256; if (gx > 148 * 4) gx = 148 * 4;
            KScope ks(c, K_POLAR_HIST, 8.0 * npx);
            hist_batch_kernel<<<dim3(gx, nb), 256, 0, c->stream>>>(a, n);
        }
    }
    c->expand_overlapped = false;
}
