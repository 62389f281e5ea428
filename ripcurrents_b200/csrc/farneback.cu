// Farneback dense optical flow for sm_100a -- kernels and launchers.
//
// Functional specification: SURVEY.md Appendix A (the algorithm cv::calcOpticalFlowFarneback runs for the
// reference's calls at RipCurrents_main/ripcurrents.cpp:215 and main.cpp:264,609,742,961,1119,1481).
// This file is compiled with -fmad=false: products and sums round separately unless fmaf() is written.
//
// HBM layout: every per-pixel coefficient set is PLANAR fp32 (5 planes for the polynomial expansion R and for
// the structure matrices M) with a row pitch rounded up to 32 floats, so that a warp reads 128 contiguous bytes
// from each plane; flow is interleaved float2 (CV_32FC2, the output format).
#include "rc_internal.h"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// cv::resize(INTER_LINEAR) source index + weight for destination index d (Appendix A.2)
__device__ __forceinline__ void resize_coef(int d, int src, int dst, double scale, int& s0, float& f)
{
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= (float)s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src - 1) { s = src - 1; fx = 0.f; }
    s0 = s; f = fx;
    (void)dst;
}

// ---------------------------------------------------------------------------------------------------
// Pyramid layer k of one frame: u8 -> fp32, Gaussian blur of the FULL-RESOLUTION image (REFLECT_101,
// rows then columns, fp32), bilinear resize to (lw, lh).  The blur is evaluated only where the resize samples.
// pass 1: horizontal blur at the (up to) two source columns each destination column samples, every source row.
// ---------------------------------------------------------------------------------------------------
__global__ void pyr_h_kernel(const uint8_t* __restrict__ img, size_t step, int W, int H, int dw, double scale_x,
                             int two, SmoothCoef sc, float* __restrict__ htmp)
{
    int X = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= dw || y >= H) return;
    int sx; float fx;
    if (two) resize_coef(X, W, dw, scale_x, sx, fx); else sx = X;
    const uint8_t* row = img + (size_t)y * step;
    const int r = sc.ksize / 2;
    float s0 = 0.f, s1 = 0.f;
    int sx1 = sx + 1 < W ? sx + 1 : W - 1;
    if (sx - r >= 0 && sx1 + r < W) {
        for (int i = 0; i < sc.ksize; i++) {
            s0 = s0 + sc.k[i] * (float)row[sx + i - r];
            if (two) s1 = s1 + sc.k[i] * (float)row[sx1 + i - r];
        }
    } else {
        for (int i = 0; i < sc.ksize; i++) {
            s0 = s0 + sc.k[i] * (float)row[reflect101(sx + i - r, W)];
            if (two) s1 = s1 + sc.k[i] * (float)row[reflect101(sx1 + i - r, W)];
        }
    }
    if (two) {
        float2* o = reinterpret_cast<float2*>(htmp) + (size_t)y * dw + X;
        *o = make_float2(s0, s1);
    } else {
        htmp[(size_t)y * dw + X] = s0;
    }
}

// pass 2: vertical blur at the two source rows each destination row samples, then the bilinear combination.
__global__ void pyr_v_kernel(const float* __restrict__ htmp, int W, int H, int dw, int dh, double scale_x,
                             double scale_y, int two, SmoothCoef sc, float* __restrict__ out, int pitch)
{
    int X = blockIdx.x * blockDim.x + threadIdx.x;
    int Y = blockIdx.y * blockDim.y + threadIdx.y;
    if (X >= dw || Y >= dh) return;
    const int r = sc.ksize / 2;
    if (!two) {
        float s = 0.f;
        for (int j = 0; j < sc.ksize; j++) s = s + sc.k[j] * htmp[(size_t)reflect101(Y + j - r, H) * dw + X];
        out[(size_t)Y * pitch + X] = s;
        return;
    }
    int sx, sy; float fx, fy;
    resize_coef(X, W, dw, scale_x, sx, fx);
    resize_coef(Y, H, dh, scale_y, sy, fy);
    int sy1 = sy + 1 < H ? sy + 1 : H - 1;
    const float2* t = reinterpret_cast<const float2*>(htmp);
    float b00 = 0.f, b01 = 0.f, b10 = 0.f, b11 = 0.f;
    for (int j = 0; j < sc.ksize; j++) {
        float2 a = t[(size_t)reflect101(sy + j - r, H) * dw + X];
        float2 b = t[(size_t)reflect101(sy1 + j - r, H) * dw + X];
        b00 = b00 + sc.k[j] * a.x; b01 = b01 + sc.k[j] * a.y;
        b10 = b10 + sc.k[j] * b.x; b11 = b11 + sc.k[j] * b.y;
    }
    float top = b00 * (1.f - fx) + b01 * fx;
    float bot = b10 * (1.f - fx) + b11 * fx;
    out[(size_t)Y * pitch + X] = top * (1.f - fy) + bot * fy;
}

// ---------------------------------------------------------------------------------------------------
// Polynomial expansion (Appendix A.3), reference tile kernel: vertical pass fp32 -> shared memory,
// horizontal pass with fp64 accumulators, replicate borders.
// ---------------------------------------------------------------------------------------------------
template <int TX, int TY>
__global__ void polyexp_ref_kernel(const float* __restrict__ I, int w, int h, int pitch, Planes R, PolyCoef pc)
{
    extern __shared__ float sm[];
    const int n = pc.n;
    const int SW = TX + 2 * n;
    float* sI = sm;
    float* sr = sm + (TY + 2 * n) * SW;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int tid = threadIdx.x, nt = blockDim.x;

    for (int idx = tid; idx < (TY + 2 * n) * SW; idx += nt) {
        int r = idx / SW, c = idx - r * SW;
        int gy = clampi(y0 + r - n, 0, h - 1), gx = clampi(x0 + c - n, 0, w - 1);
        sI[idx] = I[(size_t)gy * pitch + gx];
    }
    __syncthreads();
    for (int idx = tid; idx < TY * SW; idx += nt) {
        int r = idx / SW, c = idx - r * SW;
        const float* col = sI + (r + n) * SW + c;
        float r0 = col[0] * pc.g[0], r1 = 0.f, r2 = 0.f;
        for (int k = 1; k <= n; k++) {
            float up = col[-k * SW], dn = col[k * SW];
            float p = up + dn;
            r0 = r0 + pc.g[k] * p;
            r1 = r1 + pc.xg[k] * (dn - up);
            r2 = r2 + pc.xxg[k] * p;
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
        double b1 = q0[0] * pc.g[0], b2 = 0, b3 = q1[0] * pc.g[0], b4 = 0, b5 = q2[0] * pc.g[0], b6 = 0;
        for (int k = 1; k <= n; k++) {
            double tg = q0[k] + q0[-k];
            b1 += tg * pc.g[k];
            b4 += tg * pc.xxg[k];
            b2 += (q0[k] - q0[-k]) * pc.xg[k];
            b3 += (q1[k] + q1[-k]) * pc.g[k];
            b6 += (q1[k] - q1[-k]) * pc.xg[k];
            b5 += (q2[k] + q2[-k]) * pc.g[k];
        }
        size_t o = (size_t)y * R.pitch + x;
        R.plane(0)[o] = (float)(b3 * pc.ig11);
        R.plane(1)[o] = (float)(b2 * pc.ig11);
        R.plane(2)[o] = (float)(b1 * pc.ig03 + b5 * pc.ig33);
        R.plane(3)[o] = (float)(b1 * pc.ig03 + b4 * pc.ig33);
        R.plane(4)[o] = (float)(b6 * pc.ig55);
    }
}

// ---------------------------------------------------------------------------------------------------
// updateMatrices for one pixel (Appendix A.5).  R0/R1/M planar.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void update_matrices_px(int x, int y, float dx, float dy, int w, int h, const Planes& R0,
                                                   const Planes& R1, const Planes& M)
{
    const size_t p = (size_t)y * R0.pitch + x;
    float fx = (float)x + dx, fy = (float)y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1; fy -= (float)y1;
    const float r0_0 = R0.plane(0)[p], r0_1 = R0.plane(1)[p], r0_2 = R0.plane(2)[p], r0_3 = R0.plane(3)[p],
                r0_4 = R0.plane(4)[p];
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const size_t q = (size_t)y1 * R1.pitch + x1, qd = q + R1.pitch;
        const float* c0 = R1.plane(0); const float* c1 = R1.plane(1); const float* c2 = R1.plane(2);
        const float* c3 = R1.plane(3); const float* c4 = R1.plane(4);
        r2 = a00 * __ldg(c0 + q) + a01 * __ldg(c0 + q + 1) + a10 * __ldg(c0 + qd) + a11 * __ldg(c0 + qd + 1);
        r3 = a00 * __ldg(c1 + q) + a01 * __ldg(c1 + q + 1) + a10 * __ldg(c1 + qd) + a11 * __ldg(c1 + qd + 1);
        r4 = a00 * __ldg(c2 + q) + a01 * __ldg(c2 + q + 1) + a10 * __ldg(c2 + qd) + a11 * __ldg(c2 + qd + 1);
        r5 = a00 * __ldg(c3 + q) + a01 * __ldg(c3 + q + 1) + a10 * __ldg(c3 + qd) + a11 * __ldg(c3 + qd + 1);
        r6 = a00 * __ldg(c4 + q) + a01 * __ldg(c4 + q + 1) + a10 * __ldg(c4 + qd) + a11 * __ldg(c4 + qd + 1);
        r4 = (r0_2 + r4) * 0.5f;
        r5 = (r0_3 + r5) * 0.5f;
        r6 = (r0_4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = r0_2; r5 = r0_3; r6 = r0_4 * 0.5f;
    }
    r2 = (r0_0 - r2) * 0.5f;
    r3 = (r0_1 - r3) * 0.5f;
    r2 = r2 + (r4 * dy + r6 * dx);
    r3 = r3 + (r6 * dy + r5 * dx);
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        float scale = (x < 5 ? border[x] : 1.f) * (x >= w - 5 ? border[w - x - 1] : 1.f) *
                      (y < 5 ? border[y] : 1.f) * (y >= h - 5 ? border[h - y - 1] : 1.f);
        r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
    }
    const size_t o = (size_t)y * M.pitch + x;
    M.plane(0)[o] = r4 * r4 + r6 * r6;
    M.plane(1)[o] = (r4 + r5) * r6;
    M.plane(2)[o] = r5 * r5 + r6 * r6;
    M.plane(3)[o] = r4 * r2 + r6 * r3;
    M.plane(4)[o] = r6 * r2 + r5 * r3;
}

// updateMatrices with the flow initialisation of Appendix A.4 fused in: the upsampled flow is never stored
// (the next flow is a function of the blurred M only).
__global__ void update_matrices_kernel(Planes R0, Planes R1, Planes M, int flow_mode, const float* __restrict__ flow,
                                       int cw, int ch, double sx_scale, double sy_scale, float flow_scale)
{
    const int w = R0.w, h = R0.h;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float dx = 0.f, dy = 0.f;
    if (flow_mode == 2) {
        float2 f = reinterpret_cast<const float2*>(flow)[(size_t)y * w + x];
        dx = f.x; dy = f.y;
    } else if (flow_mode == 1) {
        int sx, sy; float fx, fy;
        resize_coef(x, cw, w, sx_scale, sx, fx);
        resize_coef(y, ch, h, sy_scale, sy, fy);
        int sx1 = sx + 1 < cw ? sx + 1 : cw - 1, sy1 = sy + 1 < ch ? sy + 1 : ch - 1;
        const float2* cf = reinterpret_cast<const float2*>(flow);
        float2 a = __ldg(cf + (size_t)sy * cw + sx), b = __ldg(cf + (size_t)sy * cw + sx1);
        float2 c = __ldg(cf + (size_t)sy1 * cw + sx), d = __ldg(cf + (size_t)sy1 * cw + sx1);
        float tx = a.x * (1.f - fx) + b.x * fx, ty = a.y * (1.f - fx) + b.y * fx;
        float bx = c.x * (1.f - fx) + d.x * fx, by = c.y * (1.f - fx) + d.y * fx;
        dx = (tx * (1.f - fy) + bx * fy) * flow_scale;
        dy = (ty * (1.f - fy) + by * fy) * flow_scale;
    }
    update_matrices_px(x, y, dx, dy, w, h, R0, R1, M);
}

// ---------------------------------------------------------------------------------------------------
// updateFlow (Appendix A.6 box / A.7 Gaussian), reference per-pixel kernels.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 solve2x2(double g11, double g12, double g22, double h1, double h2)
{
    double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
    return make_float2((float)((g11 * h2 - g12 * h1) * idet), (float)((g22 * h1 - g12 * h2) * idet));
}

template <bool FUSE_UPDATE>
__global__ void update_flow_box_ref_kernel(Planes Min, int m, double scale, Planes R0, Planes R1, Planes Mout,
                                           float* __restrict__ flow_out)
{
    const int w = Min.w, h = Min.h;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    double s[5] = {0, 0, 0, 0, 0};
    for (int i = -m; i <= m; i++) {           // columns
        int xx = clampi(x + i, 0, w - 1);
        double v[5] = {0, 0, 0, 0, 0};
        for (int j = -m; j <= m; j++) {       // vertical sum of that column
            size_t o = (size_t)clampi(y + j, 0, h - 1) * Min.pitch + xx;
#pragma unroll
            for (int c = 0; c < 5; c++) v[c] += (double)__ldg(Min.plane(c) + o);
        }
#pragma unroll
        for (int c = 0; c < 5; c++) s[c] += v[c];
    }
    float2 f = solve2x2(s[0] * scale, s[1] * scale, s[2] * scale, s[3] * scale, s[4] * scale);
    if (FUSE_UPDATE) update_matrices_px(x, y, f.x, f.y, w, h, R0, R1, Mout);
    else reinterpret_cast<float2*>(flow_out)[(size_t)y * w + x] = f;
}

template <bool FUSE_UPDATE>
__global__ void update_flow_gauss_ref_kernel(Planes Min, GaussWin gw, Planes R0, Planes R1, Planes Mout,
                                             float* __restrict__ flow_out)
{
    const int w = Min.w, h = Min.h, m = gw.m;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float hs[5];
    auto vcol = [&](int xx, float* v) {
        size_t o = (size_t)y * Min.pitch + xx;
#pragma unroll
        for (int c = 0; c < 5; c++) v[c] = __ldg(Min.plane(c) + o) * gw.k[0];
        for (int j = 1; j <= m; j++) {
            size_t ou = (size_t)clampi(y - j, 0, h - 1) * Min.pitch + xx;
            size_t od = (size_t)clampi(y + j, 0, h - 1) * Min.pitch + xx;
#pragma unroll
            for (int c = 0; c < 5; c++) v[c] = v[c] + (__ldg(Min.plane(c) + od) + __ldg(Min.plane(c) + ou)) * gw.k[j];
        }
    };
    float v0[5], va[5], vb[5];
    vcol(x, v0);
#pragma unroll
    for (int c = 0; c < 5; c++) hs[c] = v0[c] * gw.k[0];
    for (int i = 1; i <= m; i++) {
        vcol(clampi(x - i, 0, w - 1), va);
        vcol(clampi(x + i, 0, w - 1), vb);
#pragma unroll
        for (int c = 0; c < 5; c++) hs[c] = hs[c] + gw.k[i] * (va[c] + vb[c]);
    }
    float2 f = solve2x2(hs[0], hs[1], hs[2], hs[3], hs[4]);
    if (FUSE_UPDATE) update_matrices_px(x, y, f.x, f.y, w, h, R0, R1, Mout);
    else reinterpret_cast<float2*>(flow_out)[(size_t)y * w + x] = f;
}

}  // namespace

// ===================================================================================================
// launchers
// ===================================================================================================
void rc_launch_pyr_layer(rc_ctx* c, const uint8_t* d_img, size_t step, int W, int H, Layer& L)
{
    const int two = !(L.w == W && L.h == H);
    const double sx = 1.0 / ((double)L.w / (double)W), sy = 1.0 / ((double)L.h / (double)H);
    dim3 b(32, 8);
    dim3 g1((L.w + 31) / 32, (H + 7) / 8);
    {
        KScope ks(c, K_PYR_H, (double)W * H);
        pyr_h_kernel<<<g1, b, 0, c->stream>>>(d_img, step, W, H, L.w, sx, two, L.smooth, L.htmp);
    }
    dim3 g2((L.w + 31) / 32, (L.h + 7) / 8);
    const int pitch = (L.w + 31) / 32 * 32;
    {
        KScope ks(c, K_PYR_V, 4.0 * L.w * L.h);
        pyr_v_kernel<<<g2, b, 0, c->stream>>>(L.htmp, W, H, L.w, L.h, sx, sy, two, L.smooth, L.I, pitch);
    }
}

void rc_launch_polyexp(rc_ctx* c, const float* I, int w, int h, int pitch, const Planes& R)
{
    constexpr int TX = 64, TY = 32;
    const int n = c->poly.n;
    const size_t smem = sizeof(float) * ((size_t)(TY + 2 * n) * (TX + 2 * n) + 3 * (size_t)TY * (TX + 2 * n));
    static size_t configured = 0;
    if (smem > configured) {
        cudaFuncSetAttribute(polyexp_ref_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    dim3 g((w + TX - 1) / TX, (h + TY - 1) / TY);
    KScope ks(c, K_POLYEXP, 24.0 * w * h);
    polyexp_ref_kernel<TX, TY><<<g, 256, smem, c->stream>>>(I, w, h, pitch, R, c->poly);
}

void rc_launch_update_matrices(rc_ctx* c, const Planes& R0, const Planes& R1, const Planes& M, int flow_mode,
                               const float* flow, int cw, int ch, float flow_scale)
{
    dim3 b(32, 8), g((R0.w + 31) / 32, (R0.h + 7) / 8);
    double sx = 1.0, sy = 1.0;
    if (flow_mode == 1) { sx = 1.0 / ((double)R0.w / (double)cw); sy = 1.0 / ((double)R0.h / (double)ch); }
    KScope ks(c, K_UPDATE_MATRICES, (flow_mode ? 62.0 : 60.0) * R0.w * R0.h);
    update_matrices_kernel<<<g, b, 0, c->stream>>>(R0, R1, M, flow_mode, flow, cw, ch, sx, sy, flow_scale);
}

void rc_launch_update_flow(rc_ctx* c, const Planes& M_in, const Planes& R0, const Planes& R1, const Planes& M_out,
                           float* flow_out, unsigned long long* hist2d)
{
    (void)hist2d;
    dim3 b(32, 8), g((M_in.w + 31) / 32, (M_in.h + 7) / 8);
    const bool fuse = M_out.p != nullptr;
    KScope ks(c, fuse ? K_FLOW_ITER_FUSED : K_FLOW_ITER_FINAL, (fuse ? 80.0 : 28.0) * M_in.w * M_in.h);
    if (c->prm.flags & RC_FARNEBACK_GAUSSIAN) {
        if (fuse) update_flow_gauss_ref_kernel<true><<<g, b, 0, c->stream>>>(M_in, c->gwin, R0, R1, M_out, flow_out);
        else update_flow_gauss_ref_kernel<false><<<g, b, 0, c->stream>>>(M_in, c->gwin, R0, R1, M_out, flow_out);
    } else {
        const int m = c->prm.winsize / 2;
        const double scale = 1.0 / ((double)c->prm.winsize * c->prm.winsize);
        if (fuse) update_flow_box_ref_kernel<true><<<g, b, 0, c->stream>>>(M_in, m, scale, R0, R1, M_out, flow_out);
        else update_flow_box_ref_kernel<false><<<g, b, 0, c->stream>>>(M_in, m, scale, R0, R1, M_out, flow_out);
    }
}
