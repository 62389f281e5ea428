/*
 * oracle/farneback_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, no OpenCV) of the dense optical flow the reference
 * obtains from cv::calcOpticalFlowFarneback:
 *   RipCurrents_main/ripcurrents.cpp:215   (0.5,2,3,2,15,1.2,0)
 *   RipCurrents_main/main.cpp:264,742      (0.5,2,3,2,15,1.2,GAUSSIAN)
 *   RipCurrents_main/main.cpp:609,961      (0.5,2,20,3,15,1.2,GAUSSIAN)
 *   RipCurrents_main/main.cpp:1119,1481    (0.5,2,10,3,15,1.2,GAUSSIAN)
 *   RipCurrents_android/.../ripcurrents.cpp:167,171 (0.5,3,5,3,15,1.2,0)
 *
 * The arithmetic lives in OpenCV's `video` module, a third-party dependency
 * that is NOT under /root/reference and is not pinned by any lockfile (README:5
 * names 3.4.1, CMakeCache.txt:334 linked 4.1.0).  This file restates the
 * published algorithm as specified in SURVEY.md Appendix A (A.1 - A.8); it is
 * pinned against cv2 4.13.0 (the only OpenCV runnable here) by
 * tests/test_oracle_farneback.py and the fixtures in tests/golden/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product path
 * (ripcurrents_b200/csrc) never links or calls it.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: products and sums are
 * rounded separately, as in OpenCV's baseline SSE2 build; explicit fmaf() is
 * used where the oracle fuses).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RC_MAX_LAYERS 16
#define RC_FLAG_GAUSSIAN 256 /* cv::OPTFLOW_FARNEBACK_GAUSSIAN */

static int rc_round_half_even(double v) { return (int)nearbyint(v); } /* cvRound */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* ---- A.1 layer selection -------------------------------------------------- */
int rc_oracle_layers(int w, int h, double pyr_scale, int levels, int* lw, int* lh)
{
    int k;
    double scale = 1.0;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (w * scale < 32 || h * scale < 32) break;
    }
    levels = k;
    if (levels + 1 > RC_MAX_LAYERS) return -1;
    scale = 1.0;
    for (k = 0; k <= levels; k++) {
        lw[k] = rc_round_half_even(w * scale);
        lh[k] = rc_round_half_even(h * scale);
        scale *= pyr_scale;
    }
    return levels + 1; /* number of layers */
}

/* ---- A.2 layer image: u8 -> f32, Gaussian blur (REFLECT_101), bilinear resize */
int rc_oracle_smooth_kernel(double sigma, int ksize, float* kern)
{
    int i;
    if (sigma <= 0 && ksize == 3) { kern[0] = 0.25f; kern[1] = 0.5f; kern[2] = 0.25f; return 0; }
    if (sigma <= 0) sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    {
        double s2 = -0.5 / (sigma * sigma), sum = 0.0;
        double* t = (double*)malloc(sizeof(double) * ksize);
        for (i = 0; i < ksize; i++) {
            double x = i - (ksize - 1) * 0.5;
            t[i] = exp(s2 * x * x);
            sum += t[i];
        }
        sum = 1.0 / sum;
        for (i = 0; i < ksize; i++) kern[i] = (float)(t[i] * sum);
        free(t);
    }
    return 0;
}

void rc_oracle_layer_params(double pyr_scale, int k, double* sigma, int* ksize)
{
    double scale = 1.0;
    int i, ks;
    for (i = 0; i < k; i++) scale *= pyr_scale;
    *sigma = (1.0 / scale - 1.0) * 0.5;
    ks = rc_round_half_even(*sigma * 5.0) | 1;
    *ksize = ks < 3 ? 3 : ks;
}

/* separable blur of the full-resolution image: rows first, then columns, fp32 */
static void blur_full(const uint8_t* img, int w, int h, size_t step, const float* kern, int ksize, float* out)
{
    int r = ksize / 2, x, y, i;
    float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (y = 0; y < h; y++) {
        const uint8_t* row = img + (size_t)y * step;
        for (x = 0; x < w; x++) {
            float s = 0.f;
            for (i = 0; i < ksize; i++) s += kern[i] * (float)row[reflect101(x + i - r, w)];
            tmp[(size_t)y * w + x] = s;
        }
    }
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            float s = 0.f;
            for (i = 0; i < ksize; i++) s += kern[i] * tmp[(size_t)reflect101(y + i - r, h) * w + x];
            out[(size_t)y * w + x] = s;
        }
    free(tmp);
}

/* bilinear resize coefficients as cv::resize(INTER_LINEAR) computes them */
static void resize_coef(int d, int src, int dst, int* s0, float* f)
{
    double scale = 1.0 / ((double)dst / (double)src);
    float fx = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(fx);
    fx -= s;
    if (s < 0) { s = 0; fx = 0.f; }
    if (s >= src - 1) { s = src - 1; fx = 0.f; }
    *s0 = s; *f = fx;
}

static void resize_bilinear(const float* src, int sw, int sh, int cn, float* dst, int dw, int dh)
{
    int x, y, c;
    if (sw == dw && sh == dh) { memcpy(dst, src, sizeof(float) * (size_t)sw * sh * cn); return; }
    for (y = 0; y < dh; y++) {
        int sy; float fy;
        resize_coef(y, sh, dh, &sy, &fy);
        {
            int sy1 = sy + 1 < sh ? sy + 1 : sh - 1;
            for (x = 0; x < dw; x++) {
                int sx; float fx;
                resize_coef(x, sw, dw, &sx, &fx);
                {
                    int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
                    for (c = 0; c < cn; c++) {
                        float a = src[((size_t)sy * sw + sx) * cn + c], b = src[((size_t)sy * sw + sx1) * cn + c];
                        float cc = src[((size_t)sy1 * sw + sx) * cn + c], d = src[((size_t)sy1 * sw + sx1) * cn + c];
                        float top = a * (1.f - fx) + b * fx;
                        float bot = cc * (1.f - fx) + d * fx;
                        dst[((size_t)y * dw + x) * cn + c] = top * (1.f - fy) + bot * fy;
                    }
                }
            }
        }
    }
}

/* I_k for one frame: out is lw*lh f32 */
int rc_oracle_pyr_layer(const uint8_t* img, int w, int h, size_t step, double pyr_scale, int k, int lw, int lh, float* out)
{
    double sigma; int ksize;
    float kern[256];
    float* full;
    rc_oracle_layer_params(pyr_scale, k, &sigma, &ksize);
    if (ksize > 255) return -1;
    rc_oracle_smooth_kernel(sigma, ksize, kern);
    full = (float*)malloc(sizeof(float) * (size_t)w * h);
    blur_full(img, w, h, step, kern, ksize, full);
    resize_bilinear(full, w, h, 1, out, lw, lh);
    free(full);
    return 0;
}

/* ---- A.3 polynomial expansion --------------------------------------------- */
/* g, xg, xxg are arrays of n+1 entries (index 0..n); ig = {ig11, ig03, ig33, ig55} */
void rc_oracle_poly_kernels(int n, double sigma, float* g, float* xg, float* xxg, double* ig)
{
    int x, y;
    double s = 0.0, a = 0, b = 0, c = 0, d = 0;
    float* gg = (float*)malloc(sizeof(float) * (2 * n + 1));
    float* G = gg + n;
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    for (x = -n; x <= n; x++) {
        G[x] = (float)exp(-x * x / (2 * sigma * sigma));
        s += G[x];
    }
    s = 1.0 / s;
    for (x = -n; x <= n; x++) G[x] = (float)(G[x] * s);
    for (x = 0; x <= n; x++) {
        g[x] = G[x];
        xg[x] = (float)(x * G[x]);
        xxg[x] = (float)(x * x * G[x]);
    }
    /* 6x6 moment matrix: entries accumulated in double from fp32 products */
    for (y = -n; y <= n; y++)
        for (x = -n; x <= n; x++) {
            float gyx = G[y] * G[x];
            a += gyx;
            b += gyx * x * x;
            c += gyx * x * x * x * x;
            d += gyx * x * x * y * y;
        }
    /* closed-form inverse of the (1,x^2,y^2) block [[a,b,b],[b,c,d],[b,d,c]] */
    {
        double det3 = (c - d) * (a * (c + d) - 2 * b * b);
        ig[0] = 1.0 / b;                 /* invG(1,1) */
        ig[1] = -b * (c - d) / det3;     /* invG(0,3) */
        ig[2] = (a * c - b * b) / det3;  /* invG(3,3) */
        ig[3] = 1.0 / d;                 /* invG(5,5) */
    }
    free(gg);
}

/* src: w*h f32; dst: w*h*5 f32 interleaved (OpenCV channel order) */
void rc_oracle_polyexp(const float* src, int w, int h, int n, double sigma, float* dst)
{
    float *g = (float*)malloc(sizeof(float) * 3 * (n + 1)), *xg = g + n + 1, *xxg = xg + n + 1;
    double ig[4];
    float* row = (float*)malloc(sizeof(float) * 3 * (size_t)w);
    int x, y, k;
    rc_oracle_poly_kernels(n, sigma, g, xg, xxg, ig);
    for (y = 0; y < h; y++) {
        const float* c = src + (size_t)y * w;
        for (x = 0; x < w; x++) { row[x * 3] = c[x] * g[0]; row[x * 3 + 1] = 0.f; row[x * 3 + 2] = 0.f; }
        for (k = 1; k <= n; k++) {
            const float* up = src + (size_t)clampi(y - k, 0, h - 1) * w;
            const float* dn = src + (size_t)clampi(y + k, 0, h - 1) * w;
            for (x = 0; x < w; x++) {
                float p = up[x] + dn[x];
                row[x * 3] = row[x * 3] + g[k] * p;
                row[x * 3 + 1] = row[x * 3 + 1] + xg[k] * (dn[x] - up[x]);
                row[x * 3 + 2] = row[x * 3 + 2] + xxg[k] * p;
            }
        }
        for (x = 0; x < w; x++) {
            double b1 = row[x * 3] * g[0], b2 = 0, b3 = row[x * 3 + 1] * g[0], b4 = 0, b5 = row[x * 3 + 2] * g[0], b6 = 0;
            float* o = dst + ((size_t)y * w + x) * 5;
            for (k = 1; k <= n; k++) {
                const float* rp = row + 3 * clampi(x + k, 0, w - 1);
                const float* rm = row + 3 * clampi(x - k, 0, w - 1);
                double tg = rp[0] + rm[0];
                b1 += tg * g[k];
                b4 += tg * xxg[k];
                b2 += (rp[0] - rm[0]) * xg[k];
                b3 += (rp[1] + rm[1]) * g[k];
                b6 += (rp[1] - rm[1]) * xg[k];
                b5 += (rp[2] + rm[2]) * g[k];
            }
            o[1] = (float)(b2 * ig[0]);
            o[0] = (float)(b3 * ig[0]);
            o[3] = (float)(b1 * ig[1] + b4 * ig[2]);
            o[2] = (float)(b1 * ig[1] + b5 * ig[2]);
            o[4] = (float)(b6 * ig[3]);
        }
    }
    free(row); free(g);
}

/* ---- A.5 updateMatrices ---------------------------------------------------- */
/* R0,R1: w*h*5; flow: w*h*2; M: w*h*5 */
void rc_oracle_update_matrices(const float* R0, const float* R1, const float* flow, int w, int h, float* M)
{
    static const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    int x, y;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            size_t p = (size_t)y * w + x;
            float dx = flow[p * 2], dy = flow[p * 2 + 1];
            float fx = x + dx, fy = y + dy;
            int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
            const float* r0 = R0 + p * 5;
            float r2, r3, r4, r5, r6;
            fx -= x1; fy -= y1;
            if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
                float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
                const float* q = R1 + ((size_t)y1 * w + x1) * 5;
                const float* qd = q + (size_t)w * 5;
                r2 = a00 * q[0] + a01 * q[5] + a10 * qd[0] + a11 * qd[5];
                r3 = a00 * q[1] + a01 * q[6] + a10 * qd[1] + a11 * qd[6];
                r4 = a00 * q[2] + a01 * q[7] + a10 * qd[2] + a11 * qd[7];
                r5 = a00 * q[3] + a01 * q[8] + a10 * qd[3] + a11 * qd[8];
                r6 = a00 * q[4] + a01 * q[9] + a10 * qd[4] + a11 * qd[9];
                r4 = (r0[2] + r4) * 0.5f;
                r5 = (r0[3] + r5) * 0.5f;
                r6 = (r0[4] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f;
                r4 = r0[2]; r5 = r0[3]; r6 = r0[4] * 0.5f;
            }
            r2 = (r0[0] - r2) * 0.5f;
            r3 = (r0[1] - r3) * 0.5f;
            r2 += r4 * dy + r6 * dx;
            r3 += r6 * dy + r5 * dx;
            if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
                float scale = (x < 5 ? border[x] : 1.f) * (x >= w - 5 ? border[w - x - 1] : 1.f) *
                              (y < 5 ? border[y] : 1.f) * (y >= h - 5 ? border[h - y - 1] : 1.f);
                r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
            }
            {
                float* m = M + p * 5;
                m[0] = r4 * r4 + r6 * r6;
                m[1] = (r4 + r5) * r6;
                m[2] = r5 * r5 + r6 * r6;
                m[3] = r4 * r2 + r6 * r3;
                m[4] = r6 * r2 + r5 * r3;
            }
        }
}

/* ---- A.6 / A.7 updateFlow --------------------------------------------------- */
static void solve_store(double g11, double g12, double g22, double h1, double h2, float* f)
{
    double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
    f[0] = (float)((g11 * h2 - g12 * h1) * idet);
    f[1] = (float)((g22 * h1 - g12 * h2) * idet);
}

void rc_oracle_update_flow_box(const float* M, int w, int h, int winsize, float* flow)
{
    int m = winsize / 2, x, y, c, i;
    double scale = 1.0 / ((double)winsize * winsize);
    double* vs = (double*)malloc(sizeof(double) * 5 * (size_t)w);
    for (y = 0; y < h; y++) {
        for (x = 0; x < w * 5; x++) vs[x] = 0.0;
        for (i = -m; i <= m; i++) {
            const float* r = M + (size_t)clampi(y + i, 0, h - 1) * w * 5;
            for (x = 0; x < w * 5; x++) vs[x] += r[x];
        }
        for (x = 0; x < w; x++) {
            double s[5] = {0, 0, 0, 0, 0};
            for (i = -m; i <= m; i++) {
                const double* v = vs + 5 * clampi(x + i, 0, w - 1);
                for (c = 0; c < 5; c++) s[c] += v[c];
            }
            solve_store(s[0] * scale, s[1] * scale, s[2] * scale, s[3] * scale, s[4] * scale,
                        flow + ((size_t)y * w + x) * 2);
        }
    }
    free(vs);
}

/* det_mode: 0 = products of the blurred fp32 values formed in fp64;
 *           1 = g11*g22 - g12*g12 and the numerators formed in fp32 (float operands),
 *               widened only by the "+1e-3" / "*idet" (candidate reading of OpenCV's
 *               Gaussian path; see tests/test_oracle_farneback.py for which one pins). */
void rc_oracle_update_flow_gauss(const float* M, int w, int h, int winsize, int det_mode, float* flow)
{
    int m = winsize / 2, x, y, c, i;
    double sigma = m * 0.3, s = 1.0;
    float* kern = (float*)malloc(sizeof(float) * (m + 1));
    float* vs = (float*)malloc(sizeof(float) * 5 * (size_t)w);
    kern[0] = 1.f;
    for (i = 1; i <= m; i++) {
        float t = (float)exp(-i * i / (2 * sigma * sigma));
        kern[i] = t;
        s += t * 2;
    }
    s = 1.0 / s;
    for (i = 0; i <= m; i++) kern[i] = (float)(kern[i] * s);
    for (y = 0; y < h; y++) {
        const float* r0 = M + (size_t)y * w * 5;
        for (x = 0; x < w * 5; x++) vs[x] = r0[x] * kern[0];
        for (i = 1; i <= m; i++) {
            const float* up = M + (size_t)clampi(y - i, 0, h - 1) * w * 5;
            const float* dn = M + (size_t)clampi(y + i, 0, h - 1) * w * 5;
            for (x = 0; x < w * 5; x++) vs[x] += (dn[x] + up[x]) * kern[i];
        }
        for (x = 0; x < w; x++) {
            float hs[5];
            for (c = 0; c < 5; c++) hs[c] = vs[x * 5 + c] * kern[0];
            for (i = 1; i <= m; i++) {
                const float* a = vs + 5 * clampi(x - i, 0, w - 1);
                const float* b = vs + 5 * clampi(x + i, 0, w - 1);
                for (c = 0; c < 5; c++) hs[c] += kern[i] * (a[c] + b[c]);
            }
            if (det_mode == 0)
                solve_store(hs[0], hs[1], hs[2], hs[3], hs[4], flow + ((size_t)y * w + x) * 2);
            else {
                float g11 = hs[0], g12 = hs[1], g22 = hs[2], h1 = hs[3], h2 = hs[4];
                double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
                float* f = flow + ((size_t)y * w + x) * 2;
                f[0] = (float)((g11 * h2 - g12 * h1) * idet);
                f[1] = (float)((g22 * h1 - g12 * h2) * idet);
            }
        }
    }
    free(vs); free(kern);
}

/* ---- A.4 flow initialisation from the coarser layer ------------------------ */
void rc_oracle_upsample_flow(const float* coarse, int cw, int ch, float* fine, int fw, int fh, double pyr_scale)
{
    size_t i, n = (size_t)fw * fh * 2;
    float s = (float)(1.0 / pyr_scale);
    resize_bilinear(coarse, cw, ch, 2, fine, fw, fh);
    for (i = 0; i < n; i++) fine[i] = fine[i] * s;
}

/* ---- A.8 whole algorithm ---------------------------------------------------- */
static int g_gauss_det_mode = 0;
void rc_oracle_set_gauss_det_mode(int mode) { g_gauss_det_mode = mode; }

int rc_oracle_farneback(const uint8_t* prev, size_t prev_step, const uint8_t* next, size_t next_step, int w, int h,
                        float* flow_out, double pyr_scale, int levels, int winsize, int iterations, int poly_n,
                        double poly_sigma, int flags)
{
    int lw[RC_MAX_LAYERS], lh[RC_MAX_LAYERS];
    int nl = rc_oracle_layers(w, h, pyr_scale, levels, lw, lh), k, it;
    float* prev_flow = NULL; int pw = 0, ph = 0;
    if (nl < 1) return -1;
    for (k = nl - 1; k >= 0; k--) {
        int cw = lw[k], chh = lh[k];
        size_t n = (size_t)cw * chh;
        float* I = (float*)malloc(sizeof(float) * n);
        float* R0 = (float*)malloc(sizeof(float) * n * 5);
        float* R1 = (float*)malloc(sizeof(float) * n * 5);
        float* M = (float*)malloc(sizeof(float) * n * 5);
        float* flow = (k == 0) ? flow_out : (float*)malloc(sizeof(float) * n * 2);
        if (prev_flow) rc_oracle_upsample_flow(prev_flow, pw, ph, flow, cw, chh, pyr_scale);
        else memset(flow, 0, sizeof(float) * n * 2);
        rc_oracle_pyr_layer(prev, w, h, prev_step, pyr_scale, k, cw, chh, I);
        rc_oracle_polyexp(I, cw, chh, poly_n, poly_sigma, R0);
        rc_oracle_pyr_layer(next, w, h, next_step, pyr_scale, k, cw, chh, I);
        rc_oracle_polyexp(I, cw, chh, poly_n, poly_sigma, R1);
        rc_oracle_update_matrices(R0, R1, flow, cw, chh, M);
        for (it = 0; it < iterations; it++) {
            if (flags & RC_FLAG_GAUSSIAN) rc_oracle_update_flow_gauss(M, cw, chh, winsize, g_gauss_det_mode, flow);
            else rc_oracle_update_flow_box(M, cw, chh, winsize, flow);
            if (it < iterations - 1) rc_oracle_update_matrices(R0, R1, flow, cw, chh, M);
        }
        free(I); free(R0); free(R1); free(M);
        if (prev_flow) free(prev_flow);
        prev_flow = flow; pw = cw; ph = chh;
    }
    return 0;
}
