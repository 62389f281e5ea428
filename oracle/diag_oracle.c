/* CPU oracle: flow-derived diagnostics.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Restates RipCurrents_main/ripcurrents_module.cpp:
 *   :900-1015   subtructMeanMagnitude   (the printf diagnostics are not reproduced, only `meanval`)
 *   :1017-1057  vectorToColor           hue = direction, value = |v| * 255 / (previous frame's max), HSV -> BGR
 *   :1059-1138  shearRateToColor        hue = 128 - ||J||_F * 128 / (previous frame's max), J = central differences at +-10 px
 * as compiled by GCC for x86-64 and linked against this image's libm (glibc 2.39):
 *   - sqrt / atan2 on float arguments resolve to the float overloads (sqrtf, atan2f: libm's own, called directly here);
 *   - float -> unsigned char stores are `cvttss2si` + low byte: truncation toward zero, wrap modulo 256, and
 *     NaN / out-of-int-range -> 0x80000000 -> 0;
 *   - the function-local `static` maxima become explicit in/out arguments.
 * cvtColor(COLOR_HSV2BGR) on 8-bit data is pinned exhaustively (all 2^24 inputs, tests/test_oracle_diag.py) against
 * cv2 4.13.0: h *= 6/180, sector tables, tab = {v, v(1-s), v(1-s f), v(1-s(1-f))} with the two inner products as fused
 * multiply-adds, result truncated (not rounded) after * 255.  That is what cv2 does for every full block of 32 pixels of
 * a row; cv2 rounds the remaining (width mod 32) pixels of a row instead -- an inconsistency of the library that is
 * not reproduced (1920, 640 and 3840 are multiples of 32).  `fma = 0` gives cv2's setUseOptimized(False) variant.
 */
#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static uint8_t to_uchar(float f)
{
    int r = (f >= -2147483648.f && f < 2147483648.f) ? (int)f : INT_MIN;
    return (uint8_t)(r & 0xff);
}

void rc_oracle_hsv2bgr(const uint8_t* hsv, size_t n, uint8_t* bgr, int fma)
{
    static const int sector[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    const float hscale = 6.f / 180.f;
    for (size_t i = 0; i < n; i++) {
        volatile float h = (float)hsv[3 * i] * hscale;
        const float s = (float)hsv[3 * i + 1] * (1.f / 255.f), v = (float)hsv[3 * i + 2] * (1.f / 255.f);
        while (h >= 6.f) h = h - 6.f;
        int sec = (int)floorf(h);
        volatile float f = h - (float)sec;
        if ((unsigned)sec >= 6u) { sec = 0; f = 0.f; }
        volatile float omf = 1.f - f, oms = 1.f - s;
        volatile float q2, q3;
        if (fma) { q2 = fmaf(-s, f, 1.f); q3 = fmaf(-s, omf, 1.f); }
        else { volatile float p2 = s * f, p3 = s * omf; q2 = 1.f - p2; q3 = 1.f - p3; }
        volatile float tab[4];
        tab[0] = v; tab[1] = v * oms; tab[2] = v * q2; tab[3] = v * q3;
        for (int c = 0; c < 3; c++) {
            volatile float t = tab[sector[sec][c]] * 255.f;
            int r = (int)t;                                  /* truncation */
            bgr[3 * i + c] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
        }
    }
}

/* module:900-1015.  Returns meanval (sequential fp32 accumulation, as the reference's loop does it). */
float rc_oracle_subtract_mean_magnitude(float* flow, size_t n)
{
    volatile float meanval = 0.f;
    for (size_t i = 0; i < n; i++) {
        volatile float xx = flow[2 * i] * flow[2 * i], yy = flow[2 * i + 1] * flow[2 * i + 1];
        volatile float ss = xx + yy;
        meanval = meanval + sqrtf(ss);
    }
    meanval = meanval / (float)(int)n;
    const float mv = meanval;
    for (size_t i = 0; i < n; i++) {
        const float x = flow[2 * i], y = flow[2 * i + 1];
        volatile float xx = x * x, yy = y * y;
        volatile float ss = xx + yy;
        const float mag = sqrtf(ss);
        volatile float ux = 0.f, uy = 0.f;
        if (mag != 0.f) { ux = x / mag; uy = y / mag; }
        volatile float d = mag - mv;
        flow[2 * i] = ux * d; flow[2 * i + 1] = uy * d;
    }
    return mv;
}

static float theta_deg(float y, float x)
{
    /* float theta = atan2(ptr->y, ptr->x)*180/M_PI;  theta += theta < 0 ? 360 : 0; */
    volatile double t = (double)atan2f(y, x) * 180;
    volatile float theta = (float)(t / M_PI);
    if (theta < 0) theta = theta + 360.f;
    return theta;
}

/* module:1017-1057.  hsv (optional) receives the image before cvtColor; *max_displacement: in = previous frame's
 * maximum (the reference's static), out = this frame's. */
void rc_oracle_vector_to_color(const float* flow, int w, int h, uint8_t* hsv, uint8_t* bgr, float* max_displacement, int fma)
{
    const float maxd = *max_displacement;
    float newmax = 0.f;
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++) {
        const float x = flow[2 * i], y = flow[2 * i + 1];
        volatile float xx = x * x, yy = y * y;
        volatile float ss = xx + yy;
        const float mag = sqrtf(ss);
        volatile float hh = theta_deg(y, x) / 2;
        volatile float vv = mag * 255.f;
        vv = vv / maxd;
        uint8_t px[3] = {to_uchar(hh), 255, to_uchar(vv)};
        if (mag > newmax) newmax = mag;
        if (hsv) { hsv[3 * i] = px[0]; hsv[3 * i + 1] = px[1]; hsv[3 * i + 2] = px[2]; }
        rc_oracle_hsv2bgr(px, 1, bgr + 3 * i, fma);
    }
    *max_displacement = newmax;
}

/* module:1059-1138.  img is the caller's 8-bit image (in/out): interior pixels are overwritten with
 * (hue, 255, 255), then the WHOLE image goes through HSV2BGR, as in the reference. */
void rc_oracle_shear_to_color(const float* flow, int w, int h, uint8_t* img, float* max_frobenius, int fma)
{
    const int off = 10;
    const float maxf = *max_frobenius;
    float newmax = 0.f;
    for (int row = off; row < h - off; row++)
        for (int col = off; col < w - off; col++) {
            const float* above = flow + 2 * ((size_t)(row - off) * w + col);
            const float* below = flow + 2 * ((size_t)(row + off) * w + col);
            const float* left = flow + 2 * ((size_t)row * w + col - off);
            const float* right = flow + 2 * ((size_t)row * w + col + off);
            volatile float j00 = right[0] - left[0], j01 = above[0] - below[0], j10 = right[1] - left[1], j11 = above[1] - below[1];
            volatile float a = j00 * j00, b = j01 * j01, c = j10 * j10, d = j11 * j11;
            volatile float fr = a + b;
            fr = fr + c; fr = fr + d;
            const float frob = sqrtf(fr);
            volatile float t = frob * 128.f;
            t = t / maxf;
            t = 128.f - t;
            uint8_t* p = img + 3 * ((size_t)row * w + col);
            p[0] = to_uchar(t); p[1] = 255; p[2] = 255;
            if (frob > newmax) newmax = frob;              /* max(frobeniusNorm, new): NaN never enters */
        }
    *max_frobenius = newmax;
    rc_oracle_hsv2bgr(img, (size_t)w * h, img, fma);
}
