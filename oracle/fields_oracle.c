/* CPU oracle: derived fields of the per-pixel particle state.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
 *
 * Restates what the reference computes after streamline_field() with OpenCV primitives:
 *   RipCurrents_main/ripcurrents.cpp:231-279 and the same code factored out in ripcurrents_module.cpp:13-59
 *     split + magnitude(x, y)                       -> "streamfield"  (displacement length)
 *     minMaxLoc(.., &max) ; convertTo(CV_8UC1, 255/max) ; applyColorMap(COLORMAP_JET)      (three times)
 *     divide(streamfield, streamlines_distance)     -> displacement / path-length ratio
 *     streamline_positions: scatter (1,1,1) at floor(pixel + displacement)
 *
 * Pinned against cv2 4.13.0 (tests/test_oracle_fields.py):
 *   magnitude   == cv2.magnitude with cv2.setUseOptimized(False), bit-exact: sqrt(x*x + y*y) in fp32, no FMA.  (The default
 *                  cv2 build routes through IPP, whose result differs by <= 2 ulp; that path cannot be restated.)
 *   divide      == cv2.divide (IEEE: x/0 = +-inf, 0/0 = NaN), bit-exact.  div0_zero != 0 gives the OpenCV 3.x rule the
 *                  reference was written against ("Actually opencv3.2", ripcurrents_module.cpp:6): zero divisor -> 0.
 *                  [3.x rule restated from its documentation, not pinned: no 3.x build here]
 *   max         == cv2.minMaxLoc for NaN-free input (exact).  With NaNs cv2 4.13 returns a SIMD-lane dependent value
 *                  (measured: neither NaN nor the NaN-ignoring maximum); the oracle ignores NaNs and returns NaN only
 *                  when no finite-or-infinite element exists.
 *   convert     == cv2.convertScaleAbs(alpha = 255/max) for non-negative input, bit-exact: t = src * (float)alpha,
 *                  cvRound = cvtss2si (round-half-even; NaN and |t| >= 2^31 -> INT_MIN -> saturates to 0).
 *   JET         == cv2.applyColorMap(COLORMAP_JET) for all 256 levels, bit-exact: closed form
 *                  round_half_even(clamp(382.5 - |4 i - 255 c|, 0, 255)), c = 1,2,3 for B,G,R, with the one entry where
 *                  OpenCV's literal float table rounds the other way (i = 159, blue: 1).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <limits.h>

void rc_oracle_jet_lut(uint8_t* lut /* 256 x BGR */)
{
    for (int i = 0; i < 256; i++)
        for (int c = 1; c <= 3; c++) {
            int k = 4 * i - 255 * c; if (k < 0) k = -k;
            int base = 382 - k;                       /* value is base + 0.5 -> round half to even */
            int v = base + (base & 1);
            v = v < 0 ? 0 : v > 255 ? 255 : v;
            if (i == 159 && c == 1) v = 1;
            lut[i * 3 + (c - 1)] = (uint8_t)v;
        }
}

/* ripcurrents.cpp:232-233: split(streamlines_mat) ; magnitude(x, y) */
void rc_oracle_field_magnitude(const float* field, size_t n, float* mag)
{
    for (size_t i = 0; i < n; i++) {
        volatile float xx = field[2 * i] * field[2 * i], yy = field[2 * i + 1] * field[2 * i + 1];
        mag[i] = sqrtf(xx + yy);
    }
}

/* cv::divide(a, b) for CV_32F */
void rc_oracle_divide(const float* a, const float* b, size_t n, int div0_zero, float* out)
{
    for (size_t i = 0; i < n; i++) out[i] = (div0_zero && b[i] == 0.f) ? 0.f : a[i] / b[i];
}

/* minMaxLoc(src, NULL, &max): see the header for the NaN rule */
double rc_oracle_max(const float* src, size_t n)
{
    int have = 0; float m = 0.f;
    for (size_t i = 0; i < n; i++) {
        if (src[i] != src[i]) continue;
        if (!have || src[i] > m) { m = src[i]; have = 1; }
    }
    return have ? (double)m : (double)NAN;
}

static uint8_t convert_u8(float v, float alpha)
{
    volatile float t = v * alpha;
    int r;
    if (!(t >= -2147483648.f && t < 2147483648.f)) r = INT_MIN;      /* cvtss2si out of range / NaN */
    else r = (int)nearbyintf(t);
    return (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
}

/* module:13-29 (streamline_displacement / streamline_total_motion) given the already computed maximum:
 * src.convertTo(CV_8UC1, 255/max) ; applyColorMap(JET).  gray and bgr are optional. */
void rc_oracle_normalize_jet(const float* src, size_t n, double maxval, uint8_t* gray, uint8_t* bgr)
{
    uint8_t lut[768];
    rc_oracle_jet_lut(lut);
    const float alpha = (float)(255 / maxval);
    for (size_t i = 0; i < n; i++) {
        uint8_t g = convert_u8(src[i], alpha);
        if (gray) gray[i] = g;
        if (bgr) memcpy(bgr + 3 * i, lut + 3 * g, 3);
    }
}

/* module:44-59 / ripcurrents.cpp:261-276: streamline_positions.  density is CV_32FC3; zero_first mirrors the
 * Mat::zeros of ripcurrents.cpp:261. */
void rc_oracle_positions(const float* field, int w, int h, float* density, int zero_first)
{
    if (zero_first) memset(density, 0, (size_t)w * h * 12);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float* p = field + 2 * ((size_t)y * w + x);
            volatile float sx = p[0] + (float)x, sy = p[1] + (float)y;
            float fx = roundf(floorf(sx)), fy = roundf(floorf(sy));
            /* (int) of NaN / out-of-range is INT_MIN on x86 -> fails the xind < 1 test */
            if (!(fx >= 1.f && fy >= 1.f && fx <= (float)(w - 2) && fy <= (float)(h - 2))) continue;
            float* d = density + 3 * ((size_t)(int)fy * w + (int)fx);
            d[0] = d[1] = d[2] = 1.f;
        }
}
