/*
 * oracle/ingest_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the frame ingest that precedes the flow call in every loop of the reference
 * (SURVEY.md section 8(f), rank 1):
 *     resize(frame, subframe, Size(XDIM,YDIM), 0, 0, INTER_LINEAR);   ripcurrents.cpp:209, main.cpp:258,1111
 *     cvtColor(subframe, f1, COLOR_BGR2GRAY);                         ripcurrents.cpp:210, main.cpp:259,1112
 * Both are OpenCV code (imgproc), absent from /root/reference; restated from OpenCV's published fixed-point scheme
 * and pinned BIT-EXACTLY against cv2 4.13.0 (tests/golden/ingest.npz, tests/test_oracle_ingest.py):
 *   resize 8U bilinear: coefficients round((1-f)*2048), round(f*2048); horizontal pass keeps 11 fractional bits;
 *                       vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2
 *   BGR2GRAY 8U:        (B*3735 + G*19235 + R*9798 + 2^14) >> 15     (OpenCV >= 4.x; 3.4 used 14-bit 1868/9617/4899)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* horizontal: index and weight are clamped together (sx<0 -> 0,f=0; sx>=sw-1 -> sw-1,f=0);
 * vertical (clamp_f == 0): the weight is kept and the two ROW INDICES are clipped individually, as cv::resize does */
static void coef(int d, int sn, int dn, int clamp_f, int* s0, int* a0, int* a1)
{
    double scale = (double)sn / (double)dn;
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_f) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    *s0 = s;
    *a0 = (int)nearbyintf((1.f - f) * 2048.f);
    *a1 = (int)nearbyintf(f * 2048.f);
}

/* bgr: sh x sw x 3 u8 (row stride `step`); gray: dh x dw u8 (dense).  legacy14 != 0 selects OpenCV 3.4's gray weights. */
void rc_oracle_ingest_bgr(const uint8_t* bgr, size_t step, int sw, int sh, uint8_t* gray, int dw, int dh, int legacy14)
{
    int x, y, c;
    for (y = 0; y < dh; y++) {
        int sy, b0, b1, sy1;
        coef(y, sh, dh, 0, &sy, &b0, &b1);
        sy1 = sy + 1 < 0 ? 0 : (sy + 1 < sh ? sy + 1 : sh - 1);
        sy = sy < 0 ? 0 : (sy < sh ? sy : sh - 1);
        for (x = 0; x < dw; x++) {
            int sx, a0, a1, sx1, px[3];
            coef(x, sw, dw, 1, &sx, &a0, &a1);
            sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            for (c = 0; c < 3; c++) {
                int r0 = bgr[(size_t)sy * step + 3 * sx + c] * a0 + bgr[(size_t)sy * step + 3 * sx1 + c] * a1;
                int r1 = bgr[(size_t)sy1 * step + 3 * sx + c] * a0 + bgr[(size_t)sy1 * step + 3 * sx1 + c] * a1;
                px[c] = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
                if (px[c] > 255) px[c] = 255;
            }
            if (legacy14) gray[(size_t)y * dw + x] = (uint8_t)((px[0] * 1868 + px[1] * 9617 + px[2] * 4899 + (1 << 13)) >> 14);
            else gray[(size_t)y * dw + x] = (uint8_t)((px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15);
        }
    }
}

/*
 * INTER_AREA ingest of the PRIMING frame: resize(frame, subframe, Size(XDIM,YDIM), 0, 0, INTER_AREA) + cvtColor(BGR2GRAY),
 * ripcurrents.cpp:186-187, main.cpp:223,570,704,936,1067,1429.  OpenCV code, restated and pinned bit-exactly against cv2
 * 4.13.0 (tests/golden/ingest.npz `area*`, tests/test_oracle_ingest.py), downscaling only (both ratios >= 1):
 *   integer ratios  ("area fast"): 2x2 -> (a+b+c+d+2)>>2; otherwise the integer block sum times (float)(1/area), rounded
 *                    to nearest even;
 *   otherwise       per destination column / row a table of (source index, fp32 weight): a leading partial cell
 *                    (ceil(f1) - f1)/cell when > 1e-3, whole cells 1/cell, a trailing partial cell min(f2 - floor(f2), 1,
 *                    cell)/cell when > 1e-3, with f1 = d*scale, f2 = f1 + scale, cell = min(scale, n - f1) (double arithmetic);
 *                    each source row is reduced horizontally in fp32 in table order (buf += S*alpha), rows are combined as
 *                    sum = beta0*buf0; sum += beta_j*buf_j; the result is rounded to nearest even.
 */
static int area_tab(int d, int sn, double scale, int* idx, float* alpha)
{
    double f1 = d * scale, f2 = f1 + scale, cell = scale < sn - f1 ? scale : sn - f1;
    int s1 = (int)ceil(f1), s2 = (int)floor(f2), k = 0, s;
    if (s2 > sn - 1) s2 = sn - 1;
    if (s1 > s2) s1 = s2;
    if (s1 - f1 > 1e-3) { idx[k] = s1 - 1; alpha[k++] = (float)((s1 - f1) / cell); }
    for (s = s1; s < s2; s++) { idx[k] = s; alpha[k++] = (float)(1.0 / cell); }
    if (f2 - s2 > 1e-3) {
        double t = f2 - s2; if (t > 1.0) t = 1.0; if (t > cell) t = cell;
        idx[k] = s2; alpha[k++] = (float)(t / cell);
    }
    return k;
}

static uint8_t gray_of(const int* px, int legacy14)
{
    return legacy14 ? (uint8_t)((px[0] * 1868 + px[1] * 9617 + px[2] * 4899 + (1 << 13)) >> 14)
                    : (uint8_t)((px[0] * 3735 + px[1] * 19235 + px[2] * 9798 + (1 << 14)) >> 15);
}

/* returns 0, or -1 when a ratio is < 1 (cv::resize then falls back to a bilinear variant that is not restated) */
int rc_oracle_ingest_bgr_area(const uint8_t* bgr, size_t step, int sw, int sh, uint8_t* gray, int dw, int dh, int legacy14)
{
    const double fx = (double)sw / dw, fy = (double)sh / dh;
    const int ix = (int)nearbyint(fx), iy = (int)nearbyint(fy);
    int x, y, c, i, j;
    if (fx < 1.0 || fy < 1.0) return -1;
    if (fabs(fx - ix) < 2.220446049250313e-16 && fabs(fy - iy) < 2.220446049250313e-16) {
        const float sc = 1.f / (float)(ix * iy);
        for (y = 0; y < dh; y++)
            for (x = 0; x < dw; x++) {
                int px[3];
                for (c = 0; c < 3; c++) {
                    int sum = 0;
                    for (j = 0; j < iy; j++)
                        for (i = 0; i < ix; i++) sum += bgr[(size_t)(y * iy + j) * step + 3 * (x * ix + i) + c];
                    if (ix == 2 && iy == 2) px[c] = (sum + 2) >> 2;
                    else { int v = (int)nearbyintf((float)sum * sc); px[c] = v > 255 ? 255 : v; }
                }
                gray[(size_t)y * dw + x] = gray_of(px, legacy14);
            }
        return 0;
    }
    {
        int* xi = (int*)malloc(sizeof(int) * (size_t)(ix + 4)); float* xa = (float*)malloc(sizeof(float) * (size_t)(ix + 4));
        int* yi = (int*)malloc(sizeof(int) * (size_t)(iy + 4)); float* ya = (float*)malloc(sizeof(float) * (size_t)(iy + 4));
        for (y = 0; y < dh; y++) {
            const int ny = area_tab(y, sh, fy, yi, ya);
            for (x = 0; x < dw; x++) {
                const int nx = area_tab(x, sw, fx, xi, xa);
                int px[3];
                for (c = 0; c < 3; c++) {
                    float sum = 0.f;
                    for (j = 0; j < ny; j++) {
                        float buf = 0.f;
                        const uint8_t* row = bgr + (size_t)yi[j] * step;
                        for (i = 0; i < nx; i++) buf = buf + (float)row[3 * xi[i] + c] * xa[i];
                        sum = j == 0 ? ya[j] * buf : sum + ya[j] * buf;
                    }
                    { int v = (int)nearbyintf(sum); px[c] = v < 0 ? 0 : (v > 255 ? 255 : v); }
                }
                gray[(size_t)y * dw + x] = gray_of(px, legacy14);
            }
        }
        free(xi); free(xa); free(yi); free(ya);
    }
    return 0;
}

/*
 * Mask clean-up: create_edges, ripcurrents_module.cpp:216-220 == ripcurrents.cpp:494-496:
 *     morph_window = getStructuringElement(MORPH_ELLIPSE, Size(5,5));
 *     dilate(outmask, outmask, morph_window);
 *     morphologyEx(outmask, outmask, MORPH_GRADIENT, morph_window);     // dilation - erosion
 * OpenCV's 5x5 ellipse is {..X.., XXXXX, XXXXX, XXXXX, ..X..}; taps outside the image are ignored
 * (BORDER_CONSTANT with morphologyDefaultBorderValue).  Pinned bit-exactly against cv2 in tests/test_oracle_ingest.py.
 */
static const unsigned char ELL5[5][5] = {{0, 0, 1, 0, 0}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {1, 1, 1, 1, 1}, {0, 0, 1, 0, 0}};

static void morph5(const uint8_t* in, int w, int h, int dilate, uint8_t* out)
{
    int x, y, dx, dy;
    for (y = 0; y < h; y++)
        for (x = 0; x < w; x++) {
            int v = dilate ? 0 : 255;
            for (dy = -2; dy <= 2; dy++)
                for (dx = -2; dx <= 2; dx++) {
                    int xx = x + dx, yy = y + dy, p;
                    if (!ELL5[dy + 2][dx + 2] || xx < 0 || yy < 0 || xx >= w || yy >= h) continue;
                    p = in[(size_t)yy * w + xx];
                    v = dilate ? (p > v ? p : v) : (p < v ? p : v);
                }
            out[(size_t)y * w + x] = (uint8_t)v;
        }
}

void rc_oracle_edges(const uint8_t* mask, int w, int h, uint8_t* edges)
{
    size_t i, n = (size_t)w * h;
    uint8_t* d = (uint8_t*)malloc(n);
    uint8_t* dd = (uint8_t*)malloc(n);
    uint8_t* de = (uint8_t*)malloc(n);
    morph5(mask, w, h, 1, d);
    morph5(d, w, h, 1, dd);
    morph5(d, w, h, 0, de);
    for (i = 0; i < n; i++) edges[i] = (uint8_t)(dd[i] - de[i]);
    free(d); free(dd); free(de);
}
