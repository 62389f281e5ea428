/*
 * oracle/aggregate_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference-owned aggregation that follows the flow:
 *   polar conversion        RipCurrents_main/ripcurrents.cpp:305-309 (cv::cartToPolar, degrees)
 *   cumulative histograms   ripcurrents.cpp:319-330  == ripcurrents_module.cpp:94-107
 *   tail thresholds         ripcurrents.cpp:333-366  == ripcurrents_module.cpp:110-143
 *   classify                ripcurrents.cpp:376-402  == create_flow, ripcurrents_module.cpp:153-182
 *   accumulate + mask       ripcurrents.cpp:414-439  == create_accumulationbuffer, module:189-212
 *   sliding-window mean     RipCurrents_main/main.cpp:1143-1153 (W=10), :1505-1515 (W=100),
 *                           ripcurrents_module.cpp:392-400 (BUFFER_FRAME)
 *   mean subtraction        subtructAverage, ripcurrents_module.cpp:810-863
 *
 * cv::cartToPolar is OpenCV code (not in /root/reference); its default-path
 * arithmetic is restated from SURVEY.md section 8(c) and pinned bit-exactly
 * against cv2 4.13.0 in tests/test_oracle_aggregate.py.
 *
 * Counters are int64 here (the reference uses int, which overflows after
 * ~1035 cumulative 1080p frames); direction index 36 (angle == 360.0f exactly,
 * an out-of-bounds write in the reference) is counted in row 36 of a 37-row
 * hist2d / histsum2d so that "exact bin counts" stays well defined.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define HIST_BINS 50
#define HIST_DIRECTIONS 36
#define HIST_RESOLUTION 20
#define HIST_ROWS (HIST_DIRECTIONS + 1)

/* cv::cartToPolar(x, y, mag, angle, angleInDegrees=true), default (optimized) path */
void rc_oracle_cart_to_polar(const float* xs, const float* ys, size_t n, float* mag, float* ang)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    size_t i;
    for (i = 0; i < n; i++) {
        float x = xs[i], y = ys[i];
        float ax = fabsf(x), ay = fabsf(y);
        float mx = ax > ay ? ax : ay, mn = ax < ay ? ax : ay;
        float c = mn / (mx + (float)DBL_EPSILON);
        float c2 = c * c;
        float a = fmaf(fmaf(fmaf(p7, c2, p5), c2, p3), c2, p1) * c;
        if (ax < ay) a = 90.f - a;
        if (x < 0) a = 180.f - a;
        if (y < 0) a = 360.f - a;
        ang[i] = a;
        mag[i] = sqrtf(fmaf(x, x, y * y));
    }
}

/* flow: n x (dx,dy) interleaved.  hist[50], hist2d[37*50], histsum2d[37], *histsum: cumulative. */
void rc_oracle_histogram(const float* flow, size_t n, int64_t* hist, int64_t* histsum, int64_t* hist2d,
                         int64_t* histsum2d)
{
    size_t i;
    for (i = 0; i < n; i++) {
        float m, a;
        int bin, dir;
        rc_oracle_cart_to_polar(flow + 2 * i, flow + 2 * i + 1, 1, &m, &a);
        bin = (int)(m * HIST_RESOLUTION);
        dir = (int)((a * HIST_DIRECTIONS) / 360);
        if (bin < HIST_BINS && bin >= 0) {
            hist[bin]++; (*histsum)++;
            hist2d[dir * HIST_BINS + bin]++; histsum2d[dir]++;
        }
    }
}

/* thresholds from the cumulative counts; outputs UPPER, UPPER2d[36], prop_above_upper[36] */
void rc_oracle_thresholds(const int64_t* hist, int64_t histsum, const int64_t* hist2d, const int64_t* histsum2d,
                          float* UPPER, float* UPPER2d, float* prop_above_upper)
{
    int64_t threshsum = 0;
    int bin = HIST_BINS - 1, targetbin, a;
    while (threshsum < (histsum * .05)) { threshsum += hist[bin]; bin--; }
    *UPPER = bin / (float)HIST_RESOLUTION;
    targetbin = bin;
    for (a = 0; a < HIST_DIRECTIONS; a++) {
        int64_t t2 = 0, t3 = 0;
        int b = HIST_BINS - 1;
        while (t2 < (histsum2d[a] * .05)) { t2 += hist2d[a * HIST_BINS + b]; b--; }
        UPPER2d[a] = b / (float)HIST_RESOLUTION;
        if (UPPER2d[a] < 0.01) UPPER2d[a] = 0.01;
        b = HIST_BINS - 1;
        while (b > targetbin) { t3 += hist2d[a * HIST_BINS + b]; b--; }
        prop_above_upper[a] = ((float)t3) / threshsum;
    }
}

/* classify + accumulate + mask for one frame.
 * acc_x: n f32 (the .x lane of the reference's CV_32FC3 accumulator); outmask: n u8 (255 = calm);
 * waveclass: n u8 or NULL (0 calm, 1 = "val < .2*framecount" class, 2 = the other);
 * waterclass: n u8 or NULL (3 = above UPPER, 2 = above MID, 1 = above LOWER, 0 = still). */
void rc_oracle_classify_accumulate(const float* flow, size_t n, float UPPER, float MID, float LOWER, int framecount,
                                   float* acc_x, uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass)
{
    size_t i;
    for (i = 0; i < n; i++) {
        float m, a, acc2 = 0.f;
        int val;
        rc_oracle_cart_to_polar(flow + 2 * i, flow + 2 * i + 1, 1, &m, &a);
        if (m > UPPER) { acc2 = 1.f; if (waterclass) waterclass[i] = 3; }
        else if (waterclass) waterclass[i] = m > MID ? 2 : (m > LOWER ? 1 : 0);
        if (framecount > 30) acc_x[i] = acc2 + acc_x[i];
        val = (int)acc_x[i];
        if (val > .1 * framecount) {
            outmask[i] = 0;
            if (waveclass) waveclass[i] = (val < .2 * framecount) ? 1 : 2;
        } else {
            outmask[i] = 255;
            if (waveclass) waveclass[i] = 0;
        }
    }
}

/* one sliding-window update: avg -= slot/W; slot = flow; avg += slot/W  (fp32, separately rounded;
 * cv::Mat / s is evaluated as Mat * (float)(1.0/s)) */
void rc_oracle_window_update(float* avg, float* slot, const float* flow, size_t nfloats, int W)
{
    const float inv = (float)(1.0 / (double)W);
    size_t i;
    for (i = 0; i < nfloats; i++) {
        float a = avg[i] - slot[i] * inv;
        slot[i] = flow[i];
        avg[i] = a + slot[i] * inv;
    }
}

/* subtructAverage: cv::mean (double sums) then per-pixel subtraction of the double mean, stored fp32 */
void rc_oracle_subtract_mean(float* flow, size_t n, double* mean_xy)
{
    double sx = 0, sy = 0;
    size_t i;
    for (i = 0; i < n; i++) { sx += flow[2 * i]; sy += flow[2 * i + 1]; }
    sx = n ? sx / (double)n : 0.0; sy = n ? sy / (double)n : 0.0;
    for (i = 0; i < n; i++) {
        flow[2 * i] = (float)(flow[2 * i] - sx);
        flow[2 * i + 1] = (float)(flow[2 * i + 1] - sy);
    }
    mean_xy[0] = sx; mean_xy[1] = sy;
}
