/*
 * oracle/advect_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's Euler / bilinear-gather particle advection:
 *   variant 0  pathlines.cpp:9-46               streamline(pt,color,flow,overlay,dt,iterations)
 *   variant 1  ripcurrents.cpp:656-698          streamline(..., UPPER, prop)   pt += delta*dt/iterations
 *   variant 2  ripcurrents_module.cpp:486-528   streamline(..., UPPER, prop)   pt += delta*dt
 *   variant 3  ripcurrents_module.cpp:531-569   streamline_2  (cut-off r > 5)
 *   variant 4  ripcurrents_module.cpp:572-606   streamline_3  (100 steps of delta*0.1, no cut-off)
 *   variant 5  ripcurrents.cpp:611-651 == module:608-648  streamline_field (home-pixel offset, path length)
 *   variant 6  ripcurrents_module.cpp:650-679   get_delta     (one step, home-pixel offset)
 * and of the Streakline life-cycle (Streakline.cpp:11-20, 34-48) with the vertex
 * move taken from the dense flow by the variant-2 step instead of sparse LK
 * (SURVEY.md section 8(a), row A7).
 *
 * cv::Point_<float> arithmetic is restated per operator: Point*float rounds each
 * product to fp32, Point*double multiplies in double then rounds, Point/int is an
 * fp32 division, Point+Point an fp32 add; evaluation is left to right.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { RC_ADV_PATHLINE = 0, RC_ADV_LEGACY = 1, RC_ADV_MODULE = 2, RC_ADV_CUT5 = 3, RC_ADV_FIXED100 = 4,
       RC_ADV_FIELD = 5, RC_ADV_GET_DELTA = 6 };

/* returns 0 if the particle stopped (left the interior), 1 if delta is valid */
static int gather(const float* flow, int w, int h, float x, float y, float* dx, float* dy)
{
    int xi = (int)floorf(x), yi = (int)floorf(y);
    float xr = x - xi, yr = y - yi;
    const float *p00, *p01, *p10, *p11;
    if (xi < 1 || yi < 1 || xi + 2 > w || yi + 2 > h) return 0;
    p00 = flow + ((size_t)yi * w + xi) * 2; p01 = p00 + 2;
    p10 = p00 + (size_t)w * 2; p11 = p10 + 2;
    {
        float ax = 1 - xr, ay = 1 - yr;
        float t0x = p00[0] * ax * ay, t0y = p00[1] * ax * ay;
        float t1x = p01[0] * xr * ay, t1y = p01[1] * xr * ay;
        float t2x = p10[0] * ax * yr, t2y = p10[1] * ax * yr;
        float t3x = p11[0] * xr * yr, t3y = p11[1] * xr * yr;
        *dx = ((t0x + t1x) + t2x) + t3x;
        *dy = ((t0y + t1y) + t2y) + t3y;
    }
    return 1;
}

/* seeds: n x (x,y) fp32, updated in place.  dist: n fp32 path lengths (variant 5) or NULL.
 * For variants 5 and 6 the particle state is a displacement from its home pixel
 * (home_x, home_y); home = NULL means seed i has home (i % w, i / w) (the dense per-pixel field). */
void rc_oracle_advect(const float* flow, int w, int h, float* seeds, size_t n, float dt, int iterations, float upper,
                      int variant, float* dist, const int* home)
{
    size_t s;
    for (s = 0; s < n; s++) {
        float px = seeds[2 * s], py = seeds[2 * s + 1];
        int xo = 0, yo = 0, it, nit = iterations;
        if (variant == RC_ADV_FIELD || variant == RC_ADV_GET_DELTA) {
            if (home) { xo = home[2 * s]; yo = home[2 * s + 1]; }
            else { xo = (int)(s % (size_t)w); yo = (int)(s / (size_t)w); }
        }
        if (variant == RC_ADV_FIXED100) nit = 100;
        if (variant == RC_ADV_GET_DELTA) nit = 1;
        for (it = 0; it < nit; it++) {
            float dx, dy, r;
            if (!gather(flow, w, h, px + xo, py + yo, &dx, &dy)) break;
            r = sqrtf(dx * dx + dy * dy);
            switch (variant) {
            case RC_ADV_PATHLINE:
                px = px + (dx * dt) / iterations; py = py + (dy * dt) / iterations; break;
            case RC_ADV_LEGACY:
                if (r > upper) goto done;
                px = px + (dx * dt) / iterations; py = py + (dy * dt) / iterations; break;
            case RC_ADV_MODULE:
            case RC_ADV_GET_DELTA:
                if (r > upper) goto done;
                px = px + dx * dt; py = py + dy * dt; break;
            case RC_ADV_CUT5:
                if (r > 5) goto done;
                px = px + dx * dt; py = py + dy * dt; break;
            case RC_ADV_FIXED100:
                px = px + (float)(dx * 0.1); py = py + (float)(dy * 0.1); break;
            case RC_ADV_FIELD:
                if (r > upper) goto done;
                px = px + (dx * dt) / iterations; py = py + (dy * dt) / iterations;
                if (dist) dist[s] = dist[s] + r;
                break;
            }
        }
    done:
        seeds[2 * s] = px; seeds[2 * s + 1] = py;
    }
}

/* One Streakline::runLK frame for E emitters (Streakline.cpp:22-71, drawing removed).
 * vertices: E x cap x (x,y); emitter e holds count[e] vertices in REFERENCE order
 * (index 0 = newest = closest to the generation point).  Every vertex is moved by one
 * variant-2 step (dt, no speed cut-off: upper = +inf) on the dense flow; a move with
 * |dx| > 0.1*w or |dy| > 0.1*h is rejected; then the generation point is inserted at the front. */
void rc_oracle_streakline_step(const float* flow, int w, int h, const float* emitters, int E, float* vertices,
                               int* count, int cap, float dt)
{
    int e, i;
    for (e = 0; e < E; e++) {
        float* v = vertices + (size_t)e * cap * 2;
        int c = count[e];
        for (i = 0; i < c; i++) {
            float x = v[2 * i], y = v[2 * i + 1], dx, dy, nx, ny;
            if (!gather(flow, w, h, x, y, &dx, &dy)) continue;
            nx = x + dx * dt; ny = y + dy * dt;
            if (fabsf(x - nx) > w * 0.1 || fabsf(y - ny) > h * 0.1) continue;
            v[2 * i] = nx; v[2 * i + 1] = ny;
        }
        if (c < cap) {
            memmove(v + 2, v, sizeof(float) * 2 * (size_t)c);
            v[0] = emitters[2 * e]; v[1] = emitters[2 * e + 1];
            count[e] = c + 1;
        }
    }
}
