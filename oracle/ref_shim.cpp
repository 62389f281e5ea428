/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C ABI over the REFERENCE'S OWN compiled functions (oracle/ref_build.py #includes this file at the end of the
 * translation unit it generates, after the bodies sliced from /root/reference; pathlines.cpp and Streakline.cpp
 * are linked in whole).  Nothing here computes: every entry point wraps caller memory in cv::Mat headers, calls
 * the reference function, and translates the reference's display encodings back to the integer classes the
 * oracle (oracle/*.c) reports, so that tests/test_oracle_vs_ref.py can compare the two bit for bit.
 *
 * Geometry: the reference compiles XDIM = 640, YDIM = 480 into create_histogram's row loop, averageVector,
 * Streakline::runLK's rejection test and the legacy frame loop; rc_ref_xdim()/rc_ref_ydim() report them.
 */
#include <stdint.h>
#include <limits>

extern "C" {

int rc_ref_xdim(void) { return XDIM; }
int rc_ref_ydim(void) { return YDIM; }

/* create_histogram (ripcurrents_module.cpp:89-144) on a polar image (angle deg, mag, mag) of YDIM rows x w columns.
 * hist2d / histsum2d are over-allocated by the caller to 37 rows so that the reference's direction index 36
 * (angle == 360.0f) lands in row 36 instead of out of bounds. */
void rc_ref_create_histogram(float* polar3, int w, int* hist, int* histsum, int* hist2d, int* histsum2d, float* UPPER,
                             float* UPPER2d, float* prop_above_upper)
{
    cv::Mat current(YDIM, w, CV_32FC3, polar3);
    ref_module::create_histogram(current, hist, *histsum, reinterpret_cast<int(*)[HIST_BINS]>(hist2d), histsum2d, *UPPER,
                                 UPPER2d, prop_above_upper);
}

/* create_flow (module:153-182) + create_accumulationbuffer (module:189-212) for one frame, as the frame loop calls
 * them (fresh accumulator2 / waterclass / out / outmask per frame, ripcurrents.cpp:371-372,419-420).
 * polar3 is modified by create_flow (display rescale), like the reference's `current`. */
void rc_ref_classify_accumulate(float* polar3, int w, int h, float UPPER, float MID, float LOWER, float* UPPER2d,
                                int framecount, float* acc_x, uint8_t* outmask, uint8_t* waveclass, uint8_t* waterclass)
{
    cv::Mat current(h, w, CV_32FC3, polar3);
    cv::Mat accumulator = cv::Mat::zeros(h, w, CV_32FC3), accumulator2 = cv::Mat::zeros(h, w, CV_32FC3);
    cv::Mat wc = cv::Mat::zeros(h, w, CV_32FC3), out = cv::Mat::zeros(h, w, CV_32FC3), mask = cv::Mat::zeros(h, w, CV_8UC1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) accumulator.ptr<Pixel3>(y, x)->x = acc_x[(size_t)y * w + x];
    ref_module::create_flow(current, wc, accumulator2, UPPER, MID, LOWER, UPPER2d);
    ref_module::create_accumulationbuffer(accumulator, accumulator2, out, mask, framecount);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t i = (size_t)y * w + x;
            acc_x[i] = accumulator.ptr<Pixel3>(y, x)->x;
            outmask[i] = *mask.ptr<uchar>(y, x);
            const Pixel3 o = *out.ptr<Pixel3>(y, x), c = *wc.ptr<Pixel3>(y, x);
            if (waveclass) waveclass[i] = o.z == 1.f ? 1 : (o.x == 1.f ? 2 : 0);
            if (waterclass) waterclass[i] = c.x == .5f ? 3 : (c.z == 1.f ? 2 : (c.z == .5f ? 1 : 0));
        }
}

/* the seven advection variants (same numbering as oracle/advect_oracle.c); variant 7 = the LEGACY copy of
 * streamline_field (ripcurrents.cpp:611-651), which the oracle treats as identical to module:608-648 */
void rc_ref_advect(float* flow, int w, int h, float* seeds, size_t n, float dt, int iterations, float upper, int variant,
                   float* dist, const int* home)
{
    cv::Mat f(h, w, CV_32FC2, flow), overlay;
    float prop[HIST_DIRECTIONS] = {0};
    Pixel2* pts = reinterpret_cast<Pixel2*>(seeds);
    for (size_t s = 0; s < n; s++) {
        int xo = 0, yo = 0;
        if (variant >= 5) {
            if (home) { xo = home[2 * s]; yo = home[2 * s + 1]; }
            else { xo = (int)(s % (size_t)w); yo = (int)(s / (size_t)w); }
        }
        float dummy = 0.f;
        switch (variant) {
        case 0: streamline(pts + s, cv::Scalar(), f, overlay, dt, iterations); break;                       /* pathlines.cpp */
        case 1: ref_legacy::streamline(pts + s, cv::Scalar(), f, overlay, dt, iterations, upper, prop); break;
        case 2: ref_module::streamline(pts + s, cv::Scalar(), f, overlay, dt, iterations, upper, prop); break;
        case 3: ref_module::streamline_2(pts + s, cv::Scalar(), f, overlay, dt, iterations, upper, prop); break;
        case 4: ref_module::streamline_3(pts + s, cv::Scalar(), f, overlay, dt, iterations, upper, prop); break;
        case 5: ref_module::streamline_field(pts + s, dist ? dist + s : &dummy, xo, yo, f, dt, iterations, upper, prop); break;
        case 6: ref_module::get_delta(pts + s, xo, yo, f, dt, upper); break;
        case 7: ref_legacy::streamline_field(pts + s, dist ? dist + s : &dummy, xo, yo, f, dt, iterations, upper, prop); break;
        }
    }
}

/* main.cpp:1142-1153: one iteration of the sliding-window block.  ring = W slots of w*h*2 floats (the reference's
 * `buffer`), *currentBuffer is advanced as the reference does. */
void rc_ref_window_update(float* avg, float* ring, int* currentBuffer, float* flow, int w, int h, int W)
{
    const size_t nf = (size_t)w * h * 2;
    cv::Mat averageCurrent(h, w, CV_32FC2, avg), current(h, w, CV_32FC2, flow);
    std::vector<cv::Mat> buffer;
    for (int i = 0; i < W; i++) buffer.push_back(cv::Mat(h, w, CV_32FC2, ring + (size_t)i * nf));
    const int slot = *currentBuffer;
    ref_main::window_block(averageCurrent, buffer, *currentBuffer, W, current);
    memcpy(ring + (size_t)slot * nf, buffer[slot].data, nf * sizeof(float));     /* buffer[slot] = current.clone() rebound it */
}

/* averageVector's window update (module:386-400) on XDIM x YDIM data.  `buffer` is passed BY VALUE in the reference,
 * so the caller's slot is never rewritten; only `average` changes. */
void rc_ref_average_vector(float* buffer_slot, float* current_flow, float* average, float UPPER)
{
    std::vector<cv::Mat> buffer;
    buffer.push_back(cv::Mat(YDIM, XDIM, CV_32FC2, buffer_slot));
    cv::Mat current(YDIM, XDIM, CV_32FC2, current_flow), avg(YDIM, XDIM, CV_32FC2, average), color;
    ref_module::averageVector(buffer, current, 0, avg, color, nullptr, 0.f, UPPER);
}

/* One Streakline::runLK frame (Streakline.cpp:22-71) per emitter.  Vertex motion: the build replaces sparse LK by the
 * dense-flow step, so the LK hook moves every vertex with the reference's own module streamline() (one iteration,
 * no cut-off); rejection, hand-over and the insertion of the generation point are the reference's code. */
void rc_ref_streakline_step(float* flow, int w, int h, const float* emitters, int E, float* vertices, int* count, int cap,
                            float dt)
{
    cv::Mat f(h, w, CV_32FC2, flow), overlay, img;
    cv::ref_hooks::lk() = [&](const std::vector<cv::Point2f>& in, std::vector<cv::Point2f>& out) {
        float prop[HIST_DIRECTIONS] = {0};
        for (size_t i = 0; i < in.size(); i++) {
            Pixel2 p = in[i];
            ref_module::streamline(&p, cv::Scalar(), f, overlay, dt, 1, std::numeric_limits<float>::infinity(), prop);
            out[i] = p;
        }
    };
    for (int e = 0; e < E; e++) {
        float* v = vertices + (size_t)e * cap * 2;
        Streakline s(Pixel2(emitters[2 * e], emitters[2 * e + 1]));
        s.vertices.clear();
        for (int i = 0; i < count[e]; i++) s.vertices.push_back(Pixel2(v[2 * i], v[2 * i + 1]));
        s.runLK(cv::UMat(), cv::UMat(), img);
        int c = (int)s.vertices.size();
        if (c > cap) c = cap;                 /* the reference's vector is unbounded; callers keep count < cap */
        for (int i = 0; i < c; i++) { v[2 * i] = s.vertices[i].x; v[2 * i + 1] = s.vertices[i].y; }
        count[e] = c;
    }
    cv::ref_hooks::lk() = nullptr;
}

/* The legacy frame loop's aggregation (ripcurrents.cpp:305-439) with its cumulative state (:131-154), exactly as
 * written in main(): polar conversion, histograms, thresholds, classify, accumulate, mask.  XDIM x YDIM only. */
void* rc_ref_legacy_new(void) { return new ref_legacy::FrameLoopState(); }
void rc_ref_legacy_free(void* s) { delete static_cast<ref_legacy::FrameLoopState*>(s); }
void rc_ref_legacy_frame(void* sp, const float* flow, int framecount, uint8_t* outmask, float* UPPER, float* UPPER2d,
                         float* prop_above_upper, int* hist, int* histsum, int* hist2d, int* histsum2d, float* acc_x)
{
    ref_legacy::FrameLoopState* s = static_cast<ref_legacy::FrameLoopState*>(sp);
    cv::Mat current = cv::Mat(YDIM, XDIM, CV_32FC2, const_cast<float*>(flow)).clone(), mask;
    s->frame(current, framecount, mask);
    for (int y = 0; y < YDIM; y++) {
        memcpy(outmask + (size_t)y * XDIM, mask.ptr<uchar>(y), XDIM);
        for (int x = 0; x < XDIM; x++) acc_x[(size_t)y * XDIM + x] = s->accumulator.ptr<Pixel3>(y, x)->x;
    }
    *UPPER = s->UPPER; *histsum = s->histsum;
    memcpy(UPPER2d, s->UPPER2d, sizeof s->UPPER2d);
    memcpy(prop_above_upper, s->prop_above_upper, sizeof s->prop_above_upper);
    memcpy(hist, s->hist, sizeof s->hist);
    memcpy(hist2d, s->hist2d, sizeof s->hist2d);
    memcpy(histsum2d, s->histsum2d, sizeof s->histsum2d);
}

}  /* extern "C" */
