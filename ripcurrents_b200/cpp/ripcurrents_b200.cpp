// Implementation of the header-compatible entry points (ripcurrents.hpp / Streakline.hpp / pathlines.h) on top of
// the C ABI.  Host C++ only: no kernels here, no OpenCV algorithm calls; the threshold scans of create_histogram run on
// the caller's own 50 / 36x50 int arrays exactly as ripcurrents_module.cpp:110-143 does.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "ripcurrents.hpp"
#include "Streakline.hpp"
#include "pathlines.h"

namespace {
rc_ctx* g_ctx = nullptr;
int g_device = -1;

void check(int rc, const char* what)
{
    if (rc >= 0) return;
    std::fprintf(stderr, "ripcurrents_b200: %s failed: %s (%s)\n", what, rc_error_string(rc),
                 g_ctx ? rc_last_error(g_ctx) : "no context");
    std::abort();   // the reference has no error channel either (CV_Assert / exit)
}

void require(bool ok, const char* what)
{
    if (ok) return;
    std::fprintf(stderr, "ripcurrents_b200: bad argument: %s\n", what);
    std::abort();
}
}  // namespace

namespace rc {

void set_default_device(int device) { g_device = device; }

rc_ctx* default_context()
{
    if (!g_ctx) {
        int dev = g_device;
        if (dev < 0) { const char* e = std::getenv("RC_B200_DEVICE"); dev = e ? std::atoi(e) : 0; }
        int rc = rc_create(&g_ctx, dev);
        if (rc != RC_OK) {
            std::fprintf(stderr, "ripcurrents_b200: rc_create(device %d) failed: %s -- a CUDA device is required, "
                                 "there is no CPU fallback\n", dev, rc_error_string(rc));
            std::abort();
        }
    }
    return g_ctx;
}

void calcOpticalFlowFarneback(const cv::Mat& prev, const cv::Mat& next, cv::Mat& flow, double pyr_scale, int levels,
                              int winsize, int iterations, int poly_n, double poly_sigma, int flags)
{
    require(prev.type() == CV_8UC1 && next.type() == CV_8UC1 && prev.rows == next.rows && prev.cols == next.cols,
            "calcOpticalFlowFarneback: prev/next must be CV_8UC1 of equal size");
    if (flow.rows != prev.rows || flow.cols != prev.cols || flow.type() != CV_32FC2) flow.create(prev.rows, prev.cols, CV_32FC2);
    check(rc_farneback(default_context(), prev.data, prev.step, next.data, next.step, prev.cols, prev.rows,
                       flow.ptr<float>(), flow.step, pyr_scale, levels, winsize, iterations, poly_n, poly_sigma, flags),
          "rc_farneback");
}

void ingest(const cv::Mat& frame_bgr, cv::Mat& gray, cv::Size size, bool area)
{
    require(frame_bgr.type() == CV_8UC3 && size.width > 0 && size.height > 0, "ingest: CV_8UC3 frame required");
    if (gray.rows != size.height || gray.cols != size.width || gray.type() != CV_8UC1) gray.create(size.height, size.width, CV_8UC1);
    check(rc_ingest_bgr(default_context(), frame_bgr.data, frame_bgr.step, frame_bgr.cols, frame_bgr.rows, gray.data, gray.step,
                        size.width, size.height, area ? RC_INGEST_AREA : 0), "rc_ingest_bgr");
}

void flowToPolar(const cv::Mat& flow, cv::Mat& polar)
{
    require(flow.type() == CV_32FC2, "flowToPolar: flow must be CV_32FC2");
    const int w = flow.cols, h = flow.rows;
    polar.create(h, w, CV_32FC3);
    cv::Mat dense = (flow.step == (size_t)w * 8) ? flow : flow.clone();
    std::vector<float> mag((size_t)w * h), ang((size_t)w * h);
    check(rc_cart_to_polar(default_context(), dense.ptr<float>(), (size_t)w * h, mag.data(), ang.data()), "rc_cart_to_polar");
    for (int y = 0; y < h; y++) {
        Pixel3* o = polar.ptr<Pixel3>(y);
        for (int x = 0; x < w; x++) { const size_t i = (size_t)y * w + x; o[x] = Pixel3(ang[i], mag[i], mag[i]); }
    }
}

void streamline_field_all(cv::Mat& field, cv::Mat& distance, const cv::Mat& flow, float dt, int iterations, float UPPER)
{
    require(field.type() == CV_32FC2 && distance.type() == CV_32FC1 && flow.type() == CV_32FC2 &&
            field.rows == flow.rows && field.cols == flow.cols && distance.rows == flow.rows && distance.cols == flow.cols,
            "streamline_field_all: size / type mismatch");
    require(field.step == (size_t)field.cols * 8 && distance.step == (size_t)distance.cols * 4, "streamline_field_all: dense Mats required");
    check(rc_advect(default_context(), flow.ptr<float>(), flow.step, flow.cols, flow.rows, field.ptr<float>(),
                    (size_t)flow.cols * flow.rows, dt, iterations, UPPER, RC_ADV_FIELD, distance.ptr<float>(), nullptr),
          "rc_advect(field)");
}

void streamlines_all(Pixel2* pts, int n, const cv::Mat& flow, float dt, int iterations, float UPPER, int variant)
{
    require(flow.type() == CV_32FC2, "streamlines_all: flow must be CV_32FC2");
    if (n <= 0) return;
    check(rc_advect(default_context(), flow.ptr<float>(), flow.step, flow.cols, flow.rows, reinterpret_cast<float*>(pts),
                    (size_t)n, dt, iterations, UPPER, variant, nullptr, nullptr), "rc_advect");
}

void window_average(std::vector<cv::Mat>& buffer, int& currentBuffer, const cv::Mat& flow, cv::Mat& average)
{
    // main.cpp:1143-1153 with the buffer held by the caller (as there); the arithmetic runs on the device.
    rc_ctx* c = default_context();
    const int W = (int)buffer.size(), w = flow.cols, h = flow.rows;
    require(W > 0 && flow.type() == CV_32FC2 && average.type() == CV_32FC2 && average.rows == h && average.cols == w,
            "window_average: size / type mismatch");
    static int cfg_w = 0, cfg_h = 0, cfg_W = 0;
    if (cfg_w != w || cfg_h != h || cfg_W != W) {
        check(rc_window_configure(c, w, h, W), "rc_window_configure");
        cfg_w = w; cfg_h = h; cfg_W = W;
    }
    check(rc_window_update(c, flow.ptr<float>(), flow.step), "rc_window_update");
    buffer[currentBuffer] = flow.clone();
    check(rc_window_get(c, average.ptr<float>(), average.step), "rc_window_get");
    currentBuffer++;
    if (currentBuffer >= W) currentBuffer = 0;
}

}  // namespace rc

// ---- ripcurrents.hpp -------------------------------------------------------------------------------------------------
static cv::Mat dense_mat(const cv::Mat& m) { return m.isContinuous() ? m : m.clone(); }
static void one_seed(Pixel2* pt, const cv::Mat& flow, float dt, int iterations, float upper, int variant, float* dist,
                     int xo, int yo)
{
    require(flow.type() == CV_32FC2, "flow must be CV_32FC2");
    int32_t home[2] = {xo, yo};
    check(rc_advect(rc::default_context(), flow.ptr<float>(), flow.step, flow.cols, flow.rows, reinterpret_cast<float*>(pt), 1,
                    dt, iterations, upper, variant, dist, (variant == RC_ADV_FIELD || variant == RC_ADV_GET_DELTA) ? home : nullptr),
          "rc_advect");
}

void streamline_field(Pixel2* pt, float* distancetraveled, int xoffset, int yoffset, cv::Mat flow, float dt, int iterations,
                      float UPPER, float[HIST_DIRECTIONS])
{
    one_seed(pt, flow, dt, iterations, UPPER, RC_ADV_FIELD, distancetraveled, xoffset, yoffset);
}

void streamline(Pixel2* pt, cv::Scalar, cv::Mat flow, cv::Mat, float dt, int iterations, float UPPER, float[HIST_DIRECTIONS])
{
    one_seed(pt, flow, dt, iterations, UPPER, RC_ADV_MODULE, nullptr, 0, 0);       // ripcurrents_module.cpp:486-528
}

void streamline_2(Pixel2* pt, cv::Scalar, cv::Mat flow, cv::Mat, float dt, int iterations, float UPPER, float[HIST_DIRECTIONS])
{
    one_seed(pt, flow, dt, iterations, UPPER, RC_ADV_CUT5, nullptr, 0, 0);
}

void streamline_3(Pixel2* pt, cv::Scalar, cv::Mat flow, cv::Mat, float dt, int iterations, float UPPER, float[HIST_DIRECTIONS])
{
    one_seed(pt, flow, dt, iterations, UPPER, RC_ADV_FIXED100, nullptr, 0, 0);
}

void streamline(Pixel2* pt, cv::Scalar, cv::Mat flow, cv::Mat, float dt, int iterations)   // pathlines.h
{
    one_seed(pt, flow, dt, iterations, 0.f, RC_ADV_PATHLINE, nullptr, 0, 0);
}

void get_delta(Pixel2* pt, int xoffset, int yoffset, cv::Mat flow, float dt, float UPPER)
{
    one_seed(pt, flow, dt, 1, UPPER, RC_ADV_GET_DELTA, nullptr, xoffset, yoffset);
}

void get_streamlines(cv::Mat&, cv::Mat&, cv::Mat&, int streamlines, Pixel2 streampt[], int, int, cv::Mat& current,
                     float UPPER, float[])
{
    // ripcurrents_module.cpp:71-79: streamline(streampt+s, ..., 0.1, 100, UPPER, prop) for every seed; colour-map
    // compositing of the overlay is visualisation and stays with the caller.
    rc::streamlines_all(streampt, streamlines, current, 0.1f, 100, UPPER, RC_ADV_MODULE);
}

void create_histogram(cv::Mat current, int hist[HIST_BINS], int& histsum, int hist2d[HIST_DIRECTIONS][HIST_BINS],
                      int histsum2d[HIST_DIRECTIONS], float& UPPER, float UPPER2d[HIST_DIRECTIONS],
                      float prop_above_upper[HIST_DIRECTIONS])
{
    require(current.type() == CV_32FC3, "create_histogram: current must be the CV_32FC3 polar image");
    rc_ctx* c = rc::default_context();
    // counts of THIS frame on the device, then added to the caller's cumulative arrays (ripcurrents.cpp:147-153)
    check(rc_hist_reset(c), "rc_hist_reset");
    check(rc_hist_from_polar(c, current.ptr<float>(), current.step, current.cols, current.rows), "rc_hist_from_polar");
    static int64_t h1[RC_HIST_BINS], h2[RC_HIST_ROWS * RC_HIST_BINS], hs2[RC_HIST_ROWS];
    int64_t hs = 0;
    check(rc_hist_get(c, h1, &hs, h2, hs2), "rc_hist_get");
    for (int b = 0; b < HIST_BINS; b++) hist[b] += (int)h1[b];
    histsum += (int)hs;
    for (int a = 0; a < HIST_DIRECTIONS; a++) {
        for (int b = 0; b < HIST_BINS; b++) hist2d[a][b] += (int)h2[a * RC_HIST_BINS + b];
        histsum2d[a] += (int)hs2[a];
    }
    // direction index 36 (angle == 360.0f): out of bounds in the reference; its pixels are in hist/histsum only.

    // thresholds on the caller's arrays: ripcurrents_module.cpp:110-143
    int threshsum = 0;
    int bin = HIST_BINS - 1;
    while (threshsum < (histsum * .05)) { threshsum += hist[bin]; bin--; }
    UPPER = bin / float(HIST_RESOLUTION);
    const int targetbin = bin;
    for (int angle = 0; angle < HIST_DIRECTIONS; angle++) {
        int threshsum2 = 0;
        int b = HIST_BINS - 1;
        while (threshsum2 < (histsum2d[angle] * .05)) { threshsum2 += hist2d[angle][b]; b--; }
        UPPER2d[angle] = b / float(HIST_RESOLUTION);
        if (UPPER2d[angle] < 0.01) UPPER2d[angle] = 0.01;
        int threshsum3 = 0;
        b = HIST_BINS - 1;
        while (b > targetbin) { threshsum3 += hist2d[angle][b]; b--; }
        prop_above_upper[angle] = ((float)threshsum3) / threshsum;
    }
}

void averageVector(std::vector<cv::Mat> buffer, cv::Mat& current, int update_ith_buffer, cv::Mat& average, cv::Mat&, double**,
                   float, float UPPER)
{
    require(update_ith_buffer >= 0 && update_ith_buffer < (int)buffer.size(), "averageVector: buffer index");
    cv::Mat old = dense_mat(buffer[update_ith_buffer]);
    require(current.type() == CV_32FC2 && average.type() == CV_32FC2 && old.type() == CV_32FC2 && average.isContinuous() &&
            average.rows == current.rows && average.cols == current.cols && old.rows == current.rows && old.cols == current.cols,
            "averageVector: CV_32FC2 current / average / buffer of one size required");
    check(rc_average_vector(rc::default_context(), old.ptr<float>(), current.ptr<float>(), current.step, current.cols,
                            current.rows, average.ptr<float>(), nullptr, BUFFER_FRAME, 2.f, UPPER), "rc_average_vector");
}

void create_flow(cv::Mat current, cv::Mat waterclass, cv::Mat accumulator2, float UPPER, float MID, float LOWER,
                 float UPPER2d[HIST_DIRECTIONS])
{
    require(current.type() == CV_32FC3 && waterclass.type() == CV_32FC3 && accumulator2.type() == CV_32FC3,
            "create_flow: CV_32FC3 images required");
    check(rc_create_flow(rc::default_context(), current.ptr<float>(), current.step, waterclass.ptr<float>(), waterclass.step,
                         accumulator2.ptr<float>(), accumulator2.step, current.cols, current.rows, UPPER, MID, LOWER, UPPER2d),
          "rc_create_flow");
}

void create_accumulationbuffer(cv::Mat& accumulator, cv::Mat accumulator2, cv::Mat& out, cv::Mat outmask, int framecount)
{
    require(accumulator.type() == CV_32FC3 && accumulator2.type() == CV_32FC3 && out.type() == CV_32FC3 &&
            outmask.type() == CV_8UC1, "create_accumulationbuffer: image types");
    check(rc_create_accumulationbuffer(rc::default_context(), accumulator.ptr<float>(), accumulator.step,
                                       accumulator2.ptr<float>(), accumulator2.step, out.ptr<float>(), out.step,
                                       outmask.data, outmask.step, accumulator.cols, accumulator.rows, framecount),
          "rc_create_accumulationbuffer");
}

void create_edges(cv::Mat& outmask)
{
    require(outmask.type() == CV_8UC1, "create_edges: CV_8UC1 mask required");
    check(rc_mask_edges(rc::default_context(), outmask.data, outmask.step, 0, outmask.cols, outmask.rows, 1, outmask.data,
                        outmask.step, 0), "rc_mask_edges");
}

// ---- derived particle fields (ripcurrents_module.cpp:13-59) ----------------------------------------------------------
namespace {
cv::Mat dense(const cv::Mat& m) { return m.isContinuous() ? m : m.clone(); }
}

void streamline_displacement(cv::Mat& streamfield, cv::Mat& streamoverlay_color)
{
    require(streamfield.type() == CV_32FC1, "streamline_displacement: CV_32FC1 input required");
    cv::Mat src = dense(streamfield);
    streamoverlay_color.create(src.rows, src.cols, CV_8UC3);
    check(rc_normalize_jet(rc::default_context(), src.ptr<float>(), (size_t)src.rows * src.cols, nullptr,
                           streamoverlay_color.data, nullptr), "rc_normalize_jet");
}

void streamline_total_motion(cv::Mat& streamlines_distance, cv::Mat& streamoverlay_color)
{
    streamline_displacement(streamlines_distance, streamoverlay_color);        // module:23-29 is the same chain
}

void streamline_ratio(cv::Mat& streamfield, cv::Mat& streamlines_distance, cv::Mat& streamoverlay_color)
{
    require(streamfield.type() == CV_32FC1 && streamlines_distance.type() == CV_32FC1 &&
            streamfield.rows == streamlines_distance.rows && streamfield.cols == streamlines_distance.cols,
            "streamline_ratio: two CV_32FC1 images of one size required");
    cv::Mat a = dense(streamfield), b = dense(streamlines_distance);
    streamoverlay_color.create(a.rows, a.cols, CV_8UC3);
    check(rc_ratio_jet(rc::default_context(), a.ptr<float>(), b.ptr<float>(), (size_t)a.rows * a.cols, 0, nullptr, nullptr,
                       streamoverlay_color.data, nullptr), "rc_ratio_jet");
}

void streamline_positions(cv::Mat& streamlines_mat, cv::Mat& streamline_density)
{
    require(streamlines_mat.type() == CV_32FC2 && streamline_density.type() == CV_32FC3 && streamline_density.isContinuous() &&
            streamline_density.rows == streamlines_mat.rows && streamline_density.cols == streamlines_mat.cols,
            "streamline_positions: CV_32FC2 field and a dense CV_32FC3 density of the same size required");
    cv::Mat f = dense(streamlines_mat);
    check(rc_streamline_positions(rc::default_context(), f.ptr<float>(), f.cols, f.rows, streamline_density.ptr<float>(),
                                  RC_FIELDS_KEEP_DENSITY), "rc_streamline_positions");
}

void rc::particle_fields(const cv::Mat& streamlines_mat, const cv::Mat& streamlines_distance, cv::Mat* streamfield,
                         cv::Mat* displacement_color, cv::Mat* motion_color, cv::Mat* ratio_color, cv::Mat* density,
                         double maxes[3], int flags)
{
    require(streamlines_mat.type() == CV_32FC2 && streamlines_distance.type() == CV_32FC1 &&
            streamlines_mat.rows == streamlines_distance.rows && streamlines_mat.cols == streamlines_distance.cols,
            "particle_fields: CV_32FC2 displacements and CV_32FC1 distances of one size required");
    cv::Mat f = dense(streamlines_mat), d = dense(streamlines_distance);
    const int r = f.rows, c = f.cols;
    if (streamfield) streamfield->create(r, c, CV_32FC1);
    if (displacement_color) displacement_color->create(r, c, CV_8UC3);
    if (motion_color) motion_color->create(r, c, CV_8UC3);
    if (ratio_color) ratio_color->create(r, c, CV_8UC3);
    if (density && !(flags & RC_FIELDS_KEEP_DENSITY)) density->create(r, c, CV_32FC3);
    check(rc_particle_fields(rc::default_context(), f.ptr<float>(), d.ptr<float>(), c, r, flags,
                             streamfield ? streamfield->ptr<float>() : nullptr,
                             displacement_color ? displacement_color->data : nullptr, motion_color ? motion_color->data : nullptr,
                             ratio_color ? ratio_color->data : nullptr, density ? density->ptr<float>() : nullptr, maxes),
          "rc_particle_fields");
}

// ---- flow-derived diagnostics (ripcurrents_module.cpp:900-1138) -----------------------------------------------------
void subtructMeanMagnitude(cv::Mat& current)
{
    require(current.type() == CV_32FC2, "subtructMeanMagnitude: CV_32FC2 flow required");
    check(rc_subtract_mean_magnitude(rc::default_context(), current.ptr<float>(), current.step, current.cols, current.rows,
                                     RC_DIAG_SEQUENTIAL_SUM, nullptr), "rc_subtract_mean_magnitude");
}

void vectorToColor(cv::Mat& current, cv::Mat& outImg)
{
    require(current.type() == CV_32FC2 && outImg.type() == CV_8UC3 && outImg.rows == current.rows && outImg.cols == current.cols,
            "vectorToColor: CV_32FC2 flow and CV_8UC3 image of the same size required");
    check(rc_vector_to_color(rc::default_context(), current.ptr<float>(), current.step, current.cols, current.rows, outImg.data,
                             outImg.step, nullptr, 0), "rc_vector_to_color");
    check(rc_synchronize(rc::default_context()), "rc_synchronize");
}

void shearRateToColor(cv::Mat& current, cv::Mat& outImg)
{
    require(current.type() == CV_32FC2 && outImg.type() == CV_8UC3 && outImg.rows == current.rows && outImg.cols == current.cols,
            "shearRateToColor: CV_32FC2 flow and CV_8UC3 image of the same size required");
    check(rc_shear_rate_to_color(rc::default_context(), current.ptr<float>(), current.step, current.cols, current.rows,
                                 outImg.data, outImg.step, nullptr, 0), "rc_shear_rate_to_color");
    check(rc_synchronize(rc::default_context()), "rc_synchronize");
}

void subtructAverage(cv::Mat& current)
{
    require(current.type() == CV_32FC2, "subtructAverage: CV_32FC2 flow required");
    double mean[2];
    check(rc_subtract_mean(rc::default_context(), current.ptr<float>(), current.step, current.cols, current.rows, mean),
          "rc_subtract_mean");
}

// ---- Streakline.hpp --------------------------------------------------------------------------------------------------
Streakline::Streakline(Pixel2 pixel)
{
    generationPoint = pixel;
    vertices.push_back(pixel);
    numberOfVertices = 1;
    frameCount = 1;
}

void Streakline::drawLine() {}

void Streakline::runAll(std::vector<Streakline>& lines, const cv::Mat& flow)
{
    require(flow.type() == CV_32FC2, "Streakline: flow must be CV_32FC2");
    const int E = (int)lines.size();
    if (!E) return;
    size_t cap = 0;
    for (auto& s : lines) cap = s.vertices.size() + 1 > cap ? s.vertices.size() + 1 : cap;
    std::vector<float> v((size_t)E * cap * 2, 0.f), em((size_t)E * 2);
    std::vector<int32_t> cnt(E);
    for (int e = 0; e < E; e++) {
        em[2 * e] = lines[e].generationPoint.x; em[2 * e + 1] = lines[e].generationPoint.y;
        cnt[e] = (int32_t)lines[e].vertices.size();
        std::memcpy(&v[(size_t)e * cap * 2], (const void*)lines[e].vertices.data(), sizeof(float) * 2 * lines[e].vertices.size());
    }
    check(rc_streakline_step(rc::default_context(), flow.ptr<float>(), flow.step, flow.cols, flow.rows, em.data(), E, v.data(),
                             cnt.data(), (int)cap, 1.0f), "rc_streakline_step");
    for (int e = 0; e < E; e++) {
        lines[e].vertices.resize(cnt[e]);
        std::memcpy((void*)lines[e].vertices.data(), &v[(size_t)e * cap * 2], sizeof(float) * 2 * cnt[e]);
        lines[e].numberOfVertices = cnt[e];
        lines[e].frameCount++;
    }
}

void Streakline::runFlow(const cv::Mat& flow)
{
    std::vector<Streakline> one(1, *this);
    runAll(one, flow);
    *this = one[0];
}

void Streakline::runLK(cv::UMat u_prev, cv::UMat u_current, cv::Mat&)
{
    // dense flow between the two frames with the parameters of the reference's default call (ripcurrents.cpp:215)
    cv::Mat flow;
    rc::calcOpticalFlowFarneback(u_prev, u_current, flow, 0.5, 2, 3, 2, 15, 1.2, 0);
    runFlow(flow);
}
