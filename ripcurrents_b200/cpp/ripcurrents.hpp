// Header-compatible replacement for the hot-path subset of the reference's RipCurrents_main/ripcurrents.hpp
// (declarations at ripcurrents.hpp:22-25,37,39-40,48,50,52,58,77): same names, argument order and meaning, same
// in-place / aliasing behaviour, no error channel (failures abort with a message, like the reference's CV_Assert).
// Every function forwards to the C ABI of include/ripcurrents_b200.h on a process-wide context
// (rc::default_context()); nothing here computes on the CPU except the 50x37-entry threshold scans that
// create_histogram performs on the caller's own cumulative arrays.
//
// Additional, recommended entry points (namespace rc) batch what the reference does one pixel / one seed per call:
// main.cpp's forEach loops become one call (INTEGRATION.md shows the 3-line diffs).
#ifndef __RIPCURRENTS_HPP_INCLUDE__
#define __RIPCURRENTS_HPP_INCLUDE__

#include <vector>
#include "cv_compat.hpp"
#include "../../include/ripcurrents_b200.h"

#define XDIM 640  // Dimensions to resize to (kept for source compatibility; the functions take sizes from the Mats)
#define YDIM 480
#define HIST_BINS 50
#define HIST_DIRECTIONS 36
#define HIST_RESOLUTION 20
#define BUFFER_FRAME 300
#define GRID_COUNT 30

typedef cv::Point3_<uchar> Pixelc;
typedef cv::Point_<float> Pixel2;
typedef cv::Point3_<float> Pixel3;

namespace rc {
rc_ctx* default_context();                 // created on first use on device RC_B200_DEVICE (default 0)
void set_default_device(int device);       // before first use
// cv::calcOpticalFlowFarneback replacement (ripcurrents.cpp:215, main.cpp:264,609,742,961,1119,1481)
void calcOpticalFlowFarneback(const cv::Mat& prev, const cv::Mat& next, cv::Mat& flow, double pyr_scale, int levels,
                              int winsize, int iterations, int poly_n, double poly_sigma, int flags);
// ripcurrents.cpp:209-210: resize(frame, subframe, size, 0, 0, INTER_LINEAR) + cvtColor(subframe, gray, COLOR_BGR2GRAY);
// area = true: the PRIMING frame's resize(..., INTER_AREA) of ripcurrents.cpp:186-187 (downscaling only)
void ingest(const cv::Mat& frame_bgr, cv::Mat& gray, cv::Size size, bool area = false);
// ripcurrents.cpp:305-309: split + cartToPolar(deg) + merge -> CV_32FC3 (angle, mag, mag)
void flowToPolar(const cv::Mat& flow, cv::Mat& polar);
// whole-image forms of the per-pixel / per-seed loops
void streamline_field_all(cv::Mat& streamlines_mat, cv::Mat& streamlines_distance, const cv::Mat& flow, float dt,
                          int iterations, float UPPER);                                    // ripcurrents.cpp:229-231
void streamlines_all(Pixel2* pts, int n, const cv::Mat& flow, float dt, int iterations, float UPPER, int variant);
void window_average(std::vector<cv::Mat>& buffer, int& currentBuffer, const cv::Mat& flow, cv::Mat& average); // main.cpp:1143-1153
// ripcurrents.cpp:231-279 in one call: magnitude + the three normalised JET images + the position scatter (two passes
// over 12 B/px instead of the reference's twelve); any output may be nullptr
void particle_fields(const cv::Mat& streamlines_mat, const cv::Mat& streamlines_distance, cv::Mat* streamfield,
                     cv::Mat* displacement_color, cv::Mat* motion_color, cv::Mat* ratio_color, cv::Mat* density,
                     double maxes[3] = nullptr, int flags = 0);
}  // namespace rc

// ripcurrents_module.cpp:13-59 (main_old.cpp:373-386)
void streamline_displacement(cv::Mat& streamfield, cv::Mat& streamoverlay_color);
void streamline_total_motion(cv::Mat& streamlines_distance, cv::Mat& streamoverlay_color);
void streamline_ratio(cv::Mat& streamfield, cv::Mat& streamlines_distance, cv::Mat& streamoverlay_color);
void streamline_positions(cv::Mat& streamlines_mat, cv::Mat& streamline_density);
// ripcurrents_module.cpp:900-1138.  The reference's function-local statics (previous frame's maxima) live in the
// default context; they start at 0 as in the reference.
void subtructMeanMagnitude(cv::Mat& current);
void vectorToColor(cv::Mat& current, cv::Mat& outImg);
void shearRateToColor(cv::Mat& current, cv::Mat& outImg);

void streamline_field(Pixel2* pt, float* distancetraveled, int xoffset, int yoffset, cv::Mat flow, float dt,
                      int iterations, float UPPER, float prop_above_upper[HIST_DIRECTIONS]);
void streamline(Pixel2* pt, cv::Scalar color, cv::Mat flow, cv::Mat overlay, float dt, int iterations, float UPPER,
                float prop_above_upper[HIST_DIRECTIONS]);
void streamline_2(Pixel2* pt, cv::Scalar color, cv::Mat flow, cv::Mat overlay, float dt, int iterations, float UPPER,
                  float prop_above_upper[HIST_DIRECTIONS]);
void streamline_3(Pixel2* pt, cv::Scalar color, cv::Mat flow, cv::Mat overlay, float dt, int iterations, float UPPER,
                  float prop_above_upper[HIST_DIRECTIONS]);
void get_streamlines(cv::Mat& streamout, cv::Mat& streamoverlay_color, cv::Mat& streamoverlay, int streamlines,
                     Pixel2 streampt[], int framecount, int totalframes, cv::Mat& current, float UPPER,
                     float prop_above_upper[]);
void create_histogram(cv::Mat current, int hist[HIST_BINS], int& histsum, int hist2d[HIST_DIRECTIONS][HIST_BINS],
                      int histsum2d[HIST_DIRECTIONS], float& UPPER, float UPPER2d[HIST_DIRECTIONS],
                      float prop_above_upper[HIST_DIRECTIONS]);
// ripcurrents.hpp:48 / ripcurrents_module.cpp:386-400: the sliding-window update of the 300-frame flow average.  Exactly as
// in the reference `buffer` is a BY-VALUE vector, so the caller's slot is not rewritten; average_color, grid and
// max_displacement belong to the arrow / colour visualisation that follows in the reference (module:402-484) and are
// not touched.
void averageVector(std::vector<cv::Mat> buffer, cv::Mat& current, int update_ith_buffer, cv::Mat& average,
                   cv::Mat& average_color, double** grid, float max_displacement, float UPPER);
void create_flow(cv::Mat current, cv::Mat waterclass, cv::Mat accumulator2, float UPPER, float MID, float LOWER,
                 float UPPER2d[HIST_DIRECTIONS]);
void create_accumulationbuffer(cv::Mat& accumulator, cv::Mat accumulator2, cv::Mat& out, cv::Mat outmask, int framecount);
void create_edges(cv::Mat& outmask);
void get_delta(Pixel2* pt, int xoffset, int yoffset, cv::Mat flow, float dt, float UPPER);
void subtructAverage(cv::Mat& current);

#endif
