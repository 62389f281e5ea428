// Header-compatible replacement for RipCurrents_main/pathlines.h:6 (pathlines.cpp:9-46): Euler steps of
// dt/iterations with a bilinear gather of the flow, no speed cut-off.  `color` / `overlay` were only used for
// cv::line drawing and are ignored.
#ifndef __CV_PATHLINES_H
#define __CV_PATHLINES_H
#include "cv_compat.hpp"
typedef cv::Point_<float> Pixel2;
void streamline(Pixel2* pt, cv::Scalar color, cv::Mat flow, cv::Mat overlay, float dt, int iterations);
#endif
