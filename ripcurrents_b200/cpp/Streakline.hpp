// Header-compatible replacement for RipCurrents_main/Streakline.hpp:8-20.  Same public members; runLK keeps its
// name and signature but moves the vertices with the Euler/bilinear step on the dense Farneback flow between the two
// frames instead of sparse pyramidal LK (BASELINE.json north_star), then applies the
// reference's life-cycle rules (Streakline.cpp:34-48): reject moves above 10 % of the frame, insert the generation
// point at the front every frame.  Drawing is left to the caller (outImg is not touched).
#ifndef __CV_STREAKLINE_H
#define __CV_STREAKLINE_H

#include <vector>
#include "cv_compat.hpp"

typedef cv::Point_<float> Pixel2;

class Streakline {
public:
    int numberOfVertices;
    Pixel2 generationPoint;
    std::vector<Pixel2> vertices;
    int frameCount;

    Streakline(Pixel2 pixel);
    void drawLine();                                                   // declared but never defined in the reference
    void runLK(cv::UMat u_prev, cv::UMat u_current, cv::Mat& outImg);  // dense-flow step between the two frames
    void runFlow(const cv::Mat& flow);                                 // same life-cycle on a flow already computed

    // all streaklines of a scene in ONE launch per frame (what compute_streaklines' inner loop, main.cpp:150-152, becomes)
    static void runAll(std::vector<Streakline>& lines, const cv::Mat& flow);
};
#endif
