// Drop-in demonstration: the frame loop of the reference's legacy detector (RipCurrents_main/ripcurrents.cpp:184-439)
// with video decoding / display removed, written against the reference's own entry points
// (calcOpticalFlowFarneback, streamline_field, cartToPolar+merge, create_histogram, create_flow,
// create_accumulationbuffer) as provided by ripcurrents.hpp of this repository.
//
//   demo_main frames.raw W H N out.bin [--time]
// --time: only the hot path of the legacy loop (flow, per-pixel particle field, polar, histogram, create_flow,
// create_accumulationbuffer), no per-frame record; prints {"pairs_per_s": ...} -- the throughput of the SOURCE-COMPATIBLE
// route, where every intermediate crosses the host in a cv::Mat exactly as in the reference (bench.py reports it next to the
// device-resident route)
// frames.raw: N frames of W*H u8.  out.bin: per processed frame { float UPPER; int histsum; float acc_sum;
// int mask_calm; float field_sum; streak checksum; byte sums of the three JET images; density sum;
// byte sums of vectorToColor / shearRateToColor; sum of the mean-magnitude-centred flow } --
// tests/test_gpu_cpp_dropin.py compares them with the CPU oracle.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ripcurrents.hpp"
#include "Streakline.hpp"

using namespace cv;

int main(int argc, char** argv)
{
    if (argc < 6) { std::fprintf(stderr, "usage: %s frames.raw W H N out.bin\n", argv[0]); return 2; }
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), N = std::atoi(argv[4]);
    const bool timing = argc > 6 && !std::strcmp(argv[6], "--time");
    const bool timing_fused = argc > 6 && !std::strcmp(argv[6], "--time-fused");
    std::vector<uchar> raw((size_t)W * H * N);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f || std::fread(raw.data(), 1, raw.size(), f) != raw.size()) { std::fprintf(stderr, "cannot read frames\n"); return 1; }
    std::fclose(f);
    FILE* out_f = std::fopen(argv[5], "wb");
    if (!out_f) return 1;

    Mat accumulator = Mat::zeros(Size(W, H), CV_32FC3);                                  // ripcurrents.cpp:133
    float LOWER = 0.2f, MID = .5f;                                                       // :142-143
    int hist[HIST_BINS] = {0}; int histsum = 0; float UPPER = 100.0f;                    // :147-149
    static int hist2d[HIST_DIRECTIONS][HIST_BINS]; int histsum2d[HIST_DIRECTIONS] = {0}; // :151-152
    float UPPER2d[HIST_DIRECTIONS] = {0}, prop_above_upper[HIST_DIRECTIONS] = {0};       // :153-154
    Mat streamlines_mat = Mat::zeros(H, W, CV_32FC2), streamlines_distance = Mat::zeros(H, W, CV_32FC1);   // :164-165
    Mat vector_color = Mat::zeros(Size(W, H), CV_8UC3), shear_color = Mat::zeros(Size(W, H), CV_8UC3);
    std::vector<Streakline> streaks;
    streaks.push_back(Streakline(Pixel2(W * 0.3f, H * 0.4f)));
    streaks.push_back(Streakline(Pixel2(W * 0.6f, H * 0.5f)));

    Mat f2(H, W, CV_8UC1, raw.data());                                                   // preloaded frame, :184-188
    if (timing_fused) {
        // the same loop with the block between video.read and imshow replaced by ONE C-ABI call per frame (INTEGRATION.md
        // section 3): host gray frame in, outmask + thresholds out, everything else stays on the device
        rc_ctx* ctx = rc::default_context();
        if (rc_flow_configure(ctx, W, H, 0.5, 2, 3, 2, 15, 1.2, 0) || rc_hist_reset(ctx)) return 1;
        Mat outmask = Mat::zeros(Size(W, H), CV_8UC1);
        rc_frame_result res;
        rc_process_frame(ctx, raw.data(), W, 28, nullptr, nullptr);
        rc_process_frame(ctx, raw.data() + (size_t)W * H, W, 29, outmask.data, &res);
        const auto t0 = std::chrono::steady_clock::now();
        for (int framecount = 2; framecount < N; framecount++)
            if (rc_process_frame(ctx, raw.data() + (size_t)framecount * W * H, W, framecount + 28, outmask.data, &res) != 1) return 1;
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"pairs_per_s\": %.2f, \"pairs\": %d, \"seconds\": %.4f, \"UPPER\": %g}\n", (N - 2) / sec, N - 2, sec, res.UPPER);
        std::fclose(out_f);
        return 0;
    }
    if (timing) {
        Mat current, polar;
        auto body = [&](int framecount) {
            Mat f1(H, W, CV_8UC1, raw.data() + (size_t)framecount * W * H);
            rc::calcOpticalFlowFarneback(f2, f1, current, 0.5, 2, 3, 2, 15, 1.2, 0);
            f2 = f1;
            rc::streamline_field_all(streamlines_mat, streamlines_distance, current, 2, 1, UPPER);
            rc::flowToPolar(current, polar);
            create_histogram(polar, hist, histsum, hist2d, histsum2d, UPPER, UPPER2d, prop_above_upper);
            Mat accumulator2 = Mat::zeros(Size(W, H), CV_32FC3), waterclass = Mat::zeros(Size(W, H), CV_32FC3);
            create_flow(polar, waterclass, accumulator2, UPPER, MID, LOWER, UPPER2d);
            Mat out = Mat::zeros(Size(W, H), CV_32FC3), outmask = Mat::zeros(Size(W, H), CV_8UC1);
            create_accumulationbuffer(accumulator, accumulator2, out, outmask, framecount + 28);
        };
        body(1);                                                                          // warm-up (context, allocations)
        const auto t0 = std::chrono::steady_clock::now();
        for (int framecount = 2; framecount < N; framecount++) body(framecount);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::printf("{\"pairs_per_s\": %.2f, \"pairs\": %d, \"seconds\": %.4f}\n", (N - 2) / sec, N - 2, sec);
        std::fclose(out_f);
        return 0;
    }
    for (int framecount = 1; framecount < N; framecount++) {                             // :194
        Mat f1(H, W, CV_8UC1, raw.data() + (size_t)framecount * W * H);
        Mat current;
        rc::calcOpticalFlowFarneback(f2, f1, current, 0.5, 2, 3, 2, 15, 1.2, 0);         // :215
        f2 = f1;                                                                          // :216

        rc::streamline_field_all(streamlines_mat, streamlines_distance, current, 2, 1, UPPER);   // :229-231
        Streakline::runAll(streaks, current);                                             // main.cpp:150-152

        Mat streamfield, disp_color, motion_color, ratio_color;                           // :231-258 / main_old.cpp:373-386
        rc::particle_fields(streamlines_mat, streamlines_distance, &streamfield, nullptr, nullptr, nullptr, nullptr);
        streamline_displacement(streamfield, disp_color);
        streamline_total_motion(streamlines_distance, motion_color);
        streamline_ratio(streamfield, streamlines_distance, ratio_color);
        Mat streamline_density = Mat::zeros(Size(W, H), CV_32FC3);                        // :261
        streamline_positions(streamlines_mat, streamline_density);

        vectorToColor(current, vector_color);                                             // main.cpp:629,761,1156 (module:1017)
        shearRateToColor(current, shear_color);                                           // main.cpp:1518 (module:1059)
        Mat centred = current.clone();
        subtructMeanMagnitude(centred);                                                   // module:900 (main.cpp:1130)

        Mat polar;
        rc::flowToPolar(current, polar);                                                  // :305-309
        create_histogram(polar, hist, histsum, hist2d, histsum2d, UPPER, UPPER2d, prop_above_upper);   // :319-366
        Mat accumulator2 = Mat::zeros(Size(W, H), CV_32FC3), waterclass = Mat::zeros(Size(W, H), CV_32FC3);   // :371-372
        create_flow(polar, waterclass, accumulator2, UPPER, MID, LOWER, UPPER2d);         // :376-402
        Mat out = Mat::zeros(Size(W, H), CV_32FC3), outmask = Mat::zeros(Size(W, H), CV_8UC1);   // :419-420
        create_accumulationbuffer(accumulator, accumulator2, out, outmask, framecount + 28);    // :414-439 (offset: crosses 30)

        double acc_sum = 0, field_sum = 0, csum[3] = {0, 0, 0}, dsum = 0, vsum = 0, ssum = 0, msum = 0; int calm = 0;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                acc_sum += accumulator.ptr<Pixel3>(y)[x].x;
                calm += outmask.ptr<uchar>(y)[x] == 255;
                field_sum += streamlines_distance.ptr<float>(y)[x];
                for (int k = 0; k < 3; k++) {
                    csum[0] += disp_color.ptr<uchar>(y)[3 * x + k] * (k + 1);
                    csum[1] += motion_color.ptr<uchar>(y)[3 * x + k] * (k + 1);
                    csum[2] += ratio_color.ptr<uchar>(y)[3 * x + k] * (k + 1);
                }
                dsum += streamline_density.ptr<float>(y)[3 * x + 2];
                for (int k = 0; k < 3; k++) {
                    vsum += vector_color.ptr<uchar>(y)[3 * x + k] * (k + 1);
                    ssum += shear_color.ptr<uchar>(y)[3 * x + k] * (k + 1);
                }
                msum += centred.ptr<float>(y)[2 * x] + 2.0 * centred.ptr<float>(y)[2 * x + 1];
            }
        float rec[13] = {UPPER, (float)histsum, (float)acc_sum, (float)calm, (float)field_sum,
                         streaks[0].vertices.back().x + streaks[1].vertices.back().y,
                         (float)csum[0], (float)csum[1], (float)csum[2], (float)dsum,
                         (float)vsum, (float)ssum, (float)msum};
        std::fwrite(rec, sizeof rec, 1, out_f);
    }
    std::fclose(out_f);
    return 0;
}
