// Minimal stand-in for the handful of OpenCV value types that the reference's hot-path headers use
// (cv::Mat, cv::UMat, cv::Point_, cv::Point3_, cv::Scalar).  Only compiled when real OpenCV headers are absent
// (this image ships no OpenCV C++ development files); with OpenCV installed the wrappers compile
// against <opencv2/core.hpp> unchanged.  Semantics kept: a cv::Mat passed BY VALUE aliases the caller's pixels
// (reference-counted header), exactly what create_flow() etc. rely on ("Ownership").
#ifndef RC_CV_COMPAT_HPP
#define RC_CV_COMPAT_HPP

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(RC_FORCE_CV_COMPAT)
#define RC_HAVE_OPENCV 1
#include <opencv2/core.hpp>
#endif
#endif

#ifndef RC_HAVE_OPENCV
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

namespace cv {

#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8U 0
#define CV_32F 5
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)

enum { OPTFLOW_USE_INITIAL_FLOW = 4, OPTFLOW_FARNEBACK_GAUSSIAN = 256 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;      // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    // wraps caller memory (no ownership), like cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int type, void* d, size_t st = 0) : rows(r), cols(c), data((uchar*)d), type_(type)
    {
        step = st ? st : (size_t)c * elemSize();
    }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[step * (size_t)r + 64], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data, 0, m.step * (size_t)r); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    int type() const { return type_; }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return (size_t)CV_MAT_CN(type_) * (CV_MAT_DEPTH(type_) == CV_8U ? 1 : 4); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }
    template <typename T> T* ptr(int r = 0, int c = 0) { return reinterpret_cast<T*>(data + (size_t)r * step) + c; }
    template <typename T> const T* ptr(int r = 0, int c = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step) + c; }
    template <typename T> T& at(int r, int c) { return *ptr<T>(r, c); }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    void setTo(int v) { for (int r = 0; r < rows; r++) std::memset(data + (size_t)r * step, v, (size_t)cols * elemSize()); }

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

// The reference passes cv::UMat (OpenCV's T-API) to the flow call; here it is a Mat whose pixels live on the host.
class UMat : public Mat {
public:
    UMat() {}
    UMat(const Mat& m) : Mat(m) {}
    Mat getMat(int /*access*/ = 0) const { return *this; }
};
enum { ACCESS_READ = 1 << 24 };

}  // namespace cv
#endif  // !RC_HAVE_OPENCV
#endif  // RC_CV_COMPAT_HPP
