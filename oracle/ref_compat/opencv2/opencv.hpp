/*
 * oracle/ref_compat/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Stand-in for <opencv2/opencv.hpp> that lets the reference's OWN hot-path sources compile in an image without
 * OpenCV development files (oracle/ref_build.py puts this directory first on the include path and compiles
 * RipCurrents_main/pathlines.cpp and Streakline.cpp where they lie, plus function bodies sliced out of
 * ripcurrents_module.cpp, ripcurrents.cpp and main.cpp at build time).  Written from scratch; it restates the
 * documented semantics of the few OpenCV VALUE TYPES those bodies use -- nothing of the hot path's arithmetic
 * lives here except:
 *   - cv::Point_ operators (types.hpp): every operator returns saturate_cast<_Tp> of the per-component C++
 *     expression, i.e. Point*float rounds each product to fp32, Point*double multiplies in double and rounds once,
 *     Point/int is `a.x / b` in fp32;
 *   - `Mat / s` in an augmented assignment (mat.inl.hpp / matop.cpp): a MatExpr with alpha = 1./s that is
 *     materialised by convertTo(type, alpha, 0) -- fp32 data is scaled by (float)alpha -- and then added / subtracted;
 *   - cv::add on fp32 Mats, cv::split / cv::merge, and cv::cartToPolar(deg), the latter with the arithmetic that is
 *     pinned bit-exactly to cv2 4.13.0 in tests/test_oracle_aggregate.py.
 * Drawing / display / timing calls are no-ops.  calcOpticalFlowPyrLK is a HOOK (cv::ref_hooks::lk): the build moves
 * streakline vertices with the dense flow instead of sparse LK, so the test installs
 * the reference's own streamline() step there and the reference's life-cycle code runs around it unchanged.
 */
#ifndef RC_REF_COMPAT_OPENCV_HPP
#define RC_REF_COMPAT_OPENCV_HPP

#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;

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
#define CV_HSV2BGR 54
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)

namespace cv {

template <typename T> static inline T saturate_cast(float v) { return (T)v; }
template <typename T> static inline T saturate_cast(double v) { return (T)v; }
template <typename T> static inline T saturate_cast(int v) { return (T)v; }

template <typename _Tp> class Point_ {
public:
    _Tp x, y;
    Point_() : x(0), y(0) {}
    Point_(_Tp _x, _Tp _y) : x(_x), y(_y) {}
    template <typename _Tp2> operator Point_<_Tp2>() const { return Point_<_Tp2>(saturate_cast<_Tp2>(x), saturate_cast<_Tp2>(y)); }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

template <typename _Tp> static inline Point_<_Tp>& operator+=(Point_<_Tp>& a, const Point_<_Tp>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename _Tp> static inline Point_<_Tp>& operator*=(Point_<_Tp>& a, int b) { a.x = saturate_cast<_Tp>(a.x * b); a.y = saturate_cast<_Tp>(a.y * b); return a; }
template <typename _Tp> static inline Point_<_Tp>& operator/=(Point_<_Tp>& a, int b) { a.x = saturate_cast<_Tp>(a.x / b); a.y = saturate_cast<_Tp>(a.y / b); return a; }
template <typename _Tp> static inline Point_<_Tp>& operator/=(Point_<_Tp>& a, float b) { a.x = saturate_cast<_Tp>(a.x / b); a.y = saturate_cast<_Tp>(a.y / b); return a; }
template <typename _Tp> static inline Point_<_Tp>& operator/=(Point_<_Tp>& a, double b) { a.x = saturate_cast<_Tp>(a.x / b); a.y = saturate_cast<_Tp>(a.y / b); return a; }
template <typename _Tp> static inline Point_<_Tp> operator+(const Point_<_Tp>& a, const Point_<_Tp>& b) { return Point_<_Tp>(saturate_cast<_Tp>(a.x + b.x), saturate_cast<_Tp>(a.y + b.y)); }
template <typename _Tp> static inline Point_<_Tp> operator-(const Point_<_Tp>& a, const Point_<_Tp>& b) { return Point_<_Tp>(saturate_cast<_Tp>(a.x - b.x), saturate_cast<_Tp>(a.y - b.y)); }
template <typename _Tp> static inline Point_<_Tp> operator*(const Point_<_Tp>& a, int b) { return Point_<_Tp>(saturate_cast<_Tp>(a.x * b), saturate_cast<_Tp>(a.y * b)); }
template <typename _Tp> static inline Point_<_Tp> operator*(int a, const Point_<_Tp>& b) { return Point_<_Tp>(saturate_cast<_Tp>(b.x * a), saturate_cast<_Tp>(b.y * a)); }
template <typename _Tp> static inline Point_<_Tp> operator*(const Point_<_Tp>& a, float b) { return Point_<_Tp>(saturate_cast<_Tp>(a.x * b), saturate_cast<_Tp>(a.y * b)); }
template <typename _Tp> static inline Point_<_Tp> operator*(float a, const Point_<_Tp>& b) { return Point_<_Tp>(saturate_cast<_Tp>(b.x * a), saturate_cast<_Tp>(b.y * a)); }
template <typename _Tp> static inline Point_<_Tp> operator*(const Point_<_Tp>& a, double b) { return Point_<_Tp>(saturate_cast<_Tp>(a.x * b), saturate_cast<_Tp>(a.y * b)); }
template <typename _Tp> static inline Point_<_Tp> operator*(double a, const Point_<_Tp>& b) { return Point_<_Tp>(saturate_cast<_Tp>(b.x * a), saturate_cast<_Tp>(b.y * a)); }
template <typename _Tp> static inline Point_<_Tp> operator/(const Point_<_Tp>& a, int b) { Point_<_Tp> tmp(a); tmp /= b; return tmp; }
template <typename _Tp> static inline Point_<_Tp> operator/(const Point_<_Tp>& a, float b) { Point_<_Tp> tmp(a); tmp /= b; return tmp; }
template <typename _Tp> static inline Point_<_Tp> operator/(const Point_<_Tp>& a, double b) { Point_<_Tp> tmp(a); tmp /= b; return tmp; }

template <typename _Tp> class Point3_ {
public:
    _Tp x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(_Tp _x, _Tp _y, _Tp _z) : x(_x), y(_y), z(_z) {}
};

class Size {
public:
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

class Scalar {
public:
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};

class TermCriteria {
public:
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type, maxCount; double epsilon;
    TermCriteria(int t = 0, int n = 0, double e = 0) : type(t), maxCount(n), epsilon(e) {}
};

class Mat;
struct MatScaleExpr { const Mat* a; double alpha; };      // the only MatExpr the hot path forms: Mat / s

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* d, size_t st = 0) : rows(r), cols(c), data((uchar*)d), type_(type) { step = st ? st : (size_t)c * elemSize(); }
    void create(int r, int c, int type)
    {
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[step * (size_t)r + 64], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); memset(m.data, 0, m.step * (size_t)r); return m; }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    int type() const { return type_; }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return (size_t)CV_MAT_CN(type_) * (CV_MAT_DEPTH(type_) == CV_8U ? 1 : 4); }
    bool empty() const { return !data || !rows || !cols; }
    Size size() const { return Size(cols, rows); }
    template <typename T> T* ptr(int r = 0, int c = 0) { return reinterpret_cast<T*>(data + (size_t)r * step) + c; }
    template <typename T> const T* ptr(int r = 0, int c = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step) + c; }
    template <typename T> T& at(int r, int c) { return *ptr<T>(r, c); }
    // Mat::forEach: the functor sees every element once with position = {row, col}.  OpenCV runs rows under
    // parallel_for_; the hot path's functors only touch their own pixel, so sequential order is equivalent.
    template <typename T, typename F> void forEach(const F& f)
    {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) { const int pos[2] = {r, c}; f(*ptr<T>(r, c), pos); }
    }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    size_t nfloats_per_row() const { return (size_t)cols * CV_MAT_CN(type_); }

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

class UMat : public Mat {
public:
    UMat() {}
    UMat(const Mat& m) : Mat(m) {}
};

static inline MatScaleExpr operator/(const Mat& a, double s) { MatScaleExpr e; e.a = &a; e.alpha = 1. / s; return e; }

// MatOp::augAssign{Add,Subtract}: temp = convertTo(a, type, alpha, 0) (fp32 data scaled by (float)alpha, shift 0), m op= temp
static inline void rc_ref_aug(Mat& m, const MatScaleExpr& e, bool sub)
{
    const float a = (float)e.alpha, b = 0.f;
    for (int r = 0; r < m.rows; r++) {
        float* d = m.ptr<float>(r);
        const float* s = e.a->ptr<float>(r);
        const size_t n = m.nfloats_per_row();
        for (size_t i = 0; i < n; i++) {
            const float t = s[i] * a + b;
            d[i] = sub ? d[i] - t : d[i] + t;
        }
    }
}
static inline Mat& operator-=(Mat& m, const MatScaleExpr& e) { rc_ref_aug(m, e, true); return m; }
static inline Mat& operator+=(Mat& m, const MatScaleExpr& e) { rc_ref_aug(m, e, false); return m; }

// cv::add(src1, src2, dst) on fp32 Mats of equal geometry
static inline void add(const Mat& a, const Mat& b, Mat& dst)
{
    for (int r = 0; r < a.rows; r++) {
        const float* pa = a.ptr<float>(r); const float* pb = b.ptr<float>(r); float* pd = dst.ptr<float>(r);
        const size_t n = a.nfloats_per_row();
        for (size_t i = 0; i < n; i++) pd[i] = pa[i] + pb[i];
    }
}

static inline void split(const Mat& src, Mat* mv)
{
    const int cn = src.channels();
    for (int k = 0; k < cn; k++) mv[k].create(src.rows, src.cols, CV_32FC1);
    for (int r = 0; r < src.rows; r++) {
        const float* s = src.ptr<float>(r);
        for (int c = 0; c < src.cols; c++)
            for (int k = 0; k < cn; k++) mv[k].ptr<float>(r)[c] = s[c * cn + k];
    }
}

static inline void merge(const Mat* mv, size_t count, Mat& dst)
{
    Mat out(mv[0].rows, mv[0].cols, CV_MAKETYPE(CV_32F, (int)count));
    for (int r = 0; r < out.rows; r++) {
        float* d = out.ptr<float>(r);
        for (int c = 0; c < out.cols; c++)
            for (size_t k = 0; k < count; k++) d[c * count + k] = mv[k].ptr<float>(r)[c];
    }
    dst = out;
}

// cv::cartToPolar(x, y, magnitude, angle, angleInDegrees): arithmetic of cv2 4.13.0's default path (pinned bit-exactly in
// tests/test_oracle_aggregate.py): mag = sqrtf(fma(x,x,y*y)); angle = 7th-order odd polynomial, FMA Horner
static inline void cartToPolar(const Mat& x, const Mat& y, Mat& mag, Mat& ang, bool deg = false)
{
    const float scale = deg ? (float)(180.0 / 3.14159265358979323846) : 1.f;
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float q = deg ? 90.f : (float)(3.14159265358979323846 / 2);
    Mat m(x.rows, x.cols, CV_32FC1), a(x.rows, x.cols, CV_32FC1);
    for (int r = 0; r < x.rows; r++)
        for (int c = 0; c < x.cols; c++) {
            const float vx = x.ptr<float>(r)[c], vy = y.ptr<float>(r)[c];
            const float ax = fabsf(vx), ay = fabsf(vy);
            const float mx = ax > ay ? ax : ay, mn = ax < ay ? ax : ay;
            const float t = mn / (mx + (float)DBL_EPSILON), t2 = t * t;
            float v = fmaf(fmaf(fmaf(p7, t2, p5), t2, p3), t2, p1) * t;
            if (ax < ay) v = q - v;
            if (vx < 0) v = 2 * q - v;
            if (vy < 0) v = 4 * q - v;
            a.ptr<float>(r)[c] = v;
            m.ptr<float>(r)[c] = sqrtf(fmaf(vx, vx, vy * vy));
        }
    mag = m; ang = a;
}

// ---- display / drawing: no-ops ---------------------------------------------------------------------------------
template <typename P1, typename P2>
static inline void line(const Mat&, P1, P2, const Scalar&, int = 1, int = 8, int = 0) {}
template <typename P>
static inline void circle(const Mat&, P, int, const Scalar&, int = 1, int = 8, int = 0) {}
static inline void cvtColor(const Mat&, Mat&, int) {}
static inline void imshow(const std::string&, const Mat&) {}

// ---- sparse LK: hook (see the header comment) --------------------------------------------------------------------
namespace ref_hooks {
typedef std::function<void(const std::vector<Point2f>&, std::vector<Point2f>&)> LkFn;
inline LkFn& lk() { static LkFn f; return f; }
}  // namespace ref_hooks
static inline void calcOpticalFlowPyrLK(const UMat&, const UMat&, const std::vector<Point2f>& prevPts, std::vector<Point2f>& nextPts,
                                        std::vector<uchar>& status, std::vector<float>& err, Size = Size(21, 21), int = 3,
                                        TermCriteria = TermCriteria(), int = 0, double = 1e-4)
{
    nextPts = prevPts;
    status.assign(prevPts.size(), 1);
    err.assign(prevPts.size(), 0.f);
    if (ref_hooks::lk()) ref_hooks::lk()(prevPts, nextPts);
}

}  // namespace cv
#endif
