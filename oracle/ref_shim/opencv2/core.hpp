// ref_shim/opencv2/core.hpp — TEST INFRASTRUCTURE (oracle/_ref build only; never part of the product).
// A minimal stand-in for the OpenCV headers so that the reference's own translation units
// (core/operators/objDetection/OP_FtDtOrbSlam.cpp, core/operators/objAssoc/OP_FtAssocOrbSlam.cpp,
// core/sensorData/observation/FeatureGrid.cpp, core/dataTypes/frame/Frame.cpp, ...) compile UNCHANGED from
// /root/reference in a container without OpenCV C++.  Containers only (Mat with reference-counted storage and ROI views,
// KeyPoint, Point_, Size, Rect); the pixel primitives (resize / FAST / GaussianBlur / fastAtan2) forward to the oracle's
// scalar routines, which tests/test_oracle_vs_cv2.py pins bit-exact against live cv2 4.13.
#pragma once
// (the real OpenCV headers pull these standard headers in; the reference relies on that)
#include <algorithm>
#include <cassert>
#include <climits>
#include <iostream>
#include <list>
#include <map>
#include <sstream>
#include <utility>
#include <string>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

typedef unsigned char uchar;

// the cv2-pinned primitives of oracle/orb_oracle.cpp (liborb_oracle.so)
extern "C" {
void orc_resize_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep);
void orc_gauss7_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
int orc_fast_u8(const uint8_t* img, int w, int h, size_t step, int t, float* xyr, int cap);
float orc_fast_atan2(float y, float x);
}

namespace cv {

inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }
inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <class T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;

template <class T> struct Size_ { T width, height; Size_() : width(0), height(0) {} Size_(T w, T h) : width(w), height(h) {} };
typedef Size_<int> Size;
template <class T> struct Rect_ { T x, y, width, height; Rect_() : x(0), y(0), width(0), height(0) {} Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {} };
typedef Rect_<int> Rect;
typedef Rect_<float> Rect2f;

struct KeyPoint {
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1)
        : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

template <class T> using Ptr = std::shared_ptr<T>;
class FeatureDetector;
class DescriptorMatcher;

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1 };

// Matrix container: reference-counted storage, views share it (cv::Mat semantics the reference relies on:
// mvImagePyramid[level] = temp(Rect(...)) keeps temp's buffer alive, OP_FtDtOrbSlam.cpp:943).  CV_8UC1 for the ORB path;
// CV_32F (with the small dense algebra of core_algebra.hpp) for core/operators/mapInit/OP_2ViewReconstruction.cpp.
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
class MatExpr;
class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); fill(s.v[0]); }
    Mat(const Mat&) = default;
    Mat& operator=(const Mat&) = default;
    inline Mat& operator=(const MatExpr& e);      // writes INTO a view of matching size (A.row(0) = ...), else takes the result
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 64](), std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : type_ == CV_64F ? 8 : 1; }
    size_t step1() const { return step / elemSize(); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
    Mat view(int y, int x, int h, int w) const { Mat m; m.buf_ = buf_; m.type_ = type_; m.data = data + (size_t)y * step + (size_t)x * elemSize(); m.rows = h; m.cols = w; m.step = step; return m; }
    Mat operator()(const Rect& r) const { return view(r.y, r.x, r.height, r.width); }
    Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
    Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
    Mat row(int i) const { return view(i, 0, 1, cols); }
    Mat col(int i) const { return view(0, i, rows, 1); }
    Mat reshape(int /*cn = 0: unchanged*/, int r) const {      // continuous data only (vt.row(8).reshape(0, 3))
        assert(isContinuous() && r > 0 && total() % (size_t)r == 0);
        Mat m = *this; m.rows = r; m.cols = (int)(total() / (size_t)r); m.step = (size_t)m.cols * elemSize(); return m;
    }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.create(rows, cols, type_);
        copy_rows(m);
        return m;
    }
    void copyTo(Mat& m) const { m.create(rows, cols, type_); copy_rows(m); }
    void copyTo(Mat&& m) const { copy_rows(m); }      // a view of matching size (descriptors.row(i))
    template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }      // vectors: w.at<float>(2)
    template <class T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    static Mat zeros(int r, int c, int type) { Mat m; m.create(r, c, type); for (int y = 0; y < r; ++y) memset(m.ptr(y), 0, (size_t)c * m.elemSize()); return m; }
    // CV_32F algebra (core_algebra.hpp)
    static inline Mat eye(int r, int c, int type);
    static inline Mat diag(const Mat& d);
    inline MatExpr t() const;
    inline MatExpr inv() const;
    inline double dot(const Mat& m) const;
    void fill(double v) {
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) {
            if (type_ == CV_32F) at<float>(y, x) = (float)v; else if (type_ == CV_64F) at<double>(y, x) = v; else at<uchar>(y, x) = (uchar)v;
        }
    }
private:
    void copy_rows(Mat& m) const { for (int y = 0; y < rows && y < m.rows; ++y) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)(cols < m.cols ? cols : m.cols) * elemSize()); }
    int type_ = CV_8UC1;
    std::shared_ptr<uchar> buf_;
};

inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    dst.create(sz.height, sz.width, CV_8UC1);
    orc_resize_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

inline int shim_reflect101(int v, int n) { if (n == 1) return 0; while (v < 0 || v >= n) v = v < 0 ? -v : 2 * n - 2 - v; return v; }

inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*BORDER_REFLECT_101 (+ISOLATED)*/) {
    dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    for (int y = 0; y < src.rows; ++y) {
        uchar* d = dst.data + (size_t)(y + top) * dst.step + left;
        const uchar* s = src.data + (size_t)y * src.step;
        if (d != s) memmove(d, s, (size_t)src.cols);
    }
    for (int y = 0; y < dst.rows; ++y) {
        const int sy = shim_reflect101(y - top, src.rows);
        uchar* d = dst.data + (size_t)y * dst.step;
        const uchar* s = dst.data + (size_t)(sy + top) * dst.step + left;      // interior row (already in place)
        for (int x = 0; x < dst.cols; ++x) {
            const bool inside = y >= top && y < top + src.rows && x >= left && x < left + src.cols;
            if (!inside) d[x] = s[shim_reflect101(x - left, src.cols)];
        }
    }
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size, double, double, int) {      // 7x7, sigma 2, BORDER_REFLECT_101 (the only call, :891)
    Mat tmp = src.clone();
    dst.create(src.rows, src.cols, CV_8UC1);
    orc_gauss7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}

inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool /*nonmaxSuppression = true*/) {
    kps.clear();
    if (img.rows < 7 || img.cols < 7) return;
    std::vector<float> xyr((size_t)img.rows * img.cols * 3);
    const int n = orc_fast_u8(img.data, img.cols, img.rows, img.step, threshold, xyr.data(), img.rows * img.cols);
    for (int i = 0; i < n; ++i) kps.emplace_back(xyr[3 * i], xyr[3 * i + 1], 7.f, -1.f, xyr[3 * i + 2]);
}

}  // namespace cv

#include "core_algebra.hpp"

using cv::cvRound;
using cv::cvFloor;
using cv::cvCeil;
