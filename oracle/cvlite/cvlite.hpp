// cvlite.hpp -- TEST INFRASTRUCTURE ONLY (oracle). Not part of the shipped product path.
//
// An OpenCV-free restatement of the handful of OpenCV 4.x primitives that the reference's
// ORB front-end calls (OpenCV is an un-vendored third-party dependency of the reference:
// `find_package(OpenCV 4.5.1 QUIET)`, /root/reference/CMakeLists.txt:31).  The arithmetic of
// every primitive is the published OpenCV algorithm for 8-bit single-channel images
//  and is pinned against the in-container cv2 4.13 wheel by
// tests/test_oracle_cvlite.py and against the committed fixtures in tests/golden/.
//
// Call sites in the reference that these replace:
//   cv::resize          /root/reference/src/ORBextractor.cc:1848
//   cv::copyMakeBorder  /root/reference/src/ORBextractor.cc:1859,1880
//   cv::FAST            /root/reference/src/ORBextractor.cc:1126,1135
//   cv::GaussianBlur    /root/reference/src/ORBextractor.cc:1629,1793
//   cv::fastAtan2       /root/reference/src/ORBextractor.cc:160
//   cv::getStructuringElement / dilate / erode   /root/reference/src/ORBextractor.cc:1699-1704
//   cv::norm(NORM_L1), Mat::convertTo, small float Mat algebra   /root/reference/src/Frame.cc:1401-1445,
//                                                                 /root/reference/src/ORBmatcher.cc:1579-1606
//   cv::undistortPoints, Mat::reshape   /root/reference/src/Frame.cc:1087-1093, 1151-1153
//
// Two users: (1) oracle/shim/* exposes this as <opencv2/...> so the reference's own
// ORBextractor.cc compiles UNMODIFIED into oracle/_ref/; (2) oracle/port/* (our plain restatement).
#ifndef CVLITE_HPP
#define CVLITE_HPP

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <climits>
#include <cassert>
#include <memory>
#include <vector>
#include <algorithm>
#include <iostream>

namespace cv {

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_PI 3.1415926535897932384626433832795

// ---- type codes (depth | (cn-1)<<3).  The ORB path is single channel; the SLIC stage (src/cluster.cc) adds 3-channel 8U / 64F --------
enum { CV_8U = 0, CV_8S = 1, CV_16U = 2, CV_16S = 3, CV_32S = 4, CV_32F = 5, CV_64F = 6 };
#define CV_8UC1 0
#define CV_32FC1 5
#define CV_64FC1 6
#define CV_16UC1 2
#define CV_8UC3 (0 + (2 << 3))
#define CV_64FC3 (6 + (2 << 3))
static inline int cvl_channels(int type) { return ((type >> 3) & 63) + 1; }
static inline int cvl_elem_size1(int type) {
    switch (type & 7) { case CV_8U: case CV_8S: return 1; case CV_16U: case CV_16S: return 2;
                        case CV_32S: case CV_32F: return 4; default: return 8; }
}
static inline int cvl_elem_size(int type) { return cvl_elem_size1(type) * cvl_channels(type); }

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

// ---- rounding helpers: cvRound = round-half-to-even -------------------------
static inline int cvRound(double v) { return (int)lrint(v); }
static inline int cvRound(float v) { return (int)lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

template <typename T> static inline T saturate_cast(int v);
template <> inline uchar saturate_cast<uchar>(int v) { return (uchar)((unsigned)v <= 255 ? v : v > 0 ? 255 : 0); }
template <> inline short saturate_cast<short>(int v) { return (short)((unsigned)(v + 32768) <= 65535 ? v : v > 0 ? 32767 : -32768); }

// ---- small geometry types -------------------------------------------------------------------
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
};
template <typename T> static inline Point_<T> operator*(const Point_<T>& p, float s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
template <typename T> static inline Point_<T> operator*(const Point_<T>& p, double s) { return Point_<T>((T)(p.x * s), (T)(p.y * s)); }
template <typename T> static inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> static inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0]=a; val[1]=b; val[2]=c; val[3]=d; } };

// cv::KeyPoint: 28-byte POD {pt.x, pt.y, size, angle, response, octave, class_id}
struct KeyPoint {
    Point2f pt; float size; float angle; float response; int octave; int class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "KeyPoint layout must match cv::KeyPoint");
static_assert(sizeof(Point) == 8, "Point must be two ints (pattern table is reinterpreted)");

template <typename T> struct cvl_malloc_allocator {
    typedef T value_type;
    cvl_malloc_allocator() {}
    template <typename U> cvl_malloc_allocator(const cvl_malloc_allocator<U>&) {}
    T* allocate(size_t n) { return (T*)std::malloc(n * sizeof(T)); }
    void deallocate(T* p, size_t) { std::free(p); }
    template <typename U> bool operator==(const cvl_malloc_allocator<U>&) const { return true; }
    template <typename U> bool operator!=(const cvl_malloc_allocator<U>&) const { return false; }
};

// Lazy initialiser expression (cv::MatExpr for zeros/ones/eye).  Assigning it to an EXISTING Mat of the
// same size/type fills that Mat's storage in place instead of rebinding it -- the reference relies on
// this at ORBextractor.cc:1531 (`descriptors = Mat::zeros(...)` writes into rows of the output matrix).
struct MatExpr { int rows, cols, type; double fill; bool eye; };

// ---- Mat: ref-counted 2-D single-channel matrix with ROI views ------------------------------
class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;          // bytes per row
    uchar* datastart;     // start of the parent allocation
    uchar* dataend;
    std::shared_ptr<uchar> buf;
    int mtype;

    Mat() : rows(0), cols(0), data(nullptr), step(0), datastart(nullptr), dataend(nullptr), mtype(CV_8U) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : Mat() { create(r, c, type); for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) for (int k = 0; k < channels(); ++k) setd(y, x * channels() + k, s.val[k]); }
    Mat(int r, int c, int type, void* ext, size_t _step = 0) : rows(r), cols(c), data((uchar*)ext), mtype(type) {
        step = _step ? _step : (size_t)c * cvl_elem_size(type);
        datastart = data; dataend = data + step * r;
    }
    // ROI view
    Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), step(m.step), datastart(m.datastart),
                                         dataend(m.dataend), buf(m.buf), mtype(m.mtype) {
        assert(roi.x >= 0 && roi.y >= 0 && roi.x + roi.width <= m.cols && roi.y + roi.height <= m.rows);
        data = m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize();
    }
    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == mtype) return;
        rows = r; cols = c; mtype = type;
        step = (size_t)c * cvl_elem_size(type);
        size_t total = step * (size_t)r;
        // malloc, not operator new: keeps Mat storage out of the monotonic arena used by the
        // reference build for the octree tie-break (oracle/ref/arena.cpp).
        uchar* p = (uchar*)std::malloc(total ? total : 1);
        // control block via malloc too (3-arg ctor): Mats such as mvImagePyramid outlive an arena reset
        buf = std::shared_ptr<uchar>(p, [](uchar* q) { std::free(q); }, cvl_malloc_allocator<uchar>());
        data = datastart = p; dataend = p + total;
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf.reset(); data = datastart = dataend = nullptr; rows = cols = 0; step = 0; }
    int type() const { return mtype; }
    int depth() const { return mtype & 7; }
    int channels() const { return cvl_channels(mtype); }
    size_t elemSize() const { return (size_t)cvl_elem_size(mtype); }
    size_t elemSize1() const { return (size_t)cvl_elem_size1(mtype); }
    size_t step1() const { return step / elemSize(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    bool isSubmatrix() const { return data != datastart || (size_t)(dataend - datastart) != step * (size_t)rows || !isContinuous(); }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }

    template <typename T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step))[x]; }
    template <typename T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step))[x]; }
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat rowRange(int a, int b) const { return Mat(*this, Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return Mat(*this, Rect(a, 0, b - a, rows)); }
    Mat row(int y) const { return rowRange(y, y + 1); }
    Mat col(int x) const { return colRange(x, x + 1); }

    Mat clone() const {
        Mat m; if (!data) return m;
        m.create(rows, cols, mtype);
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(m.data + y * m.step, data + (size_t)y * step, rb);
        return m;
    }
    void copyTo(Mat& dst) const {
        dst.create(rows, cols, mtype);
        size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memmove(dst.data + y * dst.step, data + (size_t)y * step, rb);
    }
    void setTo(double v) {
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) setd(y, x, v);
    }
    static MatExpr zeros(int r, int c, int type) { MatExpr e = {r, c, type, 0.0, false}; return e; }
    static MatExpr zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static MatExpr ones(int r, int c, int type) { MatExpr e = {r, c, type, 1.0, false}; return e; }
    static MatExpr eye(int r, int c, int type) { MatExpr e = {r, c, type, 1.0, true}; return e; }
    Mat(const MatExpr& e) : Mat() { *this = e; }
    Mat& operator=(const MatExpr& e) {
        create(e.rows, e.cols, e.type);      // no-op (keeps storage, incl. ROI views) when size/type already match
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols * channels(); ++x) setd(y, x, e.eye ? (x == y ? e.fill : 0.0) : e.fill);
        return *this;
    }

    double getd(int y, int x) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(y, x); case CV_8S: return at<signed char>(y, x);
            case CV_16U: return at<ushort>(y, x); case CV_16S: return at<short>(y, x);
            case CV_32S: return at<int>(y, x); case CV_32F: return at<float>(y, x); default: return at<double>(y, x);
        }
    }
    void setd(int y, int x, double v) {
        switch (depth()) {
            case CV_8U: at<uchar>(y, x) = saturate_cast<uchar>(cvRound(v)); break;
            case CV_8S: at<signed char>(y, x) = (signed char)cvRound(v); break;
            case CV_16U: at<ushort>(y, x) = (ushort)cvRound(v); break;
            case CV_16S: at<short>(y, x) = (short)cvRound(v); break;
            case CV_32S: at<int>(y, x) = cvRound(v); break;
            case CV_32F: at<float>(y, x) = (float)v; break;
            default: at<double>(y, x) = v; break;
        }
    }
    // convertTo with alpha=1, beta=0.  In-place use (m.convertTo(m, T)) is supported.
    void convertTo(Mat& dst, int type) const {
        Mat out(rows, cols, type);
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) out.setd(y, x, getd(y, x));
        dst = out;
    }
    // reshape(cn): the shim has no channels; an N x 2 CV_32F matrix doubles as N two-channel points (undistortPoints below)
    Mat reshape(int /*cn*/, int /*rows*/ = 0) const { return *this; }
    // dot product of two equally sized CV_32F matrices, accumulated in double (cv::Mat::dot, dotProd_32f); the reference's uses on the
    // matcher path are a scale factor of a similarity (exactly 1 for the rigid poses the harness feeds) and the 60-degree viewing gate,
    // which the harness keeps away from its boundary
    double dot(const Mat& o) const {
        double s = 0;
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) s += (double)at<float>(y, x) * (double)o.at<float>(y, x);
        return s;
    }
    // transpose (float/double)
    Mat t() const {
        Mat m(cols, rows, mtype);
        for (int y = 0; y < rows; ++y) for (int x = 0; x < cols; ++x) m.setd(x, y, getd(y, x));
        return m;
    }
};

// ---- minimal float Mat algebra (CV_32F only; used by the matcher / stereo bodies) ------------
// Each op is evaluated element by element in float, in the same order OpenCV's scalar loops use.
static inline Mat operator-(const Mat& a, const Mat& b) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == CV_32F && b.type() == CV_32F);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = a.at<float>(y, x) - b.at<float>(y, x);
    return m;
}
static inline Mat operator+(const Mat& a, const Mat& b) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == CV_32F && b.type() == CV_32F);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = a.at<float>(y, x) + b.at<float>(y, x);
    return m;
}
static inline Mat operator*(double s, const Mat& a) {
    assert(a.type() == CV_32F);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = (float)(a.at<float>(y, x) * s);
    return m;
}
static inline Mat operator*(const Mat& a, double s) { return s * a; }
static inline Mat operator/(const Mat& a, double s) {
    assert(a.type() == CV_32F);
    Mat m(a.rows, a.cols, CV_32F);
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) m.at<float>(y, x) = (float)(a.at<float>(y, x) / s);
    return m;
}
static inline Mat operator-(const Mat& a) { return -1.0 * a; }
// matrix product of 32F matrices as cv::gemm evaluates it (modules/core/src/matmul.simd.hpp), pinned against cv2.gemm by
// tests/golden/cvlite_cv2.npz (gemm_*) -- the first version of this shim accumulated every product in double, which the real library
// only does on its general path:
//   * inner dimension 2..4 equal to the result's width or height (every pose product of the matcher bodies: 3x3 * 3x1, 3x3 * 3x3,
//     4x4 * 4x4): hand-unrolled FLOAT arithmetic, ((a0 b0 + a1 b1) + a2 b2) + a3 b3, each operation rounded to float;
//   * anything else (e.g. 1x3 * 3x1): GEMMSingleMul<float, double>, products and sum in double, one rounding to float.
// A following "+ C" is fused by cv::MatExpr into the same gemm call: on the small path that is t * alpha + c * beta in double of two
// floats = the float sum; on the general path the shim rounds twice where the library rounds once (no such call site in the bodies).
static inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.cols == b.rows && a.type() == CV_32F && b.type() == CV_32F);
    Mat m(a.rows, b.cols, CV_32F);
    const int len = a.cols;
    const bool small = len >= 2 && len <= 4 && (len == m.cols || len == m.rows);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < b.cols; ++j) {
        if (small) {
            float t = a.at<float>(i, 0) * b.at<float>(0, j);
            for (int k = 1; k < len; ++k) { const float p = a.at<float>(i, k) * b.at<float>(k, j); t = t + p; }
            m.at<float>(i, j) = t;
        } else {
            double s = 0; for (int k = 0; k < len; ++k) s += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
            m.at<float>(i, j) = (float)s;
        }
    }
    return m;
}
static inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(a.rows == b.rows && a.cols == b.cols && a.type() == b.type());
    double s = 0;
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) {
        double d = a.getd(y, x) - b.getd(y, x);
        if (normType == NORM_L1) s += std::fabs(d); else if (normType == NORM_L2) s += d * d; else s = std::max(s, std::fabs(d));
    }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}
static inline double norm(const Mat& a, int normType = NORM_L2) {
    double s = 0;
    for (int y = 0; y < a.rows; ++y) for (int x = 0; x < a.cols; ++x) {
        double d = a.getd(y, x);
        if (normType == NORM_L1) s += std::fabs(d); else if (normType == NORM_L2) s += d * d; else s = std::max(s, std::fabs(d));
    }
    return normType == NORM_L2 ? std::sqrt(s) : s;
}

// ---- InputArray / OutputArray proxies --------------------------------------------------------
class _InputArray {
public:
    const Mat* m;
    _InputArray() : m(nullptr) {}
    _InputArray(const Mat& _m) : m(&_m) {}
    bool empty() const { return !m || m->empty(); }
    Mat getMat() const { return m ? *m : Mat(); }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray() : m(nullptr) {}
    _OutputArray(Mat& _m) : m(&_m) {}
    void create(int r, int c, int type) const { if (m) m->create(r, c, type); }
    void create(Size s, int type) const { if (m) m->create(s.height, s.width, type); }
    void release() const { if (m) m->release(); }
    Mat getMat() const { return m ? *m : Mat(); }
    bool needed() const { return m != nullptr; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
static inline _InputArray noArray() { return _InputArray(); }

// =============================================================================================
// Primitive 1: resize(INTER_LINEAR) on 8UC1  
// Fixed point, INTER_RESIZE_COEF_BITS = 11.  Never reads outside `src`'s ROI.
// =============================================================================================
struct LinearCoefs { std::vector<int> ofs; std::vector<short> w0, w1; };
static inline void cvl_linear_coefs(int ssize, int dsize, LinearCoefs& c) {
    c.ofs.resize(dsize); c.w0.resize(dsize); c.w1.resize(dsize);
    double inv_scale = (double)dsize / ssize;
    double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cvFloor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        c.ofs[d] = s;
        c.w0[d] = saturate_cast<short>(cvRound((1.f - f) * 2048.f));
        c.w1[d] = saturate_cast<short>(cvRound(f * 2048.f));
    }
}
static inline void cvl_resize_linear_u8(const uchar* src, int sw, int sh, size_t sstep,
                                        uchar* dst, int dw, int dh, size_t dstep) {
    LinearCoefs cx, cy;
    cvl_linear_coefs(sw, dw, cx);
    cvl_linear_coefs(sh, dh, cy);
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        int sy0 = cy.ofs[y], sy1 = std::min(sy0 + 1, sh - 1);
        const uchar* s0 = src + (size_t)sy0 * sstep;
        const uchar* s1 = src + (size_t)sy1 * sstep;
        for (int x = 0; x < dw; ++x) {
            int sx0 = cx.ofs[x], sx1 = std::min(sx0 + 1, sw - 1);
            r0[x] = s0[sx0] * cx.w0[x] + s0[sx1] * cx.w1[x];
            r1[x] = s1[sx0] * cx.w0[x] + s1[sx1] * cx.w1[x];
        }
        int b0 = cy.w0[y], b1 = cy.w1[y];
        uchar* d = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; ++x)
            d[x] = (uchar)((((b0 * (r0[x] >> 4)) >> 16) + ((b1 * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}
static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && interpolation == INTER_LINEAR);
    if (dsize.width == 0 || dsize.height == 0) dsize = Size(cvRound(src.cols * fx), cvRound(src.rows * fy));
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    if (dsize.width == src.cols && dsize.height == src.rows) { src.copyTo(dst); return; }
    cvl_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// =============================================================================================
// Primitive 2: copyMakeBorder.  REFLECT_101: gfedcb|abcdefgh|gfedcba
// Supports src being an ROI of dst (the reference's in-place use, ORBextractor.cc:1859).
// =============================================================================================
static inline int borderInterpolate(int p, int len, int borderType) {
    borderType &= ~BORDER_ISOLATED;
    if ((unsigned)p < (unsigned)len) return p;
    if (borderType == BORDER_REPLICATE) return p < 0 ? 0 : len - 1;
    if (borderType == BORDER_REFLECT || borderType == BORDER_REFLECT_101) {
        int delta = borderType == BORDER_REFLECT_101;
        if (len == 1) return 0;
        do { if (p < 0) p = -p - 1 + delta; else p = len - 1 - (p - len) - delta; } while ((unsigned)p >= (unsigned)len);
        return p;
    }
    if (borderType == BORDER_WRAP) { if (p < 0) p -= ((p - len + 1) / len) * len; if (p >= len) p %= len; return p; }
    return -1;
}
static inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar()) {
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1);
    // NOTE: non-ISOLATED submatrix sources would read the parent in OpenCV; the reference only
    // passes whole images or uses BORDER_ISOLATED, so the ROI itself is always the source.
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    int bt = borderType & ~BORDER_ISOLATED;
    assert(bt == BORDER_REFLECT_101 || bt == BORDER_REPLICATE || bt == BORDER_REFLECT);
    // interior (memmove: src may alias dst's interior exactly)
    for (int y = 0; y < src.rows; ++y) std::memmove(dst.ptr(y + top) + left, src.ptr(y), (size_t)src.cols);
    for (int y = 0; y < src.rows; ++y) {
        uchar* row = dst.ptr(y + top);
        const uchar* in = row + left;
        for (int x = 0; x < left; ++x) row[x] = in[borderInterpolate(x - left, src.cols, bt)];
        for (int x = 0; x < right; ++x) row[left + src.cols + x] = in[borderInterpolate(src.cols + x, src.cols, bt)];
    }
    for (int y = 0; y < top; ++y) std::memcpy(dst.ptr(y), dst.ptr(top + borderInterpolate(y - top, src.rows, bt)), (size_t)dst.cols);
    for (int y = 0; y < bottom; ++y) std::memcpy(dst.ptr(top + src.rows + y), dst.ptr(top + borderInterpolate(src.rows + y, src.rows, bt)), (size_t)dst.cols);
}

// =============================================================================================
// Primitive 3: GaussianBlur 8U, fixed-point path.
// For ksize 7, sigma 2 the Q8 kernel is [18,34,48,56,48,34,18] (sum 256).
// =============================================================================================
// Q8 kernel as OpenCV's fixed-point getGaussianKernelFixedPoint_ED builds it for 8U images.
static inline std::vector<int> cvl_gauss_kernel_q8(int ksize, double sigma) {
    // float kernel (cv::getGaussianKernel): exp(-(i-c)^2/(2 sigma^2)) normalised to sum 1
    std::vector<double> k(ksize);
    double sigmaX = sigma > 0 ? sigma : ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8;
    double scale2X = -0.5 / (sigmaX * sigmaX), sum = 0;
    for (int i = 0; i < ksize; ++i) { double x = i - (ksize - 1) * 0.5; k[i] = std::exp(scale2X * x * x); sum += k[i]; }
    for (int i = 0; i < ksize; ++i) k[i] /= sum;
    // error-diffusion rounding to 8 fractional bits, from the centre outwards, symmetric
    const int fractionBits = 8; const int fractionMultiplier = 1 << fractionBits;
    std::vector<int> q(ksize);
    int n2 = ksize / 2;
    double err = 0;  // accumulated rounding error
    long long sumq = 0;
    // OpenCV: for i in [0, n2): v = k[i]*256 + err; q = round(v); err = v - q; mirrored; centre takes the remainder
    for (int i = 0; i < n2; ++i) {
        double v = k[i] * fractionMultiplier + err;
        long long qi = (long long)std::llround(v);
        err = v - (double)qi;
        q[i] = q[ksize - 1 - i] = (int)qi;
        sumq += 2 * qi;
    }
    q[n2] = (int)(fractionMultiplier - sumq);
    return q;
}
static inline void cvl_gauss_blur_u8(const uchar* src, int w, int h, size_t sstep, uchar* dst, size_t dstep,
                                     const std::vector<int>& q, int borderType) {
    // separable, exact integer: h = sum q*px (u16 range, no rounding); v = sum q*h (u32); out = (v + 2^15) >> 16
    const int ks = (int)q.size(), r = ks / 2;
    std::vector<ushort, cvl_malloc_allocator<ushort> > hbuf((size_t)w * h);
    std::vector<uchar, cvl_malloc_allocator<uchar> > prow((size_t)w + 2 * r);
    for (int y = 0; y < h; ++y) {
        const uchar* s = src + (size_t)y * sstep;
        for (int x = 0; x < r; ++x) { prow[x] = s[borderInterpolate(x - r, w, borderType)]; prow[r + w + x] = s[borderInterpolate(w + x, w, borderType)]; }
        std::memcpy(&prow[r], s, (size_t)w);
        ushort* hb = &hbuf[(size_t)y * w];
        const uchar* p = prow.data();
        if (ks == 7) {
            const unsigned q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
            for (int x = 0; x < w; ++x)
                hb[x] = (ushort)(q0 * (p[x] + p[x + 6]) + q1 * (p[x + 1] + p[x + 5]) + q2 * (p[x + 2] + p[x + 4]) + q3 * p[x + 3]);
        } else
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int k = 0; k < ks; ++k) acc += (unsigned)q[k] * p[x + k];
            hb[x] = (ushort)acc;   // <= 255*256 = 65280
        }
    }
    std::vector<const ushort*, cvl_malloc_allocator<const ushort*> > rows(ks);
    for (int y = 0; y < h; ++y) {
        for (int k = 0; k < ks; ++k) rows[k] = &hbuf[(size_t)borderInterpolate(y + k - r, h, borderType) * w];
        uchar* d = dst + (size_t)y * dstep;
        if (ks == 7) {
            const unsigned q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
            const ushort *r0 = rows[0], *r1 = rows[1], *r2 = rows[2], *r3 = rows[3], *r4 = rows[4], *r5 = rows[5], *r6 = rows[6];
            for (int x = 0; x < w; ++x) {
                unsigned acc = q0 * ((unsigned)r0[x] + r6[x]) + q1 * ((unsigned)r1[x] + r5[x]) + q2 * ((unsigned)r2[x] + r4[x]) + q3 * (unsigned)r3[x];
                d[x] = (uchar)((acc + 32768u) >> 16);
            }
        } else
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int k = 0; k < ks; ++k) acc += (unsigned)q[k] * rows[k][x];
            d[x] = (uchar)((acc + 32768u) >> 16);
        }
    }
}
static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && ksize.width == ksize.height && (ksize.width & 1));
    (void)sigmaY;
    Mat tmp = src.clone();   // allows in-place; reflection is at the ROI edge (non-submatrix clone in the reference)
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    std::vector<int> q = cvl_gauss_kernel_q8(ksize.width, sigmaX);
    cvl_gauss_blur_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step, q, borderType & ~BORDER_ISOLATED);
}

// =============================================================================================
// Primitive 4: FAST-9/16 with non-max suppression
// =============================================================================================
static const int cvl_fast_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int cvl_fast_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// S(p) = max( max_arcs min_{9 contiguous}(v - p_k), max_arcs min_{9 contiguous}(p_k - v) );
// corner at threshold t  <=>  S > t ;  cv::FAST response = S - 1.
static inline int cvl_fast_S(const uchar* p, size_t step) {
    int v = p[0], d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(ptrdiff_t)cvl_fast_dy[k] * (ptrdiff_t)step + cvl_fast_dx[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}
static inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    // Same control flow as OpenCV's scalar FAST_t<16>: table-driven early rejection on opposite
    // ring pixels, 9-contiguous run test, score only for corners, 3x3 strict NMS on the scores.
    Mat img = _img.getMat();
    assert(img.type() == CV_8UC1);
    keypoints.clear();
    threshold = std::min(std::max(threshold, 0), 255);
    const int W = img.cols, H = img.rows;
    if (W < 7 || H < 7) return;
    const ptrdiff_t step = (ptrdiff_t)img.step;
    ptrdiff_t pixel[25];
    for (int k = 0; k < 16; ++k) pixel[k] = cvl_fast_dy[k] * step + cvl_fast_dx[k];
    for (int k = 16; k < 25; ++k) pixel[k] = pixel[k - 16];
    uchar tab[512];
    for (int i = -255; i <= 255; ++i) tab[i + 255] = (uchar)(i < -threshold ? 1 : i > threshold ? 2 : 0);
    // persistent scratch: malloc-backed so it never lives in the per-frame arena of the _ref build
    static thread_local std::vector<uchar, cvl_malloc_allocator<uchar> > score;   // response (S-1) of corners, 0 elsewhere
    static thread_local std::vector<int, cvl_malloc_allocator<int> > cpos;
    score.assign((size_t)W * H, 0);
    cpos.clear();
    for (int y = 3; y < H - 3; ++y) {
        const uchar* ptr = img.ptr(y) + 3;
        for (int x = 3; x < W - 3; ++x, ++ptr) {
            int v = ptr[0];
            const uchar* t = &tab[0] - v + 255;
            int d = t[ptr[pixel[0]]] | t[ptr[pixel[8]]];
            if (d == 0) continue;
            d &= t[ptr[pixel[2]]] | t[ptr[pixel[10]]];
            d &= t[ptr[pixel[4]]] | t[ptr[pixel[12]]];
            d &= t[ptr[pixel[6]]] | t[ptr[pixel[14]]];
            if (d == 0) continue;
            d &= t[ptr[pixel[1]]] | t[ptr[pixel[9]]];
            d &= t[ptr[pixel[3]]] | t[ptr[pixel[11]]];
            d &= t[ptr[pixel[5]]] | t[ptr[pixel[13]]];
            d &= t[ptr[pixel[7]]] | t[ptr[pixel[15]]];
            bool corner = false;
            if (d & 1) {
                int vt = v - threshold, count = 0;
                for (int k = 0; k < 25; ++k) { if (ptr[pixel[k]] < vt) { if (++count > 8) { corner = true; break; } } else count = 0; }
            }
            if (!corner && (d & 2)) {
                int vt = v + threshold, count = 0;
                for (int k = 0; k < 25; ++k) { if (ptr[pixel[k]] > vt) { if (++count > 8) { corner = true; break; } } else count = 0; }
            }
            if (corner) {
                if (nonmaxSuppression) score[(size_t)y * W + x] = (uchar)(cvl_fast_S(ptr, (size_t)step) - 1);   // OpenCV scores only under NMS
                cpos.push_back(y * W + x);
            }
        }
    }
    for (size_t i = 0; i < cpos.size(); ++i) {
        int idx = cpos[i];
        const uchar* s = &score[(size_t)idx];
        int sc = s[0];
        if (nonmaxSuppression &&
            !(sc > s[-1] && sc > s[1] && sc > s[-W - 1] && sc > s[-W] && sc > s[-W + 1] &&
              sc > s[W - 1] && sc > s[W] && sc > s[W + 1])) continue;
        keypoints.push_back(KeyPoint((float)(idx % W), (float)(idx / W), 7.f, -1.f, (float)sc));
    }
}

// =============================================================================================
// Primitive 5: fastAtan2 (degrees), float polynomial, every op rounded to float, no FMA (A.5)
// =============================================================================================
#if defined(__GNUC__)
#define CVL_NOFMA __attribute__((optimize("fp-contract=off")))
#else
#define CVL_NOFMA
#endif
static inline CVL_NOFMA float fastAtan2(float y, float x) {
    static const float atan2_p1 = 0.9997878412794807f * (float)(180 / CV_PI);
    static const float atan2_p3 = -0.3258083974640975f * (float)(180 / CV_PI);
    static const float atan2_p5 = 0.1555786518463281f * (float)(180 / CV_PI);
    static const float atan2_p7 = -0.04432655554792128f * (float)(180 / CV_PI);
    volatile float ax = std::abs(x), ay = std::abs(y);
    volatile float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        volatile float t = atan2_p7 * c2; t = t + atan2_p5; t = t * c2; t = t + atan2_p3; t = t * c2; t = t + atan2_p1;
        a = t * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        volatile float t = atan2_p7 * c2; t = t + atan2_p5; t = t * c2; t = t + atan2_p3; t = t * c2; t = t + atan2_p1;
        a = 90.f - t * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// =============================================================================================
// Primitive 7: undistortPoints(src, dst, K, distCoeffs, R = noArray(), P = noArray())
// OpenCV's iterative inverse of the Brown-Conrady model (calib3d undistort, cvUndistortPointsInternal), restated from its
// published algorithm and pinned against cv2 4.13 (tests/golden/frame_cv2.npz):  everything in double; x = (u - cx) * (1/fx);
// default termination criteria = 5 iterations, no epsilon test;  icdist = (1 + ((k6 r2 + k5) r2 + k4) r2) / (1 + ((k3 r2 + k2) r2 + k1) r2);
// a negative icdist restarts from the normalised point and stops;  dX = 2 p1 x y + p2 (r2 + 2 x x) + s1 r2 + s2 r2 r2;  x = (x0 - dX) icdist;
// re-projection by RR = P * R (R empty = identity):  xx = RR00 x + RR01 y + RR02, ww = 1 / (RR20 x + RR21 y + RR22);  result cast to float.
// src / dst: N x 2 CV_32F (one point per row, see Mat::reshape); distCoeffs: 4 or 5 (k1 k2 p1 p2 [k3]) CV_32F or CV_64F, any orientation.
// =============================================================================================
static inline CVL_NOFMA void undistortPoints(InputArray _src, OutputArray _dst, InputArray _K, InputArray _D,
                                             InputArray _R = noArray(), InputArray _P = noArray()) {
    Mat src = _src.getMat(), Km = _K.getMat();
    assert(src.depth() == CV_32F && src.cols == 2 && Km.rows == 3 && Km.cols == 3);
    double k[14] = {0};
    if (!_D.empty()) {
        Mat D = _D.getMat();
        const int nd = D.rows * D.cols;
        assert(nd <= 14);
        for (int i = 0; i < nd; ++i) k[i] = D.rows == 1 ? D.getd(0, i) : D.getd(i, 0);
    }
    const double fx = Km.getd(0, 0), fy = Km.getd(1, 1), cx = Km.getd(0, 2), cy = Km.getd(1, 2);
    const double ifx = 1. / fx, ify = 1. / fy;
    double RR[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (!_R.empty()) { Mat R = _R.getMat(); for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) RR[i][j] = R.getd(i, j); }
    if (!_P.empty()) {
        Mat P = _P.getMat();
        double PP[3][3], T[3][3];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) PP[i][j] = P.getd(i, j);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { volatile double acc = 0; for (int q = 0; q < 3; ++q) acc = acc + PP[i][q] * RR[q][j]; T[i][j] = acc; }
        std::memcpy(RR, T, sizeof(RR));
    }
    const int n = src.rows;
    Mat out(n, 2, CV_32F);
    for (int i = 0; i < n; ++i) {
        volatile double x, y, x0, y0;
        const double u = src.at<float>(i, 0), v = src.at<float>(i, 1);
        x = (u - cx) * ifx; y = (v - cy) * ify;
        if (!_D.empty()) {
            x0 = x; y0 = y;
            for (int j = 0; j < 5; ++j) {
                volatile double r2 = x * x + y * y;
                volatile double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
                volatile double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                volatile double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        volatile double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2];
        volatile double yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
        volatile double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
        out.at<float>(i, 0) = (float)(xx * ww);
        out.at<float>(i, 1) = (float)(yy * ww);
    }
    if (_dst.m) *_dst.m = out;
}

// =============================================================================================
// Primitive 6: morphology with an elliptical structuring element (A.7)
// =============================================================================================
static inline Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1)) {
    if (anchor.x < 0) anchor.x = ksize.width / 2;
    if (anchor.y < 0) anchor.y = ksize.height / 2;
    Mat k = Mat::zeros(ksize.height, ksize.width, CV_8UC1);
    int r = 0, c = 0; double inv_r2 = 0;
    if (shape == MORPH_ELLIPSE) { r = ksize.height / 2; c = ksize.width / 2; inv_r2 = r ? 1. / ((double)r * r) : 0; }
    for (int i = 0; i < ksize.height; ++i) {
        int j1 = 0, j2 = 0;
        if (shape == MORPH_RECT || (shape == MORPH_CROSS && i == anchor.y)) j2 = ksize.width;
        else if (shape == MORPH_CROSS) { j1 = anchor.x; j2 = j1 + 1; }
        else {
            int dy = i - r;
            if (std::abs(dy) <= r) {
                int dx = cvRound(c * std::sqrt((r * r - dy * dy) * inv_r2));
                j1 = std::max(c - dx, 0); j2 = std::min(c + dx + 1, ksize.width);
            }
        }
        for (int j = j1; j < j2; ++j) k.at<uchar>(i, j) = 1;
    }
    return k;
}
static inline void cvl_morph(const Mat& src, Mat& dst, const Mat& kernel, bool isDilate) {
    assert(src.type() == CV_8UC1);
    Mat in = src.clone();
    dst.create(src.rows, src.cols, src.type());
    int ay = kernel.rows / 2, ax = kernel.cols / 2;
    // per-kernel-row horizontal extents (ellipse rows are contiguous runs)
    std::vector<int> j1(kernel.rows, 0), j2(kernel.rows, -1);
    for (int i = 0; i < kernel.rows; ++i) {
        int a = -1, b = -2;
        for (int j = 0; j < kernel.cols; ++j) if (kernel.at<uchar>(i, j)) { if (a < 0) a = j; b = j; }
        j1[i] = a; j2[i] = b;
    }
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            int acc = isDilate ? 0 : 255;   // default border value: ignored (min for dilate, max for erode)
            for (int i = 0; i < kernel.rows; ++i) {
                int yy = y + i - ay;
                if (yy < 0 || yy >= src.rows || j1[i] < 0) continue;
                int xa = std::max(0, x + j1[i] - ax), xb = std::min(src.cols - 1, x + j2[i] - ax);
                const uchar* row = in.ptr(yy);
                for (int xx = xa; xx <= xb; ++xx) acc = isDilate ? std::max(acc, (int)row[xx]) : std::min(acc, (int)row[xx]);
            }
            dst.at<uchar>(y, x) = (uchar)acc;
        }
}
static inline void dilate(InputArray src, OutputArray dst, InputArray kernel) { Mat s = src.getMat(), k = kernel.getMat(); dst.create(s.rows, s.cols, s.type()); Mat d = dst.getMat(); cvl_morph(s, d, k, true); }
static inline void erode(InputArray src, OutputArray dst, InputArray kernel) { Mat s = src.getMat(), k = kernel.getMat(); dst.create(s.rows, s.cols, s.type()); Mat d = dst.getMat(); cvl_morph(s, d, k, false); }

// ---- SLIC stage of the reference (src/cluster.cc:295-320): cvtColor(BGR2Lab) / Sobel(CV_64F, ksize 3) / addWeighted -----------------
// cvtColor(COLOR_BGR2Lab) on 8-bit data is OpenCV's trilinear 33^3 LUT built with its own softfloat: NOT restated.  It is the input
// boundary of the SLIC oracle: the harness computes the Lab image with the real OpenCV (cv2, fixtures under tests/golden/) and hands
// it in through this hook; a call without a hook aborts rather than inventing numbers.
enum { COLOR_BGR2Lab = 44 };
typedef void (*cvl_bgr2lab_hook_t)(const Mat& bgr, Mat& lab);
static inline cvl_bgr2lab_hook_t& cvl_bgr2lab_hook() { static thread_local cvl_bgr2lab_hook_t h = nullptr; return h; }
static inline void cvtColor(InputArray src, OutputArray dst, int code) {
    Mat s = src.getMat();
    if (code != COLOR_BGR2Lab || !cvl_bgr2lab_hook()) { std::cerr << "cvlite::cvtColor: only COLOR_BGR2Lab through the harness hook\n"; std::abort(); }
    dst.create(s.rows, s.cols, CV_8UC3);
    Mat d = dst.getMat();
    cvl_bgr2lab_hook()(s, d);
}
static inline int cvl_reflect101(int p, int len) { if (len == 1) return 0; while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p; return p; }
// cv::Sobel(src 8U, dst, CV_64F, dx, dy, 3): separable [-1 0 1] x [1 2 1], BORDER_REFLECT_101, every channel on its own; integers, exact in double
static inline void Sobel(InputArray src_, OutputArray dst_, int ddepth, int dx, int dy, int ksize = 3, double scale = 1, double delta = 0, int border = BORDER_DEFAULT) {
    Mat src = src_.getMat();
    assert(ddepth == CV_64F && ksize == 3 && src.depth() == CV_8U && scale == 1 && delta == 0 && border == BORDER_DEFAULT && dx + dy == 1);
    const int cn = src.channels();
    dst_.create(src.rows, src.cols, CV_64F + ((cn - 1) << 3));
    Mat dst = dst_.getMat();
    static const int D[3] = {-1, 0, 1}, Sm[3] = {1, 2, 1};
    const int* kx = dx ? D : Sm; const int* ky = dy ? D : Sm;
    for (int y = 0; y < src.rows; ++y) for (int x = 0; x < src.cols; ++x) for (int c = 0; c < cn; ++c) {
        int acc = 0;
        for (int i = -1; i <= 1; ++i) { const uchar* row = src.ptr(cvl_reflect101(y + i, src.rows));
            for (int j = -1; j <= 1; ++j) acc += ky[i + 1] * kx[j + 1] * (int)row[cvl_reflect101(x + j, src.cols) * cn + c]; }
        dst.ptr<double>(y)[x * cn + c] = (double)acc;
    }
}
// cv::addWeighted on CV_64F: dst = a*alpha + b*beta + gamma (OpenCV's scalar loop order)
static inline void addWeighted(InputArray a_, double alpha, InputArray b_, double beta, double gamma, OutputArray dst_) {
    Mat a = a_.getMat(), b = b_.getMat();
    assert(a.type() == b.type() && a.depth() == CV_64F && a.rows == b.rows && a.cols == b.cols);
    Mat out(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; ++y) { const double* pa = a.ptr<double>(y); const double* pb = b.ptr<double>(y); double* po = out.ptr<double>(y);
        for (int x = 0; x < n; ++x) po[x] = pa[x] * alpha + pb[x] * beta + gamma; }
    dst_.create(a.rows, a.cols, a.type());
    Mat d = dst_.getMat();
    out.copyTo(d);
}

// ---- KeyPointsFilter::retainBest: only referenced from dead code (ComputeKeyPointsOld) ------
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>& kps, int n) {
        if (n >= 0 && (int)kps.size() > n) {
            if (n == 0) { kps.clear(); return; }
            std::nth_element(kps.begin(), kps.begin() + n - 1, kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            float amb = kps[n - 1].response;
            auto it = std::partition(kps.begin() + n, kps.end(), [amb](const KeyPoint& k) { return k.response >= amb; });
            kps.resize(it - kps.begin());
        }
    }
};

}  // namespace cv

// OpenCV exports these at global scope too
using cv::cvRound; using cv::cvFloor; using cv::cvCeil; using cv::uchar; using cv::ushort;

#endif  // CVLITE_HPP
