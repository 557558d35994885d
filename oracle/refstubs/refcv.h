// TEST INFRASTRUCTURE ONLY.  Stand-in for the slice of the OpenCV C++ API that the reference's src/ORBmatcher.cc, src/Frame.cc,
// src/MapPoint.cc and src/KeyFrame.cc use, so those files compile UNMODIFIED, in place, from /root/reference (OpenCV's C++
// headers and libraries are not in this image; only the cv2 wheel).  Nothing here is derived from OpenCV source: the types model
// the members the reference touches.  The float semantics of the matrix expressions the matchers evaluate are the ones pinned
// against cv2 4.13 by tests/test_oracle_matchers.py (gemm's small-matrix branch for A(3x3)*x(3x1)+c: binary32 products and sums,
// then one binary64 multiply-add; binary64 accumulation for transposed products, norm and dot; scaling by (float)alpha).
#pragma once
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>
#include <list>
#include <set>
#include <map>
#include <string>
#include <iostream>
#include <algorithm>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_BGR2HSV 40
#define CV_PI 3.1415926535897932384626433832795

static inline int cvRound(double v) { return (int)nearbyint(v); }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T _x, T _y, T _z) : x(_x), y(_y), z(_z) {}
};
typedef Point3_<float> Point3f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

enum NormTypes { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

class Mat;
struct MatExpr;

class Mat {
    std::shared_ptr<uchar> buf;
    int tp;
public:
    int rows, cols;
    uchar* data;
    size_t step;

    Mat() : tp(0), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : tp(0), rows(0), cols(0), data(nullptr), step(0) { create(r, c, type); }
    Mat(Size s, int type) : tp(0), rows(0), cols(0), data(nullptr), step(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* ext, size_t st = 0) : tp(type), rows(r), cols(c), data((uchar*)ext), step(0) { step = st ? st : (size_t)c * elemSize(); }
    Mat(const MatExpr& e);
    Mat& operator=(const MatExpr& e);

    static int depthSize(int type) { static const int s[7] = {1, 1, 2, 2, 4, 4, 8}; return s[type & 7]; }
    int type() const { return tp; }
    int depth() const { return tp & 7; }
    int channels() const { return (tp >> 3) + 1; }
    size_t elemSize() const { return (size_t)depthSize(tp) * channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }
    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && tp == type) return;
        tp = type; rows = r; cols = c; step = (size_t)c * elemSize();
        const size_t bytes = step * (size_t)r + 64;
        buf = std::shared_ptr<uchar>((uchar*)calloc(bytes, 1), free);
        data = buf.get();
    }
    void release() { buf.reset(); rows = cols = 0; data = nullptr; step = 0; }
    template <typename T> T& at(int i, int j) { return *(T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> const T& at(int i, int j) const { return *(const T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : cols == 1 ? at<T>(i, 0) : at<T>(i / cols, i % cols); }
    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    Mat row(int i) const { Mat m(*this); m.data = data + (size_t)i * step; m.rows = 1; return m; }
    Mat col(int j) const { Mat m(*this); m.data = data + (size_t)j * elemSize(); m.cols = 1; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * elemSize(); m.cols = b - a; return m; }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat clone() const { Mat m; copyTo(m); return m; }
    // like OpenCV's, copyTo writes THROUGH an existing destination of the same shape and type (other headers of that buffer see
    // the new values) and allocates otherwise; the rvalue form takes views such as Twc.rowRange(0,3).colRange(0,3)
    void copyTo(Mat& m) const {
        if (m.data == data && m.rows == rows && m.cols == cols) return;
        if (!(m.data && m.rows == rows && m.cols == cols && m.tp == tp)) { Mat out(rows, cols, tp); m = out; }
        for (int i = 0; i < rows; i++) memmove(m.ptr(i), ptr(i), (size_t)cols * elemSize());
    }
    void copyTo(Mat&& view) const { copyTo(view); }
    // element conversion with scale: only the combinations the reference reaches (8U -> 32F, 32F -> 32F); the product runs in
    // binary32 with (float)alpha, (float)beta like OpenCV's 32f scale conversion
    void convertTo(Mat& m, int rtype, double alpha = 1, double beta = 0) const {
        if (rtype < 0) rtype = depth();
        Mat out(rows, cols, CV_MAKETYPE(rtype & 7, channels()));
        const int n = cols * channels();
        const float a = (float)alpha, b = (float)beta;
        for (int i = 0; i < rows; i++) {
            if (depth() == CV_8U && (rtype & 7) == CV_32F) { const uchar* s = ptr(i); float* d = out.ptr<float>(i); for (int j = 0; j < n; j++) d[j] = alpha == 1 && beta == 0 ? (float)s[j] : (float)s[j] * a + b; }
            else if (depth() == CV_32F && (rtype & 7) == CV_32F) { const float* s = ptr<float>(i); float* d = out.ptr<float>(i); for (int j = 0; j < n; j++) d[j] = alpha == 1 && beta == 0 ? s[j] : s[j] * a + b; }
            else if (depth() == (rtype & 7)) memcpy(out.ptr(i), ptr(i), (size_t)cols * elemSize());
            else { fprintf(stderr, "refcv: convertTo %d -> %d is not modelled\n", depth(), rtype); abort(); }
        }
        m = out;
    }
    Mat reshape(int cn, int newRows = 0) const {
        Mat c = isContinuous() ? *this : clone();
        const size_t scalars = (size_t)c.rows * c.cols * c.channels();
        if (cn == 0) cn = c.channels();
        Mat m(c);
        m.tp = CV_MAKETYPE(c.depth(), cn);
        m.rows = newRows ? newRows : c.rows;
        m.cols = (int)(scalars / ((size_t)m.rows * cn));
        m.step = (size_t)m.cols * m.elemSize();
        return m;
    }
    double dot(const Mat& o) const {          // binary64 sum of products (3-vectors stay below OpenCV's vector width)
        double s = 0;
        for (int i = 0; i < rows; i++) {
            const float* a = ptr<float>(i); const float* b = o.ptr<float>(i);
            for (int j = 0; j < cols; j++) s += (double)a[j] * (double)b[j];
        }
        return s;
    }
    MatExpr t() const;
    static MatExpr zeros(int r, int c, int type);
    static MatExpr ones(int r, int c, int type);
    static MatExpr eye(int r, int c, int type);
};

// The few lazy forms OpenCV's MatExpr keeps and the reference relies on: alpha*A, alpha*A^T, alpha*op(A)*B + beta*C,
// alpha*A + beta*B, alpha*{zeros, ones, eye}.
struct MatExpr {
    enum Op { SCALE, TRANSPOSE, GEMM, ADD, INIT } op;
    Mat a, b, c;
    double alpha, beta;
    bool transA;
    int rows, cols, type, initKind;
    MatExpr(Op o) : op(o), alpha(1), beta(0), transA(false), rows(0), cols(0), type(0), initKind(0) {}

    static void fail(const char* what) { fprintf(stderr, "refcv: %s is not modelled\n", what); abort(); }

    Mat eval() const {
        switch (op) {
        case SCALE: {
            if (a.depth() != CV_32F) fail("scaling of a non-float matrix");
            Mat out(a.rows, a.cols, a.type());
            const float s = (float)alpha;
            for (int i = 0; i < a.rows; i++)
                for (int j = 0; j < a.cols; j++) out.at<float>(i, j) = alpha == 1 ? a.at<float>(i, j) : alpha == -1 ? -a.at<float>(i, j) : a.at<float>(i, j) * s;
            return out;
        }
        case TRANSPOSE: {
            if (a.depth() != CV_32F) fail("transpose of a non-float matrix");
            Mat out(a.cols, a.rows, a.type());
            const float s = (float)alpha;
            for (int i = 0; i < a.rows; i++)
                for (int j = 0; j < a.cols; j++) out.at<float>(j, i) = alpha == 1 ? a.at<float>(i, j) : a.at<float>(i, j) * s;
            return out;
        }
        case GEMM: {
            if (a.depth() != CV_32F || b.depth() != CV_32F) fail("product of non-float matrices");
            const int M = transA ? a.cols : a.rows, K = transA ? a.rows : a.cols, N = b.cols;
            if (b.rows != K || (c.data && (c.rows != M || c.cols != N))) fail("product of mismatching shapes");
            Mat out(M, N, CV_32F);
            const bool small = !transA && K >= 2 && K <= 4 && (K == N || K == M);     // gemm's unrolled branch (flags == 0)
            for (int i = 0; i < M; i++)
                for (int j = 0; j < N; j++) {
                    const double cij = c.data ? (double)c.at<float>(i, j) * beta : 0.0;
                    if (small) {
                        float t = a.at<float>(i, 0) * b.at<float>(0, j);
                        for (int k = 1; k < K; k++) t = t + a.at<float>(i, k) * b.at<float>(k, j);
                        out.at<float>(i, j) = (float)((double)t * alpha + cij);
                    } else {
                        double s = 0;
                        for (int k = 0; k < K; k++) s += (double)(transA ? a.at<float>(k, i) : a.at<float>(i, k)) * (double)b.at<float>(k, j);
                        out.at<float>(i, j) = (float)(s * alpha + cij);
                    }
                }
            return out;
        }
        case ADD: {
            if (a.depth() != CV_32F || b.depth() != CV_32F || a.rows != b.rows || a.cols != b.cols) fail("sum of mismatching matrices");
            Mat out(a.rows, a.cols, CV_32F);
            const float fa = (float)alpha, fb = (float)beta;
            for (int i = 0; i < a.rows; i++)
                for (int j = 0; j < a.cols; j++) {
                    const float x = a.at<float>(i, j), y = b.at<float>(i, j);
                    out.at<float>(i, j) = (alpha == 1 && beta == 1) ? x + y : (alpha == 1 && beta == -1) ? x - y : x * fa + y * fb;
                }
            return out;
        }
        case INIT: {
            Mat out(rows, cols, type);
            if ((type & 7) != CV_32F) { if (initKind != 0) fail("ones / eye of a non-float type"); return out; }
            for (int i = 0; i < rows; i++)
                for (int j = 0; j < cols; j++) out.at<float>(i, j) = initKind == 1 ? (float)alpha : (initKind == 2 && i == j) ? (float)alpha : 0.f;
            return out;
        }
        }
        return Mat();
    }
    operator Mat() const { return eval(); }
    double dot(const Mat& m) const { return eval().dot(m); }
    template <typename T> T at(int i, int j) const { return eval().at<T>(i, j); }
    Mat row(int i) const { return eval().row(i); }
    Mat col(int j) const { return eval().col(j); }
    Mat rowRange(int x, int y) const { return eval().rowRange(x, y); }
    Mat colRange(int x, int y) const { return eval().colRange(x, y); }
    Mat clone() const { return eval(); }
    MatExpr t() const { if (op == SCALE) { MatExpr e(TRANSPOSE); e.a = a; e.alpha = alpha; return e; } MatExpr e(TRANSPOSE); e.a = eval(); return e; }
};

inline Mat::Mat(const MatExpr& e) : tp(0), rows(0), cols(0), data(nullptr), step(0) { *this = e.eval(); }
inline Mat& Mat::operator=(const MatExpr& e) { *this = e.eval(); return *this; }
inline MatExpr Mat::t() const { MatExpr e(MatExpr::TRANSPOSE); e.a = *this; return e; }
inline MatExpr Mat::zeros(int r, int c, int type) { MatExpr e(MatExpr::INIT); e.rows = r; e.cols = c; e.type = type; e.initKind = 0; return e; }
inline MatExpr Mat::ones(int r, int c, int type) { MatExpr e(MatExpr::INIT); e.rows = r; e.cols = c; e.type = type; e.initKind = 1; return e; }
inline MatExpr Mat::eye(int r, int c, int type) { MatExpr e(MatExpr::INIT); e.rows = r; e.cols = c; e.type = type; e.initKind = 2; return e; }

inline MatExpr operator*(const Mat& a, const Mat& b) { MatExpr e(MatExpr::GEMM); e.a = a; e.b = b; return e; }
inline MatExpr operator*(const MatExpr& x, const Mat& b) {
    MatExpr e(MatExpr::GEMM);
    if (x.op == MatExpr::TRANSPOSE) { e.a = x.a; e.transA = true; e.alpha = x.alpha; }
    else if (x.op == MatExpr::SCALE) { e.a = x.a; e.alpha = x.alpha; }
    else e.a = x.eval();
    e.b = b;
    return e;
}
inline MatExpr operator*(const Mat& a, const MatExpr& y) { MatExpr e(MatExpr::GEMM); e.a = a; e.b = y.eval(); return e; }
inline MatExpr operator*(const MatExpr& x, const MatExpr& y) { return x * y.eval(); }
inline MatExpr operator*(double s, const Mat& a) { MatExpr e(MatExpr::SCALE); e.a = a; e.alpha = s; return e; }
inline MatExpr operator*(const Mat& a, double s) { return s * a; }
inline MatExpr operator*(double s, const MatExpr& x) {
    MatExpr e = x;
    if (x.op == MatExpr::SCALE || x.op == MatExpr::TRANSPOSE || x.op == MatExpr::INIT) e.alpha *= s;
    else if (x.op == MatExpr::GEMM || x.op == MatExpr::ADD) { e.alpha *= s; e.beta *= s; }
    return e;
}
inline MatExpr operator*(const MatExpr& x, double s) { return s * x; }
inline MatExpr operator/(const Mat& a, double s) { MatExpr e(MatExpr::SCALE); e.a = a; e.alpha = 1.0 / s; return e; }
inline MatExpr operator/(const MatExpr& x, double s) { return (1.0 / s) * x; }
inline MatExpr operator-(const Mat& a) { MatExpr e(MatExpr::SCALE); e.a = a; e.alpha = -1; return e; }
inline MatExpr operator-(const MatExpr& x) { return -1.0 * x; }
inline MatExpr operator+(const Mat& a, const Mat& b) { MatExpr e(MatExpr::ADD); e.a = a; e.b = b; e.alpha = 1; e.beta = 1; return e; }
inline MatExpr operator-(const Mat& a, const Mat& b) { MatExpr e(MatExpr::ADD); e.a = a; e.b = b; e.alpha = 1; e.beta = -1; return e; }
inline MatExpr operator+(const MatExpr& x, const Mat& m) {
    if (x.op == MatExpr::GEMM && !x.c.data) { MatExpr e = x; e.c = m; e.beta = 1; return e; }
    if (x.op == MatExpr::SCALE) { MatExpr e(MatExpr::ADD); e.a = x.a; e.alpha = x.alpha; e.b = m; e.beta = 1; return e; }
    return x.eval() + m;
}
inline MatExpr operator+(const Mat& m, const MatExpr& x) {
    if (x.op == MatExpr::GEMM && !x.c.data) { MatExpr e = x; e.c = m; e.beta = 1; return e; }
    if (x.op == MatExpr::SCALE) { MatExpr e(MatExpr::ADD); e.a = m; e.alpha = 1; e.b = x.a; e.beta = x.alpha; return e; }
    return m + x.eval();
}
inline MatExpr operator-(const Mat& m, const MatExpr& x) {
    if (x.op == MatExpr::SCALE) { MatExpr e(MatExpr::ADD); e.a = m; e.alpha = 1; e.b = x.a; e.beta = -x.alpha; return e; }
    return m - x.eval();
}
inline MatExpr operator-(const MatExpr& x, const Mat& m) {
    if (x.op == MatExpr::GEMM && !x.c.data) { MatExpr e = x; e.c = m; e.beta = -1; return e; }
    return x.eval() - m;
}
inline MatExpr operator+(const MatExpr& x, const MatExpr& y) { return x + y.eval(); }
inline MatExpr operator-(const MatExpr& x, const MatExpr& y) { return x - y.eval(); }

// cv::norm: L2 of a float matrix = sqrt of the binary64 sum of squares; L1 of a difference = binary64 sum of |a - b| in binary32
inline double norm(const Mat& m, int normType = NORM_L2) {
    double s = 0;
    for (int i = 0; i < m.rows; i++) {
        const float* p = m.ptr<float>(i);
        for (int j = 0; j < m.cols; j++) { const double v = p[j]; s += normType == NORM_L1 ? std::fabs(v) : v * v; }
    }
    return normType == NORM_L1 ? s : std::sqrt(s);
}
inline double norm(const MatExpr& e, int normType = NORM_L2) { return norm(e.eval(), normType); }
inline double norm(const Mat& a, const Mat& b, int normType = NORM_L2) {
    double s = 0;
    for (int i = 0; i < a.rows; i++) {
        const float* p = a.ptr<float>(i); const float* q = b.ptr<float>(i);
        for (int j = 0; j < a.cols; j++) { const float d = p[j] - q[j]; s += normType == NORM_L1 ? (double)std::fabs(d) : (double)d * (double)d; }
    }
    return normType == NORM_L1 ? s : std::sqrt(s);
}

template <typename T> struct MatCommaInitializer_ {
    Mat m; int idx;
    MatCommaInitializer_(const Mat& mm, T v) : m(mm), idx(0) { put(v); }
    void put(T v) { m.at<T>(idx / m.cols, idx % m.cols) = v; idx++; }
    template <typename U> MatCommaInitializer_& operator,(U v) { put((T)v); return *this; }
    operator Mat() const { return m; }
};
template <typename T> struct Mat_ : public Mat {
    Mat_(int r, int c);
    Mat_(const Mat& m) : Mat(m) {}
    T& operator()(int i, int j) { return this->template at<T>(i, j); }
    const T& operator()(int i, int j) const { return this->template at<T>(i, j); }
    template <typename U> MatCommaInitializer_<T> operator<<(U v) const { return MatCommaInitializer_<T>(*this, (T)v); }
};
template <> inline Mat_<float>::Mat_(int r, int c) : Mat(r, c, CV_32F) {}
template <> inline Mat_<double>::Mat_(int r, int c) : Mat(r, c, CV_64F) {}
template <> inline Mat_<int>::Mat_(int r, int c) : Mat(r, c, CV_32S) {}
template <> inline Mat_<uchar>::Mat_(int r, int c) : Mat(r, c, CV_8U) {}

class _InputArray {
public:
    Mat m;
    _InputArray() {}
    _InputArray(const Mat& mm) : m(mm) {}
    _InputArray(const MatExpr& e) : m(e.eval()) {}
    bool empty() const { return m.empty(); }
    Mat getMat() const { return m; }
};
class _OutputArray {
public:
    Mat* m;
    _OutputArray(Mat& mm) : m(&mm) {}
    void create(int r, int c, int t) const { m->create(r, c, t); }
    void release() const { m->release(); }
    Mat getMat() const { return *m; }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// Declared only: image-level OpenCV calls of the object layer and of UndistortKeyPoints, which sit in the compiled files but
// not on the compared path (oracle/ref_matcher_harness.cpp never reaches them; the link step gives each an aborting body).
void cvtColor(const Mat& src, Mat& dst, int code);
void calcHist(const Mat* images, int nimages, const int* channels, const Mat& mask, Mat& hist, int dims, const int* histSize,
              const float** ranges, bool uniform = true, bool accumulate = false);
void hconcat(const Mat& a, const Mat& b, Mat& dst);
void normalize(const Mat& src, Mat& dst, int normType);
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R, const Mat& P);

}  // namespace cv
