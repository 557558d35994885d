// TEST INFRASTRUCTURE ONLY.  A minimal stand-in for the slice of the OpenCV C++ API that the
// reference's src/ORBextractor.cc uses, so that file can be compiled UNMODIFIED, in place,
// from /root/reference (OpenCV's C++ headers/libs are not in this image; only the cv2 wheel).
// The five image primitives forward to oracle/orc_primitives.h, which is pinned bit-for-bit
// to cv2 4.13.0 by tests/test_oracle_primitives.py.  Nothing here is derived from OpenCV
// source: the types model only the members the reference touches.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <list>
#include <iterator>
#include <algorithm>
#include "../orc_primitives.h"

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

static inline int cvRound(double v) { return orc::cv_round(v); }
static inline int cvRound(float v) { return orc::cv_round(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { return orc::cv_floor(v); }
static inline int cvCeil(double v) { return orc::cv_ceil(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

struct MatZeros { int rows, cols, type; };

// 8-bit single-channel matrix header with shared, malloc-backed storage (malloc, not operator
// new, so pixel buffers stay outside the bump arena that orders the quadtree's node addresses).
class Mat {
    struct Block { int refs; };
    Block* blk;
    void retain() { if (blk) __atomic_add_fetch(&blk->refs, 1, __ATOMIC_RELAXED); }
    void drop() {
        if (blk && __atomic_sub_fetch(&blk->refs, 1, __ATOMIC_ACQ_REL) == 0) free(blk);
        blk = nullptr;
    }
public:
    int rows, cols;
    uchar* data;
    size_t step;
    Mat() : blk(nullptr), rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int /*type*/) : blk(nullptr), rows(0), cols(0), data(nullptr), step(0) { create(r, c, 0); }
    Mat(Size s, int /*type*/) : blk(nullptr), rows(0), cols(0), data(nullptr), step(0) { create(s.height, s.width, 0); }
    Mat(int r, int c, int /*type*/, void* ext, size_t st) : blk(nullptr), rows(r), cols(c), data((uchar*)ext), step(st) {}
    Mat(const Mat& m) : blk(m.blk), rows(m.rows), cols(m.cols), data(m.data), step(m.step) { retain(); }
    Mat& operator=(const Mat& m) {
        if (this != &m) { drop(); blk = m.blk; rows = m.rows; cols = m.cols; data = m.data; step = m.step; retain(); }
        return *this;
    }
    // `descriptors = Mat::zeros(n,32,CV_8UC1)` must write through an existing view of the same
    // shape (ORBextractor.cc:1037 assigns into a rowRange of the output matrix).
    Mat& operator=(const MatZeros& z) {
        if (!(data && rows == z.rows && cols == z.cols)) create(z.rows, z.cols, z.type);
        for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, cols);
        return *this;
    }
    ~Mat() { drop(); }
    void create(int r, int c, int /*type*/) {
        if (data && rows == r && cols == c) return;
        drop();
        rows = r; cols = c; step = (size_t)c;
        size_t bytes = sizeof(Block) + 64 + (size_t)r * c + 64;
        blk = (Block*)malloc(bytes);
        blk->refs = 1;
        data = (uchar*)blk + sizeof(Block) + (64 - sizeof(Block) % 64);
    }
    void release() { drop(); rows = cols = 0; data = nullptr; step = 0; }
    static MatZeros zeros(int r, int c, int t) { return MatZeros{r, c, t}; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    size_t step1() const { return step; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const {
        Mat m(rows, cols, 0);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols);
        return m;
    }
};

class _InputArray {
public:
    const Mat* m;
    _InputArray(const Mat& mm) : m(&mm) {}
    bool empty() const { return m->empty(); }
    Mat getMat() const { return *m; }
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

enum { INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };

static inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    dst.create(sz.height, sz.width, 0);
    orc::resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// Writes `src` plus a reflect-101 border into `dst` (already sized by the caller here).
static inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int) {
    dst.create(src.rows + top + bottom, src.cols + left + right, 0);
    for (int y = 0; y < dst.rows; y++) {
        int sy = orc::reflect101(y - top, src.rows);
        const uchar* S = src.data + (size_t)sy * src.step;
        uchar* D = dst.data + (size_t)y * dst.step;
        if (D + left != S) memmove(D + left, S, src.cols);     // src may be the interior view of dst
    }
    for (int y = 0; y < dst.rows; y++) {
        uchar* D = dst.data + (size_t)y * dst.step;
        // interior rows of dst already hold src; mirrored rows were copied from src rows that
        // lie inside dst as well when src aliases dst, which is fine because they are unchanged.
        for (int x = 0; x < left; x++) D[x] = D[left + orc::reflect101(x - left, src.cols)];
        for (int x = 0; x < right; x++) D[left + src.cols + x] = D[left + orc::reflect101(src.cols + x, src.cols)];
    }
}

static inline void GaussianBlur(const Mat& src, Mat& dst, Size, double, double, int) {
    Mat tmp = (src.data == dst.data) ? src.clone() : src;
    dst.create(src.rows, src.cols, 0);
    orc::gaussian7x7_u8(tmp.data, tmp.cols, tmp.rows, tmp.step, dst.data, dst.step);
}

static inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool nms) {
    std::vector<orc::FastKp> v;
    orc::fast9_16(img.data, img.cols, img.rows, img.step, threshold, nms, v);
    kps.clear();
    for (const orc::FastKp& k : v) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1, (float)k.score));
}

static inline float fastAtan2(float y, float x) { return orc::fast_atan2(y, x); }

struct KeyPointsFilter {   // only referenced from ComputeKeyPointsOld, which is never called (ORBextractor.cc:1057)
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        if (n >= 0 && (int)k.size() > n) {
            std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            k.resize(n);
        }
    }
};

}  // namespace cv
