// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ORB front end.
// Nothing under oracle/ is part of the product path: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, load or call it.
//
// This header restates the arithmetic of the five OpenCV primitives the reference's
// ORBextractor calls (OpenCV is an external, un-vendored, un-pinned dependency of
// /root/reference: CMakeLists.txt:31-37).  The restatement is pinned to cv2 4.13.0 in this
// image by tests/test_oracle_primitives.py (bit-exact on random/smooth/binary inputs) and by
// the committed fixtures in tests/golden/.
//
// Call sites in the reference that these stand in for:
//   cv::resize(INTER_LINEAR, 8UC1)      src/ORBextractor.cc:1120
//   cv::GaussianBlur(7x7, sigma 2, REFLECT_101)   src/ORBextractor.cc:1086
//   cv::FAST(img, kps, th, nms=true)    src/ORBextractor.cc:809, :814
//   cv::fastAtan2                        src/ORBextractor.cc:103
//   cvRound / cvFloor / cvCeil           src/ORBextractor.cc:81,115,119,442,456,460,1112
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>
#include <cfloat>
#include <vector>
#include <algorithm>
#include <cstring>
#if defined(__AVX2__)
#include <immintrin.h>
#endif

namespace orc {

// cvRound(float/double): SSE cvtss2si / cvtsd2si = round-half-to-even in the default mode.
static inline int cv_round(double v) { return (int)std::nearbyint(v); }
static inline int cv_round(float v) { return (int)std::nearbyintf(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil(double v) { int i = (int)v; return i + (i < v); }

static inline short sat_short(float v) {
    int i = cv_round(v);
    return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i);
}

// ---------------------------------------------------------------------------------------
// R: cv::resize, CV_8UC1, INTER_LINEAR (OpenCV imgproc resize.cpp, HResizeLinear /
// VResizeLinear fixed-point path, INTER_RESIZE_COEF_BITS = 11).
// ---------------------------------------------------------------------------------------
struct ResizeAxis {
    std::vector<int> ofs;
    std::vector<short> a0, a1;
};

// `zero_at_edges`: the x axis zeroes the fraction when the tap leaves the image; the y axis
// keeps the fraction and clips the row index instead (resize.cpp: xofs/ialpha vs yofs/ibeta).
static inline void resize_axis(int ssize, int dsize, bool is_x, ResizeAxis& ax) {
    ax.ofs.resize(dsize); ax.a0.resize(dsize); ax.a1.resize(dsize);
    double inv_scale = (double)dsize / ssize;
    double scale = 1. / inv_scale;
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (is_x) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        ax.ofs[d] = s;
        ax.a0[d] = sat_short((1.f - f) * 2048.f);
        ax.a1[d] = sat_short(f * 2048.f);
    }
}

static inline void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                                    uint8_t* dst, int dw, int dh, size_t dstride) {
    ResizeAxis X, Y;
    resize_axis(sw, dw, true, X);
    resize_axis(sh, dh, false, Y);
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; dy++) {
        int sy = Y.ofs[dy];
        int y0 = std::min(std::max(sy, 0), sh - 1);
        int y1 = std::min(std::max(sy + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)y0 * sstride;
        const uint8_t* S1 = src + (size_t)y1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = X.ofs[dx];
            int sx1 = std::min(sx + 1, sw - 1);
            r0[dx] = S0[sx] * X.a0[dx] + S0[sx1] * X.a1[dx];
            r1[dx] = S1[sx] * X.a0[dx] + S1[sx1] * X.a1[dx];
        }
        int b0 = Y.a0[dy], b1 = Y.a1[dy];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// B: cv::GaussianBlur(8UC1, Size(7,7), sigma 2, BORDER_REFLECT_101), OpenCV >= 3.4.2/4.x
// fixed-point path (smooth.dispatch.cpp: ufixedpoint16 kernel, 8 fractional bits).
// Kernel bits for sigma=2, n=7: [18,34,48,56,48,34,18] / 256 per axis.
// ---------------------------------------------------------------------------------------
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

static const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

static inline void gaussian7x7_u8(const uint8_t* src, int w, int h, size_t sstride,
                                  uint8_t* dst, size_t dstride) {
    // horizontal pass into 16-bit rows (<= 65280: fits ufixedpoint16 without saturation), then a
    // vertical pass over those rows with one rounding at the end: (acc + 2^15) >> 16.
    std::vector<uint16_t> hbuf((size_t)w * h);
    std::vector<uint8_t> pad((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * sstride;
        for (int x = 0; x < 3; x++) pad[x] = S[reflect101(x - 3, w)];
        memcpy(pad.data() + 3, S, w);
        for (int x = 0; x < 3; x++) pad[w + 3 + x] = S[reflect101(w + x, w)];
        uint16_t* H = &hbuf[(size_t)y * w];
        const uint8_t* P = pad.data();
        for (int x = 0; x < w; x++)
            H[x] = (uint16_t)(18 * (P[x] + P[x + 6]) + 34 * (P[x + 1] + P[x + 5]) + 48 * (P[x + 2] + P[x + 4]) + 56 * P[x + 3]);
    }
    for (int y = 0; y < h; y++) {
        const uint16_t* R[7];
        for (int k = 0; k < 7; k++) R[k] = &hbuf[(size_t)reflect101(y + k - 3, h) * w];
        uint8_t* D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 18u * ((uint32_t)R[0][x] + R[6][x]) + 34u * ((uint32_t)R[1][x] + R[5][x]) +
                           48u * ((uint32_t)R[2][x] + R[4][x]) + 56u * (uint32_t)R[3][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------
// F: cv::FAST, TYPE_9_16, with the cornerScore used for non-max suppression
// (OpenCV features2d fast.cpp FAST_t<16>, fast_score.cpp cornerScore<16>).
// ---------------------------------------------------------------------------------------
struct FastKp { int x, y, score; };

static const int kRing[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                 {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// max over the 16 contiguous 9-arcs of the minimum signed contrast, both polarities.
// A pixel is a FAST-9 corner at threshold t iff fast_contrast(...) > t; its OpenCV score
// is then fast_contrast(...) - 1.
static inline int fast_contrast(const uint8_t* p, ptrdiff_t stride) {
    int d[25];
    int v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[kRing[k & 15][1] * stride + kRing[k & 15][0]];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = 256, mx = -256;
        for (int j = 0; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, mn);      // ring darker than centre
        best = std::max(best, -mx);     // ring brighter than centre
    }
    return best;
}

// Plain double loop; kept as the readable definition and cross-checked against fast9_16 below.
static inline void fast9_16_simple(const uint8_t* img, int w, int h, size_t stride, int threshold,
                                   bool nms, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    std::vector<int> sc((size_t)w * h, 0);
    std::vector<uint8_t> corner((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int c = fast_contrast(img + (size_t)y * stride + x, (ptrdiff_t)stride);
            if (c > threshold) {
                corner[(size_t)y * w + x] = 1;
                sc[(size_t)y * w + x] = c - 1;
                if (!nms) out.push_back({x, y, 0});
            }
        }
    if (!nms) return;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!corner[(size_t)y * w + x]) continue;
            int s = sc[(size_t)y * w + x];
            const int* r0 = &sc[(size_t)(y - 1) * w + x];
            const int* r1 = &sc[(size_t)y * w + x];
            const int* r2 = &sc[(size_t)(y + 1) * w + x];
            if (s > r0[-1] && s > r0[0] && s > r0[1] && s > r1[-1] && s > r1[1] &&
                s > r2[-1] && s > r2[0] && s > r2[1])
                out.push_back({x, y, s});
        }
}

// cornerScore<16> with OpenCV's pruning (fast_score.cpp); == max(threshold, contrast) - 1.
static inline int fast_corner_score(const uint8_t* p, const int* off, int threshold) {
    int d[25];
    int v = p[0];
    for (int k = 0; k < 25; k++) d[k] = v - p[off[k & 15]];
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = std::min(d[k + 1], d[k + 2]);
        a = std::min(a, d[k + 3]);
        if (a <= a0) continue;
        a = std::min(a, d[k + 4]); a = std::min(a, d[k + 5]); a = std::min(a, d[k + 6]);
        a = std::min(a, d[k + 7]); a = std::min(a, d[k + 8]);
        a0 = std::max(a0, std::min(a, d[k]));
        a0 = std::max(a0, std::min(a, d[k + 9]));
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = std::max(d[k + 1], d[k + 2]);
        b = std::max(b, d[k + 3]);
        if (b >= b0) continue;
        b = std::max(b, d[k + 4]); b = std::max(b, d[k + 5]); b = std::max(b, d[k + 6]);
        b = std::max(b, d[k + 7]); b = std::max(b, d[k + 8]);
        b0 = std::min(b0, std::max(b, d[k]));
        b0 = std::min(b0, std::max(b, d[k + 9]));
    }
    return -b0 - 1;
}

static inline bool has_arc9(uint32_t m) {      // 16-bit circular mask: nine contiguous set bits?
    m |= m << 16;
    m &= m >> 1; m &= m >> 2; m &= m >> 4;     // runs of 8
    m &= m >> 1;                                // runs of 9
    return (m & 0xffffu) != 0;
}

// Same structure as OpenCV's scalar FAST_t<16>: opposite-pixel quick reject, contiguous-arc test,
// score rows kept for the 3x3 strict non-max suppression, output in row-major order.
static inline void fast9_16(const uint8_t* img, int w, int h, size_t stride, int threshold,
                            bool nms, std::vector<FastKp>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = kRing[k][1] * (int)stride + kRing[k][0];
    std::vector<uint8_t> rows((size_t)3 * w, 0);
    std::vector<int> cpos((size_t)3 * (w + 1), 0);
    uint8_t* buf[3] = {rows.data(), rows.data() + w, rows.data() + 2 * w};
    int* cp[3] = {cpos.data() + 1, cpos.data() + (w + 1) + 1, cpos.data() + 2 * (w + 1) + 1};
    for (int y = 3; y < h - 2; y++) {
        const uint8_t* ptr = img + (size_t)y * stride + 3;
        uint8_t* curr = buf[(y - 3) % 3];
        int* cornerpos = cp[(y - 3) % 3];
        memset(curr, 0, w);
        int ncorners = 0;
        if (y < h - 3) {
            int x = 3;
#if defined(__AVX2__)
            // 32-pixel compass-point prefilter (as OpenCV's SIMD path does): a 9-arc always holds
            // one of ring pixels {0,8} and one of {4,12}, so both pairs need an outlier.
            const __m256i vt = _mm256_set1_epi8((char)std::min(threshold, 255));
            for (; x + 32 <= w - 3; x += 32) {
                const uint8_t* q = img + (size_t)y * stride + x;
                __m256i v = _mm256_loadu_si256((const __m256i*)q);
                __m256i lo = _mm256_subs_epu8(v, vt), hi = _mm256_adds_epu8(v, vt);
                auto outl = [&](int k) {
                    __m256i r = _mm256_loadu_si256((const __m256i*)(q + off[k]));
                    // r < lo  <=>  max(r,lo) != r ... use saturating subtract: lo - r > 0 ; r - hi > 0
                    __m256i d = _mm256_or_si256(_mm256_subs_epu8(lo, r), _mm256_subs_epu8(r, hi));
                    return _mm256_cmpeq_epi8(d, _mm256_setzero_si256());   // 0xff where NOT an outlier
                };
                __m256i a = _mm256_and_si256(outl(0), outl(8));
                __m256i b = _mm256_and_si256(outl(4), outl(12));
                uint32_t cand = ~(uint32_t)_mm256_movemask_epi8(_mm256_or_si256(a, b));
                while (cand) {
                    int bit = __builtin_ctz(cand);
                    cand &= cand - 1;
                    int xx = x + bit;
                    const uint8_t* pp = q + bit;
                    int vv = pp[0], l2 = vv - threshold, h2 = vv + threshold;
                    uint32_t dark = 0, bright = 0;
                    for (int k = 0; k < 16; k++) {
                        int r = pp[off[k]];
                        dark |= (uint32_t)(r < l2) << k;
                        bright |= (uint32_t)(r > h2) << k;
                    }
                    if (!has_arc9(dark) && !has_arc9(bright)) continue;
                    cornerpos[ncorners++] = xx;
                    if (nms) curr[xx] = (uint8_t)fast_corner_score(pp, off, threshold);
                }
            }
            ptr = img + (size_t)y * stride + x;
#endif
            for (; x < w - 3; x++, ptr++) {
                int v = ptr[0], lo = v - threshold, hi = v + threshold;
                int p0 = ptr[off[0]], p8 = ptr[off[8]];
                if (!((p0 < lo) | (p8 < lo) | (p0 > hi) | (p8 > hi))) continue;
                int p4 = ptr[off[4]], p12 = ptr[off[12]];
                if (!((p4 < lo) | (p12 < lo) | (p4 > hi) | (p12 > hi))) continue;
                uint32_t dark = 0, bright = 0;
                for (int k = 0; k < 16; k++) {
                    int q = ptr[off[k]];
                    dark |= (uint32_t)(q < lo) << k;
                    bright |= (uint32_t)(q > hi) << k;
                }
                if (!has_arc9(dark) && !has_arc9(bright)) continue;
                cornerpos[ncorners++] = x;
                if (nms) curr[x] = (uint8_t)fast_corner_score(ptr, off, threshold);
            }
        }
        cornerpos[-1] = ncorners;
        if (y == 3) continue;
        const uint8_t* prev = buf[(y - 4 + 3) % 3];
        const uint8_t* pprev = buf[(y - 5 + 3) % 3];
        const int* ppos = cp[(y - 4 + 3) % 3];
        int n = ppos[-1];
        for (int k = 0; k < n; k++) {
            int x = ppos[k];
            int s = prev[x];
            if (!nms || (s > prev[x + 1] && s > prev[x - 1] && s > pprev[x - 1] && s > pprev[x] &&
                         s > pprev[x + 1] && s > curr[x - 1] && s > curr[x] && s > curr[x + 1]))
                out.push_back({x, y - 1, s});
        }
    }
}

// ---------------------------------------------------------------------------------------
// A: cv::fastAtan2 (core mathfuncs_core.simd.hpp atanImpl, scalar path), degrees in [0,360).
// Every operation is a separately rounded binary32 op (compile with -ffp-contract=off).
// ---------------------------------------------------------------------------------------
static inline float fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------
// S: glibc >= 2.28 sincosf (ARM optimized-routines), restated in double so that it can be
// reproduced on a device that has no glibc.  The oracle itself calls libm cosf/sinf, as the
// reference does (src/ORBextractor.cc:113); tests check model == libm on this image.
// Valid for |y| < 120 (ORB angles are < 2*pi).
// ---------------------------------------------------------------------------------------
static inline uint32_t f32_bits(float f) { uint32_t u; __builtin_memcpy(&u, &f, 4); return u; }

static inline void sincosf_model(float y, float* sinp, float* cosp) {
    // polynomial tables (sincosf_data.c, __sincosf_table[0] and [1])
    static const double P[2][8] = {
        // c0, c1, c2, c3, c4, s1, s2, s3
        {1.0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16,
         -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
        {-1.0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16,
         -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
    static const double sgn[4] = {1.0, -1.0, -1.0, 1.0};
    double x = (double)y;
    uint32_t top = (f32_bits(y) >> 20) & 0x7ff;
    const uint32_t top_pio4 = (f32_bits(0x1.921FB6p-1f) >> 20) & 0x7ff;
    const uint32_t top_tiny = (f32_bits(0x1p-12f) >> 20) & 0x7ff;
    int n = 0;
    const double* p = P[0];
    if (top < top_pio4) {
        if (top < top_tiny) { *sinp = y; *cosp = 1.0f; return; }
    } else {
        double r = x * 0x1.45F306DC9C883p+23;          // x * 2/pi * 2^24
        n = ((int32_t)r + 0x800000) >> 24;
        x = x - n * 0x1.921FB54442D18p0;               // x - n*pi/2
        double s = sgn[n & 3];
        if (n & 2) p = P[1];
        double x2u = x * x;                            // glibc squares the unsigned x
        x = x * s;
        // polynomial with x2 from the unsigned reduction
        double x2 = x2u, x4 = x2 * x2, x3 = x2 * x;
        double c2 = p[3] + x2 * p[4], s1 = p[6] + x2 * p[7], c1 = p[0] + x2 * p[1];
        double x5 = x3 * x2, x6 = x4 * x2;
        double S = x + x3 * p[5], C = c1 + x4 * p[2];
        float sv = (float)(S + x5 * s1), cv = (float)(C + x6 * c2);
        if (n & 1) { *sinp = cv; *cosp = sv; } else { *sinp = sv; *cosp = cv; }
        return;
    }
    double x2 = x * x, x4 = x2 * x2, x3 = x2 * x;
    double c2 = p[3] + x2 * p[4], s1 = p[6] + x2 * p[7], c1 = p[0] + x2 * p[1];
    double x5 = x3 * x2, x6 = x4 * x2;
    double S = x + x3 * p[5], C = c1 + x4 * p[2];
    *sinp = (float)(S + x5 * s1);
    *cosp = (float)(C + x6 * c2);
}

}  // namespace orc
