// TEST INFRASTRUCTURE ONLY.  C entry points around the reference's own, unmodified
// ORB_SLAM2::ORBextractor (compiled in place from /root/reference/src/ORBextractor.cc against
// oracle/cvshim by oracle/Makefile; output oracle/_ref/libref_orbextractor.so, git-ignored).
//
// Allocator: the reference orders quadtree nodes of equal size by heap address
// (sort of pair<int, ExtractorNode*>, ORBextractor.cc:684), so its selected keypoints and
// their order depend on the allocator.  This library gives itself a monotonic bump
// `operator new` (hidden visibility: nothing outside this .so is affected), under which the
// address order equals node creation order -- a legitimate execution of the reference and the
// canonical one the oracle and the CUDA path reproduce.
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <sys/mman.h>

namespace {
struct Arena {
    char* base = nullptr;
    size_t cap = 0, top = 0;
    void init() {
        cap = (size_t)4 << 30;   // virtual reservation; pages are touched on demand
        void* p = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { fprintf(stderr, "ref_harness: arena mmap failed\n"); abort(); }
        base = (char*)p;
    }
    void* alloc(size_t n) {
        if (!base) init();
        size_t a = (top + 15) & ~(size_t)15;
        if (a + n > cap) { fprintf(stderr, "ref_harness: arena exhausted\n"); abort(); }
        top = a + n;
        return base + a;
    }
};
thread_local Arena g_arena;
}  // namespace

#define HIDDEN __attribute__((visibility("hidden")))
HIDDEN void* operator new(size_t n) { return g_arena.alloc(n); }
HIDDEN void* operator new[](size_t n) { return g_arena.alloc(n); }
HIDDEN void operator delete(void*) noexcept {}
HIDDEN void operator delete[](void*) noexcept {}
HIDDEN void operator delete(void*, size_t) noexcept {}
HIDDEN void operator delete[](void*, size_t) noexcept {}

#include "ORBextractor.h"

extern "C" {

// Handles are created and used on the same thread (its arena holds the object's tables).
void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new ORB_SLAM2::ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}

// Runs ORBextractor::operator() (ORBextractor.cc:1043).  kps: cap x 28 B (cv::KeyPoint layout),
// desc: cap x 32 B.  Returns the keypoint count (may exceed cap; only cap entries are written).
int ref_extract(void* h, const uint8_t* img, int w, int ht, size_t stride, void* kps, uint8_t* desc, int cap) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    const size_t mark = g_arena.top;
    int n;
    {
        cv::Mat image(ht, w, CV_8UC1, (void*)img, stride);
        cv::Mat descriptors;
        std::vector<cv::KeyPoint> keys;
        (*e)(image, cv::Mat(), keys, descriptors);
        n = (int)keys.size();
        int m = n < cap ? n : cap;
        if (kps) memcpy(kps, keys.data(), (size_t)m * sizeof(cv::KeyPoint));
        if (desc) for (int i = 0; i < m; i++) memcpy(desc + (size_t)i * 32, descriptors.ptr(i), 32);
    }
    g_arena.top = mark;      // everything allocated during the call is dead; pyramids live in malloc
    return n;
}

int ref_get_level(void* h, int level, uint8_t* dst, int* w, int* ht) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    const cv::Mat& m = e->mvImagePyramid[level];
    *w = m.cols; *ht = m.rows;
    if (dst) for (int y = 0; y < m.rows; y++) memcpy(dst + (size_t)y * m.cols, m.ptr(y), m.cols);
    return m.rows * m.cols;
}

void ref_get_scale_factors(void* h, float* out) {
    ORB_SLAM2::ORBextractor* e = (ORB_SLAM2::ORBextractor*)h;
    std::vector<float> s = e->GetScaleFactors();
    for (size_t i = 0; i < s.size(); i++) out[i] = s[i];
}

}  // extern "C"
