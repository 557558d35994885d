// Compile-and-run check of the C++ drop-in (tests/test_dropin_cpp.py builds it against
// oracle/cvshim, because OpenCV's C++ headers are not in this image, and runs it on the GPU box):
// reads a raw 8-bit image, runs ORB_SLAM2::ORBextractor::operator() and, given two images, the
// stereo matcher, and writes keypoints / descriptors / uRight / depth as raw files.
#include "ORBextractor.h"
#include <cstdio>
#include <cstdlib>
#include <vector>

static std::vector<unsigned char> readAll(const char* path, size_t n) {
    std::vector<unsigned char> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}
static void writeAll(const char* path, const void* p, size_t n) {
    FILE* f = fopen(path, "wb");
    if (!f || fwrite(p, 1, n, f) != n) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s w h nfeatures left.raw outprefix [right.raw mbf maxD]\n", argv[0]); return 2; }
    const int w = atoi(argv[1]), h = atoi(argv[2]), nf = atoi(argv[3]);
    std::vector<unsigned char> L = readAll(argv[4], (size_t)w * h);
    std::string out = argv[5];
    ORB_SLAM2::ORBextractor exL(nf, 1.2f, 8, 20, 7);
    cv::Mat imL(h, w, CV_8UC1, L.data(), (size_t)w);
    std::vector<cv::KeyPoint> kL;
    cv::Mat dL;
    exL(imL, cv::Mat(), kL, dL);
    writeAll((out + ".kp").c_str(), kL.data(), kL.size() * sizeof(cv::KeyPoint));
    std::vector<unsigned char> d(kL.size() * 32);
    for (size_t i = 0; i < kL.size(); i++) memcpy(&d[i * 32], dL.ptr((int)i), 32);
    writeAll((out + ".desc").c_str(), d.data(), d.size());
    printf("left: %zu keypoints, %d levels, level0 %dx%d\n", kL.size(), exL.GetLevels(), exL.mvImagePyramid[0].cols, exL.mvImagePyramid[0].rows);
    if (argc >= 9) {
        std::vector<unsigned char> R = readAll(argv[6], (size_t)w * h);
        ORB_SLAM2::ORBextractor exR(nf, 1.2f, 8, 20, 7);
        cv::Mat imR(h, w, CV_8UC1, R.data(), (size_t)w);
        std::vector<cv::KeyPoint> kR;
        cv::Mat dR;
        exR(imR, cv::Mat(), kR, dR);
        std::vector<float> uR, depth;
        ORB_SLAM2::ComputeStereoMatchesB200(&exL, &exR, (float)atof(argv[7]), 0.f, (float)atof(argv[8]), (int)kL.size(), uR, depth);
        writeAll((out + ".uright").c_str(), uR.data(), uR.size() * 4);
        writeAll((out + ".depth").c_str(), depth.data(), depth.size() * 4);
        int m = 0;
        for (float v : uR) m += v >= 0;
        printf("stereo: %d matches\n", m);
    }
    return 0;
}
