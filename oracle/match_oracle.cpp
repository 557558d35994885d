// TEST INFRASTRUCTURE ONLY -- CPU oracle for the matching half of the ORB front end.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, load or call this file.  The product path (object_slam_b200/csrc) never does.
//
// Restates over flat arrays (no OpenCV, no Frame/KeyFrame/MapPoint object graph -- src/Frame.cc and
// src/ORBmatcher.cc cannot be compiled here because include/Frame.h pulls in the un-vendored
// Thirdparty/DBoW2 and g2o):
//   ORBmatcher::DescriptorDistance               /root/reference/src/ORBmatcher.cc:1647-1663
//   Frame::ComputeStereoMatches                  /root/reference/src/Frame.cc:706-880
//   Frame::AssignFeaturesToGrid / PosInGrid      src/Frame.cc:455-470, :622-632
//   Frame::GetFeaturesInArea                     src/Frame.cc:567-620
//   ORBmatcher::SearchByProjection(F, MapPoints) src/ORBmatcher.cc:45-129, RadiusByViewingCos :131
//   ORBmatcher::SearchByProjection(F, LastF)     src/ORBmatcher.cc:1328-1470
//   ORBmatcher::SearchForInitialization          src/ORBmatcher.cc:405-520
//   ORBmatcher::ComputeThreeMaxima               src/ORBmatcher.cc:1601-1642
// Parity status: "parity unpinned" by upstream (the reference ships no tests, fixtures or golden
// vectors for these functions and they cannot be executed here); the restatement follows the
// source line by line and is pinned by the committed fixtures under tests/golden/ that this file
// generated, plus the property tests in tests/.  Compiled with -ffp-contract=off: every float
// expression is evaluated with individually rounded binary32 operations.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace orc {

struct KeyPt {            // layout == cv::KeyPoint (28 bytes)
    float x, y, size, angle, response;
    int octave, class_id;
};

static const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;

// ORBmatcher.cc:1647-1663 (the SWAR popcount of the reference, verbatim in meaning)
static inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4);
        memcpy(&y, b + 4 * i, 4);
        unsigned v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// Frame.cc:706-880.  pyrL/pyrR: border-less level images (pitch == width).
// Deviations, all on inputs for which the reference itself is undefined:
//   * minD/maxD are parameters (the reference reads the not-yet-assigned member mb, :736);
//   * row-table indices outside [0, nRows) are dropped (:731 would write out of bounds);
//   * SAD windows that leave the level are skipped (cv::Mat::rowRange/colRange would throw);
//   * an empty accepted list skips the median step (:867 indexes an empty vector).
static void stereo_match(const KeyPt* kL, const uint8_t* dL, int nL, const KeyPt* kR, const uint8_t* dR, int nR,
                         const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                         const float* scale, const float* invScale, float mbf, float minD, float maxD,
                         float* uRight, float* depth, int* sadOut) {
    for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; if (sadOut) sadOut[i] = -1; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = lh[0];
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < nR; iR++) {
        const float kpY = kR[iR].y;
        const float r = 2.0f * scale[kR[iR].octave];
        const int maxr = (int)ceilf(kpY + r);
        const int minr = (int)floorf(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);
    }
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nL; iL++) {
        const KeyPt& kpL = kL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const size_t row = (size_t)vL;
        if (row >= (size_t)nRows) continue;
        const std::vector<size_t>& vCandidates = vRowIndices[row];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dl = dL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const KeyPt& kpR = kR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dl, dR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kR[bestIdxR].x;
            const float scaleFactor = invScale[kpL.octave];
            const float scaleduL = roundf(kpL.x * scaleFactor);
            const float scaledvL = roundf(kpL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const int W = lw[kpL.octave], H = lh[kpL.octave];
            const uint8_t* imL = pyrL[kpL.octave];
            const uint8_t* imR = pyrR[kpL.octave];
            const int cy = (int)scaledvL, cxL = (int)scaleduL, cxR = (int)scaleduR0;
            int bestDistS = INT_MAX, bestincR = 0;
            const int L = 5;
            std::vector<float> vDists(2 * L + 1);
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= W) continue;
            if (cy - w < 0 || cy + w >= H || cxL - w < 0 || cxL + w >= W || cxR - L - w < 0 || cxR + L + w >= W) continue;
            const float centreL = (float)imL[(size_t)cy * W + cxL];
            for (int incR = -L; incR <= +L; incR++) {
                const float centreR = (float)imR[(size_t)cy * W + cxR + incR];
                double acc = 0;     // cv::norm(NORM_L1) on CV_32F accumulates in double; all terms are integers
                for (int dy = -w; dy <= w; dy++)
                    for (int dx = -w; dx <= w; dx++) {
                        const float a = (float)imL[(size_t)(cy + dy) * W + cxL + dx] - centreL;
                        const float b = (float)imR[(size_t)(cy + dy) * W + cxR + incR + dx] - centreR;
                        acc += fabsf(a - b);
                    }
                const float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                if (sadOut) sadOut[iL] = bestDistS;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return;
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
    }
}

}  // namespace orc

using namespace orc;
extern "C" {

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

int orc_stereo_match(const void* kL, const uint8_t* dL, int nL, const void* kR, const uint8_t* dR, int nR,
                     const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                     const float* scale, const float* invScale, float mbf, float minD, float maxD,
                     float* uRight, float* depth, int* sadOut) {
    stereo_match((const KeyPt*)kL, dL, nL, (const KeyPt*)kR, dR, nR, pyrL, pyrR, lw, lh, nlevels, scale, invScale,
                 mbf, minD, maxD, uRight, depth, sadOut);
    return 0;
}

}  // extern "C"
