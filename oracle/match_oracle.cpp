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
//   ORBmatcher::SearchByProjection(F, KF, ...)   src/ORBmatcher.cc:1472-1599;  (KF, Scw, ...) :290-403
//   MapPoint::PredictScale                       src/MapPoint.cc:488-519
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


// ------------------------------------------------------------------------------------------------
// Flat view of what the matchers read of a Frame (include/Frame.h): undistorted keypoints, descriptors,
// mvuRight, the image bounds and the 64x48 keypoint grid.
// ------------------------------------------------------------------------------------------------
static const int GRID_COLS = 64, GRID_ROWS = 48;      // include/Frame.h:41-42

struct FrameV {
    int n = 0;
    std::vector<KeyPt> k;          // mvKeysUn
    std::vector<uint8_t> d;        // mDescriptors, n x 32
    std::vector<float> uR;         // mvuRight (-1 = monocular point)
    float minX = 0, maxX = 0, minY = 0, maxY = 0;
    float invW = 0, invH = 0;      // mfGridElementWidthInv / HeightInv, Frame.cc:101-102
    std::vector<std::vector<size_t>> grid;   // mGrid[ix][iy] -> index ix*GRID_ROWS+iy

    // Frame.cc:622-632
    bool pos_in_grid(const KeyPt& kp, int& posX, int& posY) const {
        posX = (int)roundf((kp.x - minX) * invW);
        posY = (int)roundf((kp.y - minY) * invH);
        if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) return false;
        return true;
    }
    // Frame.cc:455-470
    void assign_features_to_grid() {
        grid.assign((size_t)GRID_COLS * GRID_ROWS, std::vector<size_t>());
        for (int i = 0; i < n; i++) {
            int gx, gy;
            if (pos_in_grid(k[i], gx, gy)) grid[(size_t)gx * GRID_ROWS + gy].push_back(i);
        }
    }
    // Frame.cc:567-620
    std::vector<size_t> features_in_area(float x, float y, float r, int minLevel = -1, int maxLevel = -1) const {
        std::vector<size_t> vIndices;
        const int nMinCellX = std::max(0, (int)floorf((x - minX - r) * invW));
        if (nMinCellX >= GRID_COLS) return vIndices;
        const int nMaxCellX = std::min(GRID_COLS - 1, (int)ceilf((x - minX + r) * invW));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floorf((y - minY - r) * invH));
        if (nMinCellY >= GRID_ROWS) return vIndices;
        const int nMaxCellY = std::min(GRID_ROWS - 1, (int)ceilf((y - minY + r) * invH));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<size_t>& vCell = grid[(size_t)ix * GRID_ROWS + iy];
                for (size_t j = 0; j < vCell.size(); j++) {
                    const KeyPt& kpUn = k[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0)
                            if (kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.x - x;
                    const float disty = kpUn.y - y;
                    if (fabsf(distx) < r && fabsf(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }
};

// ORBmatcher.cc:1601-1642 over bin sizes.
static void compute_three_maxima(const int* histoSize, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histoSize[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// rotation-histogram bin of a match, ORBmatcher.cc:1425-1433 (same text at :472-480, :1556-1564)
static inline int rot_bin(float angle1, float angle2) {
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = angle1 - angle2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// The pointer state F.mvpMapPoints[idx] is modelled by two arrays over the keypoints:
//   kpObs[idx]   = Observations() of the map point the keypoint currently holds (0 when it holds none
//                  or a point without observations) -- the only thing the matchers ask of it;
//   kpMatch[idx] = -1 untouched by this call, >= 0 index (into this call's point list) of the point
//                  assigned last, -2 reset to NULL by the rotation check.

// ORBmatcher.cc:131-137
static inline float radius_by_viewing_cos(float viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:45-129.
// Per map point: inView = mbTrackInView && !isBad(); projX/projY/projXR = mTrackProjX/Y/XR;
// level = mnTrackScaleLevel; viewCos = mTrackViewCos; mpObs = Observations().
static int search_by_projection_map(const FrameV& F, const float* scale, int nMP, const uint8_t* inView,
                                    const float* projX, const float* projY, const float* projXR, const int* level,
                                    const float* viewCos, const uint8_t* mpDesc, const int* mpObs, float th, float nnratio,
                                    int* kpObs, int* kpMatch) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int iMP = 0; iMP < nMP; iMP++) {
        if (!inView[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = radius_by_viewing_cos(viewCos[iMP]);
        if (bFactor) r *= th;
        const std::vector<size_t> vIndices =
            F.features_in_area(projX[iMP], projY[iMP], r * scale[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel);
        if (vIndices.empty()) continue;
        const uint8_t* MPdescriptor = mpDesc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t vi = 0; vi < vIndices.size(); vi++) {
            const size_t idx = vIndices[vi];
            if (kpObs[idx] > 0) continue;
            if (F.uR[idx] > 0) {
                const float er = fabsf(projXR[iMP] - F.uR[idx]);
                if (er > r * scale[nPredictedLevel]) continue;
            }
            const int dist = descriptor_distance(MPdescriptor, &F.d[idx * 32]);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist;
                bestLevel2 = bestLevel; bestLevel = F.k[idx].octave;
                bestIdx = (int)idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F.k[idx].octave;
                bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            kpMatch[bestIdx] = iMP;
            kpObs[bestIdx] = mpObs[iMP];
            nmatches++;
        }
    }
    return nmatches;
}

// cv::Mat float arithmetic of the pose expressions, pinned to cv2 4.13 (tests/test_oracle_matchers.py):
// A(3x3)*x(3x1)+c runs gemm's small-matrix branch -- binary32 products and sums left to right, then
// (float)((double)t*alpha + (double)c*beta); -A.t()*x runs the generic branch with binary64 accumulation.
static inline void mat_rx_plus_t(const float* T, const float* x, float* out) {   // T = 3x4 row-major [R|t]
    for (int r = 0; r < 3; r++) {
        const float t0 = T[4 * r + 0] * x[0] + T[4 * r + 1] * x[1] + T[4 * r + 2] * x[2];
        out[r] = (float)((double)t0 * 1.0 + (double)T[4 * r + 3] * 1.0);
    }
}
static inline void mat_minus_rt_t(const float* T, float* out) {                  // -R^T * t
    for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int kk = 0; kk < 3; kk++) s += (double)T[4 * kk + r] * (double)T[4 * kk + 3];
        out[r] = (float)(s * -1.0);
    }
}

struct Camera { float fx, fy, cx, cy, mbf, mb; };

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono), ORBmatcher.cc:1328-1470.
// Per last-frame keypoint i: lastHasPoint = mvpMapPoints[i] && !mvbOutlier[i]; lastPos = GetWorldPos();
// lastOctave = mvKeys[i].octave; lastAngle = mvKeysUn[i].angle; mpDesc = GetDescriptor(); mpObs = Observations().
static int search_by_projection_last(const FrameV& Cur, const float* scale, const Camera& cam, const float* TcwCur,
                                     const float* TcwLast, int nLast, const uint8_t* lastHasPoint, const float* lastPos,
                                     const int* lastOctave, const float* lastAngle, const uint8_t* mpDesc, const int* mpObs,
                                     float th, bool bMono, bool checkOri, int* kpObs, int* kpMatch) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    float twc[3], tlc[3];
    mat_minus_rt_t(TcwCur, twc);
    mat_rx_plus_t(TcwLast, twc, tlc);
    const bool bForward = tlc[2] > cam.mb && !bMono;
    const bool bBackward = -tlc[2] > cam.mb && !bMono;
    for (int i = 0; i < nLast; i++) {
        if (!lastHasPoint[i]) continue;
        float x3Dc[3];
        mat_rx_plus_t(TcwCur, lastPos + 3 * i, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        if (invzc < 0) continue;
        float u = cam.fx * xc * invzc + cam.cx;
        float v = cam.fy * yc * invzc + cam.cy;
        if (u < Cur.minX || u > Cur.maxX) continue;
        if (v < Cur.minY || v > Cur.maxY) continue;
        const int nLastOctave = lastOctave[i];
        const float radius = th * scale[nLastOctave];
        std::vector<size_t> vIndices2;
        if (bForward) vIndices2 = Cur.features_in_area(u, v, radius, nLastOctave);
        else if (bBackward) vIndices2 = Cur.features_in_area(u, v, radius, 0, nLastOctave);
        else vIndices2 = Cur.features_in_area(u, v, radius, nLastOctave - 1, nLastOctave + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mpDesc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t vi = 0; vi < vIndices2.size(); vi++) {
            const size_t i2 = vIndices2[vi];
            if (kpObs[i2] > 0) continue;
            if (Cur.uR[i2] > 0) {
                const float ur = u - cam.mbf * invzc;
                const float er = fabsf(ur - Cur.uR[i2]);
                if (er > radius) continue;
            }
            const int dist = descriptor_distance(dMP, &Cur.d[i2 * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
        }
        if (bestDist <= TH_HIGH) {
            kpMatch[bestIdx2] = i;
            kpObs[bestIdx2] = mpObs[i];
            nmatches++;
            if (checkOri) rotHist[rot_bin(lastAngle[i], Cur.k[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    kpMatch[rotHist[i][j]] = -2;
                    kpObs[rotHist[i][j]] = 0;
                    nmatches--;
                }
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:405-520.  prevMatched: n1 x (x,y), in/out.
static int search_for_initialization(const FrameV& F1, const FrameV& F2, float* prevMatched, int* vnMatches12,
                                     int windowSize, float nnratio, bool checkOri) {
    int nmatches = 0;
    for (int i = 0; i < F1.n; i++) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2.n, INT_MAX);
    std::vector<int> vnMatches21(F2.n, -1);
    for (int i1 = 0; i1 < F1.n; i1++) {
        const KeyPt kp1 = F1.k[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        std::vector<size_t> vIndices2 = F2.features_in_area(prevMatched[2 * i1], prevMatched[2 * i1 + 1], (float)windowSize, level1, level1);
        if (vIndices2.empty()) continue;
        const uint8_t* d1 = &F1.d[(size_t)i1 * 32];
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (size_t vi = 0; vi < vIndices2.size(); vi++) {
            const size_t i2 = vIndices2[vi];
            const int dist = descriptor_distance(d1, &F2.d[i2 * 32]);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = (int)i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    vnMatches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(F1.k[i1].angle, F2.k[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                const int idx1 = rotHist[i][j];
                if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
            }
        }
    }
    for (int i1 = 0; i1 < F1.n; i1++)
        if (vnMatches12[i1] >= 0) {
            prevMatched[2 * i1] = F2.k[vnMatches12[i1]].x;
            prevMatched[2 * i1 + 1] = F2.k[vnMatches12[i1]].y;
        }
    return nmatches;
}

// MapPoint::PredictScale, MapPoint.cc:488-519 (log and ceil resolve to the float overloads: the headers pull
// `using namespace std` into the global namespace, include/ObjectTypes.h:14).
static inline int predict_scale(float maxDistanceRaw, float currentDist, float logScaleFactor, int nLevels) {
    const float ratio = maxDistanceRaw / currentDist;
    int nScale = (int)ceilf(logf(ratio) / logScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= nLevels) nScale = nLevels - 1;
    return nScale;
}

// cv::norm(3x1 CV_32F) = sqrt of the binary64 sum of squares; Mat::dot of two 3x1 CV_32F = binary64 sum of products
static inline float norm3(const float* v) {
    double s = 0;
    for (int i = 0; i < 3; i++) { const double d = v[i]; s += d * d; }
    return (float)std::sqrt(s);
}
static inline double dot3(const float* a, const float* b) {
    double s = 0;
    for (int i = 0; i < 3; i++) s += (double)a[i] * (double)b[i];
    return s;
}

// Per map point of the two searches below: valid = pMP && !isBad() && not in the already-found set;
// pos = GetWorldPos(); minDist/maxDist = Get{Min,Max}DistanceInvariance(); maxDistRaw = mfMaxDistance (PredictScale);
// normal = GetNormal(); angle = pKF->mvKeysUn[i].angle; desc = GetDescriptor().
// kpTaken[idx] != 0 <=> mvpMapPoints[idx] / vpMatched[idx] is non-NULL (any point blocks, with or without observations).

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
// ORBmatcher.cc:1472-1599.
static int search_by_projection_keyframe(const FrameV& Cur, const float* scale, int nLevels, float logScaleFactor, const Camera& cam,
                                         const float* Tcw, int nPts, const uint8_t* valid, const float* pos, const float* minDist,
                                         const float* maxDist, const float* maxDistRaw, const float* angle, const uint8_t* desc,
                                         float th, int ORBdist, bool checkOri, int* kpTaken, int* kpMatch) {
    int nmatches = 0;
    float Ow[3];
    mat_minus_rt_t(Tcw, Ow);
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < nPts; i++) {
        if (!valid[i]) continue;
        float x3Dc[3];
        mat_rx_plus_t(Tcw, pos + 3 * i, x3Dc);
        const float xc = x3Dc[0], yc = x3Dc[1];
        const float invzc = 1.0 / x3Dc[2];
        const float u = cam.fx * xc * invzc + cam.cx;
        const float v = cam.fy * yc * invzc + cam.cy;
        if (u < Cur.minX || u > Cur.maxX) continue;
        if (v < Cur.minY || v > Cur.maxY) continue;
        const float PO[3] = {pos[3 * i] - Ow[0], pos[3 * i + 1] - Ow[1], pos[3 * i + 2] - Ow[2]};
        const float dist3D = norm3(PO);
        if (dist3D < minDist[i] || dist3D > maxDist[i]) continue;
        const int nPredictedLevel = predict_scale(maxDistRaw[i], dist3D, logScaleFactor, nLevels);
        const float radius = th * scale[nPredictedLevel];
        const std::vector<size_t> vIndices2 = Cur.features_in_area(u, v, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t vi = 0; vi < vIndices2.size(); vi++) {
            const size_t i2 = vIndices2[vi];
            if (kpTaken[i2]) continue;
            const int dist = descriptor_distance(dMP, &Cur.d[i2 * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
        }
        if (bestDist <= ORBdist) {
            kpMatch[bestIdx2] = i;
            kpTaken[bestIdx2] = 1;
            nmatches++;
            if (checkOri) rotHist[rot_bin(angle[i], Cur.k[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) {
                    kpMatch[rotHist[i][j]] = -2;
                    kpTaken[rotHist[i][j]] = 0;
                    nmatches--;
                }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, th),
// ORBmatcher.cc:290-403, after the decomposition of Scw (:299-304): Tcw = [Rcw | tcw] with Rcw = sRcw/scw, tcw = Scw.col(3)/scw.
static int search_by_projection_sim3(const FrameV& KF, const float* scale, int nLevels, float logScaleFactor, const Camera& cam,
                                     const float* Tcw, int nPts, const uint8_t* valid, const float* pos, const float* minDist,
                                     const float* maxDist, const float* maxDistRaw, const float* normal, const uint8_t* desc,
                                     int th, int* kpTaken, int* kpMatch) {
    int nmatches = 0;
    float Ow[3];
    mat_minus_rt_t(Tcw, Ow);
    for (int iMP = 0; iMP < nPts; iMP++) {
        if (!valid[iMP]) continue;
        const float* p3Dw = pos + 3 * iMP;
        float p3Dc[3];
        mat_rx_plus_t(Tcw, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0) continue;
        const float invz = 1 / p3Dc[2];
        const float x = p3Dc[0] * invz;
        const float y = p3Dc[1] * invz;
        const float u = cam.fx * x + cam.cx;
        const float v = cam.fy * y + cam.cy;
        if (!(u >= KF.minX && u < KF.maxX && v >= KF.minY && v < KF.maxY)) continue;      // KeyFrame::IsInImage, KeyFrame.cc:610-613
        const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist = norm3(PO);
        if (dist < minDist[iMP] || dist > maxDist[iMP]) continue;
        if (dot3(PO, normal + 3 * iMP) < 0.5 * dist) continue;
        const int nPredictedLevel = predict_scale(maxDistRaw[iMP], dist, logScaleFactor, nLevels);
        const float radius = th * scale[nPredictedLevel];
        const std::vector<size_t> vIndices = KF.features_in_area(u, v, radius);          // KeyFrame::GetFeaturesInArea, KeyFrame.cc:569-608
        if (vIndices.empty()) continue;
        const uint8_t* dMP = desc + (size_t)iMP * 32;
        int bestDist = 256, bestIdx = -1;
        for (size_t vi = 0; vi < vIndices.size(); vi++) {
            const size_t idx = vIndices[vi];
            if (kpTaken[idx]) continue;
            const int kpLevel = KF.k[idx].octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            const int dist2 = descriptor_distance(dMP, &KF.d[idx * 32]);
            if (dist2 < bestDist) { bestDist = dist2; bestIdx = (int)idx; }
        }
        if (bestDist <= TH_LOW) {
            kpMatch[bestIdx] = iMP;
            kpTaken[bestIdx] = 1;
            nmatches++;
        }
    }
    return nmatches;
}

// The search half of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th), ORBmatcher.cc:825-966 (sim3 = 0: Tcw =
// [GetRotation() | GetTranslation()], Ow = GetCameraCenter(), stereo / mono chi-square gates on the reprojection error) and of
// ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint), :974-1100 (sim3 = 1: after the decomposition of Scw,
// Ow = -Rcw.t()*tcw, invz = 1.0/z in binary64, no reprojection gate).  valid = "pMP && !isBad() && !IsInKeyFrame(pKF)".
// Per map point: bestIdx (-1 = none) and bestDist; the replace / add-observation bookkeeping (map mutation) stays with the caller,
// which accepts bestDist <= TH_LOW.  Nothing here depends on that bookkeeping, so the points are independent.
static void fuse_search(const FrameV& KF, const float* scale, const float* invSigma2, int nLevels, float logScaleFactor, const Camera& cam,
                        const float* Tcw, const float* OwIn, int sim3, int nPts, const uint8_t* valid, const float* pos,
                        const float* minDist, const float* maxDist, const float* maxDistRaw, const float* normal, const uint8_t* desc,
                        float th, int* bestIdxOut, int* bestDistOut) {
    float Ow[3];
    if (sim3) mat_minus_rt_t(Tcw, Ow); else memcpy(Ow, OwIn, sizeof(Ow));
    for (int i = 0; i < nPts; i++) {
        bestIdxOut[i] = -1; bestDistOut[i] = 256;
        if (!valid[i]) continue;
        const float* p3Dw = pos + 3 * i;
        float p3Dc[3];
        mat_rx_plus_t(Tcw, p3Dw, p3Dc);
        if (p3Dc[2] < 0.0f) continue;
        const float invz = sim3 ? (float)(1.0 / p3Dc[2]) : 1 / p3Dc[2];
        const float x = p3Dc[0] * invz;
        const float y = p3Dc[1] * invz;
        const float u = cam.fx * x + cam.cx;
        const float v = cam.fy * y + cam.cy;
        if (!(u >= KF.minX && u < KF.maxX && v >= KF.minY && v < KF.maxY)) continue;      // KeyFrame::IsInImage
        const float ur = u - cam.mbf * invz;
        const float PO[3] = {p3Dw[0] - Ow[0], p3Dw[1] - Ow[1], p3Dw[2] - Ow[2]};
        const float dist3D = norm3(PO);
        if (dist3D < minDist[i] || dist3D > maxDist[i]) continue;
        if (dot3(PO, normal + 3 * i) < 0.5 * dist3D) continue;
        const int nPredictedLevel = predict_scale(maxDistRaw[i], dist3D, logScaleFactor, nLevels);
        const float radius = th * scale[nPredictedLevel];
        const std::vector<size_t> vIndices = KF.features_in_area(u, v, radius);
        if (vIndices.empty()) continue;
        const uint8_t* dMP = desc + (size_t)i * 32;
        int bestDist = 256, bestIdx = -1;
        for (size_t vi = 0; vi < vIndices.size(); vi++) {
            const size_t idx = vIndices[vi];
            const KeyPt& kp = KF.k[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (!sim3) {
                if (KF.uR[idx] >= 0) {                                  // reprojection error in stereo, :909-921
                    const float ex = u - kp.x, ey = v - kp.y, er = ur - KF.uR[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * invSigma2[kpLevel] > 7.8) continue;
                } else {
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * invSigma2[kpLevel] > 5.99) continue;
                }
            }
            const int dist = descriptor_distance(dMP, &KF.d[idx * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
        }
        bestIdxOut[i] = bestIdx; bestDistOut[i] = bestDist;
    }
}

// One direction of ORBmatcher::SearchBySim3 (ORBmatcher.cc:1153-1226 resp. :1228-1302): map points of the source keyframe
// (indexed by its keypoints) are taken to the target camera by p3Dc = T2 * (T1 * p3Dw) with T1 = [R1w | t1w] and
// T2 = [sR21 | t21], projected, gated by the scale-invariance distances on cv::norm(p3Dc), and matched to the closest
// descriptor in the predicted window; vnMatch[i] = bestIdx if bestDist <= TH_HIGH else -1.
static void sim3_direction(const FrameV& Target, const float* scale, int nLevels, float logScaleFactor, const Camera& cam, const float* T1,
                           const float* T2, int nPts, const uint8_t* valid, const float* pos, const float* minDist, const float* maxDist,
                           const float* maxDistRaw, const uint8_t* desc, float th, int* vnMatch) {
    for (int i = 0; i < nPts; i++) {
        vnMatch[i] = -1;
        if (!valid[i]) continue;
        float pa[3], pc[3];
        mat_rx_plus_t(T1, pos + 3 * i, pa);
        mat_rx_plus_t(T2, pa, pc);
        if (pc[2] < 0.0) continue;
        const float invz = (float)(1.0 / pc[2]);
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float u = cam.fx * x + cam.cx, v = cam.fy * y + cam.cy;
        if (!(u >= Target.minX && u < Target.maxX && v >= Target.minY && v < Target.maxY)) continue;
        const float dist3D = norm3(pc);
        if (dist3D < minDist[i] || dist3D > maxDist[i]) continue;
        const int nPredictedLevel = predict_scale(maxDistRaw[i], dist3D, logScaleFactor, nLevels);
        const float radius = th * scale[nPredictedLevel];
        const std::vector<size_t> vIndices = Target.features_in_area(u, v, radius);
        if (vIndices.empty()) continue;
        int bestDist = INT_MAX, bestIdx = -1;
        for (size_t vi = 0; vi < vIndices.size(); vi++) {
            const size_t idx = vIndices[vi];
            const int oct = Target.k[idx].octave;
            if (oct < nPredictedLevel - 1 || oct > nPredictedLevel) continue;
            const int dist = descriptor_distance(desc + (size_t)i * 32, &Target.d[idx * 32]);
            if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
        }
        if (bestDist <= TH_HIGH) vnMatch[i] = bestIdx;
    }
}

// Brute-force best / second-best Hamming search with the ratio test: the inner loop of
// ORBmatcher::SearchByBoW (ORBmatcher.cc:200-229) over one list of candidates, without the
// "already matched" bookkeeping (every query is independent).  bestIdx = -1 when rejected.
static void hamming_knn2(const uint8_t* q, int nq, const uint8_t* db, int nd, int thLow, float nnratio,
                         int* bestIdx, int* bestDist, int* secondDist) {
    for (int i = 0; i < nq; i++) {
        int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
        for (int j = 0; j < nd; j++) {
            const int dist = descriptor_distance(q + (size_t)i * 32, db + (size_t)j * 32);
            if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = j; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        bestDist[i] = bestDist1;
        secondDist[i] = bestDist2;
        bestIdx[i] = -1;
        if (bestDist1 <= thLow)
            if ((float)bestDist1 < nnratio * (float)bestDist2) bestIdx[i] = bestIdxF;
    }
}

// ------------------------------------------------------------------------------------------------
// DBoW2-gated matchers.  A DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>, un-vendored
// Thirdparty/DBoW2) is passed as a CSR: nodeId[] ascending (the map's iteration order), nodeStart[], nodeIdx[].
// ------------------------------------------------------------------------------------------------
struct BowSide {
    int n;                       // keypoints
    const uint8_t* desc;         // n x 32
    const KeyPt* keys;           // mvKeysUn
    const uint8_t* valid;        // per keypoint, meaning depends on the caller (may be null = all)
    const float* uRight;         // mvuRight (triangulation only; may be null = all -1)
    int nNodes; const uint32_t* nodeId; const int* nodeStart; const int* nodeIdx;
};

// std::map::lower_bound over the ascending node ids
static int bow_lower_bound(const BowSide& S, uint32_t id) {
    return (int)(std::lower_bound(S.nodeId, S.nodeId + S.nNodes, id) - S.nodeId);
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), ORBmatcher.cc:159-288 (strictLow = 0: bestDist1 <= TH_LOW,
// set 1 = the keyframe with valid = "has a good map point", set 2 = the frame, result read through match21) and
// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), :522-655 (strictLow = 1: bestDist1 < TH_LOW, valid on
// both sides, result read through match12).  match12[i1] = i2 / match21[i2] = i1, -1 where there is no match.
static int search_by_bow(const BowSide& A, const BowSide& B, int thLow, int strictLow, float nnratio, int checkOri,
                         int* match12, int* match21) {
    for (int i = 0; i < A.n; i++) match12[i] = -1;
    for (int i = 0; i < B.n; i++) match21[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < A.nNodes && ib < B.nNodes) {
        if (A.nodeId[ia] == B.nodeId[ib]) {
            for (int p1 = A.nodeStart[ia]; p1 < A.nodeStart[ia + 1]; p1++) {
                const int idx1 = A.nodeIdx[p1];
                if (A.valid && !A.valid[idx1]) continue;
                const uint8_t* d1 = A.desc + (size_t)idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int p2 = B.nodeStart[ib]; p2 < B.nodeStart[ib + 1]; p2++) {
                    const int idx2 = B.nodeIdx[p2];
                    if (match21[idx2] >= 0) continue;                         // vpMapPointMatches[realIdxF] / vbMatched2[idx2]
                    if (B.valid && !B.valid[idx2]) continue;
                    const int dist = descriptor_distance(d1, B.desc + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (strictLow ? bestDist1 < thLow : bestDist1 <= thLow) {
                    if ((float)bestDist1 < nnratio * (float)bestDist2) {
                        match12[idx1] = bestIdx2;
                        match21[bestIdx2] = idx1;
                        if (checkOri) rotHist[rot_bin(A.keys[idx1].angle, B.keys[bestIdx2].angle)].push_back(idx1);
                        nmatches++;
                    }
                }
            }
            ia++; ib++;
        } else if (A.nodeId[ia] < B.nodeId[ib]) ia = bow_lower_bound(A, B.nodeId[ib]);
        else ib = bow_lower_bound(B, A.nodeId[ia]);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) {
                const int idx1 = rotHist[i][j];
                match21[match12[idx1]] = -1;
                match12[idx1] = -1;
                nmatches--;
            }
        }
    }
    return nmatches;
}

// ORBmatcher::CheckDistEpipolarLine, ORBmatcher.cc:139-156
static bool check_dist_epipolar_line(const KeyPt& kp1, const KeyPt& kp2, const float* F12, const float* levelSigma2) {
    const float a = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
    const float b = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
    const float c = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * levelSigma2[kp2.octave];
}

// ORBmatcher::SearchForTriangulation, ORBmatcher.cc:657-823.  valid = "keypoint has no map point yet" (GetMapPoint == NULL)
// on both sides.  (ex, ey): the epipole as :663-671 computes it.  The reference never sets vbMatched2 (:677 is only read),
// so every keypoint of set 1 is matched on its own; ties on the distance go to the LAST candidate (dist > bestDist rejects).
static int search_for_triangulation(const BowSide& A, const BowSide& B, const float* F12, float ex, float ey,
                                    const float* levelSigma2, const float* scaleFactors, int onlyStereo, int checkOri,
                                    int* match12) {
    for (int i = 0; i < A.n; i++) match12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    int ia = 0, ib = 0;
    while (ia < A.nNodes && ib < B.nNodes) {
        if (A.nodeId[ia] == B.nodeId[ib]) {
            for (int p1 = A.nodeStart[ia]; p1 < A.nodeStart[ia + 1]; p1++) {
                const int idx1 = A.nodeIdx[p1];
                if (A.valid && !A.valid[idx1]) continue;                      // already a map point
                const bool bStereo1 = A.uRight && A.uRight[idx1] >= 0;
                if (onlyStereo && !bStereo1) continue;
                const KeyPt& kp1 = A.keys[idx1];
                const uint8_t* d1 = A.desc + (size_t)idx1 * 32;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int p2 = B.nodeStart[ib]; p2 < B.nodeStart[ib + 1]; p2++) {
                    const int idx2 = B.nodeIdx[p2];
                    if (B.valid && !B.valid[idx2]) continue;
                    const bool bStereo2 = B.uRight && B.uRight[idx2] >= 0;
                    if (onlyStereo && !bStereo2) continue;
                    const int dist = descriptor_distance(d1, B.desc + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPt& kp2 = B.keys[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ex - kp2.x, distey = ey - kp2.y;
                        if (distex * distex + distey * distey < 100 * scaleFactors[kp2.octave]) continue;
                    }
                    if (check_dist_epipolar_line(kp1, kp2, F12, levelSigma2)) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    nmatches++;
                    if (checkOri) rotHist[rot_bin(kp1.angle, B.keys[bestIdx2].angle)].push_back(idx1);
                }
            }
            ia++; ib++;
        } else if (A.nodeId[ia] < B.nodeId[ib]) ia = bow_lower_bound(A, B.nodeId[ib]);
        else ib = bow_lower_bound(B, A.nodeId[ia]);
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1, sizes[HISTO_LENGTH];
        for (int i = 0; i < HISTO_LENGTH; i++) sizes[i] = (int)rotHist[i].size();
        compute_three_maxima(sizes, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { match12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

// MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:345-410, over the descriptors of one map point's good
// observations (n >= 1): index of the descriptor with the least median distance to the rest (first wins).
static int distinctive_descriptor(const uint8_t* desc, int n) {
    std::vector<int> D((size_t)n * n, 0);
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++) D[(size_t)i * n + j] = D[(size_t)j * n + i] = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
    int bestMedian = INT_MAX, bestIdx = 0;
    for (int i = 0; i < n; i++) {
        std::vector<int> v(D.begin() + (size_t)i * n, D.begin() + (size_t)(i + 1) * n);
        std::sort(v.begin(), v.end());
        const int median = v[(size_t)(0.5 * (n - 1))];
        if (median < bestMedian) { bestMedian = median; bestIdx = i; }
    }
    return bestIdx;
}

}  // namespace orc

using namespace orc;
extern "C" {

void* orc_frame_create(const void* keysUn, const uint8_t* desc, const float* uRight, int n,
                       float minX, float maxX, float minY, float maxY) {
    FrameV* F = new FrameV;
    F->n = n;
    F->k.assign((const KeyPt*)keysUn, (const KeyPt*)keysUn + n);
    F->d.assign(desc, desc + (size_t)n * 32);
    if (uRight) F->uR.assign(uRight, uRight + n); else F->uR.assign(n, -1.0f);
    F->minX = minX; F->maxX = maxX; F->minY = minY; F->maxY = maxY;
    F->invW = (float)GRID_COLS / (maxX - minX);
    F->invH = (float)GRID_ROWS / (maxY - minY);
    F->assign_features_to_grid();
    return F;
}
void orc_frame_destroy(void* f) { delete (FrameV*)f; }

// CSR of the grid in mGrid[ix][iy] order: cellStart has 64*48+1 entries.
void orc_frame_grid(const void* f, int* cellStart, int* cellIdx) {
    const FrameV* F = (const FrameV*)f;
    int pos = 0;
    for (int c = 0; c < GRID_COLS * GRID_ROWS; c++) {
        cellStart[c] = pos;
        for (size_t j = 0; j < F->grid[c].size(); j++) cellIdx[pos++] = (int)F->grid[c][j];
    }
    cellStart[GRID_COLS * GRID_ROWS] = pos;
}

int orc_features_in_area(const void* f, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    std::vector<size_t> v = ((const FrameV*)f)->features_in_area(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

void orc_compute_three_maxima(const int* histoSize, int L, int* ind) {
    ind[0] = ind[1] = ind[2] = -1;
    compute_three_maxima(histoSize, L, ind[0], ind[1], ind[2]);
}

int orc_search_by_projection_map(const void* f, const float* scale, int nMP, const uint8_t* inView, const float* projX,
                                 const float* projY, const float* projXR, const int* level, const float* viewCos,
                                 const uint8_t* mpDesc, const int* mpObs, float th, float nnratio, int* kpObs, int* kpMatch) {
    return search_by_projection_map(*(const FrameV*)f, scale, nMP, inView, projX, projY, projXR, level, viewCos, mpDesc, mpObs,
                                    th, nnratio, kpObs, kpMatch);
}

int orc_search_by_projection_last(const void* f, const float* scale, const float* cam6, const float* TcwCur,
                                  const float* TcwLast, int nLast, const uint8_t* lastHasPoint, const float* lastPos,
                                  const int* lastOctave, const float* lastAngle, const uint8_t* mpDesc, const int* mpObs,
                                  float th, int bMono, int checkOri, int* kpObs, int* kpMatch) {
    Camera cam{cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
    return search_by_projection_last(*(const FrameV*)f, scale, cam, TcwCur, TcwLast, nLast, lastHasPoint, lastPos, lastOctave,
                                     lastAngle, mpDesc, mpObs, th, bMono != 0, checkOri != 0, kpObs, kpMatch);
}

int orc_search_for_initialization(const void* f1, const void* f2, float* prevMatched, int* vnMatches12, int windowSize,
                                  float nnratio, int checkOri) {
    return search_for_initialization(*(const FrameV*)f1, *(const FrameV*)f2, prevMatched, vnMatches12, windowSize, nnratio,
                                     checkOri != 0);
}

void orc_hamming_knn2(const uint8_t* q, int nq, const uint8_t* db, int nd, int thLow, float nnratio, int* bestIdx,
                      int* bestDist, int* secondDist) {
    hamming_knn2(q, nq, db, nd, thLow, nnratio, bestIdx, bestDist, secondDist);
}

int orc_search_by_projection_keyframe(const void* f, const float* scale, int nLevels, float logScaleFactor, const float* cam6,
                                      const float* Tcw, int nPts, const uint8_t* valid, const float* pos, const float* minDist,
                                      const float* maxDist, const float* maxDistRaw, const float* angle, const uint8_t* desc,
                                      float th, int ORBdist, int checkOri, int* kpTaken, int* kpMatch) {
    Camera cam{cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
    return search_by_projection_keyframe(*(const FrameV*)f, scale, nLevels, logScaleFactor, cam, Tcw, nPts, valid, pos, minDist, maxDist,
                                         maxDistRaw, angle, desc, th, ORBdist, checkOri != 0, kpTaken, kpMatch);
}

int orc_search_by_projection_sim3(const void* f, const float* scale, int nLevels, float logScaleFactor, const float* cam6,
                                  const float* Tcw, int nPts, const uint8_t* valid, const float* pos, const float* minDist,
                                  const float* maxDist, const float* maxDistRaw, const float* normal, const uint8_t* desc, int th,
                                  int* kpTaken, int* kpMatch) {
    Camera cam{cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
    return search_by_projection_sim3(*(const FrameV*)f, scale, nLevels, logScaleFactor, cam, Tcw, nPts, valid, pos, minDist, maxDist,
                                     maxDistRaw, normal, desc, th, kpTaken, kpMatch);
}

void orc_fuse_search(const void* f, const float* scale, const float* invSigma2, int nLevels, float logScaleFactor, const float* cam6,
                     const float* Tcw, const float* Ow, int sim3, int nPts, const uint8_t* valid, const float* pos, const float* minDist,
                     const float* maxDist, const float* maxDistRaw, const float* normal, const uint8_t* desc, float th,
                     int* bestIdx, int* bestDist) {
    Camera cam{cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
    fuse_search(*(const FrameV*)f, scale, invSigma2, nLevels, logScaleFactor, cam, Tcw, Ow, sim3, nPts, valid, pos, minDist, maxDist,
                maxDistRaw, normal, desc, th, bestIdx, bestDist);
}

// ORBmatcher::SearchBySim3, ORBmatcher.cc:1102-1326.  valid1 / valid2 = "map point non-NULL, not bad, not already matched"
// (:1157-1162, :1232-1237 with vbAlreadyMatched of :1131-1143); T1w, T2w = [R | t] of the keyframes; T21 = [sR21 | t21],
// T12 = [sR12 | t12] as :1119-1122 builds them.  match12[i1] = idx2 where both directions agree (:1305-1319), else -1.
int orc_search_by_sim3(const void* f1, const void* f2, const float* scale, int nLevels, float logScaleFactor, const float* cam6,
                       const float* T1w, const float* T2w, const float* T21, const float* T12,
                       int n1, const uint8_t* valid1, const float* pos1, const float* minD1, const float* maxD1, const float* raw1, const uint8_t* desc1,
                       int n2, const uint8_t* valid2, const float* pos2, const float* minD2, const float* maxD2, const float* raw2, const uint8_t* desc2,
                       float th, int* match12) {
    Camera cam{cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
    std::vector<int> vn1(std::max(n1, 1)), vn2(std::max(n2, 1));
    sim3_direction(*(const FrameV*)f2, scale, nLevels, logScaleFactor, cam, T1w, T21, n1, valid1, pos1, minD1, maxD1, raw1, desc1, th, vn1.data());
    sim3_direction(*(const FrameV*)f1, scale, nLevels, logScaleFactor, cam, T2w, T12, n2, valid2, pos2, minD2, maxD2, raw2, desc2, th, vn2.data());
    int nFound = 0;
    for (int i1 = 0; i1 < n1; i1++) {
        match12[i1] = -1;
        const int idx2 = vn1[i1];
        if (idx2 >= 0 && idx2 < n2 && vn2[idx2] == i1) { match12[i1] = idx2; nFound++; }
    }
    return nFound;
}

// Head of Frame::BuildObject2DsRGBD (Frame.cc:240-311, minKeypoints = 5) / BuildObject2DsStereo (:314-385, minKeypoints = 10): the
// pool of (index, keypoint) pairs, the masks in order, erase on acceptance, Object2D numbering.  Windows that leave the image are
// undefined in the reference (cv::Mat::at) and reject here.
int orc_assign_keypoints_to_masks(const void* keysUn, const float* depth, int n, const uint8_t* masks, int nMasks, int w, int h,
                                  float thDepth, int minKeypoints, int* maskOfKp, int* objectKp, int* objectOfMask) {
    const KeyPt* kps = (const KeyPt*)keysUn;
    std::vector<std::pair<int, KeyPt>> vIndex_Kp;
    for (int i = 0; i < n; i++) { vIndex_Kp.push_back(std::make_pair(i, kps[i])); maskOfKp[i] = -1; objectKp[2 * i] = objectKp[2 * i + 1] = -1; }
    int nObjects = 0;
    for (int i = 0; i < nMasks; i++) {
        const uint8_t* mask = masks + (size_t)i * w * h;
        std::vector<int> vFrameKpIndices;
        auto it = vIndex_Kp.begin();
        while (it != vIndex_Kp.end()) {
            const KeyPt kp = it->second;
            const float kp_z = depth[it->first];
            bool semantic_flag = true;
            for (int row = -10; row < 10; row++)
                for (int col = -10; col < 10; col++) {
                    const int iy = (int)(kp.y + row), ix = (int)(kp.x + col);
                    if (iy < 0 || iy >= h || ix < 0 || ix >= w) { semantic_flag = false; continue; }
                    if ((int)mask[(size_t)iy * w + ix] != 255) semantic_flag = false;
                }
            if (semantic_flag && kp_z > 0 && kp_z <= thDepth) {
                vFrameKpIndices.push_back(it->first);
                maskOfKp[it->first] = i;
                it = vIndex_Kp.erase(it);
            } else ++it;
        }
        const int VaildKpNum = (int)vFrameKpIndices.size();
        objectOfMask[i] = -1;
        if (VaildKpNum > minKeypoints) {
            for (int j = 0; j < VaildKpNum; j++) { objectKp[2 * vFrameKpIndices[j]] = nObjects; objectKp[2 * vFrameKpIndices[j] + 1] = j; }
            objectOfMask[i] = nObjects++;
        }
    }
    return nObjects;
}

// Frame::ExtractHSVHistogramsFromMask, Frame.cc:388-414.  OpenCV is un-vendored; its 8-bit BGR->HSV (imgproc color_hsv: RGB2HSV_b,
// hsv_shift = 12, division tables by saturate_cast<int> = round half to even), calcHist's uniform bins (value * size / range for
// 8-bit data) and normalize(NORM_L1) (binary32 scale 1/sum) are restated here and pinned against cv2 4.13 by
// tests/test_bow_matchers.py (all 2^24 colours).  out: 94 floats, V (32) | S (32) | H (30).
void orc_hsv_from_bgr(const uint8_t* bgr, int n, uint8_t* hsv) {
    static int sdiv[256], hdiv[256];
    static bool init = false;
    if (!init) {
        sdiv[0] = hdiv[0] = 0;
        for (int i = 1; i < 256; i++) { sdiv[i] = (int)nearbyint((255 << 12) / (1. * i)); hdiv[i] = (int)nearbyint((180 << 12) / (6. * i)); }
        init = true;
    }
    for (int i = 0; i < n; i++) {
        const int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        int v = std::max(std::max(b, g), r);
        const int vmin = std::min(std::min(b, g), r), diff = v - vmin;
        const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
        const int s = (diff * sdiv[v] + (1 << 11)) >> 12;
        int h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
        h = (h * hdiv[diff] + (1 << 11)) >> 12;
        h += h < 0 ? 180 : 0;
        hsv[3 * i] = (uint8_t)h; hsv[3 * i + 1] = (uint8_t)s; hsv[3 * i + 2] = (uint8_t)v;
    }
}
void orc_hsv_histogram(const uint8_t* bgr, const uint8_t* mask, int w, int h, float* out) {
    std::vector<uint8_t> hsv((size_t)w * h * 3);
    orc_hsv_from_bgr(bgr, w * h, hsv.data());
    int cnt[94] = {0};
    for (int i = 0; i < w * h; i++) {
        if (!mask[i]) continue;
        cnt[hsv[3 * i + 2] * 32 / 256]++;
        cnt[32 + hsv[3 * i + 1] * 32 / 256]++;
        if (hsv[3 * i] < 180) cnt[64 + hsv[3 * i] * 30 / 180]++;
    }
    double sum = 0;
    for (int i = 0; i < 94; i++) sum += cnt[i];
    const float scale = sum > 0 ? (float)(1.0 / sum) : 0.0f;
    for (int i = 0; i < 94; i++) out[i] = (float)cnt[i] * scale;
}

// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) (OpenCV calib3d / imgproc undistort.dispatch.cpp, cvUndistortPointsInternal,
// un-vendored): binary64 scalar loop, TermCriteria(COUNT, 5, 0.01), identity tilt and rectification, P = K.  Pinned against cv2 4.13.
void orc_undistort_points(const float* pts, int n, float fxf, float fyf, float cxf, float cyf, const float* dist, int nDist, float* out) {
    const double fx = fxf, fy = fyf, cx = cxf, cy = cyf;
    double k[14] = {0};
    for (int i = 0; i < nDist && i < 14; i++) k[i] = dist[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; i++) {
        const double u = pts[2 * i], v = pts[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; j++) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        out[2 * i] = (float)(fx * x + cx);
        out[2 * i + 1] = (float)(fy * y + cy);
    }
}

// cv::distanceTransform(~mask, dst, DIST_L2, DIST_MASK_PRECISE), ObjectTypes.cc:23: OpenCV's trueDistTrans (imgproc/distransform.cpp,
// un-vendored) restated: column pass (vertical distance to the nearest zero of ~mask, squared, 2^32 beyond reach), row pass (lower
// envelope of parabolas with binary32 intersections, sqrt).  Pinned against cv2 4.13 with IPP disabled.
void orc_distance_transform(const uint8_t* mask, int w, int h, float* out) {
    const int m = h, n = w;
    const float inf = 4294967296.0f;
    std::vector<int> d(m);
    for (int x = 0; x < n; x++) {
        int dist = m - 1;
        for (int j = m - 1; j >= 0; j--) { dist = (dist + 1) & (mask[(size_t)j * n + x] == 255 ? 0 : -1); d[j] = dist; }
        dist = m - 1;
        for (int j = 0; j < m; j++) {
            dist = std::min(dist + 1, d[j]);
            out[(size_t)j * n + x] = dist < m ? (float)(dist * dist) : inf;
        }
    }
    std::vector<float> f(n), z(n + 1), sqr(n), inv(n);
    std::vector<int> v(n);
    sqr[0] = inv[0] = 0.f;
    for (int i = 1; i < n; i++) { inv[i] = (float)(0.5 / i); sqr[i] = (float)(i * i); }
    for (int y = 0; y < m; y++) {
        float* dr = out + (size_t)y * n;
        int k = 0;
        v[0] = 0; z[0] = -inf; z[1] = inf; f[0] = dr[0];
        for (int q = 1; q < n; q++) {
            const float fq = dr[q];
            f[q] = fq;
            for (;; k--) {
                const int p = v[k];
                const float s = (fq + sqr[q] - dr[p] - sqr[p]) * inv[q - p];
                if (s > z[k]) { k++; v[k] = q; z[k] = s; z[k + 1] = inf; break; }
            }
        }
        for (int q = 0, kk = 0; q < n; q++) {
            while (z[kk + 1] < q) kk++;
            const int p = v[kk];
            dr[q] = std::sqrt(sqr[std::abs(q - p)] + f[p]);
        }
    }
}

float orc_logf(float x) { return logf(x); }
float orc_norm3(const float* v) { return norm3(v); }
int orc_predict_scale(float maxDistRaw, float dist, float logScaleFactor, int nLevels) { return predict_scale(maxDistRaw, dist, logScaleFactor, nLevels); }

void orc_project_points(const float* Tcw, const float* pos, int n, float* out) {
    for (int i = 0; i < n; i++) mat_rx_plus_t(Tcw, pos + 3 * i, out + 3 * i);
}
void orc_minus_rt_t(const float* Tcw, float* out) { mat_minus_rt_t(Tcw, out); }

int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

int orc_stereo_match(const void* kL, const uint8_t* dL, int nL, const void* kR, const uint8_t* dR, int nR,
                     const uint8_t* const* pyrL, const uint8_t* const* pyrR, const int* lw, const int* lh, int nlevels,
                     const float* scale, const float* invScale, float mbf, float minD, float maxD,
                     float* uRight, float* depth, int* sadOut) {
    stereo_match((const KeyPt*)kL, dL, nL, (const KeyPt*)kR, dR, nR, pyrL, pyrR, lw, lh, nlevels, scale, invScale,
                 mbf, minD, maxD, uRight, depth, sadOut);
    return 0;
}

static BowSide bow_side(const int* hdr, const uint8_t* desc, const void* keys, const uint8_t* valid, const float* uRight,
                        const uint32_t* nodeId, const int* nodeStart, const int* nodeIdx) {
    BowSide S;
    S.n = hdr[0]; S.nNodes = hdr[1];
    S.desc = desc; S.keys = (const KeyPt*)keys; S.valid = valid; S.uRight = uRight;
    S.nodeId = nodeId; S.nodeStart = nodeStart; S.nodeIdx = nodeIdx;
    return S;
}
// hdr = {n, nNodes}
int orc_search_by_bow(const int* hdrA, const uint8_t* descA, const void* keysA, const uint8_t* validA,
                      const uint32_t* nodeIdA, const int* nodeStartA, const int* nodeIdxA,
                      const int* hdrB, const uint8_t* descB, const void* keysB, const uint8_t* validB,
                      const uint32_t* nodeIdB, const int* nodeStartB, const int* nodeIdxB,
                      int thLow, int strictLow, float nnratio, int checkOri, int* match12, int* match21) {
    return search_by_bow(bow_side(hdrA, descA, keysA, validA, nullptr, nodeIdA, nodeStartA, nodeIdxA),
                         bow_side(hdrB, descB, keysB, validB, nullptr, nodeIdB, nodeStartB, nodeIdxB),
                         thLow, strictLow, nnratio, checkOri, match12, match21);
}
int orc_search_for_triangulation(const int* hdrA, const uint8_t* descA, const void* keysA, const uint8_t* validA, const float* uRightA,
                                 const uint32_t* nodeIdA, const int* nodeStartA, const int* nodeIdxA,
                                 const int* hdrB, const uint8_t* descB, const void* keysB, const uint8_t* validB, const float* uRightB,
                                 const uint32_t* nodeIdB, const int* nodeStartB, const int* nodeIdxB,
                                 const float* F12, float ex, float ey, const float* levelSigma2, const float* scaleFactors,
                                 int onlyStereo, int checkOri, int* match12) {
    return search_for_triangulation(bow_side(hdrA, descA, keysA, validA, uRightA, nodeIdA, nodeStartA, nodeIdxA),
                                    bow_side(hdrB, descB, keysB, validB, uRightB, nodeIdB, nodeStartB, nodeIdxB),
                                    F12, ex, ey, levelSigma2, scaleFactors, onlyStereo, checkOri, match12);
}
// start: nPoints + 1 offsets into desc (descriptors of point p: [start[p], start[p+1])); best[p] = -1 for an empty list
void orc_distinctive_descriptors(const uint8_t* desc, const int* start, int nPoints, int* best) {
    for (int p = 0; p < nPoints; p++) {
        const int n = start[p + 1] - start[p];
        best[p] = n > 0 ? distinctive_descriptor(desc + (size_t)start[p] * 32, n) : -1;
    }
}

}  // extern "C"
