// TEST INFRASTRUCTURE ONLY.  C entry points around the reference's own, UNMODIFIED src/ORBmatcher.cc, src/Frame.cc,
// src/MapPoint.cc and src/KeyFrame.cc (compiled in place from /root/reference against oracle/refstubs by oracle/Makefile; output
// oracle/_ref/libref_matcher.so, git-ignored).  This file builds the reference's own objects (Frame, KeyFrame, MapPoint) from flat
// arrays, calls the reference's methods and flattens what they wrote.  It contains no matching logic.
//
// Link-level mocks (no arithmetic of the compared path lives in any of them): Map::Map, ORBextractor::ORBextractor (the extractor
// object only carries mvImagePyramid here), Map::EraseMapPoint (bookkeeping of Fuse / Replace), and aborting bodies for the
// functions of files that are not compiled (Converter, KeyFrameDatabase, Object2D, the image-level OpenCV calls of the object layer).
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>

#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "Map.h"
#include "ORBmatcher.h"
#include "ORBextractor.h"
#include "Converter.h"
#include "KeyFrameDatabase.h"
#include "ObjectTypes.h"
#undef private
#undef protected

using namespace ORB_SLAM2;

// ------------------------------------------------------------------------------------------------ mocks
static void not_on_path(const char* what) {
    fprintf(stderr, "ref_matcher_harness: %s was reached; it is not part of the compared path\n", what);
    abort();
}
namespace ORB_SLAM2 {
Map::Map() : mnMaxKFid(0), mnBigChangeIdx(0) {}
void Map::EraseMapPoint(MapPoint*) {}                    // set bookkeeping only (src/Map.cc): Fuse / Replace reach it
void Map::EraseKeyFrame(KeyFrame*) { not_on_path("Map::EraseKeyFrame"); }
ORBextractor::ORBextractor(int, float, int, int, int) {}
void ORBextractor::operator()(cv::InputArray, cv::InputArray, std::vector<cv::KeyPoint>&, cv::OutputArray) { not_on_path("ORBextractor::operator()"); }
void KeyFrameDatabase::erase(KeyFrame*) { not_on_path("KeyFrameDatabase::erase"); }
cv::Mat Converter::toCvMat(const Eigen::Matrix<float, 3, 1>&) { not_on_path("Converter::toCvMat"); return cv::Mat(); }
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat&) { not_on_path("Converter::toDescriptorVector"); return std::vector<cv::Mat>(); }
Eigen::Matrix<float, 3, 3> Converter::toEigenMat3(const cv::Mat&) { not_on_path("Converter::toEigenMat3"); return Eigen::Matrix<float, 3, 3>(); }
Eigen::Vector3f Converter::toEigenVec3(const cv::Mat&) { not_on_path("Converter::toEigenVec3"); return Eigen::Vector3f(); }
Object2D::Object2D(semantic, std::vector<cv::KeyPoint>, std::vector<cv::Mat>, std::vector<float>, std::vector<int>, cv::Mat&, ORBVocabulary*) { not_on_path("Object2D::Object2D"); }
}  // namespace ORB_SLAM2
namespace cv {
void cvtColor(const Mat&, Mat&, int) { not_on_path("cv::cvtColor"); }
void calcHist(const Mat*, int, const int*, const Mat&, Mat&, int, const int*, const float**, bool, bool) { not_on_path("cv::calcHist"); }
void hconcat(const Mat&, const Mat&, Mat&) { not_on_path("cv::hconcat"); }
void normalize(const Mat&, Mat&, int) { not_on_path("cv::normalize"); }
void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) { not_on_path("cv::undistortPoints"); }
}  // namespace cv

// ------------------------------------------------------------------------------------------------ flat descriptions
extern "C" {
struct RefFrame {
    const void* keys_un;          // n x 28 bytes (cv::KeyPoint)
    const uint8_t* desc;          // n x 32
    const float* u_right;         // n, or NULL (all -1)
    int32_t n;
    float min_x, max_x, min_y, max_y;
    float fx, fy, cx, cy, mbf, mb;
    const float* scale;           // nlevels (mvScaleFactors)
    int32_t nlevels;
    float log_scale_factor;       // mfLogScaleFactor
    const float* tcw;             // 12 floats [R | t], or NULL (identity)
    int32_t n_nodes;              // DBoW2::FeatureVector as a CSR (node ids ascending), 0 = none
    const uint32_t* node_id; const int32_t* node_start; const int32_t* node_idx;
};
struct RefPoints {                // one entry per map point (or per keypoint slot of a keyframe)
    int32_t n;
    const uint8_t* valid;         // 0 -> NULL pointer in the reference's vector
    const float* pos;             // n x 3
    const float* min_dist_raw;    // mfMinDistance
    const float* max_dist_raw;    // mfMaxDistance
    const float* normal;          // n x 3, or NULL
    const uint8_t* desc;          // n x 32
    const int32_t* obs;           // nObs, or NULL (1)
};
}

namespace {

std::mutex g_lock;                // Frame keeps its camera and grid constants in static members

void apply_statics(const RefFrame& s) {
    Frame::fx = s.fx; Frame::fy = s.fy; Frame::cx = s.cx; Frame::cy = s.cy;
    Frame::invfx = 1.0f / s.fx; Frame::invfy = 1.0f / s.fy;
    Frame::mnMinX = s.min_x; Frame::mnMaxX = s.max_x; Frame::mnMinY = s.min_y; Frame::mnMaxY = s.max_y;
    // Frame.cc:96-97 / :157-158 / :217-218
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
}

cv::Mat mat44(const float* t12) {
    cv::Mat T(4, 4, CV_32F);
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) T.at<float>(r, c) = r == c ? 1.f : 0.f;
    if (t12)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) T.at<float>(r, c) = t12[4 * r + c];
    return T;
}

cv::Mat desc_mat(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) memcpy(m.data, d, (size_t)n * 32);
    if (n == 0) m = m.rowRange(0, 0);
    return m;
}

// Fills the public members the Frame constructors would set (Frame.cc:58-227), then runs the reference's own
// AssignFeaturesToGrid and SetPose.
void build_frame(Frame& F, const RefFrame& s) {
    apply_statics(s);
    const int n = s.n;
    const cv::KeyPoint* k = reinterpret_cast<const cv::KeyPoint*>(s.keys_un);
    F.N = n;
    F.mvKeysUn.assign(k, k + n);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = desc_mat(s.desc, n);
    F.mvuRight = s.u_right ? std::vector<float>(s.u_right, s.u_right + n) : std::vector<float>(n, -1.0f);
    F.mvDepth = std::vector<float>(n, -1.0f);
    F.mvpMapPoints = std::vector<MapPoint*>(n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier = std::vector<bool>(n, false);
    F.mbf = s.mbf; F.mb = s.mb; F.mThDepth = 0.f;
    F.mnScaleLevels = s.nlevels;
    F.mvScaleFactors.assign(s.scale, s.scale + s.nlevels);
    F.mfScaleFactor = s.nlevels > 1 ? s.scale[1] : 1.0f;
    F.mfLogScaleFactor = s.log_scale_factor;
    F.mvInvScaleFactors.resize(s.nlevels); F.mvLevelSigma2.resize(s.nlevels); F.mvInvLevelSigma2.resize(s.nlevels);
    for (int i = 0; i < s.nlevels; i++) {             // ORBextractor.cc:415-431
        F.mvLevelSigma2[i] = s.scale[i] * s.scale[i];
        F.mvInvScaleFactors[i] = 1.0f / s.scale[i];
        F.mvInvLevelSigma2[i] = 1.0f / F.mvLevelSigma2[i];
    }
    F.mK = cv::Mat(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = s.fx; F.mK.at<float>(1, 1) = s.fy; F.mK.at<float>(0, 2) = s.cx; F.mK.at<float>(1, 2) = s.cy; F.mK.at<float>(2, 2) = 1.f;
    F.mDistCoef = cv::Mat(4, 1, CV_32F);
    F.mpORBvocabulary = NULL; F.mpORBextractorLeft = F.mpORBextractorRight = NULL; F.mpReferenceKF = NULL;
    F.mnId = Frame::nNextId++;
    F.mTimeStamp = 0; F.N_O = 0;
    F.mFeatVec.clear(); F.mBowVec.clear();
    for (int i = 0; i < s.n_nodes; i++) {
        std::vector<unsigned int>& v = F.mFeatVec[s.node_id[i]];
        for (int j = s.node_start[i]; j < s.node_start[i + 1]; j++) v.push_back((unsigned int)s.node_idx[j]);
    }
    for (int i = 0; i < FRAME_GRID_COLS; i++)
        for (int j = 0; j < FRAME_GRID_ROWS; j++) F.mGrid[i][j].clear();
    F.AssignFeaturesToGrid();
    F.SetPose(mat44(s.tcw));
}

struct Scene {                       // owns what a call created
    Map map;
    std::vector<Frame*> frames;
    std::vector<KeyFrame*> kfs;
    std::vector<MapPoint*> mps;
    Frame dummyFrame;
    KeyFrame* dummyKF = NULL;
    ~Scene() {
        for (MapPoint* p : mps) delete p;
        for (KeyFrame* k : kfs) delete k;
        for (Frame* f : frames) delete f;
    }
    Frame* frame(const RefFrame& s) { Frame* f = new Frame(); build_frame(*f, s); frames.push_back(f); return f; }
    KeyFrame* keyframe(const RefFrame& s) {
        Frame* f = frame(s);
        KeyFrame* k = new KeyFrame(*f, &map, NULL);
        kfs.push_back(k);
        return k;
    }
    KeyFrame* ref_kf(const RefFrame& like) {              // reference keyframe the MapPoint constructor reads two ids from
        if (!dummyKF) {
            RefFrame s = like;
            s.n = 0; s.n_nodes = 0; s.tcw = NULL;
            build_frame(dummyFrame, s);
            dummyKF = new KeyFrame(dummyFrame, &map, NULL);
            kfs.push_back(dummyKF);
        }
        return dummyKF;
    }
    MapPoint* point(const RefFrame& like, const float* pos, const uint8_t* desc, int nObs) {
        cv::Mat P(3, 1, CV_32F);
        if (pos) for (int i = 0; i < 3; i++) P.at<float>(i) = pos[i];
        MapPoint* p = new MapPoint(P, ref_kf(like), &map);
        p->nObs = nObs;
        if (desc) { cv::Mat d(1, 32, CV_8U); memcpy(d.data, desc, 32); p->mDescriptor = d; }
        mps.push_back(p);
        return p;
    }
    // vector<MapPoint*> of a RefPoints block (NULL where valid == 0)
    std::vector<MapPoint*> points(const RefFrame& like, const RefPoints& P) {
        std::vector<MapPoint*> v(P.n, static_cast<MapPoint*>(NULL));
        for (int i = 0; i < P.n; i++) {
            if (P.valid && !P.valid[i]) continue;
            MapPoint* p = point(like, P.pos ? P.pos + 3 * i : NULL, P.desc ? P.desc + 32 * (size_t)i : NULL, P.obs ? P.obs[i] : 1);
            if (P.min_dist_raw) p->mfMinDistance = P.min_dist_raw[i];
            if (P.max_dist_raw) p->mfMaxDistance = P.max_dist_raw[i];
            if (P.normal) { cv::Mat nrm(3, 1, CV_32F); for (int k = 0; k < 3; k++) nrm.at<float>(k) = P.normal[3 * i + k]; p->mNormalVector = nrm; }
            v[i] = p;
        }
        return v;
    }
};

std::map<MapPoint*, int> index_of(const std::vector<MapPoint*>& v) {
    std::map<MapPoint*, int> m;
    for (size_t i = 0; i < v.size(); i++) if (v[i]) m[v[i]] = (int)i;
    return m;
}

}  // namespace

extern "C" {
#define EXPORT __attribute__((visibility("default")))

EXPORT int refm_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    cv::Mat A(1, 32, CV_8U, (void*)a), B(1, 32, CV_8U, (void*)b);
    return ORBmatcher::DescriptorDistance(A, B);
}

EXPORT void refm_compute_three_maxima(const int* sizes, int L, int* ind) {
    std::vector<std::vector<int> > histo(L);
    for (int i = 0; i < L; i++) histo[i].assign(sizes[i], 0);
    ORBmatcher m;
    ind[0] = ind[1] = ind[2] = -1;            // every caller initialises them so (e.g. ORBmatcher.cc:112-114): the function may leave them
    m.ComputeThreeMaxima(histo.data(), L, ind[0], ind[1], ind[2]);
}

// Frame::AssignFeaturesToGrid (Frame.cc:455-470): cell = ix * 48 + iy, CSR over the 64 x 48 cells
EXPORT void refm_frame_grid(const RefFrame* s, int* cellStart, int* cellIdx) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* F = sc.frame(*s);
    int o = 0;
    for (int i = 0; i < FRAME_GRID_COLS; i++)
        for (int j = 0; j < FRAME_GRID_ROWS; j++) {
            cellStart[i * FRAME_GRID_ROWS + j] = o;
            for (size_t k : F->mGrid[i][j]) cellIdx[o++] = (int)k;
        }
    cellStart[FRAME_GRID_COLS * FRAME_GRID_ROWS] = o;
}

EXPORT int refm_features_in_area(const RefFrame* s, float x, float y, float r, int minLevel, int maxLevel, int keyframe, int* out, int cap) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    std::vector<size_t> v;
    if (keyframe) v = sc.keyframe(*s)->GetFeaturesInArea(x, y, r);
    else v = sc.frame(*s)->GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = (int)v[i];
    return (int)v.size();
}

// Frame::ComputeStereoMatches (Frame.cc:706-880).  The reference fixes minD = 0 and derives maxD = mbf / mb from its members;
// *maxD_used returns that value.  levels: border-less level images of both eyes.
EXPORT int refm_stereo_match(const RefFrame* L, const void* keysR, const uint8_t* descR, int nR,
                             const uint8_t* const* levelsL, const uint8_t* const* levelsR, const int* lw, const int* lh,
                             float* uRight, float* depth, float* maxD_used) {
    std::lock_guard<std::mutex> lk(g_lock);
    if (maxD_used) *maxD_used = L->mbf / L->mb;
    if (L->n == 0) return 0;                 // :867 indexes an empty vector there
    Scene sc;
    Frame* F = sc.frame(*L);
    const cv::KeyPoint* kr = reinterpret_cast<const cv::KeyPoint*>(keysR);
    F->mvKeysRight.assign(kr, kr + nR);
    F->mDescriptorsRight = desc_mat(descR, nR);
    ORBextractor exL(0, 0, 0, 0, 0), exR(0, 0, 0, 0, 0);
    exL.mvImagePyramid.resize(L->nlevels); exR.mvImagePyramid.resize(L->nlevels);
    for (int l = 0; l < L->nlevels; l++) {
        exL.mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8U, (void*)levelsL[l], (size_t)lw[l]);
        exR.mvImagePyramid[l] = cv::Mat(lh[l], lw[l], CV_8U, (void*)levelsR[l], (size_t)lw[l]);
    }
    F->mpORBextractorLeft = &exL; F->mpORBextractorRight = &exR;
    F->ComputeStereoMatches();
    for (int i = 0; i < L->n; i++) { uRight[i] = F->mvuRight[i]; depth[i] = F->mvDepth[i]; }
    return L->n;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), ORBmatcher.cc:45-129
EXPORT int refm_search_by_projection_map(const RefFrame* s, int nMP, const uint8_t* inView, const float* projX, const float* projY,
                                         const float* projXR, const int* level, const float* viewCos, const uint8_t* mpDesc,
                                         const int* mpObs, float th, float nnratio, const int* kpObs, int* kpMatch) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* F = sc.frame(*s);
    std::vector<MapPoint*> blockers(s->n, static_cast<MapPoint*>(NULL));
    for (int k = 0; k < s->n; k++)
        if (kpObs && kpObs[k] > 0) F->mvpMapPoints[k] = blockers[k] = sc.point(*s, NULL, NULL, kpObs[k]);
    std::vector<MapPoint*> v(nMP);
    for (int i = 0; i < nMP; i++) {
        MapPoint* p = sc.point(*s, NULL, mpDesc + 32 * (size_t)i, mpObs[i]);
        p->mbTrackInView = inView[i] != 0;
        p->mTrackProjX = projX[i]; p->mTrackProjY = projY[i]; p->mTrackProjXR = projXR[i];
        p->mnTrackScaleLevel = level[i]; p->mTrackViewCos = viewCos[i];
        v[i] = p;
    }
    ORBmatcher m(nnratio, true);
    const int n = m.SearchByProjection(*F, v, th);
    std::map<MapPoint*, int> idx = index_of(v);
    for (int k = 0; k < s->n; k++) {
        MapPoint* p = F->mvpMapPoints[k];
        kpMatch[k] = (p && p != blockers[k]) ? idx[p] : -1;
    }
    return n;
}

// ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono), ORBmatcher.cc:1328-1470
EXPORT int refm_search_by_projection_last(const RefFrame* cur, const float* tcwLast, int nLast, const uint8_t* hasPoint, const float* pos,
                                          const int* octave, const float* angle, const uint8_t* mpDesc, const int* mpObs,
                                          float th, int mono, int checkOri, const int* kpObs, int* kpMatch) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* C = sc.frame(*cur);
    std::vector<MapPoint*> blockers(cur->n, static_cast<MapPoint*>(NULL));
    for (int k = 0; k < cur->n; k++)
        if (kpObs && kpObs[k] > 0) C->mvpMapPoints[k] = blockers[k] = sc.point(*cur, NULL, NULL, kpObs[k]);
    std::vector<cv::KeyPoint> lk2(nLast);
    for (int i = 0; i < nLast; i++) { lk2[i].octave = octave[i]; lk2[i].angle = angle[i]; lk2[i].pt.x = cur->min_x; lk2[i].pt.y = cur->min_y; }
    std::vector<uint8_t> noDesc((size_t)std::max(nLast, 1) * 32, 0);
    RefFrame ls = *cur;
    ls.keys_un = lk2.data(); ls.desc = noDesc.data(); ls.u_right = NULL; ls.n = nLast; ls.tcw = tcwLast; ls.n_nodes = 0;
    Frame* Lf = sc.frame(ls);
    apply_statics(*cur);
    std::vector<MapPoint*> v(nLast, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < nLast; i++) {
        if (!hasPoint[i]) continue;
        v[i] = sc.point(*cur, pos + 3 * i, mpDesc + 32 * (size_t)i, mpObs[i]);
        Lf->mvpMapPoints[i] = v[i];
    }
    ORBmatcher m(0.9f, checkOri != 0);
    const int n = m.SearchByProjection(*C, *Lf, th, mono != 0);
    std::map<MapPoint*, int> idx = index_of(v);
    for (int k = 0; k < cur->n; k++) {
        MapPoint* p = C->mvpMapPoints[k];
        kpMatch[k] = (p && p != blockers[k]) ? idx[p] : -1;
    }
    return n;
}

// ORBmatcher::SearchForInitialization, ORBmatcher.cc:405-520
EXPORT int refm_search_for_initialization(const RefFrame* f1, const RefFrame* f2, float* prevMatched, int* matches12, int window,
                                          float nnratio, int checkOri) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* F1 = sc.frame(*f1);
    Frame* F2 = sc.frame(*f2);
    std::vector<cv::Point2f> pm(f1->n);
    for (int i = 0; i < f1->n; i++) { pm[i].x = prevMatched[2 * i]; pm[i].y = prevMatched[2 * i + 1]; }
    std::vector<int> m12;
    ORBmatcher m(nnratio, checkOri != 0);
    const int n = m.SearchForInitialization(*F1, *F2, pm, m12, window);
    for (int i = 0; i < f1->n; i++) { matches12[i] = m12[i]; prevMatched[2 * i] = pm[i].x; prevMatched[2 * i + 1] = pm[i].y; }
    return n;
}

// ORBmatcher::SearchByProjection(Frame& Current, KeyFrame*, sAlreadyFound, th, ORBdist), ORBmatcher.cc:1472-1599.
// The keyframe has one keypoint slot per entry of P (angle[i] = its mvKeysUn[i].angle); kpTaken marks mvpMapPoints of the frame.
EXPORT int refm_search_by_projection_keyframe(const RefFrame* cur, const RefPoints* P, const float* angle, float th, int ORBdist,
                                              int checkOri, const int* kpTaken, int* kpMatch) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* C = sc.frame(*cur);
    std::vector<MapPoint*> blockers(cur->n, static_cast<MapPoint*>(NULL));
    for (int k = 0; k < cur->n; k++)
        if (kpTaken && kpTaken[k]) C->mvpMapPoints[k] = blockers[k] = sc.point(*cur, NULL, NULL, 1);
    std::vector<cv::KeyPoint> kk(P->n);
    for (int i = 0; i < P->n; i++) { kk[i].angle = angle ? angle[i] : 0.f; kk[i].pt.x = cur->min_x; kk[i].pt.y = cur->min_y; }
    std::vector<uint8_t> noDesc((size_t)std::max(P->n, 1) * 32, 0);
    RefFrame ks = *cur;
    ks.keys_un = kk.data(); ks.desc = noDesc.data(); ks.u_right = NULL; ks.n = P->n; ks.tcw = NULL; ks.n_nodes = 0;
    KeyFrame* KF = sc.keyframe(ks);
    apply_statics(*cur);
    std::vector<MapPoint*> v = sc.points(*cur, *P);
    for (int i = 0; i < P->n; i++) KF->mvpMapPoints[i] = v[i];
    ORBmatcher m(0.9f, checkOri != 0);
    std::set<MapPoint*> found;
    const int n = m.SearchByProjection(*C, KF, found, th, ORBdist);
    std::map<MapPoint*, int> idx = index_of(v);
    for (int k = 0; k < cur->n; k++) {
        MapPoint* p = C->mvpMapPoints[k];
        kpMatch[k] = (p && p != blockers[k]) ? idx[p] : -1;
    }
    return n;
}

// The decomposition of Scw at ORBmatcher.cc:299-303 / :986-990 through the same matrix expressions, for callers that hold a
// restatement which starts behind it: out = [Rcw | tcw] (12 floats) and Ow (3 floats).
EXPORT void refm_decompose_scw(const float* scw12, float* rt12, float* ow3) {
    cv::Mat Scw = mat44(scw12);
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) rt12[4 * r + c] = Rcw.at<float>(r, c);
        rt12[4 * r + 3] = tcw.at<float>(r);
        ow3[r] = Ow.at<float>(r);
    }
}

// ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), ORBmatcher.cc:290-403
EXPORT int refm_search_by_projection_sim3(const RefFrame* kf, const float* scw12, const RefPoints* P, int th, const int* kpTaken, int* kpMatch) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    KeyFrame* KF = sc.keyframe(*kf);
    std::vector<MapPoint*> vpMatched(kf->n, static_cast<MapPoint*>(NULL)), blockers(kf->n, static_cast<MapPoint*>(NULL));
    for (int k = 0; k < kf->n; k++)
        if (kpTaken && kpTaken[k]) vpMatched[k] = blockers[k] = sc.point(*kf, NULL, NULL, 1);
    std::vector<MapPoint*> v = sc.points(*kf, *P);
    std::vector<MapPoint*> list;
    for (MapPoint* p : v) if (p) list.push_back(p);          // the reference's vector holds no NULLs (LoopClosing.cc:376)
    ORBmatcher m(0.75f, true);
    const int n = m.SearchByProjection(KF, mat44(scw12), list, vpMatched, th);
    std::map<MapPoint*, int> idx = index_of(v);
    for (int k = 0; k < kf->n; k++) {
        MapPoint* p = vpMatched[k];
        kpMatch[k] = (p && p != blockers[k]) ? idx[p] : -1;
    }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...), ORBmatcher.cc:159-288 (keyframePair == 0: match21[iF] = keypoint of the keyframe)
// and SearchByBoW(KeyFrame*, KeyFrame*, ...), :522-655 (keyframePair != 0: match12[i1] = keypoint of keyframe 2).  valid = map
// point present and not bad.
EXPORT int refm_search_by_bow(const RefFrame* a, const uint8_t* validA, const RefFrame* b, const uint8_t* validB, int keyframePair,
                              float nnratio, int checkOri, int* match12, int* match21) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    KeyFrame* K1 = sc.keyframe(*a);
    std::vector<MapPoint*> v1(a->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < a->n; i++) if (!validA || validA[i]) K1->mvpMapPoints[i] = v1[i] = sc.point(*a, NULL, NULL, 1);
    std::map<MapPoint*, int> idx1 = index_of(v1);
    ORBmatcher m(nnratio, checkOri != 0);
    for (int i = 0; i < a->n; i++) match12[i] = -1;
    for (int i = 0; i < b->n; i++) match21[i] = -1;
    if (!keyframePair) {
        Frame* F = sc.frame(*b);
        std::vector<MapPoint*> out;
        const int n = m.SearchByBoW(K1, *F, out);
        for (int i = 0; i < b->n; i++) if (out[i]) { match21[i] = idx1[out[i]]; match12[match21[i]] = i; }
        return n;
    }
    KeyFrame* K2 = sc.keyframe(*b);
    std::vector<MapPoint*> v2(b->n, static_cast<MapPoint*>(NULL));
    for (int i = 0; i < b->n; i++) if (!validB || validB[i]) K2->mvpMapPoints[i] = v2[i] = sc.point(*b, NULL, NULL, 1);
    std::map<MapPoint*, int> idx2 = index_of(v2);
    std::vector<MapPoint*> out;
    const int n = m.SearchByBoW(K1, K2, out);
    for (int i = 0; i < a->n; i++) if (out[i]) { match12[i] = idx2[out[i]]; match21[match12[i]] = i; }
    return n;
}

// ORBmatcher::SearchForTriangulation, ORBmatcher.cc:657-823 (+ CheckDistEpipolarLine :139-156).  hasPoint marks keypoints that
// already carry a map point (skipped).  epipole returns (ex, ey) as the reference computes it at :664-671.
EXPORT int refm_search_for_triangulation(const RefFrame* a, const uint8_t* hasPointA, const RefFrame* b, const uint8_t* hasPointB,
                                         const float* F12, int onlyStereo, int checkOri, int* match12, float* epipole) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    KeyFrame* K1 = sc.keyframe(*a);
    KeyFrame* K2 = sc.keyframe(*b);
    for (int i = 0; i < a->n; i++) if (hasPointA && hasPointA[i]) K1->mvpMapPoints[i] = sc.point(*a, NULL, NULL, 1);
    for (int i = 0; i < b->n; i++) if (hasPointB && hasPointB[i]) K2->mvpMapPoints[i] = sc.point(*b, NULL, NULL, 1);
    if (epipole) {
        cv::Mat Cw = K1->GetCameraCenter();
        cv::Mat R2w = K2->GetRotation();
        cv::Mat t2w = K2->GetTranslation();
        cv::Mat C2 = R2w * Cw + t2w;
        const float invz = 1.0f / C2.at<float>(2);
        epipole[0] = K2->fx * C2.at<float>(0) * invz + K2->cx;
        epipole[1] = K2->fy * C2.at<float>(1) * invz + K2->cy;
    }
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; i++) F.at<float>(i / 3, i % 3) = F12[i];
    std::vector<std::pair<size_t, size_t> > pairs;
    ORBmatcher m(0.6f, checkOri != 0);
    const int n = m.SearchForTriangulation(K1, K2, F, pairs, onlyStereo != 0);
    for (int i = 0; i < a->n; i++) match12[i] = -1;
    for (size_t i = 0; i < pairs.size(); i++) match12[pairs[i].first] = (int)pairs[i].second;
    return n;
}

// ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th), ORBmatcher.cc:825-966 (sim3 == 0; the pose is the keyframe's) and
// Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint), :974-1100 (sim3 != 0).  The reference does not expose the keypoint a point
// was fused into, so every point is fused ALONE into a keyframe without map points: bestIdx[i] = the slot it landed in (-1 = none).
// total returns the count of one call over all points at once.
EXPORT void refm_fuse(const RefFrame* kf, const float* scw12, int sim3, const RefPoints* P, float th, int* bestIdx, int* total) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    KeyFrame* KF = sc.keyframe(*kf);
    std::vector<MapPoint*> v = sc.points(*kf, *P);
    ORBmatcher m(0.6f, true);
    for (int i = 0; i < P->n; i++) {
        bestIdx[i] = -1;
        if (!v[i]) continue;
        std::vector<MapPoint*> one(1, v[i]);
        if (sim3) {
            std::vector<MapPoint*> rep(1, static_cast<MapPoint*>(NULL));
            m.Fuse(KF, mat44(scw12), one, th, rep);
        } else {
            m.Fuse(KF, one, th);
        }
        for (int k = 0; k < kf->n; k++)
            if (KF->mvpMapPoints[k] == v[i]) { bestIdx[i] = k; KF->mvpMapPoints[k] = NULL; }
        v[i]->mObservations.clear();
        v[i]->nObs = P->obs ? P->obs[i] : 1;
    }
    if (total) {
        std::vector<MapPoint*> list;
        for (MapPoint* p : v) if (p) list.push_back(p);
        if (sim3) {
            std::vector<MapPoint*> rep(list.size(), static_cast<MapPoint*>(NULL));
            *total = m.Fuse(KF, mat44(scw12), list, th, rep);
        } else {
            *total = m.Fuse(KF, list, th);
        }
    }
}

// ORBmatcher::SearchBySim3, ORBmatcher.cc:1102-1326.  P1 / P2: the map points of the keyframes' keypoint slots; matched12[i1] != 0
// marks entries of vpMatches12 that are already set on entry (they name `alreadyIdx2[i1]`, a slot of keyframe 2 holding a point).
EXPORT int refm_search_by_sim3(const RefFrame* k1, const RefFrame* k2, const RefPoints* P1, const RefPoints* P2, float s12,
                               const float* R12, const float* t12, float th, const int* alreadyIdx2, int* match12) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    KeyFrame* K1 = sc.keyframe(*k1);
    KeyFrame* K2 = sc.keyframe(*k2);
    std::vector<MapPoint*> v1 = sc.points(*k1, *P1), v2 = sc.points(*k2, *P2);
    for (int i = 0; i < k1->n; i++) {
        K1->mvpMapPoints[i] = v1[i];
        if (v1[i]) v1[i]->AddObservation(K1, i);
    }
    for (int i = 0; i < k2->n; i++) {
        K2->mvpMapPoints[i] = v2[i];
        if (v2[i]) v2[i]->AddObservation(K2, i);
    }
    std::vector<MapPoint*> vm(k1->n, static_cast<MapPoint*>(NULL));
    if (alreadyIdx2)
        for (int i = 0; i < k1->n; i++) if (alreadyIdx2[i] >= 0) vm[i] = v2[alreadyIdx2[i]];
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = R12[i];
    for (int i = 0; i < 3; i++) t.at<float>(i) = t12[i];
    ORBmatcher m(0.75f, true);
    const int n = m.SearchBySim3(K1, K2, vm, s12, R, t, th);
    std::map<MapPoint*, int> idx2 = index_of(v2);
    for (int i = 0; i < k1->n; i++) match12[i] = vm[i] ? idx2[vm[i]] : -1;
    return n;
}

// The [sR21 | t21] and [sR12 | t12] of ORBmatcher.cc:1119-1122 through the same matrix expressions.
EXPORT void refm_sim3_transforms(float s12, const float* R12, const float* t12, float* T21, float* T12) {
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = R12[i];
    for (int i = 0; i < 3; i++) t.at<float>(i) = t12[i];
    cv::Mat sR12 = s12 * R;
    cv::Mat sR21 = (1.0 / s12) * R.t();
    cv::Mat t21 = -sR21 * t;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) { T21[4 * r + c] = sR21.at<float>(r, c); T12[4 * r + c] = sR12.at<float>(r, c); }
        T21[4 * r + 3] = t21.at<float>(r);
        T12[4 * r + 3] = t.at<float>(r);
    }
}

// MapPoint::ComputeDistinctiveDescriptors, MapPoint.cc:345-410.  desc / start: CSR of the observations' descriptors per point.
// Observation j of a point lives in keyframe j (the keyframes are created in one block, so the std::map<KeyFrame*, size_t> of the
// reference iterates them in that order); best[p] = index of the chosen descriptor inside the point's list.
EXPORT void refm_distinctive_descriptors(const RefFrame* like, const uint8_t* desc, const int* start, int nPoints, int* best) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    int maxObs = 0;
    for (int p = 0; p < nPoints; p++) maxObs = std::max(maxObs, start[p + 1] - start[p]);
    // keyframe j holds the j-th observation's descriptor of every point in row p; placement-new in one block keeps the keyframes'
    // addresses in the order of j
    std::vector<cv::KeyPoint> keys(std::max(nPoints, 1));
    for (size_t i = 0; i < keys.size(); i++) { keys[i].pt.x = like->min_x; keys[i].pt.y = like->min_y; }
    KeyFrame* block = static_cast<KeyFrame*>(malloc(sizeof(KeyFrame) * (size_t)std::max(maxObs, 1)));
    std::vector<uint8_t> dj((size_t)std::max(nPoints, 1) * 32);
    for (int j = 0; j < maxObs; j++) {
        std::fill(dj.begin(), dj.end(), 0);
        for (int p = 0; p < nPoints; p++)
            if (j < start[p + 1] - start[p]) memcpy(&dj[32 * (size_t)p], desc + 32 * (size_t)(start[p] + j), 32);
        RefFrame s = *like;
        s.keys_un = keys.data(); s.desc = dj.data(); s.u_right = NULL; s.n = nPoints; s.n_nodes = 0; s.tcw = NULL;
        Frame* f = sc.frame(s);
        new (block + j) KeyFrame(*f, &sc.map, NULL);
    }
    for (int p = 0; p < nPoints; p++) {
        const int n = start[p + 1] - start[p];
        best[p] = -1;
        if (n == 0) continue;
        MapPoint* mp = sc.point(*like, NULL, NULL, 0);
        for (int j = 0; j < n; j++) mp->mObservations[block + j] = p;
        mp->ComputeDistinctiveDescriptors();
        cv::Mat d = mp->GetDescriptor();
        for (int j = 0; j < n; j++)
            if (!memcmp(d.ptr(0), desc + 32 * (size_t)(start[p] + j), 32)) { best[p] = j; break; }
    }
    for (int j = 0; j < maxObs; j++) (block + j)->~KeyFrame();
    free(block);
}

// MapPoint::PredictScale(currentDist, Frame*), MapPoint.cc:504-519
EXPORT int refm_predict_scale(const RefFrame* like, float maxDistRaw, float currentDist) {
    std::lock_guard<std::mutex> lk(g_lock);
    Scene sc;
    Frame* F = sc.frame(*like);
    MapPoint* p = sc.point(*like, NULL, NULL, 1);
    p->mfMaxDistance = maxDistRaw;
    return p->PredictScale(currentDist, F);
}

}  // extern "C"
