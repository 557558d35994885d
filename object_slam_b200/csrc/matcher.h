// Device-side argument blocks and launchers of the matcher kernels (matcher.cu).
#pragma once
#include "common.cuh"
#include "host_util.h"

#define OBS_GRID_COLS 64          // FRAME_GRID_COLS, include/Frame.h:41
#define OBS_GRID_ROWS 48          // FRAME_GRID_ROWS, include/Frame.h:42
#define OBS_GRID_CELLS (OBS_GRID_COLS * OBS_GRID_ROWS)
#define OBS_HISTO_LENGTH 30       // src/ORBmatcher.cc:39
#define OBS_CAND_SLOTS 8          // slots of one candidate chunk: 7 entries + link / terminator

struct FrameParamsDev {
    float minX, maxX, minY, maxY, invW, invH;
    float fx, fy, cx, cy, mbf, mb;
    int nlevels;
    float scale[OBS_MAX_LEVELS];
};

// B frames, `cap` keypoint slots each.
struct FrameSetDev {
    int cap;
    int* n;                  // [B]
    float4* kp;              // [B][cap]   x, y, uRight, octave (int bits)
    float* angle;            // [B][cap]
    uint4* desc;             // [B][cap][2]
    int* cellStart;          // [B][OBS_GRID_CELLS + 1]   CSR of mGrid[ix][iy], cell = ix*48 + iy
    uint16_t* cellIdx;       // [B][cap]
    FrameParamsDev P;
};

// Source of a frame-set build: AoS keypoints (28-byte cv::KeyPoint records) + descriptors + uRight.
struct FrameBuildArgs {
    FrameSetDev F;
    const uint8_t* keys;  size_t keysFrameStride;     // bytes between frames
    const uint8_t* desc;  size_t descFrameStride;     // bytes
    const float* uRight;  size_t uRightFrameStride;   // floats; uRight may be null
    const int* count;     size_t countStrideInts;     // keypoints of frame b = count[b * countStrideInts]
};
cudaError_t launch_frame_build(const FrameBuildArgs& a, int nFrames, cudaStream_t st);

struct MapPointDev {          // SearchByProjection(Frame&, vector<MapPoint*>, th)
    int n; size_t stride;     // stride (entries) between frames, 0 = shared
    const uint8_t* inView; const float* projX; const float* projY; const float* projXR;
    const int* level; const float* viewCos; const uint4* desc; const int* obs;
};
struct LastFrameDev {         // SearchByProjection(Frame& Current, const Frame& Last, th, bMono)
    int n; size_t stride;
    const uint8_t* hasPoint; const float* pos; const int* octave; const float* angle;
    const uint4* desc; const int* obs;
    const float* tcwLast; const float* tcwCur;      // [B][12]
    int mono, checkOri;
};
struct KeyFramePtsDev {       // SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist) and (KeyFrame*, Scw, points, matched, th)
    int n; size_t stride;
    const uint8_t* valid; const float* pos; const float* minDist; const float* maxDist; const float* maxDistRaw;
    const float* normal;      // Scw variant
    const float* angle;       // KeyFrame variant
    const uint4* desc;
    const float* tcw;         // [B][12]
    float logScaleFactor;     // mfLogScaleFactor = logf(mfScaleFactor)
    int distTh;               // ORBdist / TH_LOW
    int checkOri;
};
struct ProjSearchArgs {
    FrameSetDev F;
    MapPointDev mp;
    LastFrameDev lf;
    KeyFramePtsDev kf;
    float th, nnratio;
    const int* kpObs;         // [B][cap] or null
    uint32_t* cand;           // [B][M][OBS_CAND_SLOTS] first chunk of every point's candidate list
    uint32_t* pool;           // [B][poolChunks][OBS_CAND_SLOTS] overflow chunks
    int* poolCursor;          // [B]
    int poolChunks;
    int* choice;              // [B][M]
    int* kpMatch;             // [B][cap]
    int* nMatches;            // [B]
    int* rounds;              // [B] resolution rounds (diagnostics)
};
// variant 0: map points (ORBmatcher.cc:45-129); 1: last frame (:1328-1470); 2: keyframe (:1472-1599); 3: Sim3 keyframe (:290-403)
cudaError_t launch_proj_search(const ProjSearchArgs& a, int variant, int nFrames, cudaStream_t st);

struct InitSearchArgs {       // SearchForInitialization (:405-520)
    FrameSetDev F1, F2;
    float* prevMatched;       // [B][cap1][2]
    int* matches12;           // [B][cap1]
    int* nMatches;            // [B]
    uint32_t* list;           // [B][cap1][listCap]  dist << 16 | i2, in GetFeaturesInArea order
    int* listCount;           // [B][cap1]
    int listCap;
    int window; float nnratio; int checkOri;
};
cudaError_t launch_init_search(const InitSearchArgs& a, int nFrames, cudaStream_t st);

struct FuseSearchArgs {       // search half of ORBmatcher::Fuse (ORBmatcher.cc:825-966 and :974-1100)
    FrameSetDev F;
    KeyFramePtsDev kf;        // valid, pos, minDist, maxDist, maxDistRaw, normal, desc, tcw, logScaleFactor
    const float* ow;          // [B][3] GetCameraCenter(); null in the Sim3 variant (Ow = -Rcw.t()*tcw)
    float invSigma2[OBS_MAX_LEVELS];
    const float* tcw2;        // mode 2: second transform [B][12], p3Dc = T2 * (T1 * p3Dw)
    float th; int sim3;       // 0: Fuse(KeyFrame*, points), 1: Fuse(KeyFrame*, Scw, ...), 2: one direction of SearchBySim3 (:1153-1226)
    int* bestIdx; int* bestDist;   // [B][n]
};
cudaError_t launch_fuse_search(const FuseSearchArgs& a, int nFrames, cudaStream_t st);
// SearchBySim3 agreement (:1305-1319): vn1 [B][n1], vn2 [B][n2] = best indices with distance <= TH_HIGH applied here
cudaError_t launch_sim3_agree(const int* idx1, const int* dist1, int n1, const int* idx2, const int* dist2, int n2, int thHigh,
                              int* match12, int* nFound, int nFrames, cudaStream_t st);

cudaError_t launch_three_maxima(const int* binSizes, int nHist, int length, int* ind, cudaStream_t st);
cudaError_t launch_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int* dist, cudaStream_t st);

struct Knn2Args {
    const uint4* desc;        // [K][n][2]
    int n;
    const int2* pairs; int nPairs;
    int thLow; float nnratio;
    int* bestIdx; int* bestDist; int* secondDist;   // [P][n]; the last two may be null
};
cudaError_t launch_knn2(const Knn2Args& a, cudaStream_t st);

// ---- DBoW2-gated matchers (bow.cu).  One side = B keyframes/frames, `cap` keypoint slots and `nodeCap`
// feature-vector slots each; the DBoW2::FeatureVector is a CSR (node ids ascending like the std::map).
struct BowSideDev {
    int cap, nodeCap;
    const int* n;             // [B]
    const uint4* desc;        // [B][cap][2]
    const float* keys;        // [B][cap][7]  cv::KeyPoint words (x, y, size, angle, response, octave, class_id)
    const uint8_t* valid;     // [B][cap] or null
    const float* uRight;      // [B][cap] or null
    const int* nNodes;        // [B]
    const uint32_t* nodeId;   // [B][nodeCap]
    const int* nodeStart;     // [B][nodeCap + 1]
    const int* nodeIdx;       // [B][cap]
};
struct BowSearchArgs {        // SearchByBoW (ORBmatcher.cc:159-288, :522-655)
    BowSideDev A, B;
    int thLow, strictLow; float nnratio; int checkOri;
    int* match12;             // [B][A.cap]
    int* match21;             // [B][B.cap]
    int* nMatches;            // [B]
};
cudaError_t launch_bow_search(const BowSearchArgs& a, int nPairs, cudaStream_t st);
struct TriSearchArgs {        // SearchForTriangulation (ORBmatcher.cc:657-823)
    BowSideDev A, B;
    const float* f12;         // [B][9]
    const float* epipole;     // [B][2]
    float sigma2[OBS_MAX_LEVELS], scale[OBS_MAX_LEVELS];
    int onlyStereo, checkOri;
    int* match12;             // [B][A.cap]
    int* nMatches;            // [B]
};
cudaError_t launch_tri_search(const TriSearchArgs& a, int nPairs, cudaStream_t st);
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:345-410) over a CSR of descriptor lists
cudaError_t launch_distinctive(const uint4* desc, const int* start, int nPoints, int* best, cudaStream_t st);

// ---- object layer (objects.cu): keypoint-to-mask assignment of Frame::BuildObject2DsRGBD / BuildObject2DsStereo (src/Frame.cc:240-311, :314-385)
struct MaskAssignArgs {
    const float* keys;        // [n][7] cv::KeyPoint words (mvKeysUn)
    const float* depth;       // [n] mvDepth
    int n;
    const uint8_t* masks;     // n_masks images, rowStride bytes per row, imageStride bytes per mask
    int nMasks, w, h; size_t rowStride, imageStride;
    float thDepth; int minKeypoints;
    int* maskOfKp;            // [n] first mask whose 20x20 window is all 255 (and 0 < depth <= thDepth), else -1
    int* objectKp;            // [n][2] mvObjectKpIndices: (Object2D index, index inside the object) or (-1, -1)
    int* objectOfMask;        // [nMasks] Object2D index created for the mask, or -1
    int* nObjects;            // [1]
};
cudaError_t launch_mask_assign(const MaskAssignArgs& a, cudaStream_t st);

// Frame::ExtractHSVHistogramsFromMask (src/Frame.cc:388-414) for n_masks masks over one colour image
#define OBS_HSV_BINS 94           // 32 (V) + 32 (S) + 30 (H), in the order hconcat leaves them
cudaError_t hsv_tables_upload();
cudaError_t launch_hsv_hist(const uint8_t* bgr, size_t bgrStride, const uint8_t* masks, size_t maskStride, size_t maskImageStride,
                            int nMasks, int w, int h, int* counts, float* hist, cudaStream_t st);

// Frame::UndistortKeyPoints (src/Frame.cc:644-674) / the corner undistortion of ComputeImageBounds (:676-704):
// cv::undistortPoints(pts, pts, K, distCoef, Mat(), K) over n points; ptStride / outStride in floats (2 = packed points,
// 7 = the pt fields of cv::KeyPoint records).
struct UndistortArgs { double fx, fy, cx, cy, k[14]; };
cudaError_t launch_undistort(const UndistortArgs& a, const float* pts, int ptStride, float* out, int outStride, int n, cudaStream_t st);

// Object2D::Object2D, src/ObjectTypes.cc:23: cv::distanceTransform(~mask, dist, DIST_L2, DIST_MASK_PRECISE) for n_masks masks
cudaError_t launch_distance_transform(const uint8_t* masks, size_t maskStride, size_t maskImageStride, int nMasks, int w, int h,
                                      float* out, cudaStream_t st);
