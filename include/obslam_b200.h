/* obslam_b200 -- C ABI of the B200-native ORB front end for Object_SLAM.
 *
 * This is the drop-in boundary for the data-parallel per-frame path of the reference
 * (yangliu9527/Object_SLAM, an ORB_SLAM2 fork).  The reference has no FFI layer of its own: its
 * seam is the C++ class surface compiled into libORB_SLAM2.so.  Each entry point below names the
 * reference interface it stands behind (file:line relative to the reference tree); the thin C++
 * classes in object_slam_b200/host/ keep the reference's signatures and forward to these calls
 * (see INTEGRATION.md).
 *
 * Conventions: plain C, opaque handles, caller-owned output buffers, int status returns
 * (OBS_OK == 0), no exceptions, no exit(), no global mutable state.  A handle owns its device
 * buffers and one CUDA stream; calls on different handles may run concurrently from different
 * threads (the reference extracts the two stereo eyes in two threads on two extractor
 * instances, src/Frame.cc:78-81); calls on one handle must be serialised by the caller.
 * There is no CPU fallback: every call fails with OBS_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef OBSLAM_B200_H
#define OBSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum obs_status {
    OBS_OK = 0,
    OBS_ERR_INVALID = 1,      /* bad argument (null pointer, size out of range, shape mismatch) */
    OBS_ERR_CUDA = 2,         /* CUDA runtime error or no usable device; see obs_last_error() */
    OBS_ERR_CAPACITY = 3,     /* caller buffer / handle capacity too small */
    OBS_ERR_STATE = 4         /* call order violated (e.g. stereo match before extraction) */
} obs_status;

/* ORBextractor constructor arguments, include/ORBextractor.h:51, src/ORBextractor.cc:410. */
typedef struct obs_orb_params {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
} obs_orb_params;

/* Binary layout of cv::KeyPoint (28 bytes), so the C++ wrapper can write straight into
 * std::vector<cv::KeyPoint>::data(). */
typedef struct obs_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} obs_keypoint;

typedef struct obs_extractor obs_extractor;

/* Thread-local description of the last failure on the calling thread. */
const char* obs_last_error(void);
/* Library / build identification ("obslam_b200 <version> sm_100a"). */
const char* obs_version(void);
/* Number of CUDA devices visible, or a negative obs_status. */
int obs_device_count(void);

/* Page-locked host memory.  Host buffers handed to obs_extract*, obs_extractor_fetch and
 * obs_stereo_match that were allocated here (or with cudaMallocHost / cudaHostRegister) are moved by
 * DMA directly; any other host pointer goes through a staging copy inside the library. */
int obs_host_alloc(size_t bytes, void** out);
int obs_host_free(void* p);

/* ---------------------------------------------------------------------------------------
 * Extractor.  Replaces ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111).
 * --------------------------------------------------------------------------------------- */

/* ORBextractor::ORBextractor, src/ORBextractor.cc:410-470.  max_w/max_h bound the image size,
 * max_batch the number of images one obs_extract_batch* call may carry. */
int obs_extractor_create(const obs_orb_params* params, int max_w, int max_h, int max_batch,
                         int device, obs_extractor** out);
int obs_extractor_destroy(obs_extractor* e);

/* Getters, include/ORBextractor.h:63-83.  Each array receives nlevels floats. */
int obs_extractor_levels(const obs_extractor* e);
int obs_extractor_max_keypoints(const obs_extractor* e);   /* per-image capacity the handle was sized for */
int obs_extractor_tables(const obs_extractor* e, float* scale_factors, float* inv_scale_factors,
                         float* level_sigma2, float* inv_level_sigma2, int32_t* features_per_level);

/* ORBextractor::operator(), src/ORBextractor.cc:1043-1105, on one 8-bit single-channel host
 * image (stride in bytes).  Writes up to `cap` keypoints (level-major order, as the reference)
 * and cap x 32 descriptor bytes; *n_out is the number found.  An empty image (w or h == 0)
 * returns OBS_OK with *n_out = 0, like the reference's early return at :1046. */
int obs_extract(obs_extractor* e, const uint8_t* image, int w, int h, size_t stride,
                obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);

/* The same over n_images host images of one shape (images[i] = first byte of image i).
 * keypoints: n_images x cap records, descriptors: n_images x cap x 32 bytes, n_out: n_images. */
int obs_extract_batch(obs_extractor* e, const uint8_t* const* images, int n_images, int w, int h,
                      size_t stride, obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);

/* Device-resident form: images already in HBM (base and stride 16-byte aligned; image i starts at
 * d_images + i*image_stride).  Results stay on the device until fetched.  `stream` is a
 * cudaStream_t the work is ordered on (NULL = the handle's own stream). */
int obs_extract_batch_device(obs_extractor* e, const uint8_t* d_images, int n_images, int w, int h,
                             size_t stride, size_t image_stride, void* stream);
/* Copy the results of the last extraction to host buffers (synchronises the stream used). */
int obs_extractor_fetch(obs_extractor* e, obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out);
/* Only the per-image keypoint counts (n_images ints). */
int obs_extractor_fetch_counts(obs_extractor* e, int* n_out);
/* Device pointers of the last results: per image a record of `record_bytes`, laid out as
 * int32 header[16] (header[0] = n, header[1..nlevels] = per-level counts), then cap x 28-byte
 * keypoints, then cap x 32-byte descriptors.  Valid until the next call on the handle. */
int obs_extractor_results_device(obs_extractor* e, const void** d_records, size_t* record_bytes, int* cap);

/* ORBextractor::mvImagePyramid[level] (include/ORBextractor.h:85) of image `image_index` of the
 * last call, without the 19-px border the reference keeps around it (that border is never read
 * on this path).  which = 0: pyramid level; 1: its 7x7 sigma-2 Gaussian blur (:1085-1086).
 * dst may be NULL to query the size only. */
int obs_extractor_get_level(obs_extractor* e, int image_index, int level, int which,
                            uint8_t* dst, size_t dst_stride, int* w, int* h);

/* Stage outputs of the last call, for parity tests: FAST candidates of a level in the order
 * ComputeKeyPointsOctTree (:765-829) hands them to DistributeOctTree (x, y relative to the
 * 16-px border, response), and the keypoints DistributeOctTree (:539-763) selected, in list
 * order.  xyr receives 3 ints per entry; returns the count via *n_out. */
int obs_extractor_get_candidates(obs_extractor* e, int image_index, int level, int32_t* xyr, int cap, int* n_out);
int obs_extractor_get_selected(obs_extractor* e, int image_index, int level, int32_t* xyr, int cap, int* n_out);

/* Per-stage device timing for benchmarks.  While profiling is on, every stage of every extraction on the
 * handle (and every stereo match whose left handle it is) is bracketed by CUDA events on the stream it runs on,
 * up to 128 calls.  obs_extractor_stage_ms synchronises the device and returns the SUM over the recorded
 * calls of each stage's time in ms: stage_ms[OBS_NUM_STAGES] = {pyramid, FAST, quadtree, blur,
 * orientation+descriptors}; *stereo_ms = stereo match + outlier filter.  Without profiling the blur runs on an
 * auxiliary stream beside FAST + quadtree (both depend on the pyramid only); while profiling is on the stages
 * of a call are serialised so that their brackets do not overlap.  While profiling is on the enqueue of every stage
 * is also wrapped in an NVTX range ("obs:ComputePyramid", "obs:FAST", "obs:DistributeOctTree", "obs:GaussianBlur",
 * "obs:IC_Angle+computeOrbDescriptor", "obs:ComputeStereoMatches" inside "obs:extract") for Nsight timelines. */
#define OBS_NUM_STAGES 5
int obs_extractor_set_profiling(obs_extractor* e, int on);
int obs_extractor_stage_ms(obs_extractor* e, float* stage_ms, float* stereo_ms, int* n_calls, int* n_stereo_calls);

/* ---------------------------------------------------------------------------------------
 * Stereo.  Replaces Frame::ComputeStereoMatches (include/Frame.h:103, src/Frame.cc:706-880).
 * Uses the device-resident pyramids, keypoints and descriptors of the last extraction on the
 * two handles; image i of `left` is matched against image i of `right`.
 * min_d / max_d are the disparity limits (the reference derives them from members it has not
 * initialised yet, Frame.cc:736; the intended values are 0 and fx).
 * u_right / depth: n_images x cap floats (-1 = no match), cap >= keypoints of each left image.
 * --------------------------------------------------------------------------------------- */
int obs_stereo_match(obs_extractor* left, obs_extractor* right, float mbf, float min_d, float max_d,
                     float* u_right, float* depth, int cap);
/* Asynchronous obs_extract_batch for page-locked buffers (obs_host_alloc): image i starts at images + i * h * stride.
 * _submit returns once upload, extraction and download are enqueued (in overlapping chunks on the handle's streams);
 * _wait sleeps on a blocking-sync event until the results have landed and checks the counts.  One submission per handle
 * may be in flight; several handles driven round-robin by one host thread overlap each other. */
int obs_extract_batch_submit(obs_extractor* e, const uint8_t* images, int n_images, int w, int h, size_t stride,
                             obs_keypoint* keypoints, uint8_t* descriptors, int cap, int32_t* n_out);
int obs_extract_batch_wait(obs_extractor* e);

/* Whole stereo frames in one call -- what the stereo Frame constructor does (src/Frame.cc:78-90: ExtractORB on two threads,
 * then ComputeStereoMatches) for n_frames frames: both eyes are uploaded, extracted and downloaded in overlapping chunks on the
 * handles' own streams and the stereo match follows on the device, all enqueued by the calling thread.  Every buffer of
 * obs_stereo_io must be page-locked (obs_host_alloc); image i of an eye starts at base + i * h * stride.
 * obs_stereo_frames_submit returns once everything is enqueued; obs_stereo_frames_wait sleeps (blocking-sync event, no spinning)
 * until all results have landed and checks the counts against cap.  One submission per handle pair may be in flight; several
 * handle pairs driven round-robin by one thread overlap each other's transfers and kernels.  obs_stereo_frames = submit + wait. */
typedef struct obs_stereo_io {
    const uint8_t* left;  const uint8_t* right;                     /* n_frames x h x stride bytes each */
    obs_keypoint* kp_left;  uint8_t* desc_left;  int32_t* n_left;   /* n_frames x cap, n_frames x cap x 32, n_frames */
    obs_keypoint* kp_right; uint8_t* desc_right; int32_t* n_right;
    float* u_right; float* depth;                                   /* n_frames x cap each (-1 = no match) */
} obs_stereo_io;
int obs_stereo_frames_submit(obs_extractor* left, obs_extractor* right, const obs_stereo_io* io, int n_frames, int w, int h,
                             size_t stride, int cap, float mbf, float min_d, float max_d);
int obs_stereo_frames_wait(obs_extractor* left, obs_extractor* right);
int obs_stereo_frames(obs_extractor* left, obs_extractor* right, const obs_stereo_io* io, int n_frames, int w, int h,
                      size_t stride, int cap, float mbf, float min_d, float max_d);

/* Process-wide launch options, both on by default.
 *   "pdl"    = 1: the kernels of the per-frame chain are launched with programmatic stream serialization (each kernel releases its
 *              successor at entry -- griddepcontrol.launch_dependents -- and waits for its predecessor's results with
 *              griddepcontrol.wait), so launch latency and prologues overlap the predecessor's tail;
 *   "graphs" = 1: obs_stereo_frames_submit replays a CUDA graph once it has seen the same buffers, shape and parameters twice
 *              (one cudaGraphLaunch instead of ~60 runtime calls per stereo frame; up to 8 argument sets per handle pair;
 *              inside the graph the kernels are joined by ordinary graph edges).
 *   "knn2_cta_pair" = 1: the tensor-core engine of obs_hamming_knn2 runs as CTA pairs (tcgen05.mma.cta_group::2: two SMs share every
 *              database tile, csrc/knn2_tc.cu k_knn2_tc_pair); 0: one CTA per SM (k_knn2_tc).
 * Results are identical either way.  Returns OBS_ERR_INVALID for an unknown name. */
int obs_set_option(const char* name, int value);

/* Device-resident form: results stay in HBM (n_images x cap floats each, cap =
 * obs_extractor_max_keypoints(left)); pointers valid until the next call. */
int obs_stereo_match_device(obs_extractor* left, obs_extractor* right, float mbf, float min_d, float max_d,
                            void* stream, const float** d_u_right, const float** d_depth);


/* ---------------------------------------------------------------------------------------
 * Front-end neighbours of the extractor inside Tracking::GrabImage* and the Frame constructors, so that an RGB-D or
 * colour stereo frame only crosses PCIe once.  Device-resident images; `stream` NULL = the handle's stream.
 * --------------------------------------------------------------------------------------- */
/* cvtColor(im, im, CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY | CV_BGRA2GRAY), src/Tracking.cc:202-227, :247-258, :290-301.
 * channels: 3 or 4; rgb_order: 1 = first channel is red (mbRGB), 0 = blue.  Strides in bytes. */
int obs_gray_from_color(obs_extractor* e, const uint8_t* d_src, int n_images, int w, int h, int channels, int rgb_order,
                        size_t src_stride, size_t src_image_stride, uint8_t* d_gray, size_t gray_stride,
                        size_t gray_image_stride, void* stream);
/* imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor), src/Tracking.cc:262-263, for 16-bit depth maps. */
int obs_depth_to_float(obs_extractor* e, const uint16_t* d_src, int n_images, int w, int h, size_t src_stride,
                       size_t src_image_stride, float factor, float* d_dst, size_t dst_stride, size_t dst_image_stride,
                       void* stream);
/* Frame::ComputeStereoFromRGBD, src/Frame.cc:883-904, on the keypoints of the last extraction (taken as undistorted):
 * mvuRight / mvDepth as device arrays of n_images x obs_extractor_max_keypoints floats (-1 = no depth). */
int obs_stereo_from_rgbd(obs_extractor* e, const float* d_depth, size_t depth_stride, size_t depth_image_stride, float mbf,
                         void* stream, const float** d_u_right, const float** d_depth_out);

/* ---------------------------------------------------------------------------------------
 * Matchers.  Replace the Hamming searches of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-102)
 * that Tracking runs on every frame.  The reference walks pointer graphs (Frame, MapPoint); the
 * entry points below take the same quantities as flat arrays, named after the members they are
 * read from.  EVERY array pointer below may be host memory or device memory (the library looks the
 * pointer up; host arrays are staged, device arrays are used in place, device outputs are written
 * in place without a synchronisation).
 *
 * An obs_matcher owns one stream and its workspace; calls on different matchers may run
 * concurrently from different threads (the reference's matchers are stack-local objects used from
 * the Tracking, LocalMapping and LoopClosing threads).
 * --------------------------------------------------------------------------------------- */
typedef struct obs_matcher obs_matcher;
typedef struct obs_frame_set obs_frame_set;

int obs_matcher_create(int device, obs_matcher** out);
int obs_matcher_destroy(obs_matcher* m);
/* cudaStream_t the matcher's work is ordered on. */
void* obs_matcher_stream(obs_matcher* m);
/* Wait for everything queued on the matcher's stream (needed only after calls with device outputs). */
int obs_matcher_sync(obs_matcher* m);

/* Static members of Frame and the camera (src/Frame.cc:33-35, :93-114, :691-702). */
typedef struct obs_frame_params {
    float min_x, max_x, min_y, max_y;       /* mnMinX, mnMaxX, mnMinY, mnMaxY */
    float fx, fy, cx, cy, mbf, mb;
    int32_t nlevels;
    float scale_factors[12];                /* mvScaleFactors */
} obs_frame_params;

/* What the matchers read of one Frame. */
typedef struct obs_frame_view {
    int32_t n;                              /* N */
    const obs_keypoint* keys_un;            /* mvKeysUn (x, y, angle, octave are read) */
    const uint8_t* descriptors;             /* mDescriptors, n x 32 */
    const float* u_right;                   /* mvuRight, or NULL for a monocular frame (all -1) */
} obs_frame_view;

/* A batch of up to max_frames frames of up to max_keypoints keypoints each, resident in HBM together
 * with the 64x48 keypoint grid of each frame (Frame::AssignFeaturesToGrid, src/Frame.cc:455-470,
 * PosInGrid :622-632), which is built on the device. */
int obs_frame_set_create(obs_matcher* m, const obs_frame_params* params, int max_frames, int max_keypoints,
                         obs_frame_set** out);
int obs_frame_set_destroy(obs_frame_set* fs);
int obs_frame_set_upload(obs_frame_set* fs, const obs_frame_view* frames, int n_frames);
/* The same from the results of the last extraction on `e` (no host round trip): frame i = image i;
 * the keypoints are taken as undistorted (the reference copies mvKeys to mvKeysUn when the camera has
 * no distortion, src/Frame.cc:646-650).  d_u_right: device array n_images x obs_extractor_max_keypoints
 * (e.g. from obs_stereo_match_device) or NULL. */
int obs_frame_set_from_extractor(obs_frame_set* fs, obs_extractor* e, const float* d_u_right);
int obs_frame_set_count(const obs_frame_set* fs);
/* Grid of frame `frame` as a CSR in mGrid[ix][iy] order (cell = ix*48 + iy): cell_start receives
 * 64*48+1 ints, cell_idx the keypoint indices (ascending inside a cell, like the reference's push_back). */
int obs_frame_set_grid(obs_frame_set* fs, int frame, int32_t* cell_start, int32_t* cell_idx, int cap);

/* Fields of MapPoint read by SearchByProjection(Frame&, const vector<MapPoint*>&, th); arrays of n
 * entries per frame.  per_frame = 0: one list shared by all frames of the set; 1: n_frames lists
 * stored one after the other. */
typedef struct obs_mappoint_view {
    int32_t n;
    int32_t per_frame;
    const uint8_t* in_view;                 /* mbTrackInView && !isBad() */
    const float* proj_x;                    /* mTrackProjX */
    const float* proj_y;                    /* mTrackProjY */
    const float* proj_xr;                   /* mTrackProjXR */
    const int32_t* scale_level;             /* mnTrackScaleLevel */
    const float* view_cos;                  /* mTrackViewCos */
    const uint8_t* descriptors;             /* GetDescriptor(), n x 32 */
    const int32_t* observations;            /* Observations() */
} obs_mappoint_view;

/* ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:45-129,
 * for every frame of the set.  F.mvpMapPoints is modelled by two arrays over the keypoints
 * (n_frames x max_keypoints each): kp_observations (in, may be NULL = all free) holds Observations()
 * of the point a keypoint already carries (0 = none); kp_match (out) receives -1 where the call
 * left the keypoint alone, else the index of the map point assigned last.  n_matches: per frame,
 * the function's return value.  The result equals the reference's sequential first-come-first-served
 * loop exactly. */
int obs_search_by_projection(obs_matcher* m, obs_frame_set* frames, const obs_mappoint_view* points,
                             float th, float nnratio, const int32_t* kp_observations,
                             int32_t* kp_match, int32_t* n_matches);

/* Fields of the last Frame read by SearchByProjection(Frame& Current, const Frame& Last, th, bMono);
 * n entries per frame (per_frame as above). */
typedef struct obs_lastframe_view {
    int32_t n;                              /* LastFrame.N */
    int32_t per_frame;
    const uint8_t* has_point;               /* mvpMapPoints[i] && !mvbOutlier[i] */
    const float* world_pos;                 /* GetWorldPos(), n x 3 */
    const int32_t* octave;                  /* mvKeys[i].octave */
    const float* angle;                     /* mvKeysUn[i].angle */
    const uint8_t* descriptors;             /* pMP->GetDescriptor(), n x 32 */
    const int32_t* observations;            /* pMP->Observations() */
    const float* tcw_last;                  /* LastFrame.mTcw rows 0..2, 12 floats per frame */
    const float* tcw_current;               /* CurrentFrame.mTcw rows 0..2, 12 floats per frame */
} obs_lastframe_view;

/* src/ORBmatcher.cc:1328-1470 incl. the rotation-histogram check (:1431-1467) with
 * ComputeThreeMaxima (:1601-1642).  kp_match additionally uses -2 = reset to NULL by that check. */
int obs_search_by_projection_last(obs_matcher* m, obs_frame_set* current, const obs_lastframe_view* last,
                                  float th, int mono, int check_orientation, const int32_t* kp_observations,
                                  int32_t* kp_match, int32_t* n_matches);

/* Fields of the map points read by the two keyframe searches below; n entries per frame (per_frame as above). */
typedef struct obs_keyframe_points_view {
    int32_t n;
    int32_t per_frame;
    const uint8_t* valid;                   /* pMP && !pMP->isBad() && not in the already-found set */
    const float* world_pos;                 /* GetWorldPos(), n x 3 */
    const float* min_distance;              /* GetMinDistanceInvariance() */
    const float* max_distance;              /* GetMaxDistanceInvariance() */
    const float* max_distance_raw;          /* mfMaxDistance, read by MapPoint::PredictScale (src/MapPoint.cc:488-519) */
    const float* normal;                    /* GetNormal(), n x 3 (Sim3 variant only, else NULL) */
    const float* angle;                     /* pKF->mvKeysUn[i].angle (keyframe variant only, else NULL) */
    const uint8_t* descriptors;             /* GetDescriptor(), n x 32 */
    const float* tcw;                       /* 12 floats per frame: rows 0..2 of CurrentFrame.mTcw, resp. [Rcw | tcw] of Scw */
} obs_keyframe_points_view;

/* ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist),
 * src/ORBmatcher.cc:1472-1599 (relocalisation): projection, distance gate, MapPoint::PredictScale (glibc logf restated
 * on the device), window search on levels [l-1, l+1], best distance <= orb_dist, rotation check.  kp_taken (in, may be
 * NULL): != 0 where mvpMapPoints[k] is non-NULL (any point blocks).  kp_match as above (-2 = reset by the rotation check). */
int obs_search_by_projection_keyframe(obs_matcher* m, obs_frame_set* current, const obs_keyframe_points_view* points,
                                      float th, int orb_dist, int check_orientation, const int32_t* kp_taken,
                                      int32_t* kp_match, int32_t* n_matches);

/* ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>&
 * vpMatched, th), src/ORBmatcher.cc:290-403 (loop closing), after the decomposition of Scw (:299-304, tcw = [Rcw | tcw]):
 * the frames of `keyframes` are the keyframes' keypoints (KeyFrame::GetFeaturesInArea, src/KeyFrame.cc:569-608, is the
 * Frame function without level filter).  kp_taken: != 0 where vpMatched[k] is non-NULL. */
int obs_search_by_projection_sim3(obs_matcher* m, obs_frame_set* keyframes, const obs_keyframe_points_view* points,
                                  int th, const int32_t* kp_taken, int32_t* kp_match, int32_t* n_matches);

/* Search half of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th), src/ORBmatcher.cc:825-966 (sim3 = 0; points->tcw =
 * [GetRotation() | GetTranslation()], camera_centre = GetCameraCenter(), n_frames x 3) and of
 * ORBmatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint), :974-1100 (sim3 = 1; points->tcw = the decomposed Scw as in
 * obs_search_by_projection_sim3, camera_centre ignored).  points->valid = "pMP && !isBad() && !IsInKeyFrame(pKF)" (resp. not in
 * spAlreadyFound); points->normal is required.  Per (keyframe, point): best_idx = keypoint with the smallest descriptor distance in
 * the predicted window (-1 = none), best_dist (256 = none); both n_frames x points->n.  The caller accepts best_dist <= TH_LOW and does
 * the Replace / AddObservation bookkeeping (:935-962, :1077-1094) -- the search reads none of that state, so its results are those of
 * the reference's sequential loop. */
int obs_fuse_search(obs_matcher* m, obs_frame_set* keyframes, const obs_keyframe_points_view* points, const float* camera_centre,
                    float th, int sim3, int32_t* best_idx, int32_t* best_dist);

/* ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vector<MapPoint*>& vpMatches12, s12, R12, t12, th), src/ORBmatcher.cc:1102-1326, for
 * n_frames keyframe pairs (frame i of kf1 with frame i of kf2).  points1 / points2: the map points of the keyframes' keypoints
 * (n = number of keypoints; valid = "non-NULL, not bad, not already matched", :1131-1143, :1157-1162, :1232-1237; normal / angle unused);
 * points1->tcw = [R1w | t1w], points2->tcw = [R2w | t2w]; t21 = [sR21 | t21] and t12 = [sR12 | t12] as :1119-1122 builds them
 * (n_frames x 12 floats each).  match12[i1] = keypoint of kf2 where both directions agree (:1305-1319), else -1; n_found = the return
 * value.  Both projections use the target frame set's camera (the reference uses pKF1's for both). */
int obs_search_by_sim3(obs_matcher* m, obs_frame_set* kf1, obs_frame_set* kf2, const obs_keyframe_points_view* points1,
                       const obs_keyframe_points_view* points2, const float* t21, const float* t12, float th,
                       int32_t* match12, int32_t* n_found);

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:405-520: frame i of f1 against frame i of f2.
 * prev_matched: n_frames x max_keypoints(f1) x 2 floats, in/out (vbPrevMatched); matches12:
 * n_frames x max_keypoints(f1) (vnMatches12). */
int obs_search_for_initialization(obs_matcher* m, obs_frame_set* f1, obs_frame_set* f2, float* prev_matched,
                                  int32_t* matches12, int window_size, float nnratio, int check_orientation,
                                  int32_t* n_matches);

/* Diagnostics: resolution rounds each frame of the last projection search with host outputs took
 * (see csrc/matcher.cu); returns the number of frames. */
int obs_matcher_last_rounds(obs_matcher* m, int32_t* rounds, int cap);

/* ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1601-1642, over n_hist histograms of `length`
 * bin sizes each; ind receives 3 ints per histogram. */
int obs_compute_three_maxima(obs_matcher* m, const int32_t* bin_sizes, int n_hist, int length, int32_t* ind);

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1647-1663, over n pairs of 32-byte descriptors. */
int obs_descriptor_distance(obs_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* Brute-force best / second-best Hamming search with the ratio test (the candidate loop of
 * ORBmatcher::SearchByBoW, src/ORBmatcher.cc:200-229, over whole keyframes): `descriptors` holds
 * n_keyframes x n_desc x 32 bytes; pairs holds n_pairs x 2 ints (query keyframe, database keyframe).
 * Per pair and query descriptor: best_dist, second_dist (256 = none) and best_idx = index of the
 * nearest database descriptor if best_dist <= th_low and best_dist < nnratio * second_dist, else -1
 * (lowest index wins ties).  Outputs are n_pairs x n_desc ints each; best_dist / second_dist may be NULL. */
int obs_hamming_knn2(obs_matcher* m, const uint8_t* descriptors, int n_keyframes, int n_desc,
                     const int32_t* pairs, int n_pairs, int th_low, float nnratio,
                     int32_t* best_idx, int32_t* best_dist, int32_t* second_dist);

/* Which kernel obs_hamming_knn2 runs: the exact int8 tensor-core contraction (tcgen05.mma kind::i8 on +-1 operands,
 * hamming = (256 - a.b) / 2, csrc/knn2_tc.cu) or the POPC kernel (csrc/matcher.cu).  Both give identical results;
 * AUTO (the default) takes the tensor cores from 192 descriptors per keyframe on.  The tensor path keeps a
 * 256-byte-per-descriptor expansion of the descriptor sets in device memory (n_keyframes x n_desc x 256 bytes). */
#define OBS_KNN2_AUTO 0
#define OBS_KNN2_POPC 1
#define OBS_KNN2_TENSOR 2
int obs_matcher_set_knn2_engine(obs_matcher* m, int engine);

/* One side of a DBoW2-gated search for a batch of n_pairs keyframes / frames with `cap` keypoint slots and
 * `node_cap` feature-vector slots each.  DBoW2 is an un-vendored third-party dependency of the reference
 * (Thirdparty/DBoW2); its FeatureVector (std::map<NodeId, std::vector<unsigned> >, KeyFrame::mFeatVec /
 * Frame::mFeatVec) is passed as a CSR: node_id ascending (the map's iteration order), node_start, node_idx
 * (the vectors' contents, in order).  Host or device arrays. */
typedef struct obs_bow_side {
    int32_t cap, node_cap;
    const int32_t* n;                       /* [n_pairs] keypoints */
    const uint8_t* descriptors;             /* [n_pairs][cap][32] mDescriptors */
    const obs_keypoint* keys_un;            /* [n_pairs][cap] mvKeysUn (pt, angle, octave are read) */
    const uint8_t* valid;                   /* [n_pairs][cap] see the calls; NULL = all */
    const float* u_right;                   /* [n_pairs][cap] mvuRight (triangulation only); NULL = monocular */
    const int32_t* n_nodes;                 /* [n_pairs] mFeatVec.size() */
    const uint32_t* node_id;                /* [n_pairs][node_cap] */
    const int32_t* node_start;              /* [n_pairs][node_cap + 1] */
    const int32_t* node_idx;                /* [n_pairs][cap] */
} obs_bow_side;

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&), src/ORBmatcher.cc:159-288: side1 = the keyframe with
 * valid = "vpMapPointsKF[i] && !isBad()", side2 = the frame (valid NULL), strict_low = 0 (bestDist1 <= th_low); the
 * reference's vpMapPointMatches[iF] is the map point of keyframe keypoint match21[iF].
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&), :522-655: valid on both sides, strict_low = 1
 * (bestDist1 < th_low); vpMatches12[i1] is the map point of keypoint match12[i1] of the second keyframe.
 * match12: [n_pairs][side1.cap], match21: [n_pairs][side2.cap] (-1 = none), n_matches: [n_pairs] the return values.
 * Includes the rotation-histogram check (ComputeThreeMaxima) when check_orientation != 0. */
int obs_search_by_bow(obs_matcher* m, const obs_bow_side* side1, const obs_bow_side* side2, int n_pairs, int th_low,
                      int strict_low, float nnratio, int check_orientation, int32_t* match12, int32_t* match21,
                      int32_t* n_matches);

/* ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12, vector<pair<size_t,size_t>>&, bOnlyStereo),
 * src/ORBmatcher.cc:657-823 with CheckDistEpipolarLine (:139-156).  valid = "GetMapPoint(i) == NULL" on both sides;
 * f12: [n_pairs][9] row-major; epipole: [n_pairs][2] = (ex, ey) of :663-671 (computed by the caller with the
 * reference's cv::Mat expressions); level_sigma2 / scale_factors: pKF2->mvLevelSigma2 / mvScaleFactors (nlevels floats).
 * match12[i1] = i2 or -1 (vMatchedPairs lists the pairs in ascending i1). */
int obs_search_for_triangulation(obs_matcher* m, const obs_bow_side* side1, const obs_bow_side* side2, int n_pairs,
                                 const float* f12, const float* epipole, const float* level_sigma2,
                                 const float* scale_factors, int nlevels, int only_stereo, int check_orientation,
                                 int32_t* match12, int32_t* n_matches);

/* MapPoint::ComputeDistinctiveDescriptors, src/MapPoint.cc:345-410, batched over n_points map points: the descriptors of
 * point p's good observations are descriptors[start[p] .. start[p+1]) (32 bytes each); best[p] = index inside that list of
 * the descriptor with the least median Hamming distance to the others (first wins), -1 for an empty list. */
int obs_distinctive_descriptors(obs_matcher* m, const uint8_t* descriptors, const int32_t* start, int n_points,
                                int32_t* best);

/* Keypoint-to-mask assignment at the head of Frame::BuildObject2DsRGBD (src/Frame.cc:240-311, min_keypoints = 5) and
 * Frame::BuildObject2DsStereo (:314-385, min_keypoints = 10): the semantic masks (n_masks 8-bit images of w x h, 255 = object) are
 * visited in order; a keypoint not taken by an earlier mask goes to the mask if mask(int(y + row), int(x + col)) == 255 for all
 * row, col in [-10, 10) and 0 < depth <= th_depth (mThDepth); a mask with more than min_keypoints keypoints becomes the next
 * Object2D.  mask_of_kp (n, may be NULL): the mask that took the keypoint or -1; object_kp_indices (n x 2): mvObjectKpIndices =
 * (Object2D index, index inside the object's keypoint list) or (-1, -1); object_of_mask (n_masks, may be NULL); n_objects: N_O.
 * Windows that leave the image (undefined behaviour in the reference) reject.  Host or device arrays. */
int obs_assign_keypoints_to_masks(obs_matcher* m, const obs_keypoint* keys_un, const float* depth, int n, const uint8_t* masks,
                                  int n_masks, int w, int h, size_t mask_stride, size_t mask_image_stride, float th_depth,
                                  int min_keypoints, int32_t* mask_of_kp, int32_t* object_kp_indices, int32_t* object_of_mask,
                                  int32_t* n_objects);

/* Frame::ExtractHSVHistogramsFromMask, src/Frame.cc:388-414, for n_masks masks over one 8-bit 3-channel image (the reference
 * converts with CV_BGR2HSV whatever the channel order of imRGB): hist receives n_masks x 94 floats = the L1-normalised
 * concatenation V (32 bins) | S (32 bins) | H (30 bins), the order the reference's hconcat calls leave.  Mask pixels != 0 count. */
int obs_hsv_histograms(obs_matcher* m, const uint8_t* bgr, size_t bgr_stride, const uint8_t* masks, int n_masks, int w, int h,
                       size_t mask_stride, size_t mask_image_stride, float* hist);

/* Frame::UndistortKeyPoints, src/Frame.cc:644-674: cv::undistortPoints(mat, mat, mK, mDistCoef, cv::Mat(), mK) over the keypoints
 * (mvKeysUn[i] = mvKeys[i] with pt replaced; a plain copy when mDistCoef[0] == 0, :646-650).  dist_coef: k1, k2, p1, p2[, k3 ...]
 * (host array, n_dist <= 14).  OpenCV's scalar algorithm in binary64 (5 iterations), pinned against cv2 4.13. */
int obs_undistort_keypoints(obs_matcher* m, const obs_keypoint* keys, int n, float fx, float fy, float cx, float cy,
                            const float* dist_coef, int n_dist, obs_keypoint* keys_un);
/* The same on n packed (x, y) points, e.g. the four image corners of Frame::ComputeImageBounds (:676-704). */
int obs_undistort_points(obs_matcher* m, const float* pts, int n, float fx, float fy, float cx, float cy, const float* dist_coef,
                         int n_dist, float* out);

/* cv::distanceTransform(~mask, mDistTransImg, cv::DIST_L2, cv::DIST_MASK_PRECISE) of the Object2D constructor, src/ObjectTypes.cc:23,
 * for n_masks masks: dist (n_masks x h x w floats, packed) = exact Euclidean distance to the nearest pixel with mask == 255
 * (65536 everywhere for a mask without such a pixel, like OpenCV's own code path).  OpenCV built with IPP may differ in the last bit. */
int obs_distance_transform(obs_matcher* m, const uint8_t* masks, int n_masks, int w, int h, size_t mask_stride,
                           size_t mask_image_stride, float* dist);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU exchange step of batched keyframe-vs-keyframe matching: every rank (one process per GPU)
 * owns a contiguous shard of the keyframes as queries and needs all descriptor sets as database.
 * The all-gather runs over NCCL (bound at run time with dlopen, so the host process's own NCCL is
 * shared) on the communicator's stream, in n_chunks pieces of each rank's shard, each signalled by
 * an event: obs_hamming_knn2 calls on pairs whose database keyframes are already present overlap the
 * rest of the transfer.  Extraction, stereo and projection search shard by frame with no collective.
 * --------------------------------------------------------------------------------------- */
typedef struct obs_comm obs_comm;
int obs_comm_nccl_version(void);                                /* e.g. 22809, or -1 when NCCL cannot be loaded */
int obs_comm_unique_id(uint8_t* id128);                         /* ncclGetUniqueId; 128 bytes, to be broadcast by the host */
int obs_comm_create(const uint8_t* id128, int rank, int n_ranks, int device, obs_comm** out);
int obs_comm_destroy(obs_comm* c);
/* d_all (n_ranks x local_bytes, device) receives rank r's d_local (local_bytes, device, the same on every
 * rank, a multiple of 32) at offset r*local_bytes.  producer_stream: the cudaStream_t that wrote d_local.
 * Chunk k covers descriptors [n*k/n_chunks, n*(k+1)/n_chunks) of every shard. */
int obs_comm_allgather(obs_comm* c, const uint8_t* d_local, size_t local_bytes, uint8_t* d_all, int n_chunks,
                       void* producer_stream);
/* Make consumer_stream (e.g. obs_matcher_stream) wait until chunk `chunk` of the last all-gather has arrived. */
int obs_comm_wait(obs_comm* c, int chunk, void* consumer_stream);

/* Roofline denominator of the matchers, measured on `device`: register-resident 256-bit Hamming
 * distances per second (in 1e9/s).  mode 0: 8 xor + 8 popc + adds per distance (the instruction mix of
 * ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:1647-1663, with hardware popc); mode 1: the same
 * through three carry-save adders (5 popc); mode 3: four carry-save adders (4 popc); mode 2: popc only. */
int obs_microbench_popc(int device, int mode, double* gdist_per_s);
/* Roofline denominator of the tensor-core knn2 path: int8 tcgen05.mma throughput (TOP/s) in the kernel's own MMA shape
 * (kind::i8, M 128, N 256, K 32, both operands in shared memory), issued back to back on every SM with no loads or epilogue. */
int obs_microbench_imma(int device, double* tops);
/* Roofline denominators of the extractor / stereo stages, which are bound inside the SM, in 1e9/s for the
 * whole device: warp instructions issued per second with the ALU and FMA pipes fed 1:1, warp instructions per second of the ALU
 * pipe alone (LOP3 chains), conflict-free shared-memory load wavefronts per second. */
int obs_microbench_pipes(int device, double* issue_ginst_per_s, double* alu_ginst_per_s, double* lds_gwavefronts_per_s);

#ifdef __cplusplus
}
#endif
#endif /* OBSLAM_B200_H */
