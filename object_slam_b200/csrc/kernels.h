// Launchers of the sm_100a kernels (one translation unit per stage).  All of them enqueue on
// `st` and return the launch status; none synchronises.
#pragma once
#include "common.cuh"
#include "host_util.h"
#include <cuda.h>                  // CUtensorMap (type only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links libcuda)

// K1  pyramid.cu   -- ComputePyramid (src/ORBextractor.cc:1107-1132), levels 1..n-1
#define RESIZE_TILE_W 128          // output tile of the tiled resize kernel
#define RESIZE_TILE_H 64
cudaError_t pyramid_prepare(const Geom& g);
cudaError_t launch_pyramid(const Geom& g, PyrPtrs p, const ResizeTap* xtab, const ResizeTap* ytab, int nimg, cudaStream_t st);

cudaError_t launch_repack(const uint8_t* src, size_t srcImgStride, size_t srcPitch, uint8_t* dst, size_t dstImgStride,
                          int dstPitch, int w, int h, int nimg, cudaStream_t st);

// K2  fast.cu      -- per-cell FAST with threshold fallback (:765-829, cv::FAST :809/:814)
size_t fast_smem_bytes(int tileRows);
cudaError_t fast_prepare(int tileRows);
int fast_cells_per_cta_host(int wCell, int hCell);
cudaError_t launch_fast(const Geom& g, PyrPtrs p, const FastCta* ctaTab, uint32_t* cand, int* cellCount, int nimg, cudaStream_t st);

// K3  quadtree.cu  -- DistributeOctTree (:539-763) for every (image, level)
size_t quadtree_smem_bytes(int nodeCap);
cudaError_t quadtree_prepare(int nodeCap);
cudaError_t launch_quadtree(const Geom& g, int nodeCap, const uint32_t* cand, const int* cellCount,
                            uint32_t* keyScratch, uint16_t* nodeScratch, uint32_t* sel, int* selCount,
                            int nimg, cudaStream_t st);

// K5  blur.cu      -- GaussianBlur 7x7 sigma 2, BORDER_REFLECT_101 (:1085-1086), all levels
cudaError_t launch_blur(const Geom& g, PyrPtrs p, uint8_t* blurSlab, size_t blurStride, int nimg, cudaStream_t st);

// K4+K6+K7 describe.cu -- IC_Angle (:77-104), computeOrbDescriptor (:108-147), keypoint
// finalisation (:837-847, :1094-1103) into the per-image result record
// One TMA descriptor per level of the blur slab: 3-D tensors (pitch bytes x rows x images), box 80 x 37 x 1 -- the descriptor stage
// pulls each keypoint's 37 x 37 window into shared memory with one cp.async.bulk.tensor instead of 13 gathered loads and stores.
// The box starts at a 16-byte aligned column (the TMA unit rejects other start addresses with "illegal instruction", measured with
// tools/tma_probe), so it needs 15 + 37 = 52 bytes; it is 80 wide because the box width is the shared-memory row pitch: with 64
// (16 words) every second row falls on the same banks and the rBRIEF gathers, which cluster around the centre, collide; 20 words
// spread eight consecutive rows over all banks (A/B: +0.2 % for the whole step).
#define DESC_BOX_W 80
#define DESC_BOX_H 37
struct DescMaps { CUtensorMap m[OBS_MAX_LEVELS]; };      // host copy; the kernel reads the descriptors from global memory
cudaError_t launch_describe(const Geom& g, PyrPtrs p, const CUtensorMap* maps /* device copy of DescMaps::m */, int img0,
                            const uint32_t* sel, const int* selCount, uint8_t* records, size_t recordBytes,
                            int nimg, cudaStream_t st);

// K8  stereo.cu    -- Frame::ComputeStereoMatches (src/Frame.cc:706-880)
struct StereoArgs {
    Geom g;                  // geometry shared by both eyes
    PyrPtrs left, right;
    const uint8_t* recL; const uint8_t* recR; size_t recordBytes;
    float mbf, minD, maxD;
    float* uRight; float* depth;     // nimg x kpCap
    int* sad;                        // nimg x kpCap scratch (best SAD of accepted matches, -1 otherwise)
    int* rowStart;                   // nimg x (h + 1): CSR row table of the right keypoints
    uint2* rowIdx;                   // nimg x rowIdxCap entries of the row table: (x as float bits, octave << 16 | right keypoint index)
    int rowIdxCap;
};
cudaError_t launch_stereo(const StereoArgs& a, int nimg, cudaStream_t st);

// frontend.cu -- colour -> grey, depth scaling, Frame::ComputeStereoFromRGBD (src/Tracking.cc:202-263, src/Frame.cc:883-904)
cudaError_t launch_gray(const uint8_t* src, size_t srcStride, size_t srcImgStride, uint8_t* dst, size_t dstStride, size_t dstImgStride,
                        int w, int h, int channels, int rgbOrder, int nimg, cudaStream_t st);
cudaError_t launch_depth_scale(const uint8_t* src, size_t srcStride, size_t srcImgStride, uint8_t* dst, size_t dstStride, size_t dstImgStride,
                               int w, int h, float factor, int nimg, cudaStream_t st);
cudaError_t launch_rgbd_stereo(const uint8_t* records, size_t recordBytes, int kpCap, const uint8_t* depth, size_t depthStride,
                               size_t depthImgStride, int w, int h, float mbf, float* uRight, float* depthOut, int nimg, cudaStream_t st);
