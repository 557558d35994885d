// Internal bridge between the extractor and matcher halves of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
struct obs_frame_set;
// Build frames [0, nFrames) of the set from device-resident cv::KeyPoint records / descriptors / uRight.
// `producer` (may be null) is the stream that wrote the sources; the build is ordered after it.
int obs_frame_set_build_device(obs_frame_set* fs, const uint8_t* keys, size_t keysFrameStride, const uint8_t* desc,
                               size_t descFrameStride, const float* uRight, size_t uRightFrameStride, const int* count,
                               size_t countStrideInts, int nFrames, cudaStream_t producer);
int obs_frame_set_capacity(const obs_frame_set* fs);
int obs_frame_set_device(const obs_frame_set* fs);       // device the set lives on
