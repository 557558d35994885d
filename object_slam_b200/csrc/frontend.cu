// Front-end neighbours of the extractor inside the Frame constructors -- plain
// streaming kernels, bound by HBM bandwidth:
//   k_gray          cvtColor(RGB/BGR/RGBA/BGRA -> GRAY) of Tracking::GrabImageStereo/RGBD/Monocular
//                   (src/Tracking.cc:202-227, :247-258, :290-301).  OpenCV's 8-bit path (pinned to cv2 4.13 on all 2^24
//                   colours): gray = (R*9798 + G*19235 + B*3735 + 16384) >> 15.
//   k_depth_scale   imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) (src/Tracking.cc:262-263): (float)d * factor.
//   k_rgbd_stereo   Frame::ComputeStereoFromRGBD (src/Frame.cc:883-904) on the keypoints the extractor left in HBM.
// Each thread of k_gray turns 16 pixels (three or four 128-bit loads) into one 128-bit store; rows that are not
// 16-byte aligned take a scalar path.
#include "kernels.h"

namespace {

__device__ __forceinline__ uint32_t gray1(uint32_t r, uint32_t g, uint32_t b) {
    return (r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15;
}

template <int CH>
__global__ void __launch_bounds__(256) k_gray(const uint8_t* __restrict__ src, size_t srcStride, size_t srcImgStride,
                                             uint8_t* __restrict__ dst, size_t dstStride, size_t dstImgStride,
                                             int w, int h, int rgbOrder, int aligned) {
    const int chunksPerRow = (w + 15) >> 4;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)chunksPerRow * h) return;
    const int y = (int)(t / chunksPerRow), c = (int)(t - (long long)y * chunksPerRow);
    const uint8_t* s = src + (size_t)blockIdx.y * srcImgStride + (size_t)y * srcStride + (size_t)c * 16 * CH;
    uint8_t* d = dst + (size_t)blockIdx.y * dstImgStride + (size_t)y * dstStride + (size_t)c * 16;
    const int n = min(16, w - c * 16);
    const int ir = rgbOrder ? 0 : 2, ib = rgbOrder ? 2 : 0;
    if (aligned && n == 16) {
        uint32_t in[4 * CH];
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(s) + k);
            in[4 * k] = q.x; in[4 * k + 1] = q.y; in[4 * k + 2] = q.z; in[4 * k + 3] = q.w;
        }
        uint32_t out[4];
#pragma unroll
        for (int wd = 0; wd < 4; wd++) {
            uint32_t o = 0;
#pragma unroll
            for (int p = 0; p < 4; p++) {
                const int px = wd * 4 + p;
                uint32_t ch[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const int byte = px * CH + k;
                    ch[k] = (in[byte >> 2] >> (8 * (byte & 3))) & 0xffu;
                }
                o |= gray1(ch[ir], ch[1], ch[ib]) << (8 * p);
            }
            out[wd] = o;
        }
        *reinterpret_cast<uint4*>(d) = make_uint4(out[0], out[1], out[2], out[3]);
    } else {
        for (int p = 0; p < n; p++) d[p] = (uint8_t)gray1(s[p * CH + ir], s[p * CH + 1], s[p * CH + ib]);
    }
}

__global__ void __launch_bounds__(256) k_depth_scale(const uint8_t* __restrict__ src, size_t srcStride, size_t srcImgStride,
                                                    uint8_t* __restrict__ dst, size_t dstStride, size_t dstImgStride,
                                                    int w, int h, float factor, int aligned) {
    const int chunksPerRow = (w + 7) >> 3;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)chunksPerRow * h) return;
    const int y = (int)(t / chunksPerRow), c = (int)(t - (long long)y * chunksPerRow);
    const uint16_t* s = reinterpret_cast<const uint16_t*>(src + (size_t)blockIdx.y * srcImgStride + (size_t)y * srcStride) + c * 8;
    float* d = reinterpret_cast<float*>(dst + (size_t)blockIdx.y * dstImgStride + (size_t)y * dstStride) + c * 8;
    const int n = min(8, w - c * 8);
    if (aligned && n == 8) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(s));
        const uint32_t in[4] = {q.x, q.y, q.z, q.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            o[2 * k] = __fmul_rn((float)(in[k] & 0xffffu), factor);
            o[2 * k + 1] = __fmul_rn((float)(in[k] >> 16), factor);
        }
        reinterpret_cast<float4*>(d)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(d)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else {
        for (int p = 0; p < n; p++) d[p] = __fmul_rn((float)s[p], factor);
    }
}

__global__ void __launch_bounds__(256) k_rgbd_stereo(const uint8_t* __restrict__ records, size_t recordBytes, int kpCap,
                                                    const uint8_t* __restrict__ depth, size_t depthStride, size_t depthImgStride,
                                                    int w, int h, float mbf, float* __restrict__ uRight, float* __restrict__ depthOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
    if (i >= kpCap) return;
    const uint8_t* rec = records + (size_t)img * recordBytes;
    const int n = *reinterpret_cast<const int*>(rec);
    float ur = -1.0f, dp = -1.0f;
    if (i < n) {
        const float* kp = reinterpret_cast<const float*>(rec + OBS_HDR_INTS * 4) + (size_t)i * 7;
        const float u = kp[0], v = kp[1];
        const int row = (int)v, col = (int)u;                      // cv::Mat::at<float>(float, float): the indices truncate
        if (row >= 0 && row < h && col >= 0 && col < w) {
            const float d = *reinterpret_cast<const float*>(depth + (size_t)img * depthImgStride + (size_t)row * depthStride + (size_t)col * 4);
            if (d > 0) { dp = d; ur = __fsub_rn(u, __fdiv_rn(mbf, d)); }
        }
    }
    uRight[(size_t)img * kpCap + i] = ur;
    depthOut[(size_t)img * kpCap + i] = dp;
}

}  // namespace

cudaError_t launch_gray(const uint8_t* src, size_t srcStride, size_t srcImgStride, uint8_t* dst, size_t dstStride, size_t dstImgStride,
                        int w, int h, int channels, int rgbOrder, int nimg, cudaStream_t st) {
    const int aligned = !(((uintptr_t)src | srcStride | srcImgStride | (uintptr_t)dst | dstStride | dstImgStride) & 15);
    const long long threads = (long long)((w + 15) >> 4) * h;
    const dim3 grid((unsigned)((threads + 255) / 256), nimg);
    if (channels == 3) k_gray<3><<<grid, 256, 0, st>>>(src, srcStride, srcImgStride, dst, dstStride, dstImgStride, w, h, rgbOrder, aligned);
    else k_gray<4><<<grid, 256, 0, st>>>(src, srcStride, srcImgStride, dst, dstStride, dstImgStride, w, h, rgbOrder, aligned);
    return cudaGetLastError();
}

cudaError_t launch_depth_scale(const uint8_t* src, size_t srcStride, size_t srcImgStride, uint8_t* dst, size_t dstStride, size_t dstImgStride,
                               int w, int h, float factor, int nimg, cudaStream_t st) {
    const int aligned = !(((uintptr_t)src | srcStride | srcImgStride | (uintptr_t)dst | dstStride | dstImgStride) & 15);
    const long long threads = (long long)((w + 7) >> 3) * h;
    const dim3 grid((unsigned)((threads + 255) / 256), nimg);
    k_depth_scale<<<grid, 256, 0, st>>>(src, srcStride, srcImgStride, dst, dstStride, dstImgStride, w, h, factor, aligned);
    return cudaGetLastError();
}

cudaError_t launch_rgbd_stereo(const uint8_t* records, size_t recordBytes, int kpCap, const uint8_t* depth, size_t depthStride,
                               size_t depthImgStride, int w, int h, float mbf, float* uRight, float* depthOut, int nimg, cudaStream_t st) {
    const dim3 grid((kpCap + 255) / 256, nimg);
    k_rgbd_stereo<<<grid, 256, 0, st>>>(records, recordBytes, kpCap, depth, depthStride, depthImgStride, w, h, mbf, uRight, depthOut);
    return cudaGetLastError();
}
