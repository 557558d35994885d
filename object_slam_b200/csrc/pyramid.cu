// K1 -- image pyramid.  Replaces ORBextractor::ComputePyramid (src/ORBextractor.cc:1107-1132):
// level l = cv::resize(level l-1, INTER_LINEAR) on 8-bit pixels.  OpenCV's 8U bilinear path is
// fixed point (11-bit weights, two-stage shifts); the weights come from the host-built tap
// tables so that every output byte equals the reference's.  The 19-px reflect border the
// reference adds around each level is never read on this path and is not materialised.
#include "kernels.h"

namespace {

// 128 x 8 output pixels per CTA, 4 consecutive pixels per thread (one 32-bit store).
__global__ void __launch_bounds__(256) k_resize(const __grid_constant__ Geom g, const PyrPtrs p,
                                                const ResizeTap* __restrict__ xtab,
                                                const ResizeTap* __restrict__ ytab, int level) {
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int img = blockIdx.z;
    int spitch, dpitch;
    const uint8_t* src = level_ptr(p, g, img, level - 1, spitch);
    uint8_t* dst = const_cast<uint8_t*>(level_ptr(p, g, img, level, dpitch));
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy = blockIdx.y * 8 + threadIdx.y;
    if (dx0 >= D.w || dy >= D.h) return;

    const ResizeTap ty = ytab[D.ytab + dy];
    const int y0 = min(max(ty.ofs, 0), S.h - 1);
    const int y1 = min(max(ty.ofs + 1, 0), S.h - 1);
    const uint8_t* __restrict__ r0 = src + (size_t)y0 * spitch;
    const uint8_t* __restrict__ r1 = src + (size_t)y1 * spitch;
    const int b0 = ty.a0, b1 = ty.a1;

    uint32_t out = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int dx = dx0 + i;
        if (dx < D.w) {
            const ResizeTap tx = xtab[D.xtab + dx];
            const int sx = tx.ofs;
            const int sx1 = min(sx + 1, S.w - 1);
            const int h0 = (int)r0[sx] * tx.a0 + (int)r0[sx1] * tx.a1;
            const int h1 = (int)r1[sx] * tx.a0 + (int)r1[sx1] * tx.a1;
            int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            v = min(max(v, 0), 255);
            out |= (uint32_t)v << (8 * i);
        }
    }
    // rows are padded to the pitch, so the full word may be written
    *reinterpret_cast<uint32_t*>(dst + (size_t)dy * dpitch + dx0) = out;
}

}  // namespace

cudaError_t launch_pyramid(const Geom& g, PyrPtrs p, const ResizeTap* xtab, const ResizeTap* ytab, int nimg, cudaStream_t st) {
    for (int l = 1; l < g.nlevels; l++) {
        dim3 block(32, 8);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + 7) / 8, nimg);
        k_resize<<<grid, block, 0, st>>>(g, p, xtab, ytab, l);
    }
    return cudaGetLastError();
}
