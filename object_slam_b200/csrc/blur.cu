// K5 -- 7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, of every pyramid level.  Replaces the
// cv::GaussianBlur call at src/ORBextractor.cc:1085-1086.  OpenCV (>= 3.4.2 / 4.x) runs 8-bit
// blurs in fixed point: per-axis kernel [18,34,48,56,48,34,18]/256, a 16-bit horizontal pass and
// a vertical pass rounded once, out = (sum + 2^15) >> 16; this kernel reproduces those bytes.
// The blurred level is only read by the descriptor stage.
#include "kernels.h"

namespace {

constexpr int BT_W = 128, BT_H = 32;            // output tile
constexpr int IN_PITCH = BT_W + 8;              // 4-px aligned halo each side (3 needed)
constexpr int IN_ROWS = BT_H + 6;
constexpr int BLUR_THREADS = 256;

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(BLUR_THREADS) k_blur(const __grid_constant__ Geom g, const PyrPtrs p,
                                                       uint8_t* __restrict__ blurSlab, size_t blurStride) {
    __shared__ __align__(16) uint8_t in[IN_ROWS * IN_PITCH];
    __shared__ __align__(16) uint16_t hb[IN_ROWS * BT_W];

    const int tid = threadIdx.x, img = blockIdx.y;
    int t = blockIdx.x, level = 0;
#pragma unroll 1
    for (int l = 1; l < g.nlevels; l++) if (t >= g.lv[l].blurTileBase) level = l;
    const LevelGeom& lg = g.lv[level];
    t -= lg.blurTileBase;
    const int ty = t / lg.blurTilesX, tx = t - ty * lg.blurTilesX;
    const int x0 = tx * BT_W, y0 = ty * BT_H;
    const int W = lg.w, H = lg.h;
    int pitch;
    const uint8_t* src = level_ptr(p, g, img, level, pitch);

    // stage (BT_H+6) x (BT_W+8) source bytes; smem column c <-> x = x0 - 4 + c
    constexpr int WPR = IN_PITCH / 4;
    for (int i = tid; i < IN_ROWS * WPR; i += BLUR_THREADS) {
        const int r = i / WPR, wv = i - r * WPR;
        const int sy = reflect101(y0 - 3 + r, H);
        const int xs = x0 - 4 + 4 * wv;
        const uint8_t* row = src + (size_t)sy * pitch;
        uint32_t v;
        if (xs >= 0 && xs + 3 < W) {
            v = __ldg(reinterpret_cast<const uint32_t*>(row + xs));
        } else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int xx = xs + b;
                // columns further than the 3-px halo outside the level are never used
                const int sx = (xx >= -3 && xx < W + 3) ? reflect101(xx, W) : 0;
                v |= (uint32_t)row[sx] << (8 * b);
            }
        }
        reinterpret_cast<uint32_t*>(in)[i] = v;
    }
    __syncthreads();

    // horizontal pass: 4 outputs per task from 12 staged bytes
    for (int i = tid; i < IN_ROWS * (BT_W / 4); i += BLUR_THREADS) {
        const int r = i / (BT_W / 4), q = i - r * (BT_W / 4);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(in + r * IN_PITCH + 4 * q);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
        int b[12];
#pragma unroll
        for (int k = 0; k < 4; k++) { b[k] = (w0 >> (8 * k)) & 255; b[4 + k] = (w1 >> (8 * k)) & 255; b[8 + k] = (w2 >> (8 * k)) & 255; }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {      // output x = x0 + 4q + k  <->  centre byte index 4 + k
            const int c = 4 + k;
            o[k] = 18 * (b[c - 3] + b[c + 3]) + 34 * (b[c - 2] + b[c + 2]) + 48 * (b[c - 1] + b[c + 1]) + 56 * b[c];
        }
        uint2 st;
        st.x = o[0] | (o[1] << 16);
        st.y = o[2] | (o[3] << 16);
        *reinterpret_cast<uint2*>(hb + r * BT_W + 4 * q) = st;
    }
    __syncthreads();

    // vertical pass: each thread owns a 4-px column strip and 4 consecutive rows (sliding window)
    uint8_t* dst = blurSlab + (size_t)img * blurStride + lg.off;
    const int q = tid & 31, rg = tid >> 5;           // 32 strips x 8 row groups
    const int xo = x0 + 4 * q;
    if (xo < W) {
        uint32_t acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int k = 0; k < 4; k++) acc[a][k] = 0;
        const int wgt[7] = {18, 34, 48, 56, 48, 34, 18};
#pragma unroll
        for (int rr = 0; rr < 10; rr++) {            // staged rows rg*4 .. rg*4+9 feed output rows rg*4 .. rg*4+3
            const uint2 v = *reinterpret_cast<const uint2*>(hb + (rg * 4 + rr) * BT_W + 4 * q);
            const uint32_t e[4] = {v.x & 0xffffu, v.x >> 16, v.y & 0xffffu, v.y >> 16};
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int k7 = rr - a;               // tap index of this staged row for output row a
                if (k7 >= 0 && k7 < 7) {
#pragma unroll
                    for (int k = 0; k < 4; k++) acc[a][k] += (uint32_t)wgt[k7] * e[k];
                }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int yo = y0 + rg * 4 + a;
            if (yo < H) {
                uint32_t o = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) o |= ((acc[a][k] + 32768u) >> 16) << (8 * k);
                *reinterpret_cast<uint32_t*>(dst + (size_t)yo * lg.pitch + xo) = o;
            }
        }
    }
}

}  // namespace

cudaError_t launch_blur(const Geom& g, PyrPtrs p, uint8_t* blurSlab, size_t blurStride, int nimg, cudaStream_t st) {
    dim3 grid(g.blurTilesTotal, nimg);
    k_blur<<<grid, BLUR_THREADS, 0, st>>>(g, p, blurSlab, blurStride);
    return cudaGetLastError();
}
