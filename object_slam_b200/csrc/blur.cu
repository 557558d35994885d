// K5 -- 7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, of every pyramid level.  Replaces the
// cv::GaussianBlur call at src/ORBextractor.cc:1085-1086.  OpenCV (>= 3.4.2 / 4.x) runs 8-bit
// blurs in fixed point: per-axis kernel [18,34,48,56,48,34,18]/256, a 16-bit horizontal pass and
// a vertical pass rounded once, out = (sum + 2^15) >> 16; this kernel reproduces those bytes.
// The blurred level is only read by the descriptor stage.
//
// The stage is instruction-issue bound before it is HBM bound, so it is built around integer dot
// products: a thread owns a 4-pixel column strip and walks down the rows.  Per row it loads three
// aligned words straight from global/L1 (no shared memory, no barriers), forms the four horizontal
// sums with IDP.4A (two dot products of 4 bytes each), keeps them packed as u16 pairs of vertically
// adjacent rows, and produces each output pixel with four IDP.2A over an 8-row register window.
#include "kernels.h"

namespace {

constexpr int BL_ROWS = 32;                 // output rows per thread
constexpr int BL_BANDS = 4;                 // row bands per CTA
constexpr int BL_TW = 128, BL_TH = BL_ROWS * BL_BANDS;      // CTA tile: 128 x 128 pixels

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// 12 source bytes x-4 .. x+7 of one row as three words; columns outside the level are reflected.
__device__ __forceinline__ void load_window(const uint8_t* __restrict__ row, int x, int W, bool interior,
                                            unsigned& A, unsigned& B, unsigned& C) {
    if (interior) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(row + x);
        A = __ldg(q - 1); B = __ldg(q); C = __ldg(q + 1);
    } else {
        unsigned w[3] = {0, 0, 0};
#pragma unroll
        for (int b = 1; b < 12; b++) {                       // bytes x-3 .. x+7 (x-4 is never used)
            const int xx = x - 4 + b;
            const int sx = xx < W + 3 ? reflect101(xx, W) : 0;
            w[b >> 2] |= (unsigned)row[sx] << (8 * (b & 3));
        }
        A = w[0]; B = w[1]; C = w[2];
    }
}

// One strip (4 columns at x, output rows y0 .. y0+nOut-1) of one level.
template <bool INTERIOR>
__device__ __forceinline__ void blur_strip(const uint8_t* __restrict__ src, int pitch, uint8_t* __restrict__ dst, int dpitch,
                                           int x, int y0, int W, int H) {
    const unsigned WA = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps 0..3
    const unsigned WB = 48u | (34u << 8) | (18u << 16);                    // taps 4..6
    unsigned Q[8][4];           // Q[j & 7][k] = (H_j[k], H_{j-1}[k]) as u16 pairs, k = column in the strip
    unsigned prevLo = 0, prevHi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) Q[i][k] = 0;

    const int nOut = min(BL_ROWS, H - y0);
    const bool yInside = y0 >= 3 && y0 + nOut + 3 <= H;          // no row reflection needed
#pragma unroll 1
    for (int jb = 0; jb < BL_ROWS + 8; jb += 8) {
        if (jb >= nOut + 6) break;
        // issue the loads of all 8 rows of the block first: 24 independent words in flight per thread
        unsigned Aw[8], Bw[8], Cw[8];
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            int sy = y0 - 3 + jb + jj;
            if (!yInside) sy = reflect101(sy, H);
            load_window(src + (size_t)sy * pitch, x, W, INTERIOR, Aw[jj], Bw[jj], Cw[jj]);
        }
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int j = jb + jj;                             // staged row j <-> level row y0 - 3 + j
            const unsigned A = Aw[jj], B = Bw[jj], C = Cw[jj];
            // horizontal pass: output k needs bytes k+1 .. k+7 of the 12-byte window
            unsigned hs[4];
            hs[0] = __dp4a(__byte_perm(B, C, 0x4321), WB, __dp4a(__byte_perm(A, B, 0x4321), WA, 0u));
            hs[1] = __dp4a(__byte_perm(B, C, 0x5432), WB, __dp4a(__byte_perm(A, B, 0x5432), WA, 0u));
            hs[2] = __dp4a(__byte_perm(B, C, 0x6543), WB, __dp4a(__byte_perm(A, B, 0x6543), WA, 0u));
            hs[3] = __dp4a(C, WB, __dp4a(B, WA, 0u));
            const unsigned curLo = hs[0] | (hs[1] << 16), curHi = hs[2] | (hs[3] << 16);
            unsigned* q = Q[jj];
            q[0] = __byte_perm(curLo, prevLo, 0x5410);
            q[1] = __byte_perm(curLo, prevLo, 0x7632);
            q[2] = __byte_perm(curHi, prevHi, 0x5410);
            q[3] = __byte_perm(curHi, prevHi, 0x7632);
            prevLo = curLo; prevHi = curHi;
            // vertical pass for output row yo = y0 + j - 6 (rows yo-3 .. yo+3 are staged rows j-6 .. j)
            const int yo = j - 6;
            if (yo >= 0 && yo < nOut) {
                const unsigned* q3 = Q[jj];                    // (H_{y+3}, H_{y+2})
                const unsigned* q1 = Q[(jj + 6) & 7];          // (H_{y+1}, H_y)
                const unsigned* qm1 = Q[(jj + 4) & 7];         // (H_{y-1}, H_{y-2})
                const unsigned* qm3 = Q[(jj + 2) & 7];         // (H_{y-3}, H_{y-4}): second weight 0
                unsigned acc[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    unsigned a = __dp2a_lo(q3[k], 18u | (34u << 8), 32768u);
                    a = __dp2a_lo(q1[k], 48u | (56u << 8), a);
                    a = __dp2a_lo(qm1[k], 48u | (34u << 8), a);
                    a = __dp2a_lo(qm3[k], 18u, a);
                    acc[k] = a;                                // < 2^24: the output byte is byte 2
                }
                const unsigned o = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
                *reinterpret_cast<uint32_t*>(dst + (size_t)(y0 + yo) * dpitch) = o;
            }
        }
    }
}

// Grid: blurEdgeCtas CTAs first, then blurTilesTotal CTAs of interior strips (128 x 128 pixel tiles).  The edge CTAs are those
// whose threads each take one border strip (the first strip of a row band and the strips that touch
// the right border) -- the reflecting loads are kept out of the interior warps.
__global__ void __launch_bounds__(32 * BL_BANDS) k_blur(const __grid_constant__ Geom g, const PyrPtrs p,
                                                         uint8_t* __restrict__ blurSlab, size_t blurStride) {
    const int img = blockIdx.y;
    int t = (int)blockIdx.x - g.blurEdgeCtas;          // the (slower) border CTAs are scheduled first
    if (t >= 0) {
        int level = 0;
#pragma unroll 1
        for (int l = 1; l < g.nlevels; l++) if (t >= g.lv[l].blurTileBase) level = l;
        const LevelGeom& lg = g.lv[level];
        t -= lg.blurTileBase;
        const int ty = t / lg.blurTilesX, tx = t - ty * lg.blurTilesX;
        const int W = lg.w, H = lg.h;
        const int x = tx * BL_TW + threadIdx.x * 4;               // first column of this thread's strip
        const int y0 = ty * BL_TH + threadIdx.y * BL_ROWS;        // first output row
        if (x < 4 || x + 8 > W || y0 >= H) return;                // border strips belong to the edge CTAs
        int pitch;
        const uint8_t* src = level_ptr(p, g, img, level, pitch);
        blur_strip<true>(src, pitch, blurSlab + (size_t)img * blurStride + lg.off + x, lg.pitch, x, y0, W, H);
    } else {
        int e = (int)blockIdx.x * (32 * BL_BANDS) + threadIdx.y * 32 + threadIdx.x;
        int level = -1;
#pragma unroll 1
        for (int l = 0; l < g.nlevels; l++) if (e >= g.lv[l].blurEdgeBase) level = l;
        if (level < 0) return;
        const LevelGeom& lg = g.lv[level];
        e -= lg.blurEdgeBase;
        const int W = lg.w, H = lg.h;
        // border strips of a level: strip 0 and every strip with x + 8 > W (at most 2 of them)
        const int nStrips = (W + 3) >> 2;
        const int firstRight = max(1, (W - 8 + 4) >> 2);           // first strip index s >= 1 with 4 s + 8 > W
        const int perBand = 1 + max(0, nStrips - firstRight);
        const int band = e / perBand, k = e - band * perBand;
        const int y0 = band * BL_ROWS;
        if (y0 >= H) return;
        const int strip = k == 0 ? 0 : firstRight + k - 1;
        const int x = strip * 4;
        int pitch;
        const uint8_t* src = level_ptr(p, g, img, level, pitch);
        blur_strip<false>(src, pitch, blurSlab + (size_t)img * blurStride + lg.off + x, lg.pitch, x, y0, W, H);
    }
}

}  // namespace

cudaError_t launch_blur(const Geom& g, PyrPtrs p, uint8_t* blurSlab, size_t blurStride, int nimg, cudaStream_t st) {
    dim3 grid(g.blurTilesTotal + g.blurEdgeCtas, nimg);
    dim3 block(32, BL_BANDS);
    k_blur<<<grid, block, 0, st>>>(g, p, blurSlab, blurStride);
    return cudaGetLastError();
}
