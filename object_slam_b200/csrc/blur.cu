// K5 -- 7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, of every pyramid level.  Replaces the
// cv::GaussianBlur call at src/ORBextractor.cc:1085-1086.  OpenCV (>= 3.4.2 / 4.x) runs 8-bit
// blurs in fixed point: per-axis kernel [18,34,48,56,48,34,18]/256, a 16-bit horizontal pass and
// a vertical pass rounded once, out = (sum + 2^15) >> 16; this kernel reproduces those bytes.
// The blurred level is only read by the descriptor stage.
//
// The stage is instruction-issue bound before it is HBM bound, so it is built around integer dot
// products: a thread owns a 4-pixel column strip and walks down the rows of a tile staged in shared
// memory, forms the four horizontal sums with IDP.4A (two dot products of 4 bytes each), keeps them
// packed as u16 pairs of vertically adjacent rows, and produces each output pixel with four IDP.2A
// over an 8-row register window.
#include "kernels.h"

namespace {

#ifndef BL_OPT_STAGE
#define BL_OPT_STAGE 1
#endif
constexpr int BL_ROWS = 32;                 // output rows per thread
constexpr int BL_BANDS = 4;                 // row bands per CTA
constexpr int BL_TW = 128, BL_TH = BL_ROWS * BL_BANDS;      // CTA tile: 128 x 128 pixels

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// Tiled kernel: a CTA stages the 128 x 128 tile plus its 3-pixel apron in shared memory (16-byte cp.async copies, all in flight at once; rows
// and columns outside the level are reflected while staging, so the filter loop has no border case), then a
// thread owns a 4-pixel column strip of a 32-row band and walks down the rows: 3 LDS, 8 IDP.4A for the four
// horizontal sums, rows paired as u16x2, 4 IDP.2A per output pixel over an 8-row register window.
constexpr int BT_PITCH = BL_TW + 32;        // staged columns tx0 - 16 .. tx0 + 143
constexpr int BT_ROWS = BL_TH + 6;

__global__ void __launch_bounds__(32 * BL_BANDS) k_blur_tile(const __grid_constant__ Geom g, const PyrPtrs p,
                                                              uint8_t* __restrict__ blurSlab, size_t blurStride) {
    extern __shared__ __align__(16) uint8_t tile[];
    pdl_entry();
    const int img = blockIdx.y, tid = threadIdx.y * 32 + threadIdx.x;
    int t = (int)blockIdx.x, level = 0;
#pragma unroll 1
    for (int l = 1; l < g.nlevels; l++) if (t >= g.lv[l].blurTileBase) level = l;
    const LevelGeom& lg = g.lv[level];
    t -= lg.blurTileBase;
    const int ty = t / lg.blurTilesX, tx = t - ty * lg.blurTilesX;
    const int W = lg.w, H = lg.h;
    const int tx0 = tx * BL_TW, ty0 = ty * BL_TH;
    int pitch;
    const uint8_t* src = level_ptr(p, g, img, level, pitch);
    const int nRows = min(BT_ROWS, H + 3 - (ty0 - 3));             // staged rows: level rows ty0 - 3 .. min(ty0 + 130, H + 2)
    {
        // 10 vectors per row; vectors that start outside [0, pitch) are skipped (their bytes are either unused or
        // written by the reflection pass below)
        const int v = tid % 10, r0 = tid / 10;                      // 12 rows per sweep, threads 120..127 idle
        const int gx = tx0 - 16 + 16 * v;
        if (r0 < 12 && gx >= 0 && gx + 16 <= pitch) {
#if BL_OPT_STAGE
            // every row of the thread's column of vectors is requested before the first one is waited for (cp.async, no registers):
            // one DRAM round trip per CTA instead of one per sweep of 12 rows
            uint32_t tp = smem_u32(tile + r0 * BT_PITCH + 16 * v);
            for (int r = r0; r < nRows; r += 12, tp += 12 * BT_PITCH) {
                const uint8_t* sp = src + (size_t)reflect101(ty0 - 3 + r, H) * pitch + gx;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(tp), "l"(sp) : "memory");
            }
#else
            uint8_t* tp = tile + r0 * BT_PITCH + 16 * v;
            if (ty0 >= 3 && ty0 - 3 + nRows <= H) {                // no row reflection in this tile
                const uint8_t* sp = src + (size_t)(ty0 - 3 + r0) * pitch + gx;
                const size_t step = (size_t)pitch * 12;
                for (int r = r0; r < nRows; r += 12, sp += step, tp += 12 * BT_PITCH)
                    *reinterpret_cast<uint4*>(tp) = __ldg(reinterpret_cast<const uint4*>(sp));
            } else {
                for (int r = r0; r < nRows; r += 12, tp += 12 * BT_PITCH)
                    *reinterpret_cast<uint4*>(tp) = __ldg(reinterpret_cast<const uint4*>(src + (size_t)reflect101(ty0 - 3 + r, H) * pitch + gx));
            }
#endif
        }
#if BL_OPT_STAGE
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
    }
    const bool border = tx0 == 0 || tx0 + BL_TW + 8 > W;          // CTA-uniform
    if (border) {
        __syncthreads();
        // columns -3..-1 and W.. of the tile, byte-wise from the level
        for (int r = tid; r < nRows; r += 32 * BL_BANDS) {
            const uint8_t* row = src + (size_t)reflect101(ty0 - 3 + r, H) * pitch;
            uint8_t* trow = tile + r * BT_PITCH + 16 - tx0;
            if (tx0 == 0) for (int c = -3; c < 0; c++) trow[c] = row[reflect101(c, W)];
            for (int c = max(W, tx0); c < min(W + 7, tx0 + BL_TW + 8); c++) trow[c] = c < W + 3 ? row[reflect101(c, W)] : 0;
        }
    }
    __syncthreads();

    const int x = tx0 + 4 * threadIdx.x, y0 = ty0 + threadIdx.y * BL_ROWS;
    if (x >= W || y0 >= H) return;
    const int nOut = min(BL_ROWS, H - y0);
    const uint32_t* col = reinterpret_cast<const uint32_t*>(tile + threadIdx.y * BL_ROWS * BT_PITCH + 16 + 4 * threadIdx.x);
    const size_t dpitch = (size_t)lg.pitch;
    uint8_t* dst = blurSlab + (size_t)img * blurStride + lg.off + (size_t)y0 * dpitch + x - 6 * dpitch;   // row j - 6 of the band

    const unsigned WA = 18u | (34u << 8) | (48u << 16) | (56u << 24);     // taps 0..3
    const unsigned WB = 48u | (34u << 8) | (18u << 16);                    // taps 4..6
    unsigned Q[8][4];           // Q[j & 7][k] = (H_j[k], H_{j-1}[k]) as u16 pairs, k = column in the strip
    unsigned prevLo = 0, prevHi = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 4; k++) Q[i][k] = 0;
#pragma unroll 1
    for (int jb = 0; jb < BL_ROWS + 8; jb += 8) {
        if (jb >= nOut + 6) break;
#pragma unroll
        for (int jj = 0; jj < 8; jj++) {
            const int j = jb + jj;                             // staged row j of the band <-> level row y0 - 3 + j
            const uint32_t* q0 = col + j * (BT_PITCH / 4);
            const unsigned A = q0[-1], B = q0[0], C = q0[1];
            // horizontal pass: output k needs bytes k+1 .. k+7 of the 12-byte window
            unsigned hs[4];
            hs[0] = __dp4a(__byte_perm(B, C, 0x4321), WB, __dp4a(__byte_perm(A, B, 0x4321), WA, 0u));
            hs[1] = __dp4a(__byte_perm(B, C, 0x5432), WB, __dp4a(__byte_perm(A, B, 0x5432), WA, 0u));
            hs[2] = __dp4a(__byte_perm(B, C, 0x6543), WB, __dp4a(__byte_perm(A, B, 0x6543), WA, 0u));
            hs[3] = __dp4a(C, WB, __dp4a(B, WA, 0u));
            const unsigned curLo = __byte_perm(hs[0], hs[1], 0x5410), curHi = __byte_perm(hs[2], hs[3], 0x5410);
            unsigned* q = Q[jj];
            q[0] = __byte_perm(curLo, prevLo, 0x5410);
            q[1] = __byte_perm(curLo, prevLo, 0x7632);
            q[2] = __byte_perm(curHi, prevHi, 0x5410);
            q[3] = __byte_perm(curHi, prevHi, 0x7632);
            prevLo = curLo; prevHi = curHi;
            // vertical pass for output row yo = j - 6 of the band (rows yo-3 .. yo+3 are staged rows j-6 .. j)
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
                *reinterpret_cast<uint32_t*>(dst + dpitch * (unsigned)jj) = o;
            }
        }
        dst += 8 * dpitch;
    }
}

}  // namespace

cudaError_t launch_blur(const Geom& g, PyrPtrs p, uint8_t* blurSlab, size_t blurStride, int nimg, cudaStream_t st) {
    if (g.blurTilesTotal == 0) return cudaSuccess;
    dim3 grid(g.blurTilesTotal, nimg);
    dim3 block(32, BL_BANDS);
    /* the 8-row blocks of the loop touch two rows past the apron */
    cudaError_t le = launch_k(pdl_enabled(), k_blur_tile, grid, block, (size_t)(BT_PITCH * (BT_ROWS + 2)), st, g, p, blurSlab, blurStride);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}
