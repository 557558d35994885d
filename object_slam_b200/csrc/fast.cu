// K2 -- per-cell FAST-9/16 with threshold fallback.  Replaces the cell loop of
// ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:765-829) and the two cv::FAST
// calls inside it (:809, :814).
//
// One CTA owns one 30x30-ish cell of one level of one image: it stages the cell's sub-image
// (cell + 6 overlap px, the window the reference hands to cv::FAST) in shared memory, computes
// the exact FAST score (max over the sixteen 9-arcs of the min contrast, minus 1) of every
// interior pixel once -- one score map serves both thresholds, a pixel being a corner at t iff
// score >= t -- then applies OpenCV's strict 3x3 non-max suppression inside the cell at iniThFAST,
// falls back to minThFAST when nothing survives (the decision is made after suppression, as in
// the reference), and emits the survivors in row-major order into the cell's slot.  Interiors
// of neighbouring cells tile the level without overlap, so every pixel is scored exactly once.
#include "kernels.h"

namespace {

constexpr int TILE_PITCH = 80;       // bytes per staged row (sub-image <= 66 px + <= 3 px alignment slack, word padded)
constexpr int TILE_ROWS = 66;
constexpr int SC_PITCH = 64;         // score rows: interior <= 60 px + 1-px zero apron each side
constexpr int SC_ROWS = 62;
constexpr int FAST_THREADS = 128;

// Exact FAST-9/16 score of the pixel at c (shared memory, row pitch TILE_PITCH):
// 0 if the pixel is not a corner at threshold t, else (corner contrast - 1) >= t.
__device__ __forceinline__ int fast_score(const uint8_t* c, int t) {
    const int v = c[0];
    const int lo = v - t, hi = v + t;
    const int r0 = c[3 * TILE_PITCH], r8 = c[-3 * TILE_PITCH], r4 = c[3], r12 = c[-3];
    // every 9-arc holds one of ring pixels {0,8} and one of {4,12}
    const bool dk = ((r0 < lo) | (r8 < lo)) & ((r4 < lo) | (r12 < lo));
    const bool br = ((r0 > hi) | (r8 > hi)) & ((r4 > hi) | (r12 > hi));
    if (!(dk | br)) return 0;
    int d[16];
    d[0] = v - r0;
    d[1] = v - c[3 * TILE_PITCH + 1];
    d[2] = v - c[2 * TILE_PITCH + 2];
    d[3] = v - c[1 * TILE_PITCH + 3];
    d[4] = v - r4;
    d[5] = v - c[-1 * TILE_PITCH + 3];
    d[6] = v - c[-2 * TILE_PITCH + 2];
    d[7] = v - c[-3 * TILE_PITCH + 1];
    d[8] = v - r8;
    d[9] = v - c[-3 * TILE_PITCH - 1];
    d[10] = v - c[-2 * TILE_PITCH - 2];
    d[11] = v - c[-1 * TILE_PITCH - 3];
    d[12] = v - r12;
    d[13] = v - c[1 * TILE_PITCH - 3];
    d[14] = v - c[2 * TILE_PITCH - 2];
    d[15] = v - c[3 * TILE_PITCH - 1];
    // sliding minimum / maximum over 9 consecutive ring positions by doubling (2,4,8,+1)
    int lo2[16], hi2[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { lo2[i] = min(d[i], d[(i + 1) & 15]); hi2[i] = max(d[i], d[(i + 1) & 15]); }
    int lo4[16], hi4[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { lo4[i] = min(lo2[i], lo2[(i + 2) & 15]); hi4[i] = max(hi2[i], hi2[(i + 2) & 15]); }
    int a = -256, b = 256;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int m9 = min(min(lo4[i], lo4[(i + 4) & 15]), d[(i + 8) & 15]);
        const int x9 = max(max(hi4[i], hi4[(i + 4) & 15]), d[(i + 8) & 15]);
        a = max(a, m9);        // ring darker than the centre
        b = min(b, x9);        // ring brighter than the centre
    }
    const int contrast = max(a, -b);
    return contrast > t ? contrast - 1 : 0;
}

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(const __grid_constant__ Geom g, const PyrPtrs p,
                                                             uint32_t* __restrict__ cand, int* __restrict__ cellCount) {
    __shared__ __align__(16) uint8_t tile[TILE_ROWS * TILE_PITCH];
    __shared__ __align__(16) uint8_t score[SC_ROWS * SC_PITCH];
    __shared__ uint8_t flags[60 * 60 + 32];
    __shared__ int chunkOfs[128];
    __shared__ int sAny;

    const int tid = threadIdx.x;
    const int img = blockIdx.y;
    int cell = blockIdx.x;
    int level = 0;
#pragma unroll 1
    for (int l = 1; l < g.nlevels; l++) if (cell >= g.lv[l].cellBase) level = l;
    const LevelGeom& lg = g.lv[level];
    cell -= lg.cellBase;
    const int ci = cell / lg.nCols, cj = cell - ci * lg.nCols;
    int* countOut = cellCount + (size_t)img * g.nCellsTotal + lg.cellBase + cell;

    // cell window, :789-806 (all values are small integers, so the reference's float arithmetic is exact)
    const int maxBX = lg.w - OBS_BORDER, maxBY = lg.h - OBS_BORDER;
    const int X0 = OBS_BORDER + cj * lg.wCell, Y0 = OBS_BORDER + ci * lg.hCell;
    const int X1 = min(X0 + lg.wCell + 6, maxBX), Y1 = min(Y0 + lg.hCell + 6, maxBY);
    const int sw = X1 - X0, sh = Y1 - Y0;
    if (Y0 >= maxBY - 3 || X0 >= maxBX - 6 || sw < 7 || sh < 7) {     // skipped cell, or too small for FAST
        if (tid == 0) *countOut = 0;
        return;
    }
    const int wInt = sw - 6, hInt = sh - 6;

    // stage the window with aligned 32-bit loads
    int pitch;
    const uint8_t* base = level_ptr(p, g, img, level, pitch);
    const int xa = X0 & ~3;                       // aligned start column
    const int shift = X0 - xa;
    const int nWords = (shift + sw + 3) >> 2;     // <= 18
    for (int i = tid; i < sh * nWords; i += FAST_THREADS) {
        const int r = i / nWords, wv = i - r * nWords;
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(base + (size_t)(Y0 + r) * pitch + xa) + wv);
        *reinterpret_cast<uint32_t*>(tile + r * TILE_PITCH + 4 * wv) = v;
    }
    for (int i = tid; i < SC_ROWS * SC_PITCH / 4; i += FAST_THREADS) reinterpret_cast<uint32_t*>(score)[i] = 0;
    if (tid == 0) sAny = 0;
    __syncthreads();

    const int tLow = min(g.iniTh, g.minTh);
    const int nInt = wInt * hInt;
    for (int i = tid; i < nInt; i += FAST_THREADS) {
        const int iy = i / wInt, ix = i - iy * wInt;
        const int s = fast_score(tile + (iy + 3) * TILE_PITCH + shift + ix + 3, tLow);
        score[(iy + 1) * SC_PITCH + ix + 1] = (uint8_t)s;
    }
    __syncthreads();

    // 3x3 strict non-max suppression at both thresholds (pixels outside the interior count 0)
    const int tIni = g.iniTh, tMin = g.minTh;
    int anyIni = 0;
    for (int i = tid; i < nInt; i += FAST_THREADS) {
        const int iy = i / wInt, ix = i - iy * wInt;
        const uint8_t* sc = score + (iy + 1) * SC_PITCH + ix + 1;
        const int s = sc[0];
        int f = 0;
        if (s > 0) {
            int mIni = 0, mMin = 0;      // max neighbour score among corners at the initial / the minimum threshold
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) {
                    if (dx == 0 && dy == 0) continue;
                    const int n = sc[dy * SC_PITCH + dx];
                    mIni = max(mIni, n >= tIni ? n : 0);
                    mMin = max(mMin, n >= tMin ? n : 0);
                }
            if (s >= tIni && s > mIni) f |= 1;
            if (s >= tMin && s > mMin) f |= 2;
        }
        flags[i] = (uint8_t)f;
        anyIni |= f & 1;
    }
    if (anyIni) sAny = 1;
    __syncthreads();
    const int useBit = sAny ? 1 : 2;

    // ordered compaction (row-major inside the cell, like cv::FAST's output)
    const int lane = tid & 31, warp = tid >> 5;
    const int nChunks = (nInt + 31) >> 5;            // <= 113
    for (int c = warp; c < nChunks; c += FAST_THREADS / 32) {
        const int i = c * 32 + lane;
        const bool on = i < nInt && (flags[i] & useBit);
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (lane == 0) chunkOfs[c] = __popc(b);
    }
    __syncthreads();
    if (warp == 0) {                                  // exclusive scan of <= 128 chunk counts
        int v[4], s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int c = lane * 4 + k; v[k] = c < nChunks ? chunkOfs[c] : 0; s += v[k]; }
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        int run = incl - s;
#pragma unroll
        for (int k = 0; k < 4; k++) { const int c = lane * 4 + k; if (c < nChunks) chunkOfs[c] = run; run += v[k]; }
        if (lane == 31) *countOut = incl;
    }
    __syncthreads();
    uint32_t* slot = cand + (size_t)img * g.slotTotal + lg.slotBase + (size_t)cell * lg.cellCap;
    for (int c = warp; c < nChunks; c += FAST_THREADS / 32) {
        const int i = c * 32 + lane;
        const bool on = i < nInt && (flags[i] & useBit);
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (on) {
            const int iy = i / wInt, ix = i - iy * wInt;
            const int s = score[(iy + 1) * SC_PITCH + ix + 1];
            const int o = chunkOfs[c] + __popc(b & ((1u << lane) - 1));
            // coordinates relative to the border origin: (X0 - 16) + (ix + 3), :820-825
            slot[o] = pack_key(X0 - OBS_BORDER + ix + 3, Y0 - OBS_BORDER + iy + 3, s);
        }
    }
}

}  // namespace

cudaError_t launch_fast(const Geom& g, PyrPtrs p, uint32_t* cand, int* cellCount, int nimg, cudaStream_t st) {
    if (g.nCellsTotal == 0) return cudaSuccess;
    dim3 grid(g.nCellsTotal, nimg);
    k_fast_cells<<<grid, FAST_THREADS, 0, st>>>(g, p, cand, cellCount);
    return cudaGetLastError();
}
