// K2 -- per-cell FAST-9/16 with threshold fallback.  Replaces the cell loop of
// ORBextractor::ComputeKeyPointsOctTree (src/ORBextractor.cc:765-829) and the two cv::FAST
// calls inside it (:809, :814).
//
// Semantics kept from the reference: cv::FAST runs on each 30x30-ish cell window (cell + 6 overlap
// px); corners live in the window's interior (3-px margin), the interiors of neighbouring cells
// tile the level without overlap; OpenCV's score of a corner is (max over the sixteen 9-arcs of
// the min |contrast|) - 1, so one score map serves both thresholds (corner at t  <=>  score >= t);
// the 3x3 non-max suppression is strict and sees only the cell's own interior; a cell whose
// suppressed result at iniThFAST is empty is redone at minThFAST (decided after suppression);
// output order is row-major inside the cell.  Because a score >= iniThFAST beats every neighbour
// below iniThFAST anyway, "local maximum at T" == "score >= T and local maximum of the raw map",
// so one suppression pass serves both thresholds.
//
// Shape of the kernel (issue-bound work, so it is organised around instructions per pixel):
//   * one CTA owns a run of up to 8 cells of one cell row and stages their window with 128-bit loads;
//   * pass 1 tests 8 pixels per lane with byte-SIMD |a-b| (VABSDIFF4) against the four compass ring
//     pixels (a necessary condition for a 9-arc), leaves one bit per pixel in a bitmap, and pass 1b
//     expands the bitmap into a dense queue of pixels;
//   * pass 2 gives each queued pixel to one lane: exact score with both polarities packed as u16x2
//     (v-r+255, r-v+255) through a min/max network of VIMNMX(3).U16x2; corners are re-queued densely;
//   * pass 3 suppresses non-maxima among the corners only and re-queues the survivors; passes 4-6
//     apply the per-cell threshold fallback and write the survivors in row-major order per cell,
//     positions coming from popcounts over a bitmap of the emitted pixels.
#include "kernels.h"

namespace {

constexpr int FT_PITCH = 256;          // bytes per staged row
constexpr int FT_ROWS = 66;            // hCell <= 60, + 6
constexpr int FT_THREADS = 256;
constexpr int FT_LIST = 8192;          // queue capacity == max interior pixels per CTA (host enforces)
constexpr int FT_MAXCELLS = 8;
constexpr int FT_SEG = FT_LIST / 8 + FT_PITCH;   // capacity of one warp's queue: its share of the rows, rounded up by one row

__host__ __device__ inline int fast_cells_per_cta(int wCell, int hCell) {
    int cg = (FT_PITCH - 15 - 6) / wCell;
    const int byList = FT_LIST / (wCell * hCell);
    if (cg > byList) cg = byList;
    if (cg > FT_MAXCELLS) cg = FT_MAXCELLS;
    return cg < 1 ? 1 : cg;
}

// Exact FAST-9/16 corner contrast of the pixel at c (staged tile): max over the 16 arcs of 9
// contiguous ring pixels of the minimum of (v - r) [ring darker] and of (r - v) [ring brighter].
// Both polarities ride in one register as two 16-bit lanes, biased by +255 so that one IMAD per
// ring pixel builds the pair: r * 0xFFFF + (v + 255 | (255 - v) << 16) = (v - r + 255, r - v + 255).
__device__ __forceinline__ int fast_contrast(const uint8_t* c) {
    const unsigned v = c[0];
    const unsigned Vc = (v + 255u) | ((255u - v) << 16);
    unsigned P[16];
#define RING(k, dx, dy) P[k] = (unsigned)c[(dy) * FT_PITCH + (dx)] * 0xFFFFu + Vc
    RING(0, 0, 3); RING(1, 1, 3); RING(2, 2, 2); RING(3, 3, 1); RING(4, 3, 0); RING(5, 3, -1); RING(6, 2, -2); RING(7, 1, -3);
    RING(8, 0, -3); RING(9, -1, -3); RING(10, -2, -2); RING(11, -3, -1); RING(12, -3, 0); RING(13, -3, 1); RING(14, -2, 2); RING(15, -1, 3);
#undef RING
    unsigned m2[16], m4[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m2[i] = __vminu2(P[i], P[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; i++) m4[i] = __vminu2(m2[i], m2[(i + 2) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const unsigned a = __vimin3_u16x2(m4[i], m4[(i + 4) & 15], P[(i + 8) & 15]);
        const unsigned b = __vimin3_u16x2(m4[i + 1], m4[(i + 5) & 15], P[(i + 9) & 15]);
        best = __vimax3_u16x2(best, a, b);
    }
    return (int)max(best & 0xffffu, best >> 16) - 255;
}

// number of set bits of a 256-bit row bitmap (8 words) inside columns [a, b), a < b
__device__ __forceinline__ int row_bits(const uint32_t* bm, int a, int b) {
    const int wa = a >> 5, wb = (b - 1) >> 5;
    const uint32_t ma = 0xffffffffu << (a & 31), mb = 0xffffffffu >> (31 - ((b - 1) & 31));
    if (wa == wb) return __popc(bm[wa] & ma & mb);
    int n = __popc(bm[wa] & ma) + __popc(bm[wb] & mb);
    for (int w = wa + 1; w < wb; w++) n += __popc(bm[w]);
    return n;
}

// per byte: 0x80 where the byte of y exceeds t (t < 128, K = (127 - t) * 0x01010101)
__device__ __forceinline__ unsigned over_threshold(unsigned y, unsigned K) { return ((y & 0x7f7f7f7fu) + K) | y; }

__global__ void __launch_bounds__(FT_THREADS) k_fast_cells(const __grid_constant__ Geom g, const PyrPtrs p,
                                                           const FastCta* __restrict__ ctaTab, int tileRows,
                                                           uint32_t* __restrict__ cand, int* __restrict__ cellCount) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* tile = sm;                                   // tileRows x FT_PITCH pixels
    uint8_t* score = sm + tileRows * FT_PITCH;            // same geometry, 0 = not a corner
    uint16_t* list = reinterpret_cast<uint16_t*>(sm + 2 * tileRows * FT_PITCH);    // 8 warp-private pixel queues: row << 8 | col
    __shared__ uint32_t bitmap[FT_ROWS][8];               // emitted keypoints, one bit per tile pixel
    __shared__ uint16_t rowOfs[FT_MAXCELLS][FT_ROWS];     // per cell: keypoints in the rows above
    __shared__ uint8_t colCell[FT_PITCH], colIn[FT_PITCH];   // tile column -> cell of the run, column inside the cell
    __shared__ int sCellAny[FT_MAXCELLS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1;
    const int img = blockIdx.y;
    const FastCta cta = ctaTab[blockIdx.x];
    const int level = cta.level, ci = cta.ci, j0 = cta.j0, nCellsHere = cta.n;
    const LevelGeom& lg = g.lv[level];
    int* countOut = cellCount + (size_t)img * g.nCellsTotal + lg.cellBase + ci * lg.nCols + j0;

    // window of the cell run (:789-806): interiors tile [19, maxB-3) in both axes
    const int maxBX = lg.w - OBS_BORDER, maxBY = lg.h - OBS_BORDER;
    const int wCell = lg.wCell, hCell = lg.hCell;
    const int X0 = OBS_BORDER + j0 * wCell, Y0 = OBS_BORDER + ci * hCell;
    const int X1 = min(X0 + nCellsHere * wCell + 6, maxBX), Y1 = min(Y0 + hCell + 6, maxBY);
    const int sw = X1 - X0, sh = Y1 - Y0;
    if (sw < 7 || sh < 7) {                               // no interior pixel: every cell of the run is empty
        if (tid < nCellsHere) countOut[tid] = 0;
        return;
    }
    const int xa = X0 & ~15;                              // tile column 0 <-> level column xa
    const int cLo = X0 + 3 - xa, cHi = X1 - 3 - xa;       // interior columns in tile coordinates
    const int rLo = 3, rHi = sh - 3;                      // interior rows

    // ---- stage the window (128-bit loads), clear the maps, build the column tables
    {
        int pitch;
        const uint8_t* base = level_ptr(p, g, img, level, pitch);
        const int nVec = (X1 - xa + 15) >> 4;             // <= 16
        for (int i = tid; i < sh * 16; i += FT_THREADS) {
            const int r = i >> 4, v = i & 15;
            if (v < nVec) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(Y0 + r) * pitch + xa) + v);
                *reinterpret_cast<uint4*>(tile + r * FT_PITCH + 16 * v) = q;
            }
            *reinterpret_cast<uint4*>(score + r * FT_PITCH + 16 * v) = make_uint4(0, 0, 0, 0);
        }
        for (int i = tid; i < FT_ROWS * 8; i += FT_THREADS) (&bitmap[0][0])[i] = 0;
        const int xr = max(tid + xa - OBS_EDGE, 0);       // column relative to the level's first interior column
        const int cc = xr / wCell;
        colCell[tid] = (uint8_t)min(max(cc - j0, 0), FT_MAXCELLS - 1);
        colIn[tid] = (uint8_t)(xr - cc * wCell);
        if (tid < FT_MAXCELLS) sCellAny[tid] = 0;
    }
    __syncthreads();

    // Each warp owns the interior rows rLo + warp, + 8, ... and a private queue: no atomics, and
    // every later filtering step compacts the queue in place.
    uint16_t* seg = list + warp * FT_SEG;

    // ---- pass 1: compass-point rejection, 8 pixels per lane, one row per iteration, survivors queued.
    // Every 9-arc holds one of ring pixels {0,8} and one of {4,12}; (|d0| | |d8|) > t is implied by
    // |d0| > t or |d8| > t, so "(|d0| | |d8|) > t and (|d4| | |d12|) > t" is a necessary condition.
    const int tLow = min(g.iniTh, g.minTh);
    int nQ = 0;
    {
        const int w0 = (cLo >> 2) & ~1, w1 = (cHi + 3) >> 2;     // even-aligned word range covering the interior columns
        const unsigned K = (unsigned)(127 - min(tLow, 127)) * 0x01010101u;
        const int wA = w0 + 2 * lane;
        const int c0 = wA * 4;
        unsigned edgeMask = wA < w1 ? 0xffu : 0u;
        if (c0 < cLo) edgeMask &= 0xffu << (cLo - c0);
        if (c0 + 8 > cHi && c0 < cHi) edgeMask &= 0xffu >> (c0 + 8 - cHi);
        for (int r = rLo + warp; r < rHi; r += FT_THREADS / 32) {
            unsigned m8 = 0;
            if (edgeMask) {
                const uint32_t* row = reinterpret_cast<const uint32_t*>(tile + r * FT_PITCH);
                const uint2 C = *reinterpret_cast<const uint2*>(row + wA);
                const uint2 up = *reinterpret_cast<const uint2*>(row + wA + 3 * (FT_PITCH / 4));
                const uint2 dn = *reinterpret_cast<const uint2*>(row + wA - 3 * (FT_PITCH / 4));
                const unsigned L = row[wA - 1], R = row[wA + 2];
                unsigned f0, f1;
                if (tLow < 128) {
                    const unsigned y0 = __vabsdiffu4(C.x, up.x) | __vabsdiffu4(C.x, dn.x);
                    const unsigned z0 = __vabsdiffu4(C.x, __byte_perm(L, C.x, 0x4321)) | __vabsdiffu4(C.x, __byte_perm(C.x, C.y, 0x6543));
                    const unsigned y1 = __vabsdiffu4(C.y, up.y) | __vabsdiffu4(C.y, dn.y);
                    const unsigned z1 = __vabsdiffu4(C.y, __byte_perm(C.x, C.y, 0x4321)) | __vabsdiffu4(C.y, __byte_perm(C.y, R, 0x6543));
                    f0 = over_threshold(y0, K) & over_threshold(z0, K) & 0x80808080u;
                    f1 = over_threshold(y1, K) & over_threshold(z1, K) & 0x80808080u;
                } else {
                    f0 = f1 = 0x80808080u;                 // thresholds >= 128: no cheap rejection, score everything
                }
                // gather the 8 flag bits (bit 7 of each byte)
                m8 = ((((f0 >> 7) * 0x00204081u) >> 21 & 0xfu) | (((f1 >> 7) * 0x00204081u) >> 17 & 0xf0u)) & edgeMask;
            }
            const int n = __popc(m8);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            uint16_t* out = seg + nQ + incl - n;
            const int e0 = (r << 8) | c0;
            while (m8) {
                const int b = __ffs(m8) - 1;
                m8 &= m8 - 1;
                *out++ = (uint16_t)(e0 + b);
            }
            nQ += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncwarp();

    // ---- pass 2: exact score of the queued pixels, one lane each; corners stay queued
    int nC = 0;
    for (int i0 = 0; i0 < nQ; i0 += 32) {
        const int i = i0 + lane;
        int e = 0, s = 0;
        if (i < nQ) {
            e = seg[i];
            const int contrast = fast_contrast(tile + (e >> 8) * FT_PITCH + (e & 255));
            if (contrast > tLow && contrast > 1) s = contrast - 1;            // OpenCV: score = corner contrast - 1
        }
        const unsigned bal = __ballot_sync(0xffffffffu, s > 0);
        if (s > 0) {
            score[(e >> 8) * FT_PITCH + (e & 255)] = (uint8_t)s;
            seg[nC + __popc(bal & ltMask)] = (uint16_t)e;
        }
        nC += __popc(bal);
    }
    __syncthreads();

    // ---- pass 3: strict 3x3 non-max suppression inside each cell, over the corners only; survivors stay queued
    const int lastCol = maxBX - 4 - xa;                    // tile column of the level's last interior column
    int nS = 0;
    for (int i0 = 0; i0 < nC; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        int e = 0;
        if (i < nC) {
            e = seg[i];
            const int c = e & 255;
            const uint8_t* sc = score + (e >> 8) * FT_PITCH + c;
            const int s = sc[0];
            const int inCell = colIn[c];
            const bool hasL = inCell > 0, hasR = inCell < wCell - 1 && c < lastCol;
            int m = max(sc[-FT_PITCH], sc[FT_PITCH]);
            if (hasL) m = max(m, max(max(sc[-FT_PITCH - 1], sc[-1]), sc[FT_PITCH - 1]));
            if (hasR) m = max(m, max(max(sc[-FT_PITCH + 1], sc[1]), sc[FT_PITCH + 1]));
            keep = s > m;
            if (keep && s >= g.iniTh) sCellAny[colCell[c]] = 1;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) seg[nS + __popc(bal & ltMask)] = (uint16_t)e;
        nS += __popc(bal);
    }
    __syncthreads();

    // ---- pass 4: threshold fallback per cell (:809-816); mark what is emitted
    int nE = 0;
    for (int i0 = 0; i0 < nS; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false;
        int e = 0;
        if (i < nS) {
            e = seg[i];
            const int r = e >> 8, c = e & 255;
            const int T = sCellAny[colCell[c]] ? g.iniTh : g.minTh;
            keep = score[r * FT_PITCH + c] >= T;
            if (keep) atomicOr(&bitmap[r][c >> 5], 1u << (c & 31));
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) seg[nE + __popc(bal & ltMask)] = (uint16_t)e;
        nE += __popc(bal);
    }
    __syncthreads();

    // ---- pass 5: per cell, keypoints per row -> exclusive prefix over the rows (one warp per cell)
    for (int cl = warp; cl < nCellsHere; cl += FT_THREADS / 32) {
        const int cx0 = OBS_EDGE + (j0 + cl) * wCell - xa; // first interior tile column of the cell
        const int cx1 = min(cx0 + wCell, cHi);
        int carry = 0;
        for (int rb = rLo; rb < rHi; rb += 32) {
            const int r = rb + lane;
            const int n = (r < rHi && cx1 > cx0) ? row_bits(bitmap[r], cx0, cx1) : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (r < rHi) rowOfs[cl][r] = (uint16_t)(carry + incl - n);
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) countOut[cl] = carry;
    }
    __syncthreads();

    // ---- pass 6: write the keypoints in row-major order inside each cell (cv::FAST's order)
    for (int i = lane; i < nE; i += 32) {
        const int e = seg[i];
        const int r = e >> 8, c = e & 255;
        const int cl = colCell[c];
        const int cx0 = OBS_EDGE + (j0 + cl) * wCell - xa;
        const int pos = rowOfs[cl][r] + (c > cx0 ? row_bits(bitmap[r], cx0, c) : 0);
        uint32_t* slot = cand + (size_t)img * g.slotTotal + lg.slotBase + (size_t)(ci * lg.nCols + j0 + cl) * lg.cellCap;
        // coordinates relative to the 16-px border origin, :820-825
        slot[pos] = pack_key(xa + c - OBS_BORDER, Y0 + r - OBS_BORDER, score[r * FT_PITCH + c]);
    }
}

}  // namespace

size_t fast_smem_bytes(int tileRows) { return (size_t)2 * tileRows * FT_PITCH + (size_t)8 * FT_SEG * 2; }

cudaError_t fast_prepare(int tileRows) {
    return cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem_bytes(tileRows));
}

int fast_cells_per_cta_host(int wCell, int hCell) { return fast_cells_per_cta(wCell, hCell); }

cudaError_t launch_fast(const Geom& g, PyrPtrs p, const FastCta* ctaTab, uint32_t* cand, int* cellCount, int nimg, cudaStream_t st) {
    if (g.fastCtasTotal == 0) return cudaSuccess;
    dim3 grid(g.fastCtasTotal, nimg);
    k_fast_cells<<<grid, FT_THREADS, fast_smem_bytes(g.fastTileRows), st>>>(g, p, ctaTab, g.fastTileRows, cand, cellCount);
    return cudaGetLastError();
}
