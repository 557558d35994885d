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
//   * one CTA owns a run of up to 8 cells of one cell row and stages their window with 16-byte cp.async copies, every row in flight at once;
//   * pass 1 tests 8 pixels per lane with byte-SIMD |a-b| (VABSDIFF4) against the four compass ring
//     pixels (a necessary condition for a 9-arc), leaves one bit per pixel in a bitmap, and pass 1b
//     expands the bitmap into a dense queue of pixels;
//   * pass 2 gives each queued pixel to one lane: exact score with both polarities packed as u16x2
//     (v-r+255, r-v+255) through a min/max network of VIMNMX(3).U16x2; corners are re-queued densely
//     (a third of the kernel's instructions: its loop invariants are pinned in registers, common.cuh);
//   * pass 3 suppresses non-maxima among the corners only and re-queues the survivors; passes 4-6
//     apply the per-cell threshold fallback and write the survivors in row-major order per cell,
//     positions coming from popcounts over a bitmap of the emitted pixels.
#include "kernels.h"

namespace {

constexpr int FT_PITCH = 256;          // columns of a staged row
#ifndef FT_PAD
#define FT_PAD 16
#endif
// CTAs per SM the register allocation aims at.  Alone the kernel is ~9 % faster at 5, but the step as a whole (eyes, blur and consecutive
// batches overlapping on four streams) is 2.4 % faster at 4: the other kernels find room next to it (A/B).
#ifndef FT_MINCTAS
#define FT_MINCTAS 4
#endif
constexpr int FT_SP = FT_PITCH + FT_PAD;   // shared-memory pitch of the tile and the score map: consecutive rows start 4 banks apart, so lanes
                                       // that score pixels of neighbouring rows (same columns) do not collide
constexpr int FT_ROWS = 66;            // hCell <= 60, + 6
constexpr int FT_THREADS = 256;
constexpr int FT_LIST = 8192;          // queue capacity == max interior pixels per CTA (host enforces)
constexpr int FT_MAXCELLS = 8;

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
// c is an address in the shared window (every ring pixel is one LDS.U8 [R + imm]); returns the contrast + 255.
__device__ __forceinline__ int fast_contrast_s(uint32_t c) {
    const unsigned v = lds_u8<0>(c);
    const unsigned Vc = (v + 255u) | ((255u - v) << 16);
    unsigned P[16];
#define RING(k, dx, dy) P[k] = lds_u8<(dy) * FT_SP + (dx)>(c) * 0xFFFFu + Vc
    RING(0, 0, 3); RING(1, 1, 3); RING(2, 2, 2); RING(3, 3, 1); RING(4, 3, 0); RING(5, 3, -1); RING(6, 2, -2); RING(7, 1, -3);
    RING(8, 0, -3); RING(9, -1, -3); RING(10, -2, -2); RING(11, -3, -1); RING(12, -3, 0); RING(13, -3, 1); RING(14, -2, 2); RING(15, -1, 3);
#undef RING
    unsigned m3[16];
#pragma unroll
    for (int i = 0; i < 16; i++) m3[i] = __vimin3_u16x2(P[i], P[(i + 1) & 15], P[(i + 2) & 15]);
    unsigned best = 0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const unsigned a = __vimin3_u16x2(m3[i], m3[(i + 3) & 15], m3[(i + 6) & 15]);
        const unsigned b = __vimin3_u16x2(m3[i + 1], m3[(i + 4) & 15], m3[(i + 7) & 15]);
        best = __vimax3_u16x2(best, a, b);
    }
    return (int)max(best & 0xffffu, best >> 16);           // corner contrast + 255
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

// Pretest of one word (4 pixels): 0x80 in every byte whose pixel can be a corner at threshold t
// (t < 128, K = (127 - t) * 0x01010101) and whose column is enabled in ok.  Every 9-arc holds one of
// the ring pixels {0,8} and one of {4,12}; "(|d0| | |d8|) > t" is implied by |d0| > t or |d8| > t.
__device__ __forceinline__ unsigned pretest_word(unsigned C, unsigned up, unsigned dn, unsigned l3, unsigned r3, unsigned K, unsigned ok) {
    const unsigned a = __vabsdiffu4(C, up), b = __vabsdiffu4(C, dn);
    const unsigned c = __vabsdiffu4(C, l3), d = __vabsdiffu4(C, r3);
    const unsigned y = ((((a | b) & 0x7f7f7f7fu) + K) | a | b);
    const unsigned z = ((((c | d) & 0x7f7f7f7fu) + K) | c | d);
    return y & z & ok;
}

struct FastShared {
    // the first two members are cleared together with 128-bit stores (both 2112 bytes)
    uint32_t bitmap[FT_ROWS][8];               // emitted keypoints, one bit per tile pixel
    int rowOfs[FT_MAXCELLS][FT_ROWS];          // per cell: keypoints per row, then keypoints in the rows above
    uint8_t colCell[FT_PITCH];                 // tile column -> cell of the run
    uint8_t colFlags[FT_PITCH];                // bit 0: has a left neighbour inside its cell, bit 1: a right one
    uint8_t colOK[FT_PITCH];                   // 0x80: the column takes part in the current phase
    int cellAny[FT_MAXCELLS];
    int qCount;
    int nSurv;                                 // suppressed keypoints so far; they are listed from the top of the queue downwards
    int overflow;                              // the list ran into the queue (white-noise images): emission walks the bitmap instead
    uint32_t pinTile, pinThr;                  // loop invariants of pass 2, fetched back with volatile loads (see pass 2)
};

// One detection phase at threshold t over the columns enabled in sh.colOK:
//   pass 1  byte-SIMD pretest, 32 pixels per lane (two 16-pixel chunks, 128 columns apart so that a quarter
//           warp's LDS.128 is conflict free), 4 rows per warp; flags gathered into one word per lane, expanded
//           into the CTA's queue (one shared-memory atomic per warp and 1024 pixels);
//   pass 2  exact score, one lane per queued pixel, every warp on its own slice of the queue, corners
//           compacted in place;
//   pass 3  strict 3x3 non-max suppression inside the cell over the corners; survivors set their bitmap bit.
__device__ __forceinline__ void fast_phase(const uint8_t* tile, uint8_t* score, uint16_t* Q, FastShared& sh, int t,
                                           int rHi, bool markCells) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned ltMask = (1u << lane) - 1;
    {
        const int l8 = lane & 7, sub = lane >> 3;
        const uint4 okA = *reinterpret_cast<const uint4*>(sh.colOK + 16 * l8);
        const uint4 okB = *reinterpret_cast<const uint4*>(sh.colOK + 128 + 16 * l8);
        const bool cheap = t < 128;
        const unsigned K = (unsigned)(127 - min(t, 127)) * 0x01010101u;
        for (int rb = 3 + 4 * warp; rb < rHi; rb += 4 * (FT_THREADS / 32)) {
            const int r = rb + sub;
            unsigned acc = 0;
            if (r < rHi) {
                if (cheap) {
                    const uint8_t* rowp = tile + r * FT_SP + 16 * l8;
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint8_t* cp = rowp + 128 * q;
                        const uint4 C = *reinterpret_cast<const uint4*>(cp);
                        const uint4 U = *reinterpret_cast<const uint4*>(cp + 3 * FT_SP);
                        const uint4 D = *reinterpret_cast<const uint4*>(cp - 3 * FT_SP);
                        const unsigned L = *reinterpret_cast<const unsigned*>(cp - 4);
                        const unsigned R = *reinterpret_cast<const unsigned*>(cp + 16);
                        const uint4 ok = q ? okB : okA;
                        const unsigned f0 = pretest_word(C.x, U.x, D.x, __byte_perm(L, C.x, 0x4321), __byte_perm(C.x, C.y, 0x6543), K, ok.x);
                        const unsigned f1 = pretest_word(C.y, U.y, D.y, __byte_perm(C.x, C.y, 0x4321), __byte_perm(C.y, C.z, 0x6543), K, ok.y);
                        const unsigned f2 = pretest_word(C.z, U.z, D.z, __byte_perm(C.y, C.z, 0x4321), __byte_perm(C.z, C.w, 0x6543), K, ok.z);
                        const unsigned f3 = pretest_word(C.w, U.w, D.w, __byte_perm(C.z, C.w, 0x4321), __byte_perm(C.w, R, 0x6543), K, ok.w);
                        // flag of (chunk q, word i, byte j) -> bit 8 j + 4 q + i
                        acc |= ((f0 >> 7) | (f1 >> 6) | (f2 >> 5) | (f3 >> 4)) << (4 * q);
                    }
                } else {                                   // thresholds >= 128: no cheap rejection, score everything enabled
                    acc = (okA.x >> 7) | (okA.y >> 6) | (okA.z >> 5) | (okA.w >> 4) | (okB.x >> 3) | (okB.y >> 2) | (okB.z >> 1) | okB.w;
                }
            }
            const int n = __popc(acc);
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            int base = 0;
            // (plain PTX: for an atomicAdd under a lane predicate the compiler emits its own warp aggregation, a second scan)
            if (lane == 31 && incl > 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(base) : "r"(smem_u32(&sh.qCount)), "r"(incl) : "memory");
            base = __shfl_sync(0xffffffffu, base, 31);
            uint16_t* out = Q + base + incl - n;
            const unsigned e0 = (unsigned)(r << 8) | (unsigned)(l8 << 5);
            while (acc) {
                unsigned b;
                asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(acc));             // index of the highest flag
                acc ^= 1u << b;
                *out++ = (uint16_t)(e0 + b);
            }
        }
    }
    if (tid == 0) { sh.pinTile = smem_u32(tile); sh.pinThr = 255u + (unsigned)max(t, 1); }
    __syncthreads();

    // ---- pass 2
    const int total = sh.qCount;
    if (tid == 0 && total > FT_LIST - sh.nSurv) sh.overflow = 1;          // the queue reached the keypoints listed by the previous phase
    const int per = ((total + FT_THREADS - 1) / FT_THREADS) * 32;
    const int start = warp * per;
    int end = min(start + per, total);
    unsigned ltm = ltMask;
    end = (int)pin((unsigned)end);
    ltm = pin(ltm);
    int nC = 0;
    {
        // Uniform values ptxas would recompute in every iteration (a shuffle of a uniform value is folded away): a volatile load is not.
        uint32_t tile_s, thr;         // thr: corner at t <=> contrast > t, and OpenCV's score = contrast - 1 must be > 0; contrasts are biased by 255
        tile_s = lds_volatile_u32(&sh.pinTile);
        thr = lds_volatile_u32(&sh.pinThr);
        const uint32_t Q_s = smem_u32(Q) + 2u * (unsigned)start;
        const unsigned scoreOfs = (unsigned)(score - tile);
#pragma unroll 1
        for (int i = start + lane, i0 = start; i0 < end; i += 32, i0 += 32) {
            unsigned e = (3u << 8) | (1u << 5);                           // lanes past the end score a harmless pixel (row 3, column 16) and drop it
            if (i < end) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(Q_s + 2u * (unsigned)(i - start)));
            // entry: row << 8 | l8 << 5 | j << 3 | q << 2 | i  ->  column = q << 7 | l8 << 4 | i << 2 | j
            const unsigned c = ((e & 4u) << 5) | ((e >> 1) & 0x70u) | ((e & 3u) << 2) | ((e >> 3) & 3u);
            const unsigned pos = (e & 0xff00u) | c;
            const uint32_t a = tile_s + (e >> 8) * FT_SP + c;
            const int m = fast_contrast_s(a);
            const bool ok = i < end && (unsigned)m > thr;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            __syncwarp();                                 // every lane has read its queue entry before the corners are compacted over them
            if (ok) {
                asm volatile("st.shared.u8 [%0], %1;" :: "r"(a + scoreOfs), "r"(m - 256) : "memory");   // OpenCV: score = corner contrast - 1
                asm volatile("st.shared.u16 [%0], %1;" :: "r"(Q_s + 2u * (unsigned)(nC + __popc(bal & ltm))), "r"(pos) : "memory");
            }
            nC += __popc(bal);
        }
    }
    __syncthreads();

    // ---- pass 3
    for (int i = lane; i < nC; i += 32) {
        const int pos = Q[start + i];
        const int c = pos & 255;
        const uint8_t* sc = score + (pos >> 8) * FT_SP + c;
        const int fl = sh.colFlags[c];
        int m = max(sc[-FT_SP], sc[FT_SP]);
        const int mL = max(max(sc[-FT_SP - 1], sc[-1]), sc[FT_SP - 1]);
        const int mR = max(max(sc[-FT_SP + 1], sc[1]), sc[FT_SP + 1]);
        if (fl & 1) m = max(m, mL);
        if (fl & 2) m = max(m, mR);
        if (sc[0] > m) {
            atomicOr(&sh.bitmap[pos >> 8][c >> 5], 1u << (c & 31));
            const int cl = sh.colCell[c];
            atomicAdd(&sh.rowOfs[cl][pos >> 8], 1);
            const int slot = FT_LIST - 1 - atomicAdd(&sh.nSurv, 1);
            if (slot >= total) Q[slot] = (uint16_t)pos; else sh.overflow = 1;
            if (markCells) sh.cellAny[cl] = 1;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FT_THREADS, FT_MINCTAS) k_fast_cells(const __grid_constant__ Geom g, const PyrPtrs p,
                                                           const FastCta* __restrict__ ctaTab, int tileRows,
                                                           uint32_t* __restrict__ cand, int* __restrict__ cellCount) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* tile = sm;                                   // tileRows x FT_PITCH pixels
    uint8_t* score = sm + tileRows * FT_SP;            // same geometry, 0 = not a corner
    uint16_t* Q = reinterpret_cast<uint16_t*>(sm + 2 * tileRows * FT_SP);    // pixel queue of the CTA
    __shared__ __align__(16) FastShared sh;

    pdl_entry();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    const int sw = X1 - X0, shh = Y1 - Y0;
    if (sw < 7 || shh < 7) {                              // no interior pixel: every cell of the run is empty
        if (tid < nCellsHere) countOut[tid] = 0;
        return;
    }
    const int xa = X0 & ~15;                              // tile column 0 <-> level column xa
    const int cLo = X0 + 3 - xa, cHi = X1 - 3 - xa;       // interior columns in tile coordinates
    const int rLo = 3, rHi = shh - 3;                     // interior rows

    // ---- stage the window (16-byte cp.async copies, 16 lanes per row, all rows requested before anything is waited for), clear the maps,
    // build the column tables
    {
        int pitch;
        const uint8_t* base = level_ptr(p, g, img, level, pitch);
        const int nVec = (X1 - xa + 15) >> 4;             // <= 16
        const int v = tid & 15;
        const uint8_t* src = base + (size_t)(Y0 + (tid >> 4)) * pitch + xa + 16 * v;
        const size_t srcStep = (size_t)pitch * (FT_THREADS / 16);
        const bool ld = v < nVec;
        {
            uint32_t tp = smem_u32(tile + (tid >> 4) * FT_SP + 16 * v);
            const uint8_t* sp = src;
            for (int r = tid >> 4; r < shh; r += FT_THREADS / 16, sp += srcStep, tp += (FT_THREADS / 16) * FT_SP) {
                if (ld) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(tp), "l"(sp) : "memory");
                *reinterpret_cast<uint4*>(score + r * FT_SP + 16 * v) = make_uint4(0, 0, 0, 0);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        static_assert((sizeof(sh.bitmap) + sizeof(sh.rowOfs)) % 16 == 0 && (sizeof(sh.bitmap) + sizeof(sh.rowOfs)) / 16 <= 2 * FT_THREADS, "clear");
        for (int i = tid; i < (int)((sizeof(sh.bitmap) + sizeof(sh.rowOfs)) / 16); i += FT_THREADS)
            reinterpret_cast<uint4*>(&sh.bitmap[0][0])[i] = make_uint4(0, 0, 0, 0);
        const int lastCol = maxBX - 4 - xa;               // tile column of the level's last interior column
        int inCell = max(tid - (X0 + 3 - xa), 0);         // column relative to the first interior column of the run
        int cc = 0;                                       // cell of the run (< 8): three compare-subtract steps
        if (inCell >= 4 * wCell) { inCell -= 4 * wCell; cc += 4; }
        if (inCell >= 2 * wCell) { inCell -= 2 * wCell; cc += 2; }
        if (inCell >= wCell) { inCell -= wCell; cc += 1; }
        sh.colCell[tid] = (uint8_t)cc;
        sh.colFlags[tid] = (uint8_t)((inCell > 0 ? 1 : 0) | ((inCell < wCell - 1 && tid < lastCol) ? 2 : 0));
        sh.colOK[tid] = (tid >= cLo && tid < cHi) ? 0x80 : 0;
        if (tid < FT_MAXCELLS) sh.cellAny[tid] = 0;
        if (tid == 0) { sh.qCount = 0; sh.nSurv = 0; sh.overflow = 0; }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // ---- phase A: cv::FAST(cell, iniThFAST) for every cell of the run (:809).  A corner at iniTh only
    // competes with neighbours that are corners at iniTh as well, so nothing below iniTh is scored here.
    fast_phase(tile, score, Q, sh, g.iniTh, rHi, true);

    // ---- phase B: cells whose suppressed result is empty are redone at minThFAST (:811-816)
    if (g.minTh < g.iniTh) {
        if (__syncthreads_or(tid < nCellsHere && sh.cellAny[tid] == 0)) {
            sh.colOK[tid] = (tid >= cLo && tid < cHi && sh.cellAny[sh.colCell[tid]] == 0) ? 0x80 : 0;
            if (tid == 0) sh.qCount = 0;
            __syncthreads();
            fast_phase(tile, score, Q, sh, g.minTh, rHi, false);
        }
    }

    // ---- per cell, keypoints per row -> exclusive prefix over the rows (one warp per cell)
    for (int cl = warp; cl < nCellsHere; cl += FT_THREADS / 32) {
        int carry = 0;
        for (int rb = rLo; rb < rHi; rb += 32) {
            const int r = rb + lane;
            const int n = r < rHi ? sh.rowOfs[cl][r] : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (r < rHi) sh.rowOfs[cl][r] = carry + incl - n;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) countOut[cl] = carry;
    }
    __syncthreads();

    // ---- write the keypoints in row-major order inside each cell (cv::FAST's order): position = keypoints of the cell in the rows
    // above + keypoints of the cell to the left in the same row
    uint32_t* slots = cand + (size_t)img * g.slotTotal + lg.slotBase + (size_t)(ci * lg.nCols + j0) * lg.cellCap;
    auto emit = [&](int r, int c) {
        const int cl = sh.colCell[c];
        const int cx0 = OBS_EDGE + (j0 + cl) * wCell - xa;
        const int pos = sh.rowOfs[cl][r] + (c > cx0 ? row_bits(sh.bitmap[r], cx0, c) : 0);
        // coordinates relative to the 16-px border origin, :820-825
        slots[cl * lg.cellCap + pos] = pack_key(xa + c - OBS_BORDER, Y0 + r - OBS_BORDER, score[r * FT_SP + c]);
    };
    if (!sh.overflow) {
        const int K = sh.nSurv;
        for (int i = tid; i < K; i += FT_THREADS) {
            const int e = Q[FT_LIST - 1 - i];
            emit(e >> 8, e & 255);
        }
    } else {
        for (int i = tid; i < shh * 8; i += FT_THREADS) {
            const int r = i >> 3;
            uint32_t word = sh.bitmap[r][i & 7];
            while (word) {
                const int c = ((i & 7) << 5) + __ffs(word) - 1;
                word &= word - 1;
                emit(r, c);
            }
        }
    }
}

}  // namespace

size_t fast_smem_bytes(int tileRows) { return (size_t)2 * tileRows * FT_SP + (size_t)FT_LIST * 2; }

cudaError_t fast_prepare(int tileRows) {
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_fast_cells);
    if (e != cudaSuccess) return e;
    return (int)fast_smem_bytes(tileRows) <= obsdetail::max_dynamic_smem((const void*)k_fast_cells) ? cudaSuccess : cudaErrorInvalidValue;
}

int fast_cells_per_cta_host(int wCell, int hCell) { return fast_cells_per_cta(wCell, hCell); }

cudaError_t launch_fast(const Geom& g, PyrPtrs p, const FastCta* ctaTab, uint32_t* cand, int* cellCount, int nimg, cudaStream_t st) {
    if (g.fastCtasTotal == 0) return cudaSuccess;
    dim3 grid(g.fastCtasTotal, nimg);
    return launch_k(pdl_enabled(), k_fast_cells, grid, dim3(FT_THREADS), fast_smem_bytes(g.fastTileRows), st, g, p, ctaTab, g.fastTileRows, cand, cellCount);
}
