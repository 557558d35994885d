// K1 -- image pyramid.  Replaces ORBextractor::ComputePyramid (src/ORBextractor.cc:1107-1132):
// level l = cv::resize(level l-1, INTER_LINEAR) on 8-bit pixels.  OpenCV's 8U bilinear path is
// fixed point: 11-bit weights per axis, horizontal sums kept at full precision, then
// out = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2.  The weights come from
// host-built tap tables so that every output byte equals the reference's.  The 19-px reflect
// border the reference adds around each level is never read on this path and is not materialised.
//
// A thread owns four consecutive output columns and walks down the output rows: its x taps, byte
// selectors and weights live in registers; per source row it loads the (at most) 12 bytes under
// its taps as three aligned words, forms each horizontal sum with one PRMT + one IDP.2A
// (a0*S[x] + a1*S[x+1]) and reuses the lower source row of one output row as the upper row of the
// next when the vertical taps allow; the vertical blend is two IMAD.HI per pixel.
#include "kernels.h"
#include <algorithm>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

constexpr int RS_ROWS = 8;           // output rows per thread
constexpr int RS_BANDS = 4;          // row bands per CTA (CTA tile: 128 x 32 output pixels)

struct XTaps {
    unsigned w[4];       // a0 | a1 << 16
    unsigned sel[4];     // PRMT selector picking (S[sx], S[sx+1]) out of the 8-byte shifted window
    int word0;           // first source word of the window
    int shift;           // bit shift aligning the window to the first tap
    bool fast;           // all four taps fit the 8-byte window
    int ofs[4];
};

struct RowWords { unsigned w0, w1, w2; };

__device__ __forceinline__ RowWords load_row(const uint8_t* __restrict__ row, int lastWord, const XTaps& t) {
    RowWords r;
    if (t.fast) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(row);
        r.w0 = __ldg(q + t.word0); r.w1 = __ldg(q + min(t.word0 + 1, lastWord)); r.w2 = __ldg(q + min(t.word0 + 2, lastWord));
    } else {
        r.w0 = r.w1 = r.w2 = 0;
    }
    return r;
}

__device__ __forceinline__ void hrow(const RowWords& rw, const uint8_t* __restrict__ row, int srcW, const XTaps& t, unsigned hs[4]) {
    if (t.fast) {
        const unsigned U0 = __funnelshift_r(rw.w0, rw.w1, t.shift), U1 = __funnelshift_r(rw.w1, rw.w2, t.shift);
#pragma unroll
        for (int i = 0; i < 4; i++) hs[i] = __dp2a_lo(t.w[i], __byte_perm(U0, U1, t.sel[i]), 0u) >> 4;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int sx = t.ofs[i];
            const unsigned s0 = row[sx], s1 = row[min(sx + 1, srcW - 1)];
            hs[i] = (s0 * (t.w[i] & 0xffffu) + s1 * (t.w[i] >> 16)) >> 4;
        }
    }
}

__global__ void __launch_bounds__(32 * RS_BANDS) k_resize(const __grid_constant__ Geom g, const PyrPtrs p,
                                                          const ResizeTap* __restrict__ xtab,
                                                          const ResizeTap* __restrict__ ytab, int level) {
    pdl_entry();
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int img = blockIdx.z;
    const int dx0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int dy0 = (blockIdx.y * RS_BANDS + threadIdx.y) * RS_ROWS;
    if (dx0 >= D.w || dy0 >= D.h) return;
    int spitch, dpitch;
    const uint8_t* src = level_ptr(p, g, img, level - 1, spitch);
    uint8_t* dst = const_cast<uint8_t*>(level_ptr(p, g, img, level, dpitch)) + dx0;
    const int lastWord = (spitch >> 2) - 1;

    XTaps t;
    {
        int ofs[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const ResizeTap tx = xtab[D.xtab + min(dx0 + i, D.w - 1)];
            ofs[i] = tx.ofs;
            t.ofs[i] = tx.ofs;
            t.w[i] = (unsigned)(unsigned short)tx.a0 | ((unsigned)(unsigned short)tx.a1 << 16);
        }
        t.word0 = ofs[0] >> 2;
        t.shift = 8 * (ofs[0] & 3);
        t.fast = true;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int o = ofs[i] - ofs[0];
            if (o < 0 || o > 6) t.fast = false;
            t.sel[i] = (unsigned)(o & 7) | ((unsigned)((o + 1) & 7) << 4);
        }
    }

    // Walk the source rows the band needs, two rows of loads ahead of the arithmetic; an output row is
    // produced when its lower source row arrives (the vertical taps advance by >= 1 source row per output row).
    const int nOut = min(RS_ROWS, D.h - dy0);
    const ResizeTap* yt = ytab + D.ytab + dy0;
    const int sA = min(max(yt[0].ofs, 0), S.h - 1);
    const int sB = min(max(yt[nOut - 1].ofs + 1, 0), S.h - 1);
    RowWords wa = load_row(src + (size_t)sA * spitch, lastWord, t);
    RowWords wb = load_row(src + (size_t)min(sA + 1, sB) * spitch, lastWord, t);
    unsigned hprev[4] = {0, 0, 0, 0}, hcur[4];
    int r = 0;
    ResizeTap ty = yt[0];
#pragma unroll 1
    for (int s = sA; s <= sB; s++) {
        const RowWords wc = load_row(src + (size_t)min(s + 2, sB) * spitch, lastWord, t);
        hrow(wa, src + (size_t)s * spitch, S.w, t, hcur);
        while (r < nOut) {
            const int y0 = min(max(ty.ofs, 0), S.h - 1);
            const int y1 = min(max(ty.ofs + 1, 0), S.h - 1);
            if (y1 != s) break;
            const unsigned b0 = (unsigned)ty.a0 << 16, b1 = (unsigned)ty.a1 << 16;      // (b * h) >> 16 == umulhi(b << 16, h)
            unsigned o = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned h0 = (y0 == s) ? hcur[i] : hprev[i];
                unsigned v = (__umulhi(b0, h0) + __umulhi(b1, hcur[i]) + 2u) >> 2;
                v = min(v, 255u);
                o |= v << (8 * i);
            }
            *reinterpret_cast<uint32_t*>(dst + (size_t)(dy0 + r) * dpitch) = o;      // rows are padded to the pitch
            r++;
            if (r < nOut) ty = yt[r];
        }
#pragma unroll
        for (int i = 0; i < 4; i++) hprev[i] = hcur[i];
        wa = wb; wb = wc;
    }
}

// Tiled variant for the usual scale factors (every group of four output columns reads within 7 source bytes,
// LevelGeom::rsPitch > 0).  A CTA stages the source footprint of a 128 x 64 output tile in shared memory with
// 128-bit loads; a thread owns four output columns of a 16-row band.  No clamping is left in the loop: rows
// are rebased in a shared copy of the y taps, columns by the staged pitch.
// (b * h) >> 16 == umulhi(b << 16, h); the result cannot exceed 255 because a pair of weights sums to at most 2049.
#ifndef RS_OPT_STAGE
#define RS_OPT_STAGE 1
#endif
constexpr int RT_ROWS = 16;          // output rows per thread
constexpr int RT_BANDS = RESIZE_TILE_H / RT_ROWS;

struct __align__(16) YTap { int off0, off1; unsigned b0, b1; };   // byte offsets of the two source rows inside the staged tile, weights << 16

__device__ __forceinline__ void hrow_tile(const uint8_t* row, int shift, const unsigned (&w)[4], const unsigned (&sel)[4], unsigned (&hs)[4]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(row);
    const unsigned w0 = q[0], w1 = q[1], w2 = q[2];
    const unsigned U0 = __funnelshift_r(w0, w1, shift), U1 = __funnelshift_r(w1, w2, shift);
#pragma unroll
    for (int i = 0; i < 4; i++) hs[i] = __dp2a_lo(w[i], __byte_perm(U0, U1, sel[i]), 0u) >> 4;
}

// One 128 x 64 output tile of level `level` of image `img`.  CHAIN: the source level may have been written by another CTA of
// the same cluster moments ago (k_resize_chain), so it is read through L2 (ld.global.cg) instead of the non-coherent path, and the
// function ends with a barrier because the caller reuses the shared buffers for its next tile.
// source index of destination index d (OpenCV resize INTER_LINEAR: floor((d + 0.5) * scale - 0.5) on the float value, x clamped)
__device__ __forceinline__ int tap_ofs(int d, double sc, bool isX, int ssize) {
    const float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, sc), 0.5);
    int s = __float2int_rd(f);
    if (isX) { if (s < 0) s = 0; if (s >= ssize - 1) s = ssize - 1; }
    return s;
}

template <bool CHAIN>
__device__ __forceinline__ void resize_tile_body(const Geom& g, const PyrPtrs& p, const ResizeTap* __restrict__ xtab,
                                                 const ResizeTap* __restrict__ ytab, int level, int img, int tileX, int tileY,
                                                 uint8_t* tile, YTap* sY) {
    const LevelGeom& D = g.lv[level];
    const LevelGeom& S = g.lv[level - 1];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int tx0 = tileX * RESIZE_TILE_W, ty0 = tileY * RESIZE_TILE_H;
    const int tw = min(RESIZE_TILE_W, D.w - tx0), th = min(RESIZE_TILE_H, D.h - ty0);
    const ResizeTap* xt = xtab + D.xtab;
    const ResizeTap* yt = ytab + D.ytab;
    int spitch, dpitch;
    const uint8_t* src = level_ptr(p, g, img, level - 1, spitch);
    uint8_t* dst = const_cast<uint8_t*>(level_ptr(p, g, img, level, dpitch));
    // source footprint of the tile from the same arithmetic the host built the tap tables with (api.cu resize_taps) -- no table
    // load in front of the copies
    const int xs0 = tap_ofs(tx0, D.rsScaleX, true, S.w) & ~15, xs1 = tap_ofs(tx0 + tw - 1, D.rsScaleX, true, S.w) + 1;
    const int ys0 = min(max(tap_ofs(ty0, D.rsScaleY, false, S.h), 0), S.h - 1);
    const int ys1 = min(max(tap_ofs(ty0 + th - 1, D.rsScaleY, false, S.h) + 1, 0), S.h - 1);
    const int pitch = D.rsPitch;
    auto ld = [](const uint4* q) { return CHAIN ? __ldcg(q) : __ldg(q); };
    {
        // rows of the source are padded to a multiple of 16 bytes (slab pitch 128; level 0: 16-byte aligned stride)
        const int nVec = min((xs1 - xs0) / 16 + 1, (spitch - xs0) >> 4), nRows = ys1 - ys0 + 1;
        constexpr int RSTEP = 32 * RT_BANDS / 16;
#if RS_OPT_STAGE
        // every row of the thread's column of vectors is requested before anything is waited for (cp.async through L2, which is
        // also what the chained levels need): one DRAM / L2 round trip per tile instead of one per four sweeps
        for (int v = tid & 15; v < nVec; v += 16) {
            const uint8_t* gp = src + (size_t)(ys0 + (tid >> 4)) * spitch + xs0 + 16 * v;
            uint32_t tp = smem_u32(tile + (tid >> 4) * pitch + 16 * v);
            for (int r = tid >> 4; r < nRows; r += RSTEP, gp += (size_t)spitch * RSTEP, tp += RSTEP * pitch)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(tp), "l"(gp) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        for (int v = tid & 15; v < nVec; v += 16) {
            const uint4* gp = reinterpret_cast<const uint4*>(src + (size_t)(ys0 + (tid >> 4)) * spitch + xs0) + v;
            uint8_t* tp = tile + (tid >> 4) * pitch + 16 * v;
            const size_t gstep = (size_t)spitch * RSTEP / 16;
            int r = tid >> 4;
            for (; r + 3 * RSTEP < nRows; r += 4 * RSTEP, gp += 4 * gstep, tp += 4 * RSTEP * pitch) {
                const uint4 a = ld(gp), b = ld(gp + gstep), c = ld(gp + 2 * gstep), d = ld(gp + 3 * gstep);
                *reinterpret_cast<uint4*>(tp) = a;
                *reinterpret_cast<uint4*>(tp + RSTEP * pitch) = b;
                *reinterpret_cast<uint4*>(tp + 2 * RSTEP * pitch) = c;
                *reinterpret_cast<uint4*>(tp + 3 * RSTEP * pitch) = d;
            }
            for (; r < nRows; r += RSTEP, gp += gstep, tp += RSTEP * pitch) *reinterpret_cast<uint4*>(tp) = ld(gp);
        }
#endif
        if (tid < th) {
            const ResizeTap t = yt[ty0 + tid];
            YTap y;
            y.off0 = (min(max(t.ofs, 0), S.h - 1) - ys0) * pitch;
            y.off1 = (min(max(t.ofs + 1, 0), S.h - 1) - ys0) * pitch;
            y.b0 = (unsigned)t.a0 << 16; y.b1 = (unsigned)t.a1 << 16;
            sY[tid] = y;
        }
    }
    // the thread's four column taps are fetched while the tile is in flight
    const int dx0 = tx0 + 4 * threadIdx.x, r0 = threadIdx.y * RT_ROWS;
    const bool active = dx0 < D.w && r0 < th;
    unsigned w[4] = {0, 0, 0, 0}, sel[4] = {0, 0, 0, 0};
    int ofs0 = 0;
    if (active) {
        const ResizeTap t0 = xt[dx0];
        ofs0 = t0.ofs;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const ResizeTap tx = i ? xt[min(dx0 + i, D.w - 1)] : t0;
            const int o = tx.ofs - ofs0;                    // 0..6 (host-checked)
            w[i] = (unsigned)(unsigned short)tx.a0 | ((unsigned)(unsigned short)tx.a1 << 16);
            sel[i] = (unsigned)o | ((unsigned)(o + 1) << 4);
        }
    }
#if RS_OPT_STAGE
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
    __syncthreads();
    if (active) {
        const int nOut = min(RT_ROWS, th - r0);
        const int rel = ofs0 - xs0;
        const uint8_t* col = tile + (rel & ~3);
        const int shift = 8 * (rel & 3);
        uint8_t* out = dst + (size_t)(ty0 + r0) * dpitch + dx0;
        // Every output row filters its two source rows afresh: at scale 1.2 that is 2 row filters per output row
        // instead of 1.2, but the loop carries no state, no branches and no register shuffling.
#pragma unroll 4
        for (int r = 0; r < nOut; r++, out += dpitch) {
            const YTap y = sY[r0 + r];
            unsigned hA[4], hB[4];
            hrow_tile(col + y.off0, shift, w, sel, hA);
            hrow_tile(col + y.off1, shift, w, sel, hB);
            unsigned o = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) o += ((__umulhi(y.b0, hA[i]) + __umulhi(y.b1, hB[i]) + 2u) >> 2) << (8 * i);
            *reinterpret_cast<uint32_t*>(out) = o;          // rows are padded to the pitch
        }
    }
    if (CHAIN) __syncthreads();
}

__global__ void __launch_bounds__(32 * RT_BANDS) k_resize_tile(const __grid_constant__ Geom g, const PyrPtrs p,
                                                               const ResizeTap* __restrict__ xtab,
                                                               const ResizeTap* __restrict__ ytab, int level) {
    extern __shared__ __align__(16) uint8_t tile[];
    __shared__ YTap sY[RESIZE_TILE_H];
    pdl_entry();
    resize_tile_body<false>(g, p, xtab, ytab, level, blockIdx.z, blockIdx.x, blockIdx.y, tile, sY);
}

// The small levels of the pyramid in one launch: a thread-block cluster of RC_CTAS CTAs owns one image and walks the levels
// firstLevel .. n-1; the tiles of a level are dealt round-robin to the CTAs of the cluster, and a cluster barrier (release /
// acquire at cluster scope) separates the levels.  Replaces 4 dependent launches of a few dozen CTAs each per image batch.
constexpr int RC_CTAS = 8;

__global__ void __launch_bounds__(32 * RT_BANDS) k_resize_chain(const __grid_constant__ Geom g, const PyrPtrs p,
                                                                const ResizeTap* __restrict__ xtab,
                                                                const ResizeTap* __restrict__ ytab, int firstLevel) {
    extern __shared__ __align__(16) uint8_t tile[];
    __shared__ YTap sY[RESIZE_TILE_H];
    pdl_entry();
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), img = blockIdx.y;
    for (int level = firstLevel; level < g.nlevels; level++) {
        const LevelGeom& D = g.lv[level];
        const int tilesX = (D.w + RESIZE_TILE_W - 1) / RESIZE_TILE_W, tilesY = (D.h + RESIZE_TILE_H - 1) / RESIZE_TILE_H;
        for (int t = rank; t < tilesX * tilesY; t += RC_CTAS)
            resize_tile_body<true>(g, p, xtab, ytab, level, img, t % tilesX, t / tilesX, tile, sY);
        if (level + 1 < g.nlevels) cluster.sync();
    }
}

// Host uploads arrive as one contiguous block (rows of w bytes); level 0 lives in the slab with a
// 128-byte row pitch so that every later stage can use aligned vector loads.
__global__ void __launch_bounds__(256) k_repack(const uint8_t* __restrict__ src, size_t srcImgStride, size_t srcPitch,
                                                uint8_t* __restrict__ dst, size_t dstImgStride, int dstPitch, int w, int h) {
    pdl_entry();
    const int img = blockIdx.z;
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (y >= h) return;
    const uint8_t* s = src + (size_t)img * srcImgStride + (size_t)y * srcPitch;
    uint8_t* d = dst + (size_t)img * dstImgStride + (size_t)y * dstPitch;
    const int lane = threadIdx.x & 31;
    // destination words are aligned; source bytes are fetched individually only at unaligned rows
    const int nWords = (w + 3) >> 2;
    const uintptr_t mis = reinterpret_cast<uintptr_t>(s) & 3;
    for (int i = blockIdx.x * 32 + lane; i < nWords; i += gridDim.x * 32) {
        uint32_t v;
        if (mis == 0 && 4 * i + 3 < w) {
            v = __ldg(reinterpret_cast<const uint32_t*>(s) + i);
        } else if (4 * i + 3 < w) {
            const uint32_t* a = reinterpret_cast<const uint32_t*>(s - mis) + i;      // aligned pair straddling the word
            v = __funnelshift_r(__ldg(a), __ldg(a + 1), 8 * (int)mis);
        } else {
            v = 0;
            for (int b = 0; b < 4; b++) if (4 * i + b < w) v |= (uint32_t)s[4 * i + b] << (8 * b);
        }
        reinterpret_cast<uint32_t*>(d)[i] = v;
    }
}

}  // namespace

cudaError_t launch_repack(const uint8_t* src, size_t srcImgStride, size_t srcPitch, uint8_t* dst, size_t dstImgStride,
                          int dstPitch, int w, int h, int nimg, cudaStream_t st) {
    dim3 grid((w + 4 * 32 * 4 - 1) / (4 * 32 * 4), (h + 7) / 8, nimg);
    return launch_k(pdl_enabled(), k_repack, grid, dim3(256), 0, st, src, srcImgStride, srcPitch, dst, dstImgStride, dstPitch, w, h);
}

// First level of the cluster-chained tail: the smallest l >= 1 from which every level uses the tiled kernel and has at most
// 2 * RC_CTAS tiles, provided at least two levels remain; 0 = no chain.
static int chain_first_level(const Geom& g) {
    int first = 0;
    for (int l = g.nlevels - 1; l >= 1; l--) {
        const int tiles = ((g.lv[l].w + RESIZE_TILE_W - 1) / RESIZE_TILE_W) * ((g.lv[l].h + RESIZE_TILE_H - 1) / RESIZE_TILE_H);
        if (g.lv[l].rsPitch <= 0 || tiles > 2 * RC_CTAS) break;
        first = l;
    }
    return (first > 0 && g.nlevels - first >= 2) ? first : 0;
}

cudaError_t pyramid_prepare(const Geom& g) {
    size_t need = 0;
    for (int l = 1; l < g.nlevels; l++) need = std::max(need, (size_t)g.lv[l].rsPitch * g.lv[l].rsRows);
    if (need <= 48 * 1024) return cudaSuccess;
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_resize_tile);
    if (e != cudaSuccess) return e;
    return OBS_ALLOW_MAX_SMEM(k_resize_chain);
}

cudaError_t launch_pyramid(const Geom& g, PyrPtrs p, const ResizeTap* xtab, const ResizeTap* ytab, int nimg, cudaStream_t st) {
    const int chain = chain_first_level(g);
    for (int l = 1; l < g.nlevels; l++) {
        if (chain && l == chain) {
            size_t smem = 0;
            for (int k = chain; k < g.nlevels; k++) smem = std::max(smem, (size_t)g.lv[k].rsPitch * g.lv[k].rsRows);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(RC_CTAS, nimg); cfg.blockDim = dim3(32, RT_BANDS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = RC_CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_resize_chain, g, p, xtab, ytab, chain);
            if (e != cudaSuccess) return e;
            break;
        }
        if (g.lv[l].rsPitch > 0) {
            dim3 block(32, RT_BANDS);
            dim3 grid((g.lv[l].w + RESIZE_TILE_W - 1) / RESIZE_TILE_W, (g.lv[l].h + RESIZE_TILE_H - 1) / RESIZE_TILE_H, nimg);
            cudaError_t e = launch_k(pdl_enabled(), k_resize_tile, grid, block, (size_t)g.lv[l].rsPitch * g.lv[l].rsRows, st, g, p, xtab, ytab, l);
            if (e != cudaSuccess) return e;
            continue;
        }
        dim3 block(32, RS_BANDS);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + RS_ROWS * RS_BANDS - 1) / (RS_ROWS * RS_BANDS), nimg);
        cudaError_t e = launch_k(pdl_enabled(), k_resize, grid, block, 0, st, g, p, xtab, ytab, l);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}
