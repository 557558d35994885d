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

// Host uploads arrive as one contiguous block (rows of w bytes); level 0 lives in the slab with a
// 128-byte row pitch so that every later stage can use aligned vector loads.
__global__ void __launch_bounds__(256) k_repack(const uint8_t* __restrict__ src, size_t srcImgStride, size_t srcPitch,
                                                uint8_t* __restrict__ dst, size_t dstImgStride, int dstPitch, int w, int h) {
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
    k_repack<<<grid, 256, 0, st>>>(src, srcImgStride, srcPitch, dst, dstImgStride, dstPitch, w, h);
    return cudaGetLastError();
}

cudaError_t launch_pyramid(const Geom& g, PyrPtrs p, const ResizeTap* xtab, const ResizeTap* ytab, int nimg, cudaStream_t st) {
    for (int l = 1; l < g.nlevels; l++) {
        dim3 block(32, RS_BANDS);
        dim3 grid((g.lv[l].w + 127) / 128, (g.lv[l].h + RS_ROWS * RS_BANDS - 1) / (RS_ROWS * RS_BANDS), nimg);
        k_resize<<<grid, block, 0, st>>>(g, p, xtab, ytab, l);
    }
    return cudaGetLastError();
}
