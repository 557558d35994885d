// K8 -- stereo matching.  Replaces Frame::ComputeStereoMatches (src/Frame.cc:706-880).
//
//   k_stereo_rows  : one CTA per frame builds the reference's row table (:716-733) as a CSR in HBM (entries carry x, octave and index):
//     right keypoint i is listed under every row of its band [floor(y-r), ceil(y+r)], r = 2*scale[octave].
//   k_stereo_match : one warp per left keypoint scans the candidates of row int(vL), keeps those whose
//     octave is within +-1 and whose x lies in [uL-maxD, uL-minD] (:773-778), and takes the smallest
//     256-bit Hamming distance below TH_HIGH, the lowest index winning ties like the reference's
//     in-order scan with a strict < (:781-787; key = dist<<16 | index, warp min).  If that
//     distance is < 75 the warp refines it with the 11x11 centre-subtracted SAD slid over +-5 px on
//     the keypoint's pyramid level (:802-832), fits the parabola (:838-845) and converts to
//     disparity / depth (:848-862).
//   k_stereo_filter: one CTA per frame: median of the accepted SADs by a two-level 256-bin histogram selection and
//     the 1.5*1.4*median outlier cut (:866-879).
// All float expressions are evaluated with individually rounded binary32 operations.
#include "kernels.h"

namespace {

#ifndef ST_NW
#define ST_NW 4
#endif
constexpr int ST_WARPS = ST_NW;
constexpr int TH_HIGH = 100, TH_LOW = 50;

__device__ __forceinline__ int hamming256(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

// Row table of the right keypoints (:716-733) as a CSR: rowStart[H+1], rowIdx[...].  One CTA per frame.
// The order inside a row does not matter here: ties are broken by the keypoint index in the match kernel.
constexpr int SR_THREADS = 1024;       // one CTA per frame on the critical path of a live frame: as many threads as a CTA takes
__global__ void __launch_bounds__(SR_THREADS) k_stereo_rows(const __grid_constant__ StereoArgs A) {
    extern __shared__ int sRow[];                 // H + 1 counters, then cursors
    __shared__ int sWarp[SR_THREADS / 32 + 1];
    pdl_entry();
    const Geom& g = A.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, img = blockIdx.x;
    const int H = g.lv[0].h;
    const uint8_t* recR = A.recR + (size_t)img * A.recordBytes;
    const int nR = min(*reinterpret_cast<const int*>(recR), g.kpCap);
    const float* kpR = reinterpret_cast<const float*>(recR + OBS_HDR_INTS * 4);
    int* rowStart = A.rowStart + (size_t)img * (g.h + 1);
    uint2* rowIdx = A.rowIdx + (size_t)img * A.rowIdxCap;

    for (int i = tid; i <= H; i += SR_THREADS) sRow[i] = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += SR_THREADS) {
        const float y = kpR[i * 7 + 1];
        const int oct = reinterpret_cast<const int*>(kpR)[i * 7 + 5];
        const float r = __fmul_rn(2.0f, g.lv[oct].scale);
        const int maxr = min((int)ceilf(__fadd_rn(y, r)), H - 1);
        const int minr = max((int)floorf(__fsub_rn(y, r)), 0);
        for (int yi = minr; yi <= maxr; yi++) atomicAdd(&sRow[yi], 1);
    }
    __syncthreads();
    // exclusive scan of the H row counts (chunks of SR_THREADS)
    int carry = 0;
    for (int base = 0; base < H; base += SR_THREADS) {
        const int i = base + tid;
        const int v = i < H ? sRow[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = sWarp[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            sWarp[lane] = wi - w;
            if (lane == 31) sWarp[32] = wi;
        }
        __syncthreads();
        if (i < H) {
            const int ex = carry + sWarp[warp] + incl - v;
            sRow[i] = ex;                           // becomes the fill cursor
            rowStart[i] = ex;
        }
        carry += sWarp[32];
        __syncthreads();
    }
    if (tid == 0) rowStart[H] = carry;
    __syncthreads();
    for (int i = tid; i < nR; i += SR_THREADS) {
        const float y = kpR[i * 7 + 1];
        const int oct = reinterpret_cast<const int*>(kpR)[i * 7 + 5];
        const float r = __fmul_rn(2.0f, g.lv[oct].scale);
        const int maxr = min((int)ceilf(__fadd_rn(y, r)), H - 1);
        const int minr = max((int)floorf(__fsub_rn(y, r)), 0);
        // the entry carries what the match kernel filters on (x, octave), so that a candidate costs one dependent load, not two
        const uint2 ent = make_uint2(__float_as_uint(kpR[i * 7]), ((unsigned)oct << 16) | (unsigned)i);
        for (int yi = minr; yi <= maxr; yi++) {
            const int pos = atomicAdd(&sRow[yi], 1);
            if (pos < A.rowIdxCap) rowIdx[pos] = ent;
        }
    }
}

#ifndef ST_MINCTAS
#define ST_MINCTAS 16       // 32 registers (a few spilled words): the kernel waits on dependent gathers, so resident warps are what counts
#endif
__global__ void __launch_bounds__(ST_WARPS * 32, ST_MINCTAS) k_stereo_match(const __grid_constant__ StereoArgs A) {
    pdl_entry();
    const Geom& g = A.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.y;
    const int iL = blockIdx.x * ST_WARPS + warp;
    if (iL >= g.kpCap) return;
    const uint8_t* recL = A.recL + (size_t)img * A.recordBytes;
    const uint8_t* recR = A.recR + (size_t)img * A.recordBytes;
    const int nL = *reinterpret_cast<const int*>(recL);
    const float* kpL = reinterpret_cast<const float*>(recL + OBS_HDR_INTS * 4);
    const uint8_t* descL = recL + OBS_HDR_INTS * 4 + (size_t)g.kpCap * 28;
    const uint8_t* descR = recR + OBS_HDR_INTS * 4 + (size_t)g.kpCap * 28;
    const int* rowStart = A.rowStart + (size_t)img * (g.h + 1);
    const uint2* rowIdx = A.rowIdx + (size_t)img * A.rowIdxCap;

    float outU = -1.0f, outD = -1.0f;
    int outSad = -1;
    if (iL < nL) {
        const float uL = kpL[iL * 7], vL = kpL[iL * 7 + 1];
        const int levelL = reinterpret_cast<const int*>(kpL)[iL * 7 + 5];
        const int row = (int)vL;                                  // vRowIndices[vL], :758
        const float minU = __fsub_rn(uL, A.maxD), maxU = __fsub_rn(uL, A.minD);
        uint32_t best = ((uint32_t)TH_HIGH << 16);                // dist << 16 | iR ; only dist < TH_HIGH can win
        float bestX = 0.f;
        if (!(maxU < 0) && row >= 0 && row < g.lv[0].h) {
            uint32_t dl[8];
            {
                const uint4 d0 = *reinterpret_cast<const uint4*>(descL + (size_t)iL * 32);
                const uint4 d1 = *reinterpret_cast<const uint4*>(descL + (size_t)iL * 32 + 16);
                dl[0] = d0.x; dl[1] = d0.y; dl[2] = d0.z; dl[3] = d0.w; dl[4] = d1.x; dl[5] = d1.y; dl[6] = d1.z; dl[7] = d1.w;
            }
            const int c0 = rowStart[row], c1 = min(rowStart[row + 1], A.rowIdxCap);
            for (int ci = c0 + lane; ci < c1; ci += 32) {
                const uint2 ent = __ldg(rowIdx + ci);
                const float xR = __uint_as_float(ent.x);
                const int octR = (int)(ent.y >> 16), iR = (int)(ent.y & 0xffffu);
                if (octR < levelL - 1 || octR > levelL + 1) continue;
                if (!(xR >= minU && xR <= maxU)) continue;
                const uint4 b0 = *reinterpret_cast<const uint4*>(descR + (size_t)iR * 32);
                const uint4 b1 = *reinterpret_cast<const uint4*>(descR + (size_t)iR * 32 + 16);
                const uint32_t key = ((uint32_t)hamming256(dl, b0, b1) << 16) | (uint32_t)iR;
                if (key < best) { best = key; bestX = xR; }
            }
        }
        const uint32_t mine = best;
        best = __reduce_min_sync(0xffffffffu, best);
        const int bestDist = (int)(best >> 16);
        if (bestDist < (TH_HIGH + TH_LOW) / 2) {
            // ---- sub-pixel refinement by correlation on level `levelL` (:794-862)
            // (keys are unique -- they hold the keypoint index -- so exactly one lane owns the minimum and its x)
            const float uR0 = __shfl_sync(0xffffffffu, bestX, __ffs(__ballot_sync(0xffffffffu, mine == best)) - 1);
            const float sf = g.lv[levelL].invScale;
            const float scaleduL = roundf(__fmul_rn(uL, sf));
            const float scaledvL = roundf(__fmul_rn(vL, sf));
            const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
            const int w = 5, Ls = 5;
            const LevelGeom& lg = g.lv[levelL];
            const float iniu = scaleduR0 + Ls - w, endu = scaleduR0 + Ls + w + 1;
            const int cxL = (int)scaleduL, cy = (int)scaledvL, cxR = (int)scaleduR0;
            // the reference would throw inside cv::Mat::rowRange/colRange for windows that leave the level
            const bool inside = cy - w >= 0 && cy + w < lg.h && cxL - w >= 0 && cxL + w < lg.w &&
                                cxR - Ls - w >= 0 && cxR + Ls + w < lg.w;
            if (!(iniu < 0 || endu >= (float)lg.w) && inside) {
                int pl, pr;
                const uint8_t* imL = level_ptr(A.left, g, img, levelL, pl);
                const uint8_t* imR = level_ptr(A.right, g, img, levelL, pr);
                const uint8_t* cL = imL + (size_t)cy * pl + cxL;
                const uint8_t* cR = imR + (size_t)cy * pr + cxR;
                const int centreL = cL[0];
                // Lane's (up to 4) patch pixels e = lane + 32 t; for each it needs the 11 right-image bytes under the 11 shifts,
                // fetched as four aligned words and shifted into place (three words A0..A2 = shifts -5..-2, -1..2, 3..5).
                int lv[4];
                unsigned A0[4], A1[4], A2[4];
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const int e = lane + 32 * t;
                    const int py = e / 11, px = e - py * 11;
                    const bool on = e < 121;
                    lv[t] = on ? (int)cL[(py - w) * pl + (px - w)] - centreL : 0;
                    A0[t] = A1[t] = A2[t] = 0;
                    if (on) {
                        const uint8_t* rp = cR + (py - w) * pr + (px - w) - Ls;         // byte under shift -5
                        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(rp) & 3);
                        const uint32_t* q = reinterpret_cast<const uint32_t*>(rp - mis);
                        const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = mis >= 2 ? __ldg(q + 3) : 0u;   // 11 bytes from offset mis
                        A0[t] = __funnelshift_r(w0, w1, 8 * mis);
                        A1[t] = __funnelshift_r(w1, w2, 8 * mis);
                        A2[t] = __funnelshift_r(w2, w3, 8 * mis);
                    }
                }
                const unsigned onMask = (lane + 96 < 121) ? 0xfu : 0x7u;                 // which t are patch pixels for this lane
                unsigned c0, c1, c2;                                                      // the centre row's bytes: cR[-5..5]
                {
                    const uint8_t* rp = cR - Ls;
                    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(rp) & 3);
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(rp - mis);
                    const unsigned w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = mis >= 2 ? __ldg(q + 3) : 0u;   // 11 bytes from offset mis
                    c0 = __funnelshift_r(w0, w1, 8 * mis); c1 = __funnelshift_r(w1, w2, 8 * mis); c2 = __funnelshift_r(w2, w3, 8 * mis);
                }
                int bestSad = 0x7fffffff, bestInc = 0;
                int mySad = 0;                                                            // lane k keeps the SAD of shift k - 5
#pragma unroll
                for (int inc = -Ls; inc <= Ls; inc++) {
                    const int k = inc + Ls;
                    const unsigned sel = 0x4440u | (unsigned)(k & 3);                     // byte k & 3 of the word, zero extended
                    const int centreR = (int)__byte_perm(k < 4 ? c0 : k < 8 ? c1 : c2, 0u, sel);
                    int s = 0;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const int r = (int)__byte_perm(k < 4 ? A0[t] : k < 8 ? A1[t] : A2[t], 0u, sel);
                        if (t < 3 || onMask == 0xfu) s = (int)__sad(lv[t] + centreR, r, (unsigned)s);
                    }
                    s = __reduce_add_sync(0xffffffffu, s);
                    if (lane == k) mySad = s;
                    if (s < bestSad) { bestSad = s; bestInc = inc; }
                }
                if (bestInc != -Ls && bestInc != Ls) {
                    const float d1 = (float)__shfl_sync(0xffffffffu, mySad, bestInc + Ls - 1);
                    const float d2 = (float)__shfl_sync(0xffffffffu, mySad, bestInc + Ls);
                    const float d3 = (float)__shfl_sync(0xffffffffu, mySad, bestInc + Ls + 1);
                    const float deltaR = __fdiv_rn(__fsub_rn(d1, d3),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = __fmul_rn(lg.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestInc), deltaR));
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= A.minD && disparity < A.maxD) {
                            if (disparity <= 0) {
                                disparity = 0.01f;
                                bestuR = (float)__dsub_rn((double)uL, 0.01);
                            }
                            outD = __fdiv_rn(A.mbf, disparity);
                            outU = bestuR;
                            outSad = bestSad;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) {
        A.uRight[(size_t)img * g.kpCap + iL] = outU;
        A.depth[(size_t)img * g.kpCap + iL] = outD;
        A.sad[(size_t)img * g.kpCap + iL] = outSad;
    }
}

// median of the accepted SADs (element [m/2] of the sorted list, :867) and the outlier cut.  The k-th smallest of the 16-bit
// values (SAD <= 121 * 510 < 65536) is selected by two 256-bin histograms -- high byte, then low byte inside the selected bin --
// each followed by one block scan: six barriers instead of the fifty of a bit-by-bit radix selection (the kernel is one CTA per
// frame and sits on the critical path of a live frame).
__device__ __forceinline__ void select_bin(int* hist, int* sWarp, int* sSel, int k, int tid) {
    // thread t owns bin t: exclusive prefix over the 256 bins, the bin whose range holds rank k publishes (bin, rank inside the bin)
    const int lane = tid & 31, warp = tid >> 5;
    const int v = hist[tid];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) if (w < warp) base += sWarp[w];
    const int excl = base + incl - v;
    if (k >= excl && k < excl + v) { sSel[0] = tid; sSel[1] = k - excl; }
    __syncthreads();
}

__global__ void __launch_bounds__(256) k_stereo_filter(const __grid_constant__ StereoArgs A) {
    __shared__ int hist[256];
    __shared__ int sWarp[8];
    __shared__ int sSel[2];
    __shared__ int sCnt;
    pdl_entry();
    const int tid = threadIdx.x, img = blockIdx.x;
    const int cap = A.g.kpCap;
    const int nL = min(*reinterpret_cast<const int*>(A.recL + (size_t)img * A.recordBytes), cap);
    const int* sad = A.sad + (size_t)img * cap;
    float* uRightOut = A.uRight + (size_t)img * cap;
    float* depthOut = A.depth + (size_t)img * cap;

    hist[tid] = 0;
    if (tid == 0) sCnt = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0) { mine++; atomicAdd(&hist[(s >> 8) & 255], 1); }
    }
    if (mine) atomicAdd(&sCnt, mine);
    __syncthreads();
    const int m = sCnt;
    if (m == 0) return;                          // the reference indexes an empty vector here (undefined); nothing to filter
    select_bin(hist, sWarp, sSel, m / 2, tid);
    const int hiByte = sSel[0], kIn = sSel[1];
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && (s >> 8) == hiByte) atomicAdd(&hist[s & 255], 1);
    }
    __syncthreads();
    select_bin(hist, sWarp, sSel, kIn, tid);
    const float median = (float)((hiByte << 8) | sSel[0]);
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && !((float)s < thDist)) { uRightOut[i] = -1.0f; depthOut[i] = -1.0f; }
    }
}

}  // namespace

cudaError_t launch_stereo(const StereoArgs& a, int nimg, cudaStream_t st) {
    const size_t smemRows = (size_t)(a.g.lv[0].h + 1) * sizeof(int);
    if (smemRows > 48 * 1024) {
        cudaError_t e = OBS_ALLOW_MAX_SMEM(k_stereo_rows);
        if (e != cudaSuccess) return e;
    }
    cudaError_t le = launch_k(pdl_enabled(), k_stereo_rows, dim3(nimg), dim3(SR_THREADS), smemRows, st, a);
    if (le != cudaSuccess) return le;
    dim3 grid((a.g.kpCap + ST_WARPS - 1) / ST_WARPS, nimg);
    le = launch_k(pdl_enabled(), k_stereo_match, grid, dim3(ST_WARPS * 32), 0, st, a);
    if (le != cudaSuccess) return le;
    le = launch_k(pdl_enabled(), k_stereo_filter, dim3(nimg), dim3(256), 0, st, a);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}
