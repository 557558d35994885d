// K9-K11 -- Hamming matchers.  Replace the per-frame searches of ORB_SLAM2::ORBmatcher
// (src/ORBmatcher.cc) and the keypoint grid of Frame they walk (src/Frame.cc:455-470, :567-632).
//
//   k_frame_build       one CTA per frame: unpacks cv::KeyPoint records into (x, y, uRight, octave),
//                       assigns keypoints to the 64x48 grid with the reference's round() rule
//                       (PosInGrid, Frame.cc:622-632) and writes the grid as a CSR in mGrid[ix][iy]
//                       order, ascending keypoint index inside a cell (AssignFeaturesToGrid :455-470).
//   k_proj_candidates   one thread per projected point: GetFeaturesInArea (Frame.cc:567-620) in the
//                       reference's cell order, the static filters of SearchByProjection, 256-bit
//                       Hamming distances (__popc); the candidates (dist|level|index) go to a list of
//                       8-slot chunks: 7 inline per point, further chunks from a per-frame pool.
//   k_proj_resolve      one CTA per frame: the reference assigns map points to keypoints first come,
//                       first served (a keypoint that received a point with Observations()>0 is
//                       skipped by every later point, ORBmatcher.cc:87-89 / :1391-1393).  Point i
//                       therefore sees keypoint k iff no earlier point locked it: T[k] = index of the
//                       first locking point.  T is found by fixed-point iteration over blocks of 1024
//                       consecutive points: all points of a block re-evaluate their candidates against
//                       the previous claims in parallel and the new claims are the atomicMin of the
//                       locking ones.  By induction on the point index the fixed point is unique and
//                       equals the sequential result (point i is final once points 0..i-1 are); a
//                       converged block's locks are final for all later blocks.  Then: match counts,
//                       rotation histogram, ComputeThreeMaxima (:1601-1642), reset of the inconsistent bins.
//   k_init_search       SearchForInitialization (:405-520): candidate lists in parallel, then one warp
//                       replays the frame-1 keypoints in order (vMatchedDistance / vnMatches21 state in
//                       shared memory; candidates of one keypoint spread over the lanes).
//   k_knn2              brute-force best / second-best with ratio test: one query descriptor per thread
//                       held in registers, database descriptors broadcast from shared memory.
// Float expressions use individually rounded binary32 operations (__fmul_rn & co), as the oracle.
#include "matcher.h"
#include <algorithm>
#include <cstdlib>

namespace {

constexpr int TH_HIGH = 100, TH_LOW = 50;
constexpr uint32_t CAND_EMPTY = 0xffffffffu;
constexpr uint32_t CAND_LINK = 0x80000000u;      // slot 7 of a chunk: index of the next chunk in the frame's pool
constexpr uint32_t CAND_TRUNC = 0x40000000u;     // on the first entry: the pool ran out, the list is incomplete
constexpr int T_FREE = 0x7fffffff;

__device__ __forceinline__ int hamming8(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}

__device__ __forceinline__ void load_desc(const uint4* p, uint32_t* d) {
    const uint4 a = __ldg(p), b = __ldg(p + 1);
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
}

// ------------------------------------------------------------------------------------------------
// Frame set build
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_frame_build(const __grid_constant__ FrameBuildArgs A) {
    __shared__ int sCnt[OBS_GRID_CELLS];
    __shared__ int sWarp[9];
    const FrameSetDev& F = A.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const int n = max(0, min(A.count[(size_t)b * A.countStrideInts], F.cap));
    const uint8_t* keys = A.keys + (size_t)b * A.keysFrameStride;
    const uint4* dsrc = reinterpret_cast<const uint4*>(A.desc + (size_t)b * A.descFrameStride);
    const float* ur = A.uRight ? A.uRight + (size_t)b * A.uRightFrameStride : nullptr;
    float4* kp = F.kp + (size_t)b * F.cap;
    float* ang = F.angle + (size_t)b * F.cap;
    uint4* dd = F.desc + (size_t)b * F.cap * 2;
    int* cellStart = F.cellStart + (size_t)b * (OBS_GRID_CELLS + 1);
    uint16_t* cellIdx = F.cellIdx + (size_t)b * F.cap;
    if (tid == 0) F.n[b] = n;
    for (int c = tid; c < OBS_GRID_CELLS; c += 256) sCnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const float* k = reinterpret_cast<const float*>(keys + (size_t)i * 28);
        const float x = k[0], y = k[1];
        const int oct = reinterpret_cast<const int*>(k)[5];
        kp[i] = make_float4(x, y, ur ? ur[i] : -1.0f, __int_as_float(oct));
        ang[i] = k[3];
        dd[2 * i] = dsrc[2 * i];
        dd[2 * i + 1] = dsrc[2 * i + 1];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(x, F.P.minX), F.P.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(y, F.P.minY), F.P.invH));
        if (px >= 0 && px < OBS_GRID_COLS && py >= 0 && py < OBS_GRID_ROWS) atomicAdd(&sCnt[px * OBS_GRID_ROWS + py], 1);
    }
    __syncthreads();
    // exclusive scan over the 3072 cells: 12 consecutive cells per thread
    constexpr int PER = OBS_GRID_CELLS / 256;
    int loc[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) { loc[j] = sCnt[tid * PER + j]; sum += loc[j]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) sWarp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = lane < 8 ? sWarp[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
        if (lane < 8) sWarp[lane] = wi - w;
        if (lane == 7) sWarp[8] = wi;
    }
    __syncthreads();
    int run = sWarp[warp] + incl - sum;
#pragma unroll
    for (int j = 0; j < PER; j++) { sCnt[tid * PER + j] = run; cellStart[tid * PER + j] = run; run += loc[j]; }
    if (tid == 255) cellStart[OBS_GRID_CELLS] = sWarp[8];
    __syncthreads();
    for (int i = tid; i < n; i += 256) {
        const float4 k = kp[i];
        const int px = (int)roundf(__fmul_rn(__fsub_rn(k.x, F.P.minX), F.P.invW));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(k.y, F.P.minY), F.P.invH));
        if (px >= 0 && px < OBS_GRID_COLS && py >= 0 && py < OBS_GRID_ROWS)
            cellIdx[atomicAdd(&sCnt[px * OBS_GRID_ROWS + py], 1)] = (uint16_t)i;
    }
    __syncthreads();
    // ascending keypoint index inside every cell (the reference appends in index order)
    for (int c = tid; c < OBS_GRID_CELLS; c += 256) {
        const int s = cellStart[c], e = sCnt[c];
        for (int i = s + 1; i < e; i++) {
            const uint16_t v = cellIdx[i];
            int j = i - 1;
            while (j >= s && cellIdx[j] > v) { cellIdx[j + 1] = cellIdx[j]; j--; }
            cellIdx[j + 1] = v;
        }
    }
}

// Frame::GetFeaturesInArea (Frame.cc:567-620): calls fn(idx, kp) for every keypoint it would return, in its order.
// A column ix of the window is one contiguous run of the CSR (cells ix*48+minY .. ix*48+maxY).
template <typename Fn>
__device__ __forceinline__ void for_each_in_area(const FrameSetDev& F, int b, float x, float y, float r, int minLevel,
                                                 int maxLevel, Fn&& fn) {
    const FrameParamsDev& P = F.P;
    const float dx = __fsub_rn(x, P.minX), dy = __fsub_rn(y, P.minY);
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(dx, r), P.invW)));
    if (nMinCellX >= OBS_GRID_COLS) return;
    const int nMaxCellX = min(OBS_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dx, r), P.invW)));
    if (nMaxCellX < 0) return;
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(dy, r), P.invH)));
    if (nMinCellY >= OBS_GRID_ROWS) return;
    const int nMaxCellY = min(OBS_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(dy, r), P.invH)));
    if (nMaxCellY < 0) return;
    const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    const int* cs = F.cellStart + (size_t)b * (OBS_GRID_CELLS + 1);
    const uint16_t* ci = F.cellIdx + (size_t)b * F.cap;
    const float4* kp = F.kp + (size_t)b * F.cap;
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
        if (nMinCellY > nMaxCellY) break;
        const int s = __ldg(cs + ix * OBS_GRID_ROWS + nMinCellY), e = __ldg(cs + ix * OBS_GRID_ROWS + nMaxCellY + 1);
        for (int j = s; j < e; j++) {
            const int idx = __ldg(ci + j);
            const float4 k = __ldg(kp + idx);
            const int oct = __float_as_int(k.w);
            if (bCheckLevels) {
                if (oct < minLevel) continue;
                if (maxLevel >= 0 && oct > maxLevel) continue;
            }
            if (fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r) fn(idx, k);
        }
    }
}

// ORBmatcher.cc:1601-1642
__device__ void three_maxima(const int* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    ind1 = ind2 = ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
}

// ORBmatcher.cc:1425-1433
__device__ __forceinline__ int rot_bin(float a1, float a2) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / OBS_HISTO_LENGTH));
    if (bin == OBS_HISTO_LENGTH) bin = 0;
    return bin;
}

// ------------------------------------------------------------------------------------------------
// Projection searches
// ------------------------------------------------------------------------------------------------
struct Query {
    float u, v, r, ur, rthr;
    int minLevel, maxLevel;
    bool valid, locks, stereo;
};

// cv::Mat A(3x3)*x+c on floats: binary32 products and sums left to right, the final addition in binary64
// (cv::gemm's small-matrix branch; see oracle/match_oracle.cpp mat_rx_plus_t)
__device__ __forceinline__ float row_rx_plus_t(const float* T, int r, float x0, float x1, float x2) {
    const float t0 = __fadd_rn(__fadd_rn(__fmul_rn(T[4 * r], x0), __fmul_rn(T[4 * r + 1], x1)), __fmul_rn(T[4 * r + 2], x2));
    return __double2float_rn(__dadd_rn((double)t0, (double)T[4 * r + 3]));
}

// bForward / bBackward of ORBmatcher.cc:1339-1350: bit 0 forward, bit 1 backward
__device__ int motion_direction(const float* Tc, const float* Tl, float mb, int mono) {
    float twc[3];
    for (int r = 0; r < 3; r++) {       // -Rcw^T * tcw with binary64 accumulation (generic gemm branch)
        double s = __dmul_rn((double)Tc[r], (double)Tc[3]);
        s = __dadd_rn(s, __dmul_rn((double)Tc[4 + r], (double)Tc[7]));
        s = __dadd_rn(s, __dmul_rn((double)Tc[8 + r], (double)Tc[11]));
        twc[r] = __double2float_rn(-s);
    }
    const float tlcz = row_rx_plus_t(Tl, 2, twc[0], twc[1], twc[2]);
    const int fwd = (tlcz > mb) && !mono;
    const int bwd = (-tlcz > mb) && !mono;
    return fwd | (bwd << 1);
}

// glibc >= 2.27 logf (ARM optimized-routines; table __logf_data of this image's libm): binary64 evaluation, result
// rounded to binary32.  Equals libm's logf on all 2,130,706,432 positive normal floats (with and without FMA
// contraction of the binary64 operations; tests/test_oracle_matchers.py checks a strided sweep).
__device__ const double d_logf_tab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2}, {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2},
    {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3}, {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4}, {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5},
    {0x1p+0, 0x0p+0}, {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5}, {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3}, {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},
    {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};

__device__ float logf_glibc(float x) {
    uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        if (ix * 2 == 0) return -__int_as_float(0x7f800000);                  // log(0) = -inf
        if (ix == 0x7f800000u) return x;                                       // log(inf) = inf
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __int_as_float(0x7fc00000);   // log(negative), log(nan)
        ix = __float_as_uint(__fmul_rn(x, 8388608.0f));                        // subnormal: normalise
        ix -= 23u << 23;
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double z = (double)__uint_as_float(iz);
    const double r = __dsub_rn(__dmul_rn(z, d_logf_tab[i][0]), 1.0);
    const double y0 = __dadd_rn(d_logf_tab[i][1], __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
    y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// MapPoint::PredictScale, MapPoint.cc:488-519.  The float -> int conversion follows x86 (cvttss2si): values that do not
// fit (incl. NaN and +-inf) become INT_MIN, which the clamp below turns into level 0.
__device__ __forceinline__ int predict_scale(float maxDistRaw, float dist, float logScaleFactor, int nLevels) {
    const float ratio = __fdiv_rn(maxDistRaw, dist);
    const float q = ceilf(__fdiv_rn(logf_glibc(ratio), logScaleFactor));
    int nScale = (q >= -2147483648.0f && q < 2147483648.0f) ? (int)q : (int)0x80000000;
    if (nScale < 0) nScale = 0;
    else if (nScale >= nLevels) nScale = nLevels - 1;
    return nScale;
}

// camera centre -Rcw^T * tcw (binary64 accumulation, the generic gemm branch)
__device__ __forceinline__ void camera_centre(const float* Tc, float* Ow) {
    for (int r = 0; r < 3; r++) {
        double s = __dmul_rn((double)Tc[r], (double)Tc[3]);
        s = __dadd_rn(s, __dmul_rn((double)Tc[4 + r], (double)Tc[7]));
        s = __dadd_rn(s, __dmul_rn((double)Tc[8 + r], (double)Tc[11]));
        Ow[r] = __double2float_rn(-s);
    }
}

template <int V>
__device__ __forceinline__ Query make_query(const ProjSearchArgs& A, int b, int i, int dir) {
    Query q;
    q.valid = false; q.locks = false; q.stereo = V < 2; q.u = q.v = q.r = q.ur = q.rthr = 0.f; q.minLevel = q.maxLevel = -1;
    const FrameParamsDev& P = A.F.P;
    if (V == 0) {
        const size_t o = (size_t)b * A.mp.stride + i;
        if (!A.mp.inView[o]) return q;
        const int lvl = A.mp.level[o];
        if (lvl < 0 || lvl >= P.nlevels) return q;
        float r = ((double)A.mp.viewCos[o] > 0.998) ? 2.5f : 4.0f;          // RadiusByViewingCos, :131-137
        if (A.th != 1.0f) r = __fmul_rn(r, A.th);
        q.r = __fmul_rn(r, P.scale[lvl]);
        q.rthr = q.r;
        q.u = A.mp.projX[o]; q.v = A.mp.projY[o]; q.ur = A.mp.projXR[o];
        q.minLevel = lvl - 1; q.maxLevel = lvl;
        q.locks = A.mp.obs[o] > 0;
        q.valid = true;
    } else if (V >= 2) {
        const size_t o = (size_t)b * A.kf.stride + i;
        if (!A.kf.valid[o]) return q;
        const float* Tc = A.kf.tcw + (size_t)b * 12;
        const float* p = A.kf.pos + o * 3;
        const float x0 = p[0], x1 = p[1], x2 = p[2];
        const float xc = row_rx_plus_t(Tc, 0, x0, x1, x2);
        const float yc = row_rx_plus_t(Tc, 1, x0, x1, x2);
        const float zc = row_rx_plus_t(Tc, 2, x0, x1, x2);
        float u, v;
        if (V == 2) {                                                   // :1500-1510 (no depth-sign test here)
            const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));
            u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, xc), invzc), P.cx);
            v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, yc), invzc), P.cy);
            if (u < P.minX || u > P.maxX) return q;
            if (v < P.minY || v > P.maxY) return q;
        } else {                                                        // :320-335
            if (zc < 0.0f) return q;
            const float invz = __fdiv_rn(1.0f, zc);
            u = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(xc, invz)), P.cx);
            v = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(yc, invz)), P.cy);
            if (!(u >= P.minX && u < P.maxX && v >= P.minY && v < P.maxY)) return q;
        }
        float Ow[3];
        camera_centre(Tc, Ow);
        const float po0 = __fsub_rn(x0, Ow[0]), po1 = __fsub_rn(x1, Ow[1]), po2 = __fsub_rn(x2, Ow[2]);
        const double ss = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)),
                                    __dmul_rn((double)po2, (double)po2));
        const float dist = __double2float_rn(__dsqrt_rn(ss));            // cv::norm
        if (dist < A.kf.minDist[o] || dist > A.kf.maxDist[o]) return q;
        if (V == 3) {                                                   // viewing angle below 60 degrees, :347-350
            const float* nrm = A.kf.normal + o * 3;
            const double dt = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)nrm[0]), __dmul_rn((double)po1, (double)nrm[1])),
                                        __dmul_rn((double)po2, (double)nrm[2]));
            if (dt < __dmul_rn(0.5, (double)dist)) return q;
        }
        const int lvl = predict_scale(A.kf.maxDistRaw[o], dist, A.kf.logScaleFactor, P.nlevels);
        q.r = __fmul_rn(A.th, P.scale[lvl]);
        q.u = u; q.v = v;
        q.minLevel = lvl - 1; q.maxLevel = V == 2 ? lvl + 1 : lvl;
        q.locks = true;                                                 // any assigned point blocks the keypoint (:1536, :367)
        q.valid = true;
    } else {
        const size_t o = (size_t)b * A.lf.stride + i;
        if (!A.lf.hasPoint[o]) return q;
        const float* Tc = A.lf.tcwCur + (size_t)b * 12;
        const float* p = A.lf.pos + o * 3;
        const float x0 = p[0], x1 = p[1], x2 = p[2];
        const float xc = row_rx_plus_t(Tc, 0, x0, x1, x2);
        const float yc = row_rx_plus_t(Tc, 1, x0, x1, x2);
        const float zc = row_rx_plus_t(Tc, 2, x0, x1, x2);
        const float invzc = __double2float_rn(__ddiv_rn(1.0, (double)zc));
        if (invzc < 0) return q;
        const float u = __fadd_rn(__fmul_rn(__fmul_rn(P.fx, xc), invzc), P.cx);
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(P.fy, yc), invzc), P.cy);
        if (u < P.minX || u > P.maxX) return q;
        if (v < P.minY || v > P.maxY) return q;
        const int oct = A.lf.octave[o];
        if (oct < 0 || oct >= P.nlevels) return q;
        q.r = __fmul_rn(A.th, P.scale[oct]);
        q.rthr = q.r;
        q.u = u; q.v = v;
        q.ur = __fsub_rn(u, __fmul_rn(P.mbf, invzc));
        if (dir & 1) { q.minLevel = oct; q.maxLevel = -1; }
        else if (dir & 2) { q.minLevel = 0; q.maxLevel = oct; }
        else { q.minLevel = oct - 1; q.maxLevel = oct + 1; }
        q.locks = A.lf.obs[o] > 0;
        q.valid = true;
    }
    return q;
}

template <int V> __device__ __forceinline__ int num_points(const ProjSearchArgs& A) { return V == 0 ? A.mp.n : V == 1 ? A.lf.n : A.kf.n; }
template <int V> __device__ __forceinline__ const uint4* point_desc(const ProjSearchArgs& A, int b, int i) {
    return V == 0 ? A.mp.desc + ((size_t)b * A.mp.stride + i) * 2
         : V == 1 ? A.lf.desc + ((size_t)b * A.lf.stride + i) * 2 : A.kf.desc + ((size_t)b * A.kf.stride + i) * 2;
}
template <int V> __device__ __forceinline__ bool point_locks(const ProjSearchArgs& A, int b, int i) {
    return V == 0 ? A.mp.obs[(size_t)b * A.mp.stride + i] > 0 : V == 1 ? A.lf.obs[(size_t)b * A.lf.stride + i] > 0 : true;
}
template <int V> __device__ __forceinline__ float point_angle(const ProjSearchArgs& A, int b, int i) {
    return V == 1 ? A.lf.angle[(size_t)b * A.lf.stride + i] : A.kf.angle[(size_t)b * A.kf.stride + i];
}

// the keypoint-side filters that do not change during the call (ORBmatcher.cc:87-96 / :1391-1402)
__device__ __forceinline__ bool static_ok(const ProjSearchArgs& A, int b, int idx, const float4& k, const Query& q) {
    if (A.kpObs && A.kpObs[(size_t)b * A.F.cap + idx] > 0) return false;
    if (q.stereo && k.z > 0) {
        const float er = fabsf(__fsub_rn(q.ur, k.z));
        if (er > q.rthr) return false;
    }
    return true;
}

struct Best {
    int d1 = 256, l1 = -1, d2 = 256, l2 = -1, idx = -1;
    __device__ __forceinline__ void upd(int dist, int lvl, int i) {
        if (dist < d1) { d2 = d1; d1 = dist; l2 = l1; l1 = lvl; idx = i; }
        else if (dist < d2) { l2 = lvl; d2 = dist; }
    }
};

template <int V>
__global__ void __launch_bounds__(128) k_proj_candidates(const __grid_constant__ ProjSearchArgs A) {
    __shared__ int sDir;
    const int b = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (V == 1) {
        if (threadIdx.x == 0)
            sDir = motion_direction(A.lf.tcwCur + (size_t)b * 12, A.lf.tcwLast + (size_t)b * 12, A.F.P.mb, A.lf.mono);
        __syncthreads();
    }
    const int M = num_points<V>(A);
    if (i >= M) return;
    uint32_t* chunk = A.cand + ((size_t)b * M + i) * OBS_CAND_SLOTS;
    const Query q = make_query<V>(A, b, i, V == 1 ? sDir : 0);
    if (!q.valid) { chunk[0] = CAND_EMPTY; return; }
    uint32_t d[8];
    load_desc(point_desc<V>(A, b, i), d);
    const uint4* fd = A.F.desc + (size_t)b * A.F.cap * 2;
    uint32_t* const head = chunk;
    uint32_t* pool = A.pool + (size_t)b * A.poolChunks * OBS_CAND_SLOTS;
    int pos = 0;                 // next slot in `chunk`; slot 7 is reserved for the link / terminator
    bool truncated = false;
    uint32_t first = CAND_EMPTY;
    for_each_in_area(A.F, b, q.u, q.v, q.r, q.minLevel, q.maxLevel, [&](int idx, const float4& k) {
        if (truncated || !static_ok(A, b, idx, k, q)) return;
        const int dist = hamming8(d, __ldg(fd + 2 * idx), __ldg(fd + 2 * idx + 1));
        const uint32_t e = ((uint32_t)dist << 20) | ((uint32_t)(__float_as_int(k.w) & 15) << 16) | (uint32_t)idx;
        if (pos == OBS_CAND_SLOTS - 1) {         // chunk full: continue in a chunk of the frame's pool
            const int c = atomicAdd(A.poolCursor + b, 1);
            if (c >= A.poolChunks) { truncated = true; return; }
            chunk[OBS_CAND_SLOTS - 1] = CAND_LINK | (uint32_t)c;
            chunk = pool + (size_t)c * OBS_CAND_SLOTS;
            pos = 0;
        }
        if (chunk == head && pos == 0) first = e; else chunk[pos] = e;
        pos++;
    });
    if (truncated) first |= CAND_TRUNC;          // pool exhausted: this point traverses the grid again on demand
    else chunk[pos] = CAND_EMPTY;                // pos <= 7
    if (chunk == head && pos == 0) { head[0] = CAND_EMPTY; return; }
    head[0] = first;
}

// Best candidate of point i; returns the keypoint index or -1.  A keypoint is hidden from point i when an
// earlier block of points locked it (T, final) or an earlier point of the current block claims it (Tc).
template <int V>
__device__ __forceinline__ int evaluate(const ProjSearchArgs& A, int b, int i, int M, const int* T, const int* Tc, int dir) {
    const uint4* c4 = reinterpret_cast<const uint4*>(A.cand + ((size_t)b * M + i) * OBS_CAND_SLOTS);
    uint4 lo = c4[0];
    if (lo.x == CAND_EMPTY) return -1;
    Best B;
    if (!(lo.x & CAND_TRUNC)) {
        const uint4* pool = reinterpret_cast<const uint4*>(A.pool + (size_t)b * A.poolChunks * OBS_CAND_SLOTS);
        for (;;) {
            const uint4 hi = c4[1];
            const uint32_t e[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            bool done = false;
#pragma unroll
            for (int s = 0; s < OBS_CAND_SLOTS - 1; s++) {
                if (done || e[s] == CAND_EMPTY) { done = true; continue; }
                const int k = (int)(e[s] & 0xffffu);
                if (T[k] != T_FREE || Tc[k] < i) continue;
                B.upd((int)((e[s] >> 20) & 0x1ffu), (int)((e[s] >> 16) & 15u), k);
            }
            if (done || e[7] == CAND_EMPTY) break;
            c4 = pool + (size_t)(e[7] & 0x3fffffffu) * 2;
            lo = c4[0];
        }
    } else {
        const Query q = make_query<V>(A, b, i, dir);
        uint32_t d[8];
        load_desc(point_desc<V>(A, b, i), d);
        const uint4* fd = A.F.desc + (size_t)b * A.F.cap * 2;
        for_each_in_area(A.F, b, q.u, q.v, q.r, q.minLevel, q.maxLevel, [&](int idx, const float4& k) {
            if (T[idx] != T_FREE || Tc[idx] < i) return;
            if (!static_ok(A, b, idx, k, q)) return;
            B.upd(hamming8(d, __ldg(fd + 2 * idx), __ldg(fd + 2 * idx + 1)), __float_as_int(k.w) & 15, idx);
        });
    }
    if (B.d1 > (V >= 2 ? A.kf.distTh : TH_HIGH)) return -1;
    if (V == 0) {
        if (B.l1 == B.l2 && (float)B.d1 > __fmul_rn(A.nnratio, (float)B.d2)) return -1;     // :118-121
    }
    return B.idx;
}

template <int V>
__global__ void __launch_bounds__(1024) k_proj_resolve(const __grid_constant__ ProjSearchArgs A) {
    extern __shared__ int sm[];
    __shared__ int sHist[OBS_HISTO_LENGTH];
    __shared__ int sInd[3];
    __shared__ int sN, sDir;
    const int cap = A.F.cap, b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    int* T = sm;                // first locking point per keypoint, final for the blocks already processed
    int* Tc = sm + cap;         // claims of the current block (previous iteration)
    int* Tn = sm + 2 * cap;     // claims of the current block (this iteration)
    int* last = sm + 3 * cap;   // last non-locking point assigned per keypoint; afterwards: "reset by the rotation check"
    const int M = num_points<V>(A);
    const int n = A.F.n[b];
    const bool ori = (V == 1 && A.lf.checkOri) || (V == 2 && A.kf.checkOri);
    if (tid == 0) {
        sN = 0;
        sDir = V == 1 ? motion_direction(A.lf.tcwCur + (size_t)b * 12, A.lf.tcwLast + (size_t)b * 12, A.F.P.mb, A.lf.mono) : 0;
    }
    if (tid < OBS_HISTO_LENGTH) sHist[tid] = 0;
    for (int k = tid; k < cap; k += nt) { T[k] = T_FREE; Tc[k] = T_FREE; Tn[k] = T_FREE; last[k] = -1; }
    __syncthreads();
    const int dir = sDir;
    int iters = 0, mine = 0;
    // Points are taken in blocks of blockDim consecutive indices.  Inside a block the claims are iterated to
    // their fixed point (point i only depends on points < i, so the fixed point is the sequential result);
    // then the block's locks become final and the next block starts.
    for (int base = 0; base < M; base += nt) {
        const int i = base + tid;
        const bool active = i < M;
        const bool locks = active && point_locks<V>(A, b, i);
        int k = -1;
        for (;;) {
            k = active ? evaluate<V>(A, b, i, M, T, Tc, dir) : -1;
            if (locks && k >= 0) atomicMin(&Tn[k], i);
            iters++;
            __syncthreads();
            int changed = 0;
            for (int kk = tid; kk < cap; kk += nt) {
                const int v = Tn[kk];
                changed |= (v != Tc[kk]);
                Tc[kk] = v;
                Tn[kk] = T_FREE;
            }
            if (!__syncthreads_or(changed)) break;
        }
        // the block has converged: k is final
        if (k >= 0) {
            mine++;
            if (!locks) atomicMax(&last[k], i);
            if (ori) {
                const int bin = rot_bin(point_angle<V>(A, b, i), A.F.angle[(size_t)b * cap + k]);
                atomicAdd(&sHist[bin], 1);
                k |= bin << 16;
            }
        }
        if (active) A.choice[(size_t)b * M + i] = k;
        __syncthreads();
        for (int kk = tid; kk < cap; kk += nt) {
            if (Tc[kk] != T_FREE) { T[kk] = Tc[kk]; Tc[kk] = T_FREE; }       // a keypoint is claimed in one block only
        }
        __syncthreads();
    }
    if (mine) atomicAdd(&sN, mine);
    __syncthreads();
    // final owner per keypoint: the locking point if there is one (nothing is assigned after it), else the last
    // non-locking point
    for (int k = tid; k < cap; k += nt) {
        int out = -1;
        if (k < n) out = T[k] != T_FREE ? T[k] : last[k];
        Tn[k] = out;
        last[k] = 0;
    }
    __syncthreads();
    if (ori) {
        if (tid == 0) three_maxima(sHist, OBS_HISTO_LENGTH, sInd[0], sInd[1], sInd[2]);
        __syncthreads();
        int drop = 0;
        for (int i = tid; i < M; i += nt) {
            const int c = A.choice[(size_t)b * M + i];
            if (c < 0) continue;
            const int bin = c >> 16;
            if (bin != sInd[0] && bin != sInd[1] && bin != sInd[2]) { last[c & 0xffff] = 1; drop++; }
        }
        if (drop) atomicSub(&sN, drop);
        __syncthreads();
    }
    for (int k = tid; k < cap; k += nt) A.kpMatch[(size_t)b * cap + k] = last[k] ? -2 : Tn[k];
    if (tid == 0) { A.nMatches[b] = sN; if (A.rounds) A.rounds[b] = iters; }
}

// ------------------------------------------------------------------------------------------------
// SearchForInitialization
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) k_init_search(const __grid_constant__ InitSearchArgs A) {
    extern __shared__ int sm[];
    __shared__ int sHist[OBS_HISTO_LENGTH];
    __shared__ int sInd[3];
    __shared__ int sN;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, nt = blockDim.x;
    const int cap1 = A.F1.cap, cap2 = A.F2.cap;
    int* vMD = sm;                    // vMatchedDistance [cap2]
    int* m21 = sm + cap2;             // vnMatches21 [cap2]
    int* m12 = sm + 2 * cap2;         // vnMatches12 [cap1]
    int* binOf = sm + 2 * cap2 + cap1;   // rotation bin of the match pushed for i1, -1 = none [cap1]
    const int n1 = A.F1.n[b], n2 = A.F2.n[b];
    const float4* kp1 = A.F1.kp + (size_t)b * cap1;
    const float4* kp2 = A.F2.kp + (size_t)b * cap2;
    const uint4* d1 = A.F1.desc + (size_t)b * cap1 * 2;
    const uint4* d2 = A.F2.desc + (size_t)b * cap2 * 2;
    float* prev = A.prevMatched + (size_t)b * cap1 * 2;
    uint32_t* list = A.list + (size_t)b * cap1 * A.listCap;
    int* cnt = A.listCount + (size_t)b * cap1;
    if (tid == 0) sN = 0;
    if (tid < OBS_HISTO_LENGTH) sHist[tid] = 0;
    for (int k = tid; k < cap2; k += nt) { vMD[k] = 0x7fffffff; m21[k] = -1; }
    for (int k = tid; k < cap1; k += nt) { m12[k] = -1; binOf[k] = -1; }
    // candidate lists (:417-425, :434-438), all frame-1 keypoints in parallel
    for (int i1 = tid; i1 < n1; i1 += nt) {
        int c = 0;
        if (__float_as_int(kp1[i1].w) <= 0) {
            uint32_t q[8];
            load_desc(d1 + 2 * i1, q);
            uint32_t* li = list + (size_t)i1 * A.listCap;
            for_each_in_area(A.F2, b, prev[2 * i1], prev[2 * i1 + 1], (float)A.window, 0, 0, [&](int idx, const float4&) {
                if (c < A.listCap) li[c] = ((uint32_t)hamming8(q, __ldg(d2 + 2 * idx), __ldg(d2 + 2 * idx + 1)) << 16) | (uint32_t)idx;
                c++;
            });
        }
        cnt[i1] = min(c, A.listCap);
    }
    __syncthreads();
    // ordered replay by one warp (:414-487)
    if (tid < 32) {
        for (int i1 = 0; i1 < n1; i1++) {
            const int c = cnt[i1];
            if (c == 0) continue;
            const uint32_t* li = list + (size_t)i1 * A.listCap;
            uint32_t a1 = 0xffffffffu, a2 = 0xffffffffu;           // lane-local two smallest keys dist<<16 | slot
            for (int s = lane; s < c; s += 32) {
                const uint32_t e = li[s];
                const int dist = (int)(e >> 16);
                if (vMD[e & 0xffffu] <= dist) continue;            // :444
                const uint32_t key = ((uint32_t)dist << 16) | (uint32_t)s;
                if (key < a1) { a2 = a1; a1 = key; } else if (key < a2) a2 = key;
            }
            const uint32_t k1 = __reduce_min_sync(0xffffffffu, a1);
            if (a1 == k1) a1 = a2;                                 // keys are distinct (slot), except the empty key
            const uint32_t k2 = __reduce_min_sync(0xffffffffu, a1);
            if (lane == 0 && k1 != 0xffffffffu) {
                const int bestDist = (int)(k1 >> 16);
                const float second = k2 != 0xffffffffu ? (float)(int)(k2 >> 16) : 2147483648.0f;   // (float)INT_MAX
                if (bestDist <= TH_LOW && (float)bestDist < __fmul_rn(second, A.nnratio)) {
                    const int i2 = (int)(li[k1 & 0xffffu] & 0xffffu);
                    if (m21[i2] >= 0) { m12[m21[i2]] = -1; sN--; }
                    m12[i1] = i2;
                    m21[i2] = i1;
                    vMD[i2] = bestDist;
                    sN++;
                    if (A.checkOri) {
                        const int bin = rot_bin(A.F1.angle[(size_t)b * cap1 + i1], A.F2.angle[(size_t)b * cap2 + i2]);
                        sHist[bin]++;
                        binOf[i1] = bin;
                    }
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (A.checkOri) {
        if (tid == 0) three_maxima(sHist, OBS_HISTO_LENGTH, sInd[0], sInd[1], sInd[2]);
        __syncthreads();
        int drop = 0;
        for (int i1 = tid; i1 < n1; i1 += nt) {
            const int bin = binOf[i1];
            if (bin < 0 || bin == sInd[0] || bin == sInd[1] || bin == sInd[2]) continue;
            if (m12[i1] >= 0) { m12[i1] = -1; drop++; }
        }
        if (drop) atomicSub(&sN, drop);
        __syncthreads();
    }
    for (int i1 = tid; i1 < cap1; i1 += nt) {
        const int mt = i1 < n1 ? m12[i1] : -1;
        A.matches12[(size_t)b * cap1 + i1] = mt;
        if (mt >= 0) { prev[2 * i1] = kp2[mt].x; prev[2 * i1 + 1] = kp2[mt].y; }     // :515-517
    }
    if (tid == 0) A.nMatches[b] = sN;
    (void)n2;
}

// ------------------------------------------------------------------------------------------------
// Fuse: projection of every map point into a keyframe and the best keypoint in the window.  One thread per
// (keyframe, map point); unlike the projection searches nothing is locked during the search (the reference
// mutates the map after each point, but the search itself never reads that state), so the points are independent.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_fuse_search(const __grid_constant__ FuseSearchArgs A) {
    const int b = blockIdx.y, i = blockIdx.x * 128 + threadIdx.x;
    if (i >= A.kf.n) return;
    const FrameParamsDev& P = A.F.P;
    const size_t o = (size_t)b * A.kf.stride + i;
    const size_t out = (size_t)b * A.kf.n + i;
    int bestDist = 256, bestIdx = -1;
    do {
        if (!A.kf.valid[o]) break;
        const float* Tc = A.kf.tcw + (size_t)b * 12;
        const float* p = A.kf.pos + o * 3;
        const float x0 = p[0], x1 = p[1], x2 = p[2];
        float xc = row_rx_plus_t(Tc, 0, x0, x1, x2);
        float yc = row_rx_plus_t(Tc, 1, x0, x1, x2);
        float zc = row_rx_plus_t(Tc, 2, x0, x1, x2);
        if (A.sim3 == 2) {                                               // p3Dc2 = sR21*p3Dc1 + t21, :1169
            const float* T2 = A.tcw2 + (size_t)b * 12;
            const float a0 = xc, a1 = yc, a2 = zc;
            xc = row_rx_plus_t(T2, 0, a0, a1, a2); yc = row_rx_plus_t(T2, 1, a0, a1, a2); zc = row_rx_plus_t(T2, 2, a0, a1, a2);
        }
        if (zc < 0.0f) break;                                            // :857, :1012, :1172
        const float invz = A.sim3 ? __double2float_rn(__ddiv_rn(1.0, (double)zc)) : __fdiv_rn(1.0f, zc);      // :1016, :1175 vs :860
        const float u = __fadd_rn(__fmul_rn(P.fx, __fmul_rn(xc, invz)), P.cx);
        const float v = __fadd_rn(__fmul_rn(P.fy, __fmul_rn(yc, invz)), P.cy);
        if (!(u >= P.minX && u < P.maxX && v >= P.minY && v < P.maxY)) break;
        const float ur = __fsub_rn(u, __fmul_rn(P.mbf, invz));
        float po0 = xc, po1 = yc, po2 = zc;                              // SearchBySim3: dist3D = cv::norm(p3Dc2), :1188
        if (A.sim3 != 2) {
            float Ow[3];
            if (A.ow) { Ow[0] = A.ow[b * 3]; Ow[1] = A.ow[b * 3 + 1]; Ow[2] = A.ow[b * 3 + 2]; }
            else camera_centre(Tc, Ow);
            po0 = __fsub_rn(x0, Ow[0]); po1 = __fsub_rn(x1, Ow[1]); po2 = __fsub_rn(x2, Ow[2]);
        }
        const double ss = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)),
                                    __dmul_rn((double)po2, (double)po2));
        const float dist3D = __double2float_rn(__dsqrt_rn(ss));          // cv::norm
        if (dist3D < A.kf.minDist[o] || dist3D > A.kf.maxDist[o]) break;
        if (A.sim3 != 2) {
            const float* nrm = A.kf.normal + o * 3;
            const double dt = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)nrm[0]), __dmul_rn((double)po1, (double)nrm[1])),
                                        __dmul_rn((double)po2, (double)nrm[2]));
            if (dt < __dmul_rn(0.5, (double)dist3D)) break;              // viewing angle below 60 degrees
        }
        const int lvl = predict_scale(A.kf.maxDistRaw[o], dist3D, A.kf.logScaleFactor, P.nlevels);
        const float radius = __fmul_rn(A.th, P.scale[lvl]);
        uint32_t q[8];
        load_desc(A.kf.desc + o * 2, q);
        const uint4* fd = A.F.desc + (size_t)b * A.F.cap * 2;
        const bool sim3 = A.sim3 != 0;                                  // no reprojection gate in the Scw / Sim3 variants
        for_each_in_area(A.F, b, u, v, radius, lvl - 1, lvl, [&](int idx, const float4& k) {
            if (!sim3) {
                const int oct = __float_as_int(k.w);
                const float ex = __fsub_rn(u, k.x), ey = __fsub_rn(v, k.y);
                if (k.z >= 0.0f) {                                        // reprojection error in stereo, :909-921
                    const float er = __fsub_rn(ur, k.z);
                    const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
                    if ((double)__fmul_rn(e2, A.invSigma2[oct]) > 7.8) return;
                } else {
                    const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
                    if ((double)__fmul_rn(e2, A.invSigma2[oct]) > 5.99) return;
                }
            }
            const int dist = hamming8(q, fd[idx * 2], fd[idx * 2 + 1]);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        });
    } while (false);
    A.bestIdx[out] = bestIdx;
    A.bestDist[out] = bestDist;
}

// SearchBySim3 agreement, :1305-1319
__global__ void __launch_bounds__(256) k_sim3_agree(const int* __restrict__ idx1, const int* __restrict__ dist1, int n1,
                                                    const int* __restrict__ idx2, const int* __restrict__ dist2, int n2, int thHigh,
                                                    int* __restrict__ match12, int* __restrict__ nFound) {
    __shared__ int sCount;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) sCount = 0;
    __syncthreads();
    int mine = 0;
    for (int i1 = threadIdx.x; i1 < n1; i1 += 256) {
        int m = -1;
        const int i2 = dist1[(size_t)b * n1 + i1] <= thHigh ? idx1[(size_t)b * n1 + i1] : -1;
        if (i2 >= 0 && i2 < n2) {
            const int back = dist2[(size_t)b * n2 + i2] <= thHigh ? idx2[(size_t)b * n2 + i2] : -1;
            if (back == i1) { m = i2; mine++; }
        }
        match12[(size_t)b * n1 + i1] = m;
    }
    if (mine) atomicAdd(&sCount, mine);
    __syncthreads();
    if (threadIdx.x == 0) nFound[b] = sCount;
}

__global__ void k_three_maxima(const int* binSizes, int nHist, int length, int* ind) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nHist) return;
    int i1, i2, i3;
    three_maxima(binSizes + (size_t)h * length, length, i1, i2, i3);
    ind[3 * h] = i1; ind[3 * h + 1] = i2; ind[3 * h + 2] = i3;
}

__global__ void k_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int* dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a) + (size_t)i * 8;
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(b) + (size_t)i * 8;
    int d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) d += __popc(pa[k] ^ pb[k]);
    dist[i] = d;
}

// ------------------------------------------------------------------------------------------------
// Brute-force best / second best (K11)
// ------------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 256;
constexpr int KNN_TILE = 256;          // database descriptors staged per step (8 KB)

// The POPC pipe issues 16 lanes/clk/SM against 64 for LOP3: carry-save adders (2 LOP3 each) fold the eight xor
// words into fewer population counts.  sum(x0..x6) = s3 + 2*(c1+c2+c3) = s3 + 2*(s4 + 2*c4).
__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}
template <int MODE>
__device__ __forceinline__ int hamming_knn(const uint32_t* q, const uint4 b0, const uint4 b1) {
    if (MODE == 0) return hamming8(q, b0, b1);
    const uint32_t x0 = q[0] ^ b0.x, x1 = q[1] ^ b0.y, x2 = q[2] ^ b0.z, x3 = q[3] ^ b0.w;
    const uint32_t x4 = q[4] ^ b1.x, x5 = q[5] ^ b1.y, x6 = q[6] ^ b1.z, x7 = q[7] ^ b1.w;
    uint32_t s1, c1, s2, c2, s3, c3;
    csa(x0, x1, x2, s1, c1);
    csa(x3, x4, x5, s2, c2);
    csa(s1, s2, x6, s3, c3);
    if (MODE == 1) return __popc(s3) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
    uint32_t s4, c4;
    csa(c1, c2, c3, s4, c4);
    return __popc(s3) + __popc(x7) + 2 * __popc(s4) + 4 * __popc(c4);
}

template <int MODE>
__global__ void __launch_bounds__(KNN_THREADS) k_knn2(const __grid_constant__ Knn2Args A) {
    __shared__ uint4 sDb[KNN_TILE * 2];
    const int tid = threadIdx.x;
    const int qTiles = (A.n + KNN_THREADS - 1) / KNN_THREADS;
    const int pair = blockIdx.x / qTiles;
    const int2 pr = A.pairs[pair];
    const int qi = (blockIdx.x - pair * qTiles) * KNN_THREADS + tid;
    const uint4* Q = A.desc + (size_t)pr.x * A.n * 2;
    const uint4* D = A.desc + (size_t)pr.y * A.n * 2;
    uint32_t q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (qi < A.n) load_desc(Q + 2 * qi, q);
    // keys dist << 16 | j are distinct, so the two smallest keys are the reference's (best, second best)
    // with its strict-< / first-wins order (ORBmatcher.cc:213-222)
    uint32_t best = (256u << 16) | 0xffffu, second = (256u << 16) | 0xffffu;
    for (int base = 0; base < A.n; base += KNN_TILE) {
        const int m = min(KNN_TILE, A.n - base);
        __syncthreads();
        for (int t = tid; t < m * 2; t += KNN_THREADS) sDb[t] = __ldg(D + (size_t)base * 2 + t);
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < m; j++) {
            const uint32_t key = ((uint32_t)hamming_knn<MODE>(q, sDb[2 * j], sDb[2 * j + 1]) << 16) | (uint32_t)(base + j);
            const uint32_t hi = max(key, best);
            best = min(key, best);
            second = min(second, hi);
        }
    }
    if (qi < A.n) {
        const int bd = (int)(best >> 16), sd = (int)(second >> 16);
        const size_t o = (size_t)pair * A.n + qi;
        int idx = -1;
        if (bd <= A.thLow && (float)bd < __fmul_rn(A.nnratio, (float)sd)) idx = (int)(best & 0xffffu);
        A.bestIdx[o] = idx;
        if (A.bestDist) A.bestDist[o] = bd;
        if (A.secondDist) A.secondDist[o] = sd;
    }
}

}  // namespace

cudaError_t launch_frame_build(const FrameBuildArgs& a, int nFrames, cudaStream_t st) {
    k_frame_build<<<nFrames, 256, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_proj_search(const ProjSearchArgs& a, int variant, int nFrames, cudaStream_t st) {
    const int M = variant == 0 ? a.mp.n : variant == 1 ? a.lf.n : a.kf.n;
    if (nFrames <= 0) return cudaSuccess;
    const dim3 grid((std::max(M, 1) + 127) / 128, nFrames);
    const size_t smem = (size_t)a.F.cap * 4 * sizeof(int);
    cudaError_t e;
    e = cudaMemsetAsync(a.poolCursor, 0, (size_t)nFrames * sizeof(int), st);
    if (e != cudaSuccess) return e;
#define OBS_PROJ_CASE(VV)                                                                                                   \
    case VV:                                                                                                                \
        e = OBS_ALLOW_MAX_SMEM(k_proj_resolve<VV>);                                                                         \
        if (e != cudaSuccess) return e;                                                                                     \
        k_proj_candidates<VV><<<grid, 128, 0, st>>>(a);                                                                     \
        k_proj_resolve<VV><<<nFrames, 1024, smem, st>>>(a);                                                                 \
        break;
    switch (variant) {
        OBS_PROJ_CASE(0)
        OBS_PROJ_CASE(1)
        OBS_PROJ_CASE(2)
        OBS_PROJ_CASE(3)
        default: return cudaErrorInvalidValue;
    }
#undef OBS_PROJ_CASE
    return cudaGetLastError();
}

cudaError_t launch_fuse_search(const FuseSearchArgs& a, int nFrames, cudaStream_t st) {
    if (nFrames <= 0 || a.kf.n <= 0) return cudaSuccess;
    k_fuse_search<<<dim3((a.kf.n + 127) / 128, nFrames), 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_sim3_agree(const int* idx1, const int* dist1, int n1, const int* idx2, const int* dist2, int n2, int thHigh,
                              int* match12, int* nFound, int nFrames, cudaStream_t st) {
    if (nFrames <= 0) return cudaSuccess;
    k_sim3_agree<<<nFrames, 256, 0, st>>>(idx1, dist1, n1, idx2, dist2, n2, thHigh, match12, nFound);
    return cudaGetLastError();
}

cudaError_t launch_init_search(const InitSearchArgs& a, int nFrames, cudaStream_t st) {
    const size_t smem = ((size_t)a.F2.cap * 2 + (size_t)a.F1.cap * 2) * sizeof(int);
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_init_search);
    if (e != cudaSuccess) return e;
    k_init_search<<<nFrames, 512, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_three_maxima(const int* binSizes, int nHist, int length, int* ind, cudaStream_t st) {
    k_three_maxima<<<(nHist + 127) / 128, 128, 0, st>>>(binSizes, nHist, length, ind);
    return cudaGetLastError();
}

cudaError_t launch_descriptor_distance(const uint8_t* a, const uint8_t* b, int n, int* dist, cudaStream_t st) {
    k_descriptor_distance<<<(n + 255) / 256, 256, 0, st>>>(a, b, n, dist);
    return cudaGetLastError();
}

cudaError_t launch_knn2(const Knn2Args& a, cudaStream_t st) {
    if (a.nPairs <= 0 || a.n <= 0) return cudaSuccess;
    const long long blocks = (long long)((a.n + KNN_THREADS - 1) / KNN_THREADS) * a.nPairs;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    k_knn2<1><<<(unsigned)blocks, KNN_THREADS, 0, st>>>(a);          // carry-save variant (5 POPC per distance)
    return cudaGetLastError();
}
