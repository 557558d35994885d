// K12 -- DBoW2-gated matchers: ORBmatcher::SearchByBoW (src/ORBmatcher.cc:159-288 and :522-655),
// ORBmatcher::SearchForTriangulation (:657-823, CheckDistEpipolarLine :139-156) and
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:345-410).
//
// The reference walks two DBoW2::FeatureVectors (std::map<NodeId, vector<unsigned>>) in step and matches only
// keypoints under the same vocabulary node.  A keypoint sits under exactly one node, so node pairs are independent
// problems: one CTA per keyframe pair, one warp per common node.  Inside a node SearchByBoW is sequential over the
// keypoints of side 1 (a matched keypoint of side 2 is skipped by the later ones), which the warp keeps, spreading
// the candidates of one keypoint over its lanes (best and second best by two warp minima on dist << 16 | position,
// so ties go to the first candidate like the reference's strict <).  SearchForTriangulation never marks side 2
// (the reference only reads vbMatched2), so its keypoints are independent: a lane per keypoint of side 1.
#include "matcher.h"

namespace {

__device__ __forceinline__ int hamming8(const uint32_t* a, const uint4 b0, const uint4 b1) {
    return __popc(a[0] ^ b0.x) + __popc(a[1] ^ b0.y) + __popc(a[2] ^ b0.z) + __popc(a[3] ^ b0.w) +
           __popc(a[4] ^ b1.x) + __popc(a[5] ^ b1.y) + __popc(a[6] ^ b1.z) + __popc(a[7] ^ b1.w);
}
__device__ __forceinline__ void load_desc(const uint4* p, uint32_t* d) {
    const uint4 a = p[0], b = p[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
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

// ORBmatcher.cc:240-247 (same text at :611-618, :775-782)
__device__ __forceinline__ int rot_bin(float a1, float a2) {
    float rot = __fsub_rn(a1, a2);
    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, 1.0f / OBS_HISTO_LENGTH));
    if (bin == OBS_HISTO_LENGTH) bin = 0;
    return bin;
}

// std::map::lower_bound over the ascending node ids of one frame
__device__ __forceinline__ int node_lower_bound(const uint32_t* ids, int n, uint32_t id) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ids[mid] < id) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct SideView {
    int n, nNodes;
    const uint4* desc; const float* keys; const uint8_t* valid; const float* uRight;
    const uint32_t* nodeId; const int* nodeStart; const int* nodeIdx;
};
__device__ __forceinline__ SideView side_of(const BowSideDev& S, int b) {
    SideView v;
    v.n = min(S.n[b], S.cap); v.nNodes = min(S.nNodes[b], S.nodeCap);
    v.desc = S.desc + (size_t)b * S.cap * 2;
    v.keys = S.keys + (size_t)b * S.cap * 7;
    v.valid = S.valid ? S.valid + (size_t)b * S.cap : nullptr;
    v.uRight = S.uRight ? S.uRight + (size_t)b * S.cap : nullptr;
    v.nodeId = S.nodeId + (size_t)b * S.nodeCap;
    v.nodeStart = S.nodeStart + (size_t)b * (S.nodeCap + 1);
    v.nodeIdx = S.nodeIdx + (size_t)b * S.cap;
    return v;
}

// rotation-consistency filter shared by the three searches (:260-281): matches outside the three largest bins go
__device__ void rotation_filter(const int* sHist, int* sInd, const SideView& A, const SideView& B, int* m12, int* m21, int checkOri,
                                int* sCount, int* nOut) {
    const int tid = threadIdx.x;
    if (tid == 0) { three_maxima(sHist, OBS_HISTO_LENGTH, sInd[0], sInd[1], sInd[2]); *sCount = 0; }
    __syncthreads();
    int mine = 0;
    for (int i1 = tid; i1 < A.n; i1 += blockDim.x) {
        const int i2 = m12[i1];
        if (i2 < 0) continue;
        if (checkOri) {
            const int bin = rot_bin(A.keys[i1 * 7 + 3], B.keys[i2 * 7 + 3]);
            if (bin != sInd[0] && bin != sInd[1] && bin != sInd[2]) {
                m12[i1] = -1;
                if (m21) m21[i2] = -1;
                continue;
            }
        }
        mine++;
    }
    if (mine) atomicAdd(sCount, mine);
    __syncthreads();
    if (tid == 0) *nOut = *sCount;
}

__global__ void __launch_bounds__(256) k_bow_search(const __grid_constant__ BowSearchArgs P) {
    __shared__ int sHist[OBS_HISTO_LENGTH], sInd[3], sCount;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const SideView A = side_of(P.A, b), B = side_of(P.B, b);
    int* m12 = P.match12 + (size_t)b * P.A.cap;
    int* m21 = P.match21 + (size_t)b * P.B.cap;
    for (int i = tid; i < P.A.cap; i += 256) m12[i] = -1;
    for (int i = tid; i < P.B.cap; i += 256) m21[i] = -1;
    if (tid < OBS_HISTO_LENGTH) sHist[tid] = 0;
    __syncthreads();

    for (int ia = warp; ia < A.nNodes; ia += 8) {
        const uint32_t id = A.nodeId[ia];
        const int ib = node_lower_bound(B.nodeId, B.nNodes, id);
        if (ib >= B.nNodes || B.nodeId[ib] != id) continue;
        const int s2 = B.nodeStart[ib], e2 = B.nodeStart[ib + 1];
        for (int p1 = A.nodeStart[ia]; p1 < A.nodeStart[ia + 1]; p1++) {
            const int idx1 = A.nodeIdx[p1];
            if (A.valid && !A.valid[idx1]) continue;
            uint32_t d1[8];
            load_desc(A.desc + (size_t)idx1 * 2, d1);
            const uint32_t NONE = (256u << 16) | 0xffffu;
            uint32_t k1 = NONE, k2 = NONE;                      // the lane's two smallest keys
            for (int c = s2 + lane; c < e2; c += 32) {
                const int idx2 = B.nodeIdx[c];
                if (m21[idx2] >= 0) continue;                   // vpMapPointMatches[realIdxF] / vbMatched2[idx2]
                if (B.valid && !B.valid[idx2]) continue;
                const uint4* q = B.desc + (size_t)idx2 * 2;
                const uint32_t key = ((uint32_t)hamming8(d1, q[0], q[1]) << 16) | (uint32_t)(c - s2);
                if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
            }
            const uint32_t best = __reduce_min_sync(0xffffffffu, k1);
            const uint32_t second = __reduce_min_sync(0xffffffffu, k1 == best ? k2 : k1);
            const int bestDist1 = (int)(best >> 16), bestDist2 = (int)(second >> 16);
            if (P.strictLow ? bestDist1 < P.thLow : bestDist1 <= P.thLow) {
                if ((float)bestDist1 < __fmul_rn(P.nnratio, (float)bestDist2)) {
                    const int idx2 = B.nodeIdx[s2 + (int)(best & 0xffffu)];
                    if (lane == 0) {
                        m12[idx1] = idx2;
                        m21[idx2] = idx1;
                        if (P.checkOri) atomicAdd(&sHist[rot_bin(A.keys[idx1 * 7 + 3], B.keys[idx2 * 7 + 3])], 1);
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
    rotation_filter(sHist, sInd, A, B, m12, m21, P.checkOri, &sCount, P.nMatches + b);
}

__global__ void __launch_bounds__(256) k_tri_search(const __grid_constant__ TriSearchArgs P) {
    __shared__ int sHist[OBS_HISTO_LENGTH], sInd[3], sCount;
    __shared__ float sF[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
    const SideView A = side_of(P.A, b), B = side_of(P.B, b);
    int* m12 = P.match12 + (size_t)b * P.A.cap;
    for (int i = tid; i < P.A.cap; i += 256) m12[i] = -1;
    if (tid < OBS_HISTO_LENGTH) sHist[tid] = 0;
    if (tid < 9) sF[tid] = P.f12[(size_t)b * 9 + tid];
    __syncthreads();
    const float ex = P.epipole[2 * b], ey = P.epipole[2 * b + 1];
    const int TH_LOW = 50;

    for (int ia = warp; ia < A.nNodes; ia += 8) {
        const uint32_t id = A.nodeId[ia];
        const int ib = node_lower_bound(B.nodeId, B.nNodes, id);
        if (ib >= B.nNodes || B.nodeId[ib] != id) continue;
        const int s1 = A.nodeStart[ia], e1 = A.nodeStart[ia + 1];
        const int s2 = B.nodeStart[ib], e2 = B.nodeStart[ib + 1];
        for (int p1 = s1 + lane; p1 < e1; p1 += 32) {
            const int idx1 = A.nodeIdx[p1];
            if (A.valid && !A.valid[idx1]) continue;                          // GetMapPoint(idx1) != NULL
            const bool stereo1 = A.uRight && A.uRight[idx1] >= 0;
            if (P.onlyStereo && !stereo1) continue;
            const float x1 = A.keys[idx1 * 7], y1 = A.keys[idx1 * 7 + 1];
            // epipolar line in the second image l = x1' F12 = [a b c], :142-144
            const float la = __fadd_rn(__fadd_rn(__fmul_rn(x1, sF[0]), __fmul_rn(y1, sF[3])), sF[6]);
            const float lb = __fadd_rn(__fadd_rn(__fmul_rn(x1, sF[1]), __fmul_rn(y1, sF[4])), sF[7]);
            const float lc = __fadd_rn(__fadd_rn(__fmul_rn(x1, sF[2]), __fmul_rn(y1, sF[5])), sF[8]);
            const float den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
            uint32_t d1[8];
            load_desc(A.desc + (size_t)idx1 * 2, d1);
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int p2 = s2; p2 < e2; p2++) {
                const int idx2 = B.nodeIdx[p2];
                if (B.valid && !B.valid[idx2]) continue;                      // vbMatched2 is never set by the reference
                const bool stereo2 = B.uRight && B.uRight[idx2] >= 0;
                if (P.onlyStereo && !stereo2) continue;
                const uint4* q = B.desc + (size_t)idx2 * 2;
                const int dist = hamming8(d1, q[0], q[1]);
                if (dist > TH_LOW || dist > bestDist) continue;
                const float x2 = B.keys[idx2 * 7], y2 = B.keys[idx2 * 7 + 1];
                const int oct2 = __float_as_int(B.keys[idx2 * 7 + 5]);
                if (!stereo1 && !stereo2) {
                    const float dx = __fsub_rn(ex, x2), dy = __fsub_rn(ey, y2);
                    if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, P.scale[oct2])) continue;
                }
                if (den == 0) continue;
                const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, x2), __fmul_rn(lb, y2)), lc);
                const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
                if ((double)dsqr < __dmul_rn(3.84, (double)P.sigma2[oct2])) { bestIdx2 = idx2; bestDist = dist; }
            }
            if (bestIdx2 >= 0) {
                m12[idx1] = bestIdx2;
                if (P.checkOri) atomicAdd(&sHist[rot_bin(A.keys[idx1 * 7 + 3], B.keys[bestIdx2 * 7 + 3])], 1);
            }
        }
    }
    __syncthreads();
    rotation_filter(sHist, sInd, A, B, m12, nullptr, P.checkOri, &sCount, P.nMatches + b);
}

// One warp per map point; lane i owns row i of the distance matrix (rows i, i + 32, ...).  The median of a row,
// sorted[int(0.5 (N - 1))], is found by bisection on the distance value (0..256) with the row's distances
// recomputed per step -- N is a handful of observations, and nothing is stored.
__global__ void __launch_bounds__(256) k_distinctive(const uint4* __restrict__ desc, const int* __restrict__ start, int nPoints,
                                                     int* __restrict__ best) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (p >= nPoints) return;
    const int s = start[p], N = start[p + 1] - s;
    if (N <= 0) { if (lane == 0) best[p] = -1; return; }
    const int k = (N - 1) / 2;                                  // int(0.5 * (N - 1))
    uint32_t bestKey = 0xffffffffu;
    for (int i = lane; i < N; i += 32) {
        uint32_t di[8];
        load_desc(desc + (size_t)(s + i) * 2, di);
        int lo = 0, hi = 256;                                   // smallest v with #{j : d_ij <= v} >= k + 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            int cnt = 0;
            for (int j = 0; j < N; j++) {
                const uint4* q = desc + (size_t)(s + j) * 2;
                cnt += hamming8(di, q[0], q[1]) <= mid ? 1 : 0;
            }
            if (cnt >= k + 1) hi = mid; else lo = mid + 1;
        }
        bestKey = min(bestKey, ((uint32_t)lo << 16) | (uint32_t)min(i, 0xffff));
    }
    bestKey = __reduce_min_sync(0xffffffffu, bestKey);
    if (lane == 0) best[p] = (int)(bestKey & 0xffffu);
}

}  // namespace

cudaError_t launch_bow_search(const BowSearchArgs& a, int nPairs, cudaStream_t st) {
    k_bow_search<<<nPairs, 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_tri_search(const TriSearchArgs& a, int nPairs, cudaStream_t st) {
    k_tri_search<<<nPairs, 256, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_distinctive(const uint4* desc, const int* start, int nPoints, int* best, cudaStream_t st) {
    if (nPoints <= 0) return cudaSuccess;
    k_distinctive<<<(nPoints + 7) / 8, 256, 0, st>>>(desc, start, nPoints, best);
    return cudaGetLastError();
}
