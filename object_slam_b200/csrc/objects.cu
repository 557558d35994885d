// Object-layer pixel ops that read the same frame as the ORB front end: the keypoint-to-mask assignment at the head of
// Frame::BuildObject2DsRGBD (src/Frame.cc:240-311) and Frame::BuildObject2DsStereo (:314-385).
//
// The reference walks the semantic masks in order and, for each, the keypoints not yet taken by an earlier mask: a keypoint goes to
// the mask if the 20 x 20 window mask(int(y + row), int(x + col)), row, col in [-10, 10), is 255 everywhere and 0 < depth <= mThDepth.
// Taken keypoints are erased from the pool whether or not the mask ends up with enough keypoints to become an Object2D
// (> 5 for RGB-D, > 10 for stereo).  So "first mask that accepts the keypoint" is a per-keypoint question (k_mask_first, one
// warp per keypoint), and the Object2D numbering / in-object indices are counts and ranks over that answer (k_mask_number).
// Windows that leave the image are undefined behaviour in the reference (cv::Mat::at without a bounds check); they reject here.
#include "matcher.h"
#include <cmath>

namespace {

__global__ void __launch_bounds__(256) k_mask_first(const __grid_constant__ MaskAssignArgs A) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= A.n) return;
    const float x = A.keys[k * 7], y = A.keys[k * 7 + 1];
    const float z = A.depth[k];
    int first = -1;
    if (z > 0.0f && z <= A.thDepth) {
        // the lane's window pixels (row, col) = (e / 20 - 10, e % 20 - 10), e = lane + 32 t; indices as the reference forms them:
        // float addition, then truncation
        int off[13];
#pragma unroll
        for (int t = 0; t < 13; t++) {
            const int e = lane + 32 * t;
            off[t] = -1;
            if (e < 400) {
                const int row = e / 20 - 10, col = e - (e / 20) * 20 - 10;
                const int iy = (int)__fadd_rn(y, (float)row), ix = (int)__fadd_rn(x, (float)col);
                if (iy >= 0 && iy < A.h && ix >= 0 && ix < A.w) off[t] = iy * (int)A.rowStride + ix;
                else off[t] = -2;                                  // outside the image: rejects
            }
        }
        for (int m = 0; m < A.nMasks && first < 0; m++) {
            const uint8_t* mk = A.masks + (size_t)m * A.imageStride;
            bool ok = true;
#pragma unroll
            for (int t = 0; t < 13; t++) {
                if (off[t] == -2) ok = false;
                else if (off[t] >= 0 && __ldg(mk + off[t]) != 255) ok = false;
            }
            if (__all_sync(0xffffffffu, ok)) first = m;
        }
    }
    if (lane == 0) A.maskOfKp[k] = first;
}

// One CTA: per mask, the keypoints assigned to it in ascending index order -> (object, j); objects are numbered in mask order.
__global__ void __launch_bounds__(1024) k_mask_number(const __grid_constant__ MaskAssignArgs A) {
    __shared__ int sWarp[33];
    __shared__ int sObj;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sObj = 0;
    for (int i = tid; i < A.n; i += 1024) { A.objectKp[2 * i] = -1; A.objectKp[2 * i + 1] = -1; }
    __syncthreads();
    for (int m = 0; m < A.nMasks; m++) {
        // pass 1: count
        int mine = 0;
        for (int i = tid; i < A.n; i += 1024) mine += A.maskOfKp[i] == m;
        mine = __reduce_add_sync(0xffffffffu, mine);
        if (lane == 0) sWarp[warp] = mine;
        __syncthreads();
        if (warp == 0) {
            int v = sWarp[lane];
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) sWarp[32] = v;
        }
        __syncthreads();
        const int total = sWarp[32];
        const int obj = sObj;
        __syncthreads();
        if (total > A.minKeypoints) {
            // pass 2: rank in ascending keypoint order (chunks of 1024 with a running carry)
            int carry = 0;
            for (int base = 0; base < A.n; base += 1024) {
                const int i = base + tid;
                const int f = (i < A.n && A.maskOfKp[i] == m) ? 1 : 0;
                const unsigned bal = __ballot_sync(0xffffffffu, f);
                if (lane == 0) sWarp[warp] = __popc(bal);
                __syncthreads();
                if (warp == 0) {
                    const int v = sWarp[lane];
                    int incl = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    sWarp[lane] = incl - v;
                    if (lane == 31) sWarp[32] = incl;
                }
                __syncthreads();
                if (f) {
                    A.objectKp[2 * i] = obj;
                    A.objectKp[2 * i + 1] = carry + sWarp[warp] + __popc(bal & ((1u << lane) - 1));
                }
                carry += sWarp[32];
                __syncthreads();
            }
            if (tid == 0) { A.objectOfMask[m] = obj; sObj = obj + 1; }
        } else if (tid == 0) A.objectOfMask[m] = -1;
        __syncthreads();
    }
    if (tid == 0) *A.nObjects = sObj;
}

}  // namespace

cudaError_t launch_mask_assign(const MaskAssignArgs& a, cudaStream_t st) {
    if (a.n > 0) k_mask_first<<<(a.n + 7) / 8, 256, 0, st>>>(a);
    k_mask_number<<<1, 1024, 0, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Frame::ExtractHSVHistogramsFromMask, src/Frame.cc:388-414: cvtColor(CV_BGR2HSV) on 8-bit pixels, three calcHist
// calls under the mask (H: 30 bins over [0,180), S and V: 32 bins over [0,256)), hconcat in the order V | S | H,
// normalize(NORM_L1).  OpenCV's 8-bit BGR->HSV is fixed point (hsv_shift = 12, division tables rounded half to even);
// reproduced exactly (oracle/orc_primitives.h hsv_from_bgr, pinned on all 2^24 colours against cv2 4.13).
// ------------------------------------------------------------------------------------------------
namespace {

__constant__ int c_sdiv[256];
__constant__ int c_hdiv[256];

__device__ __forceinline__ void hsv_from_bgr(int b, int g, int r, int& h, int& s, int& v) {
    v = max(max(b, g), r);
    const int vmin = min(min(b, g), r), diff = v - vmin;
    const int vr = v == r ? -1 : 0, vg = v == g ? -1 : 0;
    s = (diff * c_sdiv[v] + (1 << 11)) >> 12;
    h = (vr & (g - b)) + (~vr & ((vg & (b - r + 2 * diff)) + ((~vg) & (r - g + 4 * diff))));
    h = (h * c_hdiv[diff] + (1 << 11)) >> 12;
    h += h < 0 ? 180 : 0;
}

// grid (row chunks, masks): shared-memory histogram of the CTA's rows, then one global atomic per non-empty bin
__global__ void __launch_bounds__(256) k_hsv_hist(const uint8_t* __restrict__ bgr, size_t bgrStride, const uint8_t* __restrict__ masks,
                                                  size_t maskStride, size_t maskImageStride, int w, int h, int rowsPerCta,
                                                  int* __restrict__ counts) {
    __shared__ int sH[OBS_HSV_BINS];
    const int m = blockIdx.y;
    if (threadIdx.x < OBS_HSV_BINS) sH[threadIdx.x] = 0;
    __syncthreads();
    const int y0 = blockIdx.x * rowsPerCta, y1 = min(y0 + rowsPerCta, h);
    const uint8_t* mk = masks + (size_t)m * maskImageStride;
    for (int y = y0; y < y1; y++) {
        const uint8_t* mrow = mk + (size_t)y * maskStride;
        const uint8_t* prow = bgr + (size_t)y * bgrStride;
        for (int x = threadIdx.x; x < w; x += 256) {
            if (mrow[x] == 0) continue;
            int hh, ss, vv;
            hsv_from_bgr(prow[3 * x], prow[3 * x + 1], prow[3 * x + 2], hh, ss, vv);
            atomicAdd(&sH[vv >> 3], 1);                       // V: 32 bins over [0, 256)
            atomicAdd(&sH[32 + (ss >> 3)], 1);                // S
            if (hh < 180) atomicAdd(&sH[64 + hh / 6], 1);      // H: 30 bins over [0, 180)
        }
    }
    __syncthreads();
    if (threadIdx.x < OBS_HSV_BINS && sH[threadIdx.x]) atomicAdd(&counts[m * OBS_HSV_BINS + threadIdx.x], sH[threadIdx.x]);
}

// normalize(NORM_L1): scale = (float)(1.0 / sum) (binary64 sum and reciprocal, binary32 scale and products)
__global__ void __launch_bounds__(128) k_hsv_normalize(const int* __restrict__ counts, float* __restrict__ hist) {
    __shared__ int sSum[4];
    const int m = blockIdx.x, t = threadIdx.x;
    int v = t < OBS_HSV_BINS ? counts[m * OBS_HSV_BINS + t] : 0;
    const int ws = __reduce_add_sync(0xffffffffu, v);
    if ((t & 31) == 0) sSum[t >> 5] = ws;
    __syncthreads();
    const int total = sSum[0] + sSum[1] + sSum[2] + sSum[3];
    if (t < OBS_HSV_BINS) {
        // cv::normalize leaves an all-zero histogram untouched (norm < DBL_EPSILON -> scale 0)
        const float scale = total > 0 ? __double2float_rn(__ddiv_rn(1.0, (double)total)) : 0.0f;
        hist[m * OBS_HSV_BINS + t] = __fmul_rn((float)v, scale);
    }
}

}  // namespace

cudaError_t hsv_tables_upload() {
    int sdiv[256], hdiv[256];
    sdiv[0] = hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        sdiv[i] = (int)nearbyint((255 << 12) / (1. * i));     // saturate_cast<int>(double) = cvRound: half to even
        hdiv[i] = (int)nearbyint((180 << 12) / (6. * i));
    }
    cudaError_t e = cudaMemcpyToSymbol(c_sdiv, sdiv, sizeof(sdiv));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_hdiv, hdiv, sizeof(hdiv));
}

cudaError_t launch_hsv_hist(const uint8_t* bgr, size_t bgrStride, const uint8_t* masks, size_t maskStride, size_t maskImageStride,
                            int nMasks, int w, int h, int* counts, float* hist, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)nMasks * OBS_HSV_BINS * sizeof(int), st);
    if (e != cudaSuccess) return e;
    const int rowsPerCta = 8;
    k_hsv_hist<<<dim3((h + rowsPerCta - 1) / rowsPerCta, nMasks), 256, 0, st>>>(bgr, bgrStride, masks, maskStride, maskImageStride, w, h, rowsPerCta, counts);
    k_hsv_normalize<<<nMasks, 128, 0, st>>>(counts, hist);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// cv::undistortPoints(src, dst, K, distCoeffs, noArray(), K) as Frame::UndistortKeyPoints calls it: normalise, five fixed-point
// iterations of the inverse distortion (TermCriteria(COUNT, 5, 0.01)), reproject with P = K; all in binary64 with every
// operation rounded on its own (OpenCV's scalar loop; pinned against cv2 4.13 on 2 M points, tests/test_bow_matchers.py).
// ------------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256) k_undistort(const __grid_constant__ UndistortArgs A, const float* __restrict__ pts, int ptStride,
                                                   float* __restrict__ out, int outStride, int n) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double* k = A.k;
    const double u = (double)pts[(size_t)i * ptStride], v = (double)pts[(size_t)i * ptStride + 1];
    const double ifx = __ddiv_rn(1.0, A.fx), ify = __ddiv_rn(1.0, A.fy);
    double x = __dmul_rn(__dsub_rn(u, A.cx), ifx), y = __dmul_rn(__dsub_rn(v, A.cy), ify);
    const double x0 = x, y0 = y;
#define MUL __dmul_rn
#define ADD __dadd_rn
    for (int j = 0; j < 5; j++) {
        const double r2 = ADD(MUL(x, x), MUL(y, y));
        const double num = ADD(1.0, MUL(ADD(MUL(ADD(MUL(k[7], r2), k[6]), r2), k[5]), r2));
        const double den = ADD(1.0, MUL(ADD(MUL(ADD(MUL(k[4], r2), k[1]), r2), k[0]), r2));
        const double icdist = __ddiv_rn(num, den);
        if (icdist < 0) { x = x0; y = y0; break; }
        const double dX = ADD(ADD(ADD(MUL(MUL(MUL(2.0, k[2]), x), y), MUL(k[3], ADD(r2, MUL(MUL(2.0, x), x)))), MUL(k[8], r2)), MUL(MUL(k[9], r2), r2));
        const double dY = ADD(ADD(ADD(MUL(k[2], ADD(r2, MUL(MUL(2.0, y), y))), MUL(MUL(MUL(2.0, k[3]), x), y)), MUL(k[10], r2)), MUL(MUL(k[11], r2), r2));
        x = MUL(__dsub_rn(x0, dX), icdist);
        y = MUL(__dsub_rn(y0, dY), icdist);
    }
    out[(size_t)i * outStride] = __double2float_rn(ADD(MUL(A.fx, x), A.cx));
    out[(size_t)i * outStride + 1] = __double2float_rn(ADD(MUL(A.fy, y), A.cy));
#undef MUL
#undef ADD
}

}  // namespace

cudaError_t launch_undistort(const UndistortArgs& a, const float* pts, int ptStride, float* out, int outStride, int n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_undistort<<<(n + 255) / 256, 256, 0, st>>>(a, pts, ptStride, out, outStride, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// cv::distanceTransform(~mask, dst, DIST_L2, DIST_MASK_PRECISE) (Object2D constructor, src/ObjectTypes.cc:23): exact Euclidean
// distance to the nearest pixel with mask == 255, OpenCV's trueDistTrans (imgproc/distransform.cpp, un-vendored) in two stages:
// per column the vertical distance d to the nearest zero of ~mask, stored as (float)(d*d) (2^32 where the column has none within
// reach); per row the lower envelope of the parabolas f[p] + (q - p)^2 with OpenCV's binary32 intersection arithmetic, then
// sqrt.  Pinned against cv2 4.13 with IPP disabled (IPP builds differ in the last bit on some sizes).
// ------------------------------------------------------------------------------------------------
namespace {

constexpr float DT_INF = 4294967296.0f;

// stage 1: one thread per (mask, column)
__global__ void __launch_bounds__(128) k_dt_columns(const uint8_t* __restrict__ masks, size_t maskStride, size_t maskImageStride,
                                                    int w, int h, float* __restrict__ out) {
    const int x = blockIdx.x * 128 + threadIdx.x, m = blockIdx.y;
    if (x >= w) return;
    const uint8_t* src = masks + (size_t)m * maskImageStride + x;
    float* dst = out + (size_t)m * w * h + x;
    int dist = h - 1;
    for (int j = h - 1; j >= 0; j--) {                          // distance to the nearest object pixel below
        dist = src[(size_t)j * maskStride] == 255 ? 0 : dist + 1;
        dst[(size_t)j * w] = __int_as_float(dist);
    }
    dist = h - 1;
    for (int j = 0; j < h; j++) {
        const int below = __float_as_int(dst[(size_t)j * w]);
        dist = min(dist + 1, below);
        dst[(size_t)j * w] = dist < h ? (float)(dist * dist) : DT_INF;
    }
}

// stage 2: one warp per (mask, row); lane 0 builds the envelope in shared memory, all lanes evaluate it
__global__ void __launch_bounds__(128) k_dt_rows(int w, int h, int nRows, float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t dtsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 4 + warp;
    if (row >= nRows) return;
    float* f = reinterpret_cast<float*>(dtsm) + (size_t)warp * (3 * w + 2);
    float* z = f + w;                                           // w + 1 entries
    int* v = reinterpret_cast<int*>(z + w + 1);
    float* d = out + (size_t)row * w;
    for (int q = lane; q < w; q += 32) f[q] = d[q];
    __syncwarp();
    int K = 0;
    if (lane == 0) {
        int k = 0;
        v[0] = 0; z[0] = -DT_INF; z[1] = DT_INF;
        for (int q = 1; q < w; q++) {
            const float fq = f[q];
            const float a = __fadd_rn(fq, (float)(q * q));
            for (;; k--) {
                const int p = v[k];
                const float s = __fmul_rn(__fsub_rn(__fsub_rn(a, f[p]), (float)(p * p)), (float)(0.5 / (q - p)));
                if (s > z[k]) { k++; v[k] = q; z[k] = s; z[k + 1] = DT_INF; break; }
            }
        }
        K = k;
    }
    K = __shfl_sync(0xffffffffu, K, 0);
    __syncwarp();
    for (int q = lane; q < w; q += 32) {
        // smallest k with z[k + 1] >= q (the reference's running "while (z[k+1] < q) k++")
        int lo = 0, hi = K;
        const float fq = (float)q;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (z[mid + 1] < fq) lo = mid + 1; else hi = mid;
        }
        const int p = v[lo];
        const int dq = abs(q - p);
        d[q] = __fsqrt_rn(__fadd_rn((float)(dq * dq), f[p]));
    }
}

}  // namespace

cudaError_t launch_distance_transform(const uint8_t* masks, size_t maskStride, size_t maskImageStride, int nMasks, int w, int h,
                                      float* out, cudaStream_t st) {
    k_dt_columns<<<dim3((w + 127) / 128, nMasks), 128, 0, st>>>(masks, maskStride, maskImageStride, w, h, out);
    const size_t smem = (size_t)4 * (3 * w + 2) * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = OBS_ALLOW_MAX_SMEM(k_dt_rows);
        if (e != cudaSuccess) return e;
    }
    const int nRows = nMasks * h;
    k_dt_rows<<<(nRows + 3) / 4, 128, smem, st>>>(w, h, nRows, out);
    return cudaGetLastError();
}
