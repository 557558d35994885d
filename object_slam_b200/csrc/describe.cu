// K4 + K6 + K7 -- orientation, rBRIEF descriptor and keypoint finalisation, one warp per keypoint; the
// keypoint's window of the blurred level arrives in shared memory by one TMA box load per keypoint.
// Replaces IC_Angle / computeOrientation (src/ORBextractor.cc:77-104, :472-479),
// computeOrbDescriptor / computeDescriptors (:108-147, :1034-1041) and the keypoint bookkeeping of
// ComputeKeyPointsOctTree (:837-847) and operator() (:1094-1103).
//
// Float semantics are part of the contract (descriptor bits depend on them): cv::fastAtan2's
// polynomial and the pattern rotation are evaluated with individually rounded binary32 operations
// (no FMA contraction), cvRound is round-half-to-even, and cosf/sinf are glibc's sincosf algorithm
// evaluated in binary64 -- CUDA's cosf/sinf round differently and are not used.
#include "kernels.h"
#include <float.h>

namespace {

__device__ const int8_t d_pattern[256 * 4] = {
#include "orb_pattern.inc"
};

#ifndef DESC_NW
#define DESC_NW 4
#endif
constexpr int DESC_WARPS = DESC_NW;
#ifndef DESC_MINCTAS
#define DESC_MINCTAS 12          // 40 registers: the kernel lives on resident warps (its loads are gathers), not on instruction count
#endif

// cv::fastAtan2 (OpenCV core, atanImpl scalar path): degrees in [0, 360).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)DBL_EPSILON));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc >= 2.28 sincosf (ARM optimized-routines), |y| < 120, every binary64 op rounded on its own.
__device__ __forceinline__ void sincosf_glibc(float y, float* sinp, float* cosp) {
    const double c0 = 1.0, c1 = -0x1.ffffffd0c621cp-2, c2c = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10,
                 c4 = 0x1.99343027bf8c3p-16, s1c = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    double x = (double)y;
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;
    const uint32_t top_pio4 = (__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ff;
    const uint32_t top_tiny = (__float_as_uint(0x1p-12f) >> 20) & 0x7ff;
    int n = 0;
    double x2;
    if (top < top_pio4) {
        if (top < top_tiny) { *sinp = y; *cosp = 1.0f; return; }
        x2 = __dmul_rn(x, x);
    } else {
        const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
        n = (__double2int_rz(r) + 0x800000) >> 24;
        x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
        x2 = __dmul_rn(x, x);                        // squared before the sign is applied
        if ((n & 3) == 1 || (n & 3) == 2) x = -x;    // sign table {1,-1,-1,1}
    }
    const double sg = (n & 2) ? -1.0 : 1.0;          // second table = first with the cosine coefficients negated
    const double x4 = __dmul_rn(x2, x2), x3 = __dmul_rn(x2, x);
    const double pc2 = __dadd_rn(sg * c3, __dmul_rn(x2, sg * c4));
    const double ps1 = __dadd_rn(s2, __dmul_rn(x2, s3));
    const double pc1 = __dadd_rn(sg * c0, __dmul_rn(x2, sg * c1));
    const double x5 = __dmul_rn(x3, x2), x6 = __dmul_rn(x4, x2);
    const double Sv = __dadd_rn(x, __dmul_rn(x3, s1c));
    const double Cv = __dadd_rn(pc1, __dmul_rn(x4, sg * c2c));
    const float sv = (float)__dadd_rn(Sv, __dmul_rn(x5, ps1));
    const float cv = (float)__dadd_rn(Cv, __dmul_rn(x6, pc2));
    if (n & 1) { *sinp = cv; *cosp = sv; } else { *sinp = sv; *cosp = cv; }
}

#ifndef DESC_NPW
#define DESC_NPW 8
#endif
constexpr int DESC_PER_WARP = DESC_NPW;     // keypoints per warp: tables and the level offsets are set up once per 64 keypoints
constexpr int PR = 18, PP = DESC_BOX_W;      // rBRIEF window radius (pattern radius <= 18.39), row pitch of the TMA box
constexpr int PATCH_BYTES = (DESC_BOX_W * DESC_BOX_H + 127) / 128 * 128;
static_assert(DESC_BOX_H == 2 * PR + 1 && DESC_BOX_W >= 2 * PR + 1 + 15 && DESC_BOX_W % 16 == 0, "TMA box");

__global__ void __launch_bounds__(DESC_WARPS * 32, DESC_MINCTAS) k_describe(const __grid_constant__ Geom g, const PyrPtrs p,
                                                              const CUtensorMap* __restrict__ maps, int img0,
                                                              const uint32_t* __restrict__ sel, const int* __restrict__ selCount,
                                                              uint8_t* __restrict__ records, size_t recordBytes) {
    // (x0, y0 | x1, y1) of point pair 8*lane + k at [k * 32 + lane] as four bfloat16 (|coordinate| <= 13 is exact; a bfloat16 is the upper
    // half of the binary32): one LDS.64 instead of an LDS.128 per pair -- the kernel is bound by shared-memory wavefronts
    __shared__ uint2 pat[256];
    // rowMask[|v|][i]: 0xff in the bytes of patch word i (columns u = -15 + 4 i + k) that lie inside the circular patch
    // (row OBS_HALF_PATCH + 1 is empty: the 32nd row slot of the warp's sweep)
    __shared__ uint32_t rowMask[OBS_HALF_PATCH + 2][8];
    __shared__ int levelEnd[OBS_MAX_LEVELS + 1];  // keypoints up to and including level l (level-major output order)
    __shared__ __align__(128) uint8_t patch[DESC_WARPS][PATCH_BYTES];     // one TMA box per warp
    __shared__ __align__(8) uint64_t mbar[DESC_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.y;
    uint8_t* rec = records + (size_t)img * recordBytes;
    for (int i = tid; i < 256; i += DESC_WARPS * 32) {
        const int pr = (i & 31) * 8 + (i >> 5);
        const unsigned f0 = __float_as_uint((float)d_pattern[4 * pr]) >> 16, f1 = __float_as_uint((float)d_pattern[4 * pr + 1]) & 0xffff0000u;
        const unsigned f2 = __float_as_uint((float)d_pattern[4 * pr + 2]) >> 16, f3 = __float_as_uint((float)d_pattern[4 * pr + 3]) & 0xffff0000u;
        pat[i] = make_uint2(f0 | f1, f2 | f3);
    }
    for (int t = tid; t < (OBS_HALF_PATCH + 2) * 8; t += DESC_WARPS * 32) {
        const int av = t >> 3, wi = t & 7;
        uint32_t m = 0;
        if (av <= OBS_HALF_PATCH)
            for (int k = 0; k < 4; k++) if (abs(-OBS_HALF_PATCH + 4 * wi + k) <= g.umax[av]) m |= 0xffu << (8 * k);
        rowMask[av][wi] = m;
    }
    const uint32_t mbar_s = smem_u32(&mbar[warp]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar_s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned phase = 0;
    pdl_entry();              // the pattern / mask tables above come from constant memory and the parameters: built while the quadtree finishes
    if (warp == DESC_WARPS - 1) {
        // lanes 0..nlevels-1 hold the per-level counts, a warp scan gives the offsets
        const int myCnt = lane < g.nlevels ? selCount[(size_t)img * g.nlevels + lane] : 0;
        int incl = myCnt;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane <= OBS_MAX_LEVELS) levelEnd[lane] = incl;
        if (blockIdx.x == 0 && lane < g.nlevels) {      // header: n, per-level counts
            reinterpret_cast<int32_t*>(rec)[1 + lane] = myCnt;
            if (lane == g.nlevels - 1) reinterpret_cast<int32_t*>(rec)[0] = incl;
        }
    }
    __syncthreads();
    const int total = levelEnd[g.nlevels - 1];
    uint8_t* kpOut = rec + OBS_HDR_INTS * 4;
    uint8_t* descOut = kpOut + (size_t)g.kpCap * 28;
    uint8_t* mp = patch[warp];
    unsigned rm[8];                               // the lane's eight row masks of the moment sweep, fixed for all keypoints (pinned in registers)
#pragma unroll
    for (int t = 0; t < 8; t++) rm[t] = pin(rowMask[abs(4 * t + (lane >> 3) - OBS_HALF_PATCH)][lane & 7]);

    for (int it = 0; it < DESC_PER_WARP; it++) {
        const int j = (blockIdx.x * DESC_PER_WARP + it) * DESC_WARPS + warp;     // output index of the keypoint (level-major)
        if (j >= total || j >= g.kpCap) break;            // (the capacity covers the extractor's worst case; never write past it)
        const int level = __popc(__ballot_sync(0xffffffffu, lane < g.nlevels - 1 && j >= levelEnd[min(lane, OBS_MAX_LEVELS)]));
        const int local = j - (level ? levelEnd[level - 1] : 0);
        const LevelGeom& lg = g.lv[level];
        const uint32_t key = sel[((size_t)img * g.nlevels + level) * g.selCap + local];
        const int cx = key_x(key) + OBS_BORDER, cy = key_y(key) + OBS_BORDER;       // :837-838

        // ---- the 37 x 37 window of the blurred level around the keypoint (the rotated pattern stays within 18 px), fetched by one
        // TMA box load (64 x 37 bytes from the 16-byte aligned column at or before cx - 18, row cy - 18, image img0 + img): no
        // per-lane addresses, no LSU traffic, and the copy runs while the warp computes the moments; the warp waits on its mbarrier
        // before the gathers.  The previous keypoint's gathers are behind the __syncwarp at the loop's end.
        const int bmis = (cx - PR) & 15;
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar_s), "r"(DESC_BOX_W * DESC_BOX_H) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         :: "r"(smem_u32(mp)), "l"(reinterpret_cast<uint64_t>(maps + level)), "r"((cx - PR) & ~15), "r"(cy - PR), "r"(img0 + img), "r"(mbar_s)
                         : "memory");
        }

        // ---- IC_Angle on the unblurred level (:77-104): integer moments over the circular patch.
        // A quarter warp owns a patch row per step (4 rows per step, 8 steps); a lane fetches two neighbouring aligned
        // words of its row, shifts its four patch bytes u = -15 + 4 wi .. +3 into place, masks them to the row's extent
        // |u| <= umax[|v|] and sums I and (u + 15) I by integer dot products.  Rows are word aligned (pitch % 4 == 0),
        // so the shift is the same for all lanes; a warp-wide load touches four 36-byte row segments.
        int pitch;
        const uint8_t* im = level_ptr(p, g, img, level, pitch);
        int m10 = 0, m01 = 0;
        {
            const int wi = lane & 7, r4 = lane >> 3;
            const uint8_t* rowp = im + (size_t)(cy - OBS_HALF_PATCH + r4) * pitch + (cx - OBS_HALF_PATCH);
            const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(rowp) & 3);
            const uint32_t* q = reinterpret_cast<const uint32_t*>(rowp - mis) + wi;
            const unsigned uw = 0x03020100u + 0x04040404u * wi;
            const uint64_t step = 4ull * (unsigned)pitch;
            int s0 = 0, s1 = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                const int v = 4 * t + r4 - OBS_HALF_PATCH;
                if (t) asm("add.u64 %0, %0, %1;" : "+l"(q) : "l"(step));
                const unsigned W = __funnelshift_r(__ldg(q), __ldg(q + 1), 8 * mis) & rm[t];
                const int rs = (int)__dp4a(W, 0x01010101u, 0u);
                s0 += rs;
                s1 = (int)__dp4a(W, uw, (unsigned)s1);
                m01 += v * rs;
            }
            m10 = s1 - OBS_HALF_PATCH * s0;
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- rBRIEF on the blurred level (:108-147): lane i produces descriptor byte i
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        float a, b;
        sincosf_glibc(__fmul_rn(angle, factorPI), &b, &a);
        {
            unsigned ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(mbar_s), "r"(phase) : "memory");
            } while (!ok);
            phase ^= 1u;
        }
        const uint8_t* cb = mp + PR * PP + PR + bmis;
        unsigned val = 0;
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            const uint2 qw = pat[k * 32 + lane];
            const float4 q = make_float4(__uint_as_float(qw.x << 16), __uint_as_float(qw.x & 0xffff0000u),
                                         __uint_as_float(qw.y << 16), __uint_as_float(qw.y & 0xffff0000u));
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(q.x, b), __fmul_rn(q.y, a)));
            const int q0 = __float2int_rn(__fsub_rn(__fmul_rn(q.x, a), __fmul_rn(q.y, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(q.z, b), __fmul_rn(q.w, a)));
            const int q1 = __float2int_rn(__fsub_rn(__fmul_rn(q.z, a), __fmul_rn(q.w, b)));
            const int d = (int)cb[r0 * PP + q0] - (int)cb[r1 * PP + q1];        // t0 < t1  <=>  d < 0
            val = __funnelshift_l((unsigned)d, val, 1);                           // shift the sign bit in; bit k last for k = 0
        }
        descOut[(size_t)j * 32 + lane] = (uint8_t)val;
        __syncwarp();                                    // the window is restaged by the next keypoint

        // ---- keypoint record (cv::KeyPoint layout); pt scaled to level-0 coordinates (:1094-1101)
        if (lane < 7) {
            float fx = (float)cx, fy = (float)cy;
            if (level != 0) { fx = __fmul_rn(fx, lg.scale); fy = __fmul_rn(fy, lg.scale); }
            uint32_t w = 0xffffffffu;                    // class_id = -1
            w = lane == 5 ? (uint32_t)level : w;
            w = lane == 4 ? __float_as_uint((float)key_r(key)) : w;
            w = lane == 3 ? __float_as_uint(angle) : w;
            w = lane == 2 ? __float_as_uint(lg.patchSize) : w;
            w = lane == 1 ? __float_as_uint(fy) : w;
            w = lane == 0 ? __float_as_uint(fx) : w;
            reinterpret_cast<uint32_t*>(kpOut + (size_t)j * 28)[lane] = w;
        }
    }
}

}  // namespace

cudaError_t launch_describe(const Geom& g, PyrPtrs p, const CUtensorMap* maps, int img0,
                            const uint32_t* sel, const int* selCount, uint8_t* records, size_t recordBytes,
                            int nimg, cudaStream_t st) {
    dim3 grid((g.kpCap + DESC_WARPS * DESC_PER_WARP - 1) / (DESC_WARPS * DESC_PER_WARP), nimg);
    cudaError_t le = launch_k(pdl_enabled(), k_describe, grid, dim3(DESC_WARPS * 32), 0, st, g, p, maps, img0, sel, selCount, records, recordBytes);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}
