// K11 on the 5th-generation tensor cores: brute-force keyframe-vs-keyframe Hamming search
// (the candidate loop of ORBmatcher::SearchByBoW, /root/reference/src/ORBmatcher.cc:205-226, on
// DescriptorDistance, :1647-1663) as an exact integer GEMM.
//
// With every descriptor bit b stored as the int8 value 8 * (1 - 2b), the dot product of two 256-element rows is
//     a . b = 64 * ((#equal bits) - (#different bits)) = 64 * (256 - 2 * hamming(a, b)),
// exact in the int32 accumulators of tcgen05.mma kind::i8 (|a.b| <= 16384).
//
//   k_knn2_expand   bits -> +-8 int8 rows of 256 bytes (one pass over the descriptor sets per call), so a.b = 64 * (256 - 2 hamming)
//   k_knn2_tc_pair / k_knn2_tc   persistent, warp specialised (CTA pairs with tcgen05.mma.cta_group::2, or one CTA per SM):
//       warps 0-15  epilogue.  Every accumulator column is PRE-LOADED (tcgen05.st ... unpack::16b) with 16384 + 63 - (column in the
//                   warp's 64-column part); the MMAs accumulate onto it, so a finished column holds the 16-bit key
//                   V = 128 * (256 - dist) + 63 - column, and tcgen05.ld ... pack::16b returns two keys per register with no
//                   arithmetic.  The running (best, second best) of every 16-bit lane is three packed VIMNMX.U16x2 per register --
//                   1.5 ALU instructions per distance and nothing else.  Keys are distinct, so the two largest keys are the
//                   reference's best / second best under its strict-< first-wins update; parts, tiles and threads merge exactly.
//       warp 16     TMA producer: query tile (128 rows) and database tiles as 128-byte swizzle atoms
//                   (cp.async.bulk.tensor.3d on a [keyframe][descriptor][256 B] tensor map, SWIZZLE_128B; rows past the end
//                   of a keyframe are zero-filled by the TMA unit)
//       warp 17     TMEM allocation; one lane issues 8 x tcgen05.mma.kind::i8 (K 32 each) per database tile into one of two
//                   256-column accumulators and commits to the mbarriers of the pipeline
// The POPC kernel (matcher.cu k_knn2) remains the path for tiny descriptor sets; obs_hamming_knn2 picks per call.
#include "matcher.h"
#include "knn2_tc.h"

#include <cuda.h>
#include <algorithm>
#include <atomic>

namespace {

std::atomic<int> g_knn2Pair{1};       // obs_set_option("knn2_cta_pair"): k_knn2_tc_pair (default) or k_knn2_tc

constexpr int TC_M = 128;                 // query rows of a work item (= TMEM lanes)
constexpr int TC_N = 256;                 // database rows of one accumulator
constexpr int TC_KB = 256;                // bytes of an expanded descriptor = K of the contraction
constexpr int ATOM_B = 128;               // swizzle atom width in bytes
constexpr int A_ATOM_BYTES = TC_M * ATOM_B;       // 16 KB
constexpr int B_ATOM_BYTES = TC_N * ATOM_B;       // 32 KB
constexpr int A_BYTES = 2 * A_ATOM_BYTES;         // 32 KB
constexpr int B_BYTES = 2 * B_ATOM_BYTES;         // 64 KB
constexpr int EPI2_PARTS = 4;                     // epilogue warps per TMEM lane quarter: each takes 256 / 4 accumulator columns
constexpr int EPI2_COLS = TC_N / EPI2_PARTS;
constexpr int EPI2_THREADS = 128 * EPI2_PARTS, TC2_THREADS = 64 + EPI2_THREADS;
constexpr uint32_t SENT32 = (256u << 16) | 0xffffu;
constexpr size_t TC_SMEM = 2 * A_BYTES + 2 * B_BYTES + 1024;

// ---- PTX wrappers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// (suspend-time hint: a waiting thread sleeps in the barrier unit until the phase completes instead of re-issuing try_wait every few
// cycles -- the producer and MMA lanes wait most of the time and would otherwise take ~15 % of the issue slots of the epilogue warps
// they share a scheduler with)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(200000u) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
// K-major operand tile of 128-byte rows under SWIZZLE_128B: groups of 8 rows 1024 bytes apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p; }"
                 :: "r"(tmemD), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&d)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                   "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]),
                   "=r"(d[16]), "=r"(d[17]), "=r"(d[18]), "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]),
                   "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]), "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// instruction descriptor of kind::i8: D = S32 (bits 4-5 = 2), A and B signed 8 bit (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC_I8 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// ---- bits -> +-1 int8 -------------------------------------------------------------------------------------------------
// thread = 16 descriptor bits -> 16 bytes; bit k of the descriptor (bit k%8 of byte k/8) becomes element k: 0 -> +8, 1 -> -8.
// (+-8 instead of +-1: a.b = 64 * (256 - 2 * hamming), so the six low bits of an accumulator are free for a column index -- the
// CTA-pair kernel lets the tensor core add the MMA result onto pre-stored (bias + column) constants and reads finished keys.)
__device__ __forceinline__ uint32_t expand4(uint32_t b) {
    const uint32_t x = (b * 0x00204081u) & 0x01010101u;       // bit i of b -> byte i
    return x * 0xf0u + 0x08080808u;                             // 1 -> 0xf8 (-8), 0 -> 0x08 (+8)
}
// only the keyframes a pair of this call names are expanded: k_knn2_mark flags them, one CTA per flagged keyframe expands it
__global__ void __launch_bounds__(256) k_knn2_mark(const int2* __restrict__ pairs, int nPairs, int nKeyframes, uint8_t* __restrict__ used) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nPairs) return;
    const int2 p = pairs[i];
    if ((unsigned)p.x < (unsigned)nKeyframes) used[p.x] = 1;
    if ((unsigned)p.y < (unsigned)nKeyframes) used[p.y] = 1;
}
__global__ void __launch_bounds__(256) k_knn2_expand(const uint16_t* __restrict__ bits, uint4* __restrict__ out, int chunksPerKf,
                                                     const uint8_t* __restrict__ used) {
    if (!used[blockIdx.x]) return;
    const size_t base = (size_t)blockIdx.x * chunksPerKf;
    for (int i = threadIdx.x; i < chunksPerKf; i += 256) {
        const uint32_t v = bits[base + i];
        out[base + i] = make_uint4(expand4(v & 15u), expand4((v >> 4) & 15u), expand4((v >> 8) & 15u), expand4(v >> 12));
    }
}

// ---- epilogue helpers ---------------------------------------------------------------------------------------------------
// accumulator key V = 128 * (256 - dist) + 63 - column-in-part, 0 = nothing (masked column / no candidate: distance 256)
__device__ __forceinline__ uint32_t vkey_to_32(uint32_t v, uint32_t colBase) {
    return v == 0u ? SENT32 : (((256u - (v >> 7)) << 16) | (colBase + 63u - (v & 127u)));
}
__device__ __forceinline__ void merge2(uint32_t& best, uint32_t& second, uint32_t b, uint32_t s) {
    const uint32_t hi = max(best, b);
    best = min(best, b);
    second = min(hi, min(second, s));
}

struct Knn2TcArgs {
    const int2* pairs;
    int nPairs, n, mTiles, nTiles;
    int thLow; float nnratio;
    int* bestIdx; int* bestDist; int* secondDist;
    unsigned long long* workCounter;      // zeroed before the launch: next (pair, query tile) work item
};

// ---- CTA-pair variant (cta_group::2) -------------------------------------------------------------------------------------
// k_knn2_tc pulls every 64 KB database tile through L2 for ONE 128-row query tile: ~53 B per cycle per SM, more than L2 delivers
// to an SM (~42 B/clk), so the tensor pipe waits.  Here a cluster of two CTAs (the two SMs of a TPC) works on TWO query tiles
// against the same database tile: one tcgen05.mma.cta_group::2 (M 256, N 256, K 32) per K step, issued by the leader CTA, reads
// each CTA's own 128 query rows and each CTA's HALF of the database tile (128 rows, 32 KB) from both shared memories and leaves
// every CTA its own 128 x 256 accumulator in its own TMEM.  Per SM the L2 traffic halves; the smaller stage buys a 4-deep ring.
//   * work items = (pair, pair of query tiles) from the global counter, fetched by the leader's producer and published to both
//     CTAs' queues (st.shared::cluster + cluster-scope release / acquire on the queue barriers);
//   * both producers issue their own TMA loads with .cta_group::2 onto the LEADER's full barriers (expect_tx there counts both);
//   * tcgen05.commit ... multicast::cluster frees the stages in / signals the accumulators of both CTAs;
//   * both CTAs' epilogue warps hand accumulators and queue slots back by (remote) arrives on the leader's barriers.
constexpr int B2_ATOM_BYTES = 128 * ATOM_B;        // 128 database rows x 128 bytes
constexpr int B2_BYTES = 2 * B2_ATOM_BYTES;        // 32 KB: this CTA's half of a database tile
constexpr int B2_STAGES = 4;
constexpr size_t TC2_SMEM = 2 * A_BYTES + B2_STAGES * B2_BYTES + 1024;
constexpr uint32_t IDESC_I8_PAIR = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA).  _release: cluster-scope release, for the one barrier
// that publishes data written with ordinary stores (the work queue); the others only hand back resources whose reads are complete
// (TMEM accumulators behind tcgen05.fence, queue slots already read), and a cluster-scope fence per arrive would sit on the
// accumulator hand-off path.
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cbar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cbar) : "memory");
}
// wait with acquire at cluster scope (the data behind the barrier may have been written by the peer CTA); bounded: a protocol error
// must not hang the device
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done;
    const long long t0 = clock64();
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(200000u) : "memory");
        if (!done && clock64() - t0 > 20000000000ll) __trap();
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t leaderBar) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(leaderBar) : "memory");
}
__device__ __forceinline__ void umma_i8_pair(uint32_t tmemD, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p; }"
                 :: "r"(tmemD), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {       // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

// ---- epilogue of both tensor-core kernels (16 warps: warp w reads the TMEM lanes 32 * (w % 4) and the 64-column part w / 4 of every
// accumulator).  qEmptyL0 / accEmptyL0: shared::cluster addresses of the barriers the MMA lane waits on (the leader CTA's in a pair).
template <bool PAIR>
__device__ __forceinline__ void knn2_epilogue(const Knn2TcArgs& A, uint32_t tmemBase, int warp, int lane, uint32_t rank, int itemsPerPair,
                                              uint32_t qFull0, uint32_t qEmptyL0, uint32_t accFull0, uint32_t accEmptyL0,
                                              const long long* workQ, uint2 (*sMerge)[EPI2_PARTS - 1][TC_M]) {
    const int ew = warp;
    const int quad = warp & 3;                  // a warp reads the TMEM lanes 32 * (warp id % 4) ...
    const int part = ew >> 2;                   // ... and this 64-column quarter of every accumulator
    const int row = quad * 32 + lane;
    // The accumulator columns of this thread's part are pre-loaded with  16384 + 63 - (column in part)  (tcgen05.st ...
    // unpack::16b: one register per two columns, zero extended); the MMAs ADD 64 a.b = 16384 - 128 dist, so a finished column holds
    //     V = 128 * (256 - dist) + 63 - column      (0 <= V < 32832: larger = closer, then the lower column)
    // and tcgen05.ld ... pack::16b returns two such keys per register with no arithmetic at all.  Best / second best are the two
    // LARGEST keys of every 16-bit lane: three packed VIMNMX per register.
    uint32_t cst[32];
#pragma unroll
    for (int j = 0; j < 32; j++) cst[j] = (uint32_t)(16384 + 63 - 2 * j) | ((uint32_t)(16384 + 63 - (2 * j + 1)) << 16);
    const uint32_t tpart = tmemBase + ((uint32_t)(quad * 32) << 16) + part * EPI2_COLS;
    auto store_constants = [&](uint32_t taddr) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.unpack::16b.b32 [%0], "
                     "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                     "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                     :: "r"(taddr), "r"(cst[0]), "r"(cst[1]), "r"(cst[2]), "r"(cst[3]), "r"(cst[4]), "r"(cst[5]), "r"(cst[6]), "r"(cst[7]),
                        "r"(cst[8]), "r"(cst[9]), "r"(cst[10]), "r"(cst[11]), "r"(cst[12]), "r"(cst[13]), "r"(cst[14]), "r"(cst[15]),
                        "r"(cst[16]), "r"(cst[17]), "r"(cst[18]), "r"(cst[19]), "r"(cst[20]), "r"(cst[21]), "r"(cst[22]), "r"(cst[23]),
                        "r"(cst[24]), "r"(cst[25]), "r"(cst[26]), "r"(cst[27]), "r"(cst[28]), "r"(cst[29]), "r"(cst[30]), "r"(cst[31])
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    };
    // both accumulators start out holding the constants; the first arrive on accEmpty is that release
    for (uint32_t acc = 0; acc < 2; acc++) {
        store_constants(tpart + acc * TC_N);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(accEmptyL0 + 8u * acc);
    }
    uint32_t bi = 0;
    for (uint32_t ai = 0;; ai++) {
        const uint32_t sa = ai & 1u;
        if (PAIR) mbar_wait_cluster(qFull0 + 8u * sa, (ai >> 1) & 1u); else mbar_wait(qFull0 + 8u * sa, (ai >> 1) & 1u);
        const long long w = *reinterpret_cast<const volatile long long*>(&workQ[sa]);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(qEmptyL0 + 8u * sa);
        if (w < 0) break;
        const int pair = (int)(w / itemsPerPair), it = (int)(w - (long long)pair * itemsPerPair), mt = PAIR ? 2 * it + (int)rank : it;
        uint32_t best = SENT32, second = SENT32;
        for (int nt = 0; nt < A.nTiles; nt++, bi++) {
            const uint32_t acc = bi & 1u;
            mbar_wait(accFull0 + 8u * acc, (bi >> 1) & 1u);
            tc_fence_after();
            const int colBase = nt * TC_N + part * EPI2_COLS;
            const int valid = A.n - colBase;                       // columns of this part that exist
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(tpart + acc * TC_N) : "memory");
            tmem_ld_wait();
            // the keys are in registers: put the constants back and hand the accumulator to the leader's MMA lane
            store_constants(tpart + acc * TC_N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(accEmptyL0 + 8u * acc);
            if (valid <= 0) continue;
            if (valid < EPI2_COLS) {                               // ragged end of the keyframe: columns past it (zero rows) drop to key 0
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (2 * j >= valid) v[j] &= 0xffff0000u;
                    if (2 * j + 1 >= valid) v[j] &= 0x0000ffffu;
                }
            }
            uint32_t bq[4] = {0u, 0u, 0u, 0u}, sq[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                uint32_t lo[4];
#pragma unroll
                for (int u = 0; u < 4; u++) lo[u] = __vminu2(v[j + u], bq[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) bq[u] = __vmaxu2(v[j + u], bq[u]);
#pragma unroll
                for (int u = 0; u < 4; u++) sq[u] = __vmaxu2(sq[u], lo[u]);
            }
            // four chains -> one, 16-bit lanes -> one, part keys -> global (dist << 16 | column) keys
            const uint32_t b01 = __vmaxu2(bq[0], bq[1]), s01 = __vmaxu2(__vminu2(bq[0], bq[1]), __vmaxu2(sq[0], sq[1]));
            const uint32_t b23 = __vmaxu2(bq[2], bq[3]), s23 = __vmaxu2(__vminu2(bq[2], bq[3]), __vmaxu2(sq[2], sq[3]));
            const uint32_t bb = __vmaxu2(b01, b23), ss = __vmaxu2(__vminu2(b01, b23), __vmaxu2(s01, s23));
            const uint32_t bl = bb & 0xffffu, bh = bb >> 16, sl = ss & 0xffffu, sh = ss >> 16;
            const uint32_t kb = max(bl, bh), ks = max(min(bl, bh), max(sl, sh));
            merge2(best, second, vkey_to_32(kb, (uint32_t)colBase), vkey_to_32(ks, (uint32_t)colBase));
        }
        if (part) sMerge[ai & 1u][part - 1][row] = make_uint2(best, second);
        asm volatile("bar.sync 1, %0;" :: "n"(EPI2_THREADS) : "memory");
        if (part == 0) {
#pragma unroll
            for (int q = 0; q < EPI2_PARTS - 1; q++) {
                const uint2 o = sMerge[ai & 1u][q][row];
                merge2(best, second, o.x, o.y);
            }
            const int qi = mt * TC_M + row;
            if (qi < A.n) {
                const int bd = (int)(best >> 16), sd = (int)(second >> 16);
                const size_t o2 = (size_t)pair * A.n + qi;
                int idx = -1;
                if (bd <= A.thLow && (float)bd < __fmul_rn(A.nnratio, (float)sd)) idx = (int)(best & 0xffffu);
                A.bestIdx[o2] = idx;
                if (A.bestDist) A.bestDist[o2] = bd;
                if (A.secondDist) A.secondDist[o2] = sd;
            }
        }
    }
}

__global__ void __launch_bounds__(TC2_THREADS, 1)
k_knn2_tc_pair(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ Knn2TcArgs A) {
    extern __shared__ uint8_t smemRaw[];
    // aFull[2] aEmpty[2] bFull[4] bEmpty[4] accFull[2] accEmpty[2] qFull[2] qEmpty[2]
    __shared__ __align__(8) uint64_t bars[20];
    __shared__ __align__(8) long long workQ[2];
    __shared__ uint32_t tmemBaseS;
    __shared__ uint2 sMerge[2][EPI2_PARTS - 1][TC_M];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the epilogue warps come first: the scheduler favours the higher warp id, and the two single-lane warps that feed the tensor
    // pipe (TMA producer, MMA issuer) must not wait behind the epilogue warps they share a scheduler with
    constexpr int W_PRODUCER = EPI2_THREADS / 32, W_MMA = W_PRODUCER + 1;
    const uint32_t rank = cluster_rank();
    const bool leader = rank == 0;
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    const uint32_t sA0 = base, sB0 = base + 2 * A_BYTES;
    const uint32_t bar0 = smem_u32(bars);
    auto aFull = [&](int s) { return bar0 + 8u * (0 + s); };
    auto aEmpty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto bFull = [&](int s) { return bar0 + 8u * (4 + s); };
    auto bEmpty = [&](int s) { return bar0 + 8u * (8 + s); };
    auto accFull = [&](int s) { return bar0 + 8u * (12 + s); };
    auto accEmpty = [&](int s) { return bar0 + 8u * (14 + s); };
    auto qFull = [&](int s) { return bar0 + 8u * (16 + s); };
    auto qEmpty = [&](int s) { return bar0 + 8u * (18 + s); };
    constexpr uint32_t EPI_WARPS = EPI2_THREADS / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(aFull(s), 1); mbar_init(aEmpty(s), 1);
            mbar_init(accFull(s), 1); mbar_init(accEmpty(s), 2 * EPI_WARPS);           // the epilogue warps of both CTAs
            mbar_init(qFull(s), 1); mbar_init(qEmpty(s), 2 * EPI_WARPS + 2);           // ... + the leader's MMA lane + the peer's producer
        }
        for (int s = 0; s < B2_STAGES; s++) { mbar_init(bFull(s), 1); mbar_init(bEmpty(s), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBaseS)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmemBase = tmemBaseS;
    const int mPairs = (A.mTiles + 1) >> 1;
    const long long total = (long long)A.nPairs * mPairs;

    if (warp == W_PRODUCER) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&mapT) : "memory");
            uint32_t bi = 0, ai = 0;
            for (;; ai++) {
                const uint32_t sa = ai & 1u, pa = (ai >> 1) & 1u;
                long long w;
                if (leader) {
                    mbar_wait(qEmpty(sa), pa ^ 1u);                         // both CTAs have read the slot's previous item
                    w = (long long)atomicAdd(A.workCounter, 1ull);
                    if (w >= total) w = -1;
                    workQ[sa] = w;
                    asm volatile("st.shared::cluster.b64 [%0], %1;" :: "r"(map_to_cta(smem_u32(&workQ[sa]), 1)), "l"(w) : "memory");
                    mbar_arrive_cluster_release(map_to_cta(qFull(sa), 0));
                    mbar_arrive_cluster_release(map_to_cta(qFull(sa), 1));
                } else {
                    mbar_wait_cluster(qFull(sa), pa);
                    w = *reinterpret_cast<volatile long long*>(&workQ[sa]);
                    mbar_arrive_cluster(map_to_cta(qEmpty(sa), 0));
                }
                if (w < 0) break;
                const int pair = (int)(w / mPairs), mt = 2 * (int)(w - (long long)pair * mPairs) + (int)rank;
                const int2 pr = A.pairs[pair];
                const uint32_t aBar = map_to_cta(aFull(sa), 0);
                mbar_wait(aEmpty(sa), pa ^ 1u);
                if (leader) mbar_expect_tx(aFull(sa), 2 * A_BYTES);        // this CTA's query tile and the peer's
                tma_load_3d_pair(sA0 + sa * A_BYTES, &mapT, 0, mt * TC_M, pr.x, aBar);
                tma_load_3d_pair(sA0 + sa * A_BYTES + A_ATOM_BYTES, &mapT, ATOM_B, mt * TC_M, pr.x, aBar);
                for (int nt = 0; nt < A.nTiles; nt++, bi++) {
                    const uint32_t sb = bi % B2_STAGES, pb = (bi / B2_STAGES) & 1u;
                    const uint32_t bBar = map_to_cta(bFull(sb), 0);
                    mbar_wait(bEmpty(sb), pb ^ 1u);
                    if (leader) mbar_expect_tx(bFull(sb), 2 * B2_BYTES);   // both halves of the database tile
                    const int y = nt * TC_N + (int)rank * 128;
                    tma_load_3d_pair(sB0 + sb * B2_BYTES, &mapT, 0, y, pr.y, bBar);
                    tma_load_3d_pair(sB0 + sb * B2_BYTES + B2_ATOM_BYTES, &mapT, ATOM_B, y, pr.y, bBar);
                }
            }
            // every multicast commit aimed at this CTA's stage barriers has landed before the CTA may leave
            for (uint32_t s = 0; s < (uint32_t)B2_STAGES; s++) {
                const uint32_t uses = bi > s ? (bi - s + B2_STAGES - 1) / B2_STAGES : 0;
                if (uses) mbar_wait(bEmpty(s), (uses - 1) & 1u);
            }
            for (uint32_t s = 0; s < 2; s++) {
                const uint32_t uses = ai > s ? (ai - s + 1) / 2 : 0;      // items that used stage s (the terminating slot loaded nothing)
                if (uses) mbar_wait(aEmpty(s), (uses - 1) & 1u);
            }
        }
    } else if (warp == W_MMA) {
        if (lane == 0 && leader) {
            uint32_t bi = 0;
            for (uint32_t ai = 0;; ai++) {
                const uint32_t sa = ai & 1u, pa = (ai >> 1) & 1u;
                mbar_wait_cluster(qFull(sa), pa);
                const long long w = *reinterpret_cast<volatile long long*>(&workQ[sa]);
                mbar_arrive_cluster(map_to_cta(qEmpty(sa), 0));
                if (w < 0) break;
                mbar_wait(aFull(sa), pa);
                for (int nt = 0; nt < A.nTiles; nt++, bi++) {
                    const uint32_t sb = bi % B2_STAGES, pb = (bi / B2_STAGES) & 1u;
                    const uint32_t acc = bi & 1u, pacc = (bi >> 1) & 1u;
                    mbar_wait(accEmpty(acc), pacc);          // phase u = the accumulator holds its constants for use u (phase 0: the initial store)
                    mbar_wait(bFull(sb), pb);
                    tc_fence_after();
                    const uint32_t d = tmemBase + acc * TC_N;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint32_t ka = (uint32_t)(k >> 2) * A_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                        const uint32_t kb = (uint32_t)(k >> 2) * B2_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                        umma_i8_pair(d, umma_desc(sA0 + sa * A_BYTES + ka), umma_desc(sB0 + sb * B2_BYTES + kb), IDESC_I8_PAIR, 1u);   // onto the stored constants
                    }
                    umma_commit_pair(bEmpty(sb));       // both CTAs' database stages are free once these MMAs have read them
                    umma_commit_pair(accFull(acc));     // ... and both accumulators are complete
                }
                umma_commit_pair(aEmpty(sa));
            }
        }
    } else {
        knn2_epilogue<true>(A, tmemBase, warp, lane, rank, mPairs, qFull(0), map_to_cta(qEmpty(0), 0), accFull(0), map_to_cta(accEmpty(0), 0),
                            workQ, sMerge);
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmemBase), "r"(512u) : "memory");
}

// ---- one CTA per SM (cta_group::1): same pipeline without the peer; every database tile serves one 128-row query tile
__global__ void __launch_bounds__(TC2_THREADS, 1)
k_knn2_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ Knn2TcArgs A) {
    extern __shared__ uint8_t smemRaw[];
    __shared__ __align__(8) uint64_t bars[16];     // aFull[2] aEmpty[2] bFull[2] bEmpty[2] accFull[2] accEmpty[2] qFull[2] qEmpty[2]
    __shared__ __align__(8) long long workQ[2];    // work items handed from the producer to the MMA and epilogue warps (-1 = done)
    __shared__ uint32_t tmemBaseS;
    __shared__ uint2 sMerge[2][EPI2_PARTS - 1][TC_M];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int W_PRODUCER = EPI2_THREADS / 32, W_MMA = W_PRODUCER + 1;      // (the epilogue warps come first, see k_knn2_tc_pair)
    constexpr uint32_t EPI_WARPS = EPI2_THREADS / 32;
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    const uint32_t sA0 = base, sB0 = base + 2 * A_BYTES;
    const uint32_t bar0 = smem_u32(bars);
    auto aFull = [&](int s) { return bar0 + 8u * (0 + s); };
    auto aEmpty = [&](int s) { return bar0 + 8u * (2 + s); };
    auto bFull = [&](int s) { return bar0 + 8u * (4 + s); };
    auto bEmpty = [&](int s) { return bar0 + 8u * (6 + s); };
    auto accFull = [&](int s) { return bar0 + 8u * (8 + s); };
    auto accEmpty = [&](int s) { return bar0 + 8u * (10 + s); };
    auto qFull = [&](int s) { return bar0 + 8u * (12 + s); };
    auto qEmpty = [&](int s) { return bar0 + 8u * (14 + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) {
            mbar_init(aFull(s), 1); mbar_init(aEmpty(s), 1);
            mbar_init(bFull(s), 1); mbar_init(bEmpty(s), 1);
            mbar_init(accFull(s), 1); mbar_init(accEmpty(s), EPI_WARPS);
            mbar_init(qFull(s), 1); mbar_init(qEmpty(s), 1 + EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBaseS)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = tmemBaseS;

    const long long total = (long long)A.nPairs * A.mTiles;
    if (warp == W_PRODUCER) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&mapB) : "memory");
            // Work items come from a global counter, not from a fixed stride: a CTA that starts late (its SM was held by another
            // kernel, e.g. NCCL's while the descriptor gather is in flight) simply takes fewer of them.
            uint32_t bi = 0;
            for (uint32_t ai = 0;; ai++) {
                const uint32_t sa = ai & 1u;
                mbar_wait(qEmpty(sa), ((ai >> 1) & 1u) ^ 1u);
                long long w = (long long)atomicAdd(A.workCounter, 1ull);
                if (w >= total) w = -1;
                workQ[sa] = w;
                mbar_arrive(qFull(sa));
                if (w < 0) break;
                const int pair = (int)(w / A.mTiles), mt = (int)(w - (long long)pair * A.mTiles);
                const int2 pr = A.pairs[pair];
                mbar_wait(aEmpty(sa), ((ai >> 1) & 1u) ^ 1u);
                mbar_expect_tx(aFull(sa), A_BYTES);
                tma_load_3d(sA0 + sa * A_BYTES, &mapA, 0, mt * TC_M, pr.x, aFull(sa));
                tma_load_3d(sA0 + sa * A_BYTES + A_ATOM_BYTES, &mapA, ATOM_B, mt * TC_M, pr.x, aFull(sa));
                for (int nt = 0; nt < A.nTiles; nt++, bi++) {
                    const uint32_t sb = bi & 1u;
                    mbar_wait(bEmpty(sb), ((bi >> 1) & 1u) ^ 1u);
                    mbar_expect_tx(bFull(sb), B_BYTES);
                    tma_load_3d(sB0 + sb * B_BYTES, &mapB, 0, nt * TC_N, pr.y, bFull(sb));
                    tma_load_3d(sB0 + sb * B_BYTES + B_ATOM_BYTES, &mapB, ATOM_B, nt * TC_N, pr.y, bFull(sb));
                }
            }
        }
    } else if (warp == W_MMA) {
        if (lane == 0) {
            uint32_t bi = 0;
            for (uint32_t ai = 0;; ai++) {
                const uint32_t sa = ai & 1u;
                mbar_wait(qFull(sa), (ai >> 1) & 1u);
                const long long w = *reinterpret_cast<volatile long long*>(&workQ[sa]);
                mbar_arrive(qEmpty(sa));
                if (w < 0) break;
                mbar_wait(aFull(sa), (ai >> 1) & 1u);
                for (int nt = 0; nt < A.nTiles; nt++, bi++) {
                    const uint32_t sb = bi & 1u, ph = (bi >> 1) & 1u;
                    mbar_wait(accEmpty(sb), ph);    // phase u = the accumulator holds its constants for use u (phase 0: the initial store)
                    mbar_wait(bFull(sb), ph);
                    tc_fence_after();
                    const uint32_t d = tmemBase + sb * TC_N;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const uint32_t ka = (uint32_t)(k >> 2) * A_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                        const uint32_t kb = (uint32_t)(k >> 2) * B_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                        umma_i8(d, umma_desc(sA0 + sa * A_BYTES + ka), umma_desc(sB0 + sb * B_BYTES + kb), IDESC_I8, 1u);   // onto the stored constants
                    }
                    umma_commit(bEmpty(sb));        // the database stage is free once these MMAs have read it
                    umma_commit(accFull(sb));       // ... and the accumulator is complete
                }
                umma_commit(aEmpty(sa));
            }
        }
    } else {
        knn2_epilogue<false>(A, tmemBase, warp, lane, 0u, A.mTiles, qFull(0), qEmpty(0), accFull(0), accEmpty(0), workQ, sMerge);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmemBase), "r"(512u) : "memory");
}

// ---- measured denominator of the knn2 roofline: the same tcgen05.mma shape (kind::i8, M 128, N 256, K 32, operands in
// SWIZZLE_128B shared memory, cta_group::1) issued back to back with no loads and no epilogue, one CTA per SM
__global__ void __launch_bounds__(128, 1) k_imma_peak(int iters, int* sink) {
    extern __shared__ uint8_t smemRaw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBaseS;
    const int warp = threadIdx.x >> 5;
    const uint32_t base = (smem_u32(smemRaw) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smemRaw + (base - smem_u32(smemRaw)))[i] = 0x01ff01ffu;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmemBaseS)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores above -> async-proxy reads of the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmemBase = tmemBaseS;
    if (threadIdx.x == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t ka = (uint32_t)(k >> 2) * A_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                const uint32_t kb = (uint32_t)(k >> 2) * B_ATOM_BYTES + (uint32_t)(k & 3) * 32u;
                umma_i8(tmemBase + (uint32_t)(it & 1) * TC_N, umma_desc(base + ka), umma_desc(base + A_BYTES + kb), IDESC_I8, k != 0);
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        if (sink && iters < 0) *sink = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmemBase), "r"(512u) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tma_encoder() {
    static const EncodeTiledFn enc = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    return enc;
}

}  // namespace

// int8 tensor-core throughput of this device in the shape k_knn2_tc uses: *tops = 2 * M * N * K operations per tcgen05.mma / time
cudaError_t knn2_tc_peak(double* tops) {
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_imma_peak);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    const int iters = 4096;
    cudaEvent_t e0, e1;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_imma_peak<<<sms, 128, A_BYTES + B_BYTES + 1024>>>(iters, nullptr);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    *tops = 2.0 * TC_M * TC_N * 32 * 8.0 * iters * sms / (best * 1e-3) / 1e12;
    return cudaGetLastError();
}

size_t knn2_tc_expanded_bytes(int nKeyframes, int n) { return (size_t)nKeyframes * (size_t)n * TC_KB; }
size_t knn2_tc_scratch_bytes(int nKeyframes) { return (((size_t)nKeyframes + 7) & ~(size_t)7) + 8; }

cudaError_t launch_knn2_tc(const Knn2Args& a, int nKeyframes, uint8_t* expanded, uint8_t* used, cudaStream_t st) {
    if (a.nPairs <= 0 || a.n <= 0) return cudaSuccess;
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return cudaErrorNotSupported;
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_knn2_tc);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;

    // `used`: nKeyframes flags, then (8-byte aligned) the work counter of the tensor-core kernel
    const size_t counterOff = ((size_t)nKeyframes + 7) & ~(size_t)7;
    e = cudaMemsetAsync(used, 0, counterOff + 8, st);
    if (e != cudaSuccess) return e;
    k_knn2_mark<<<(a.nPairs + 255) / 256, 256, 0, st>>>(a.pairs, a.nPairs, nKeyframes, used);
    k_knn2_expand<<<nKeyframes, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(a.desc), reinterpret_cast<uint4*>(expanded), a.n * 16, used);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;

    CUtensorMap mapA, mapB;
    const cuuint64_t dims[3] = {(cuuint64_t)TC_KB, (cuuint64_t)a.n, (cuuint64_t)nKeyframes};
    const cuuint64_t strides[2] = {(cuuint64_t)TC_KB, (cuuint64_t)a.n * TC_KB};
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t boxA[3] = {ATOM_B, TC_M, 1}, boxB[3] = {ATOM_B, TC_N, 1};
    if (enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, expanded, dims, strides, boxA, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;
    if (enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, expanded, dims, strides, boxB, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorInvalidValue;

    Knn2TcArgs k;
    k.pairs = a.pairs; k.nPairs = a.nPairs; k.n = a.n;
    k.mTiles = (a.n + TC_M - 1) / TC_M; k.nTiles = (a.n + TC_N - 1) / TC_N;
    k.thLow = a.thLow; k.nnratio = a.nnratio;
    k.bestIdx = a.bestIdx; k.bestDist = a.bestDist; k.secondDist = a.secondDist;
    k.workCounter = reinterpret_cast<unsigned long long*>(used + counterOff);
    if (g_knn2Pair.load(std::memory_order_relaxed) && sms >= 2) {
        // CTA pairs (cta_group::2): one cluster of two CTAs per TPC, each cluster takes (pair, two query tiles) work items
        e = OBS_ALLOW_MAX_SMEM(k_knn2_tc_pair);
        if (e != cudaSuccess) return e;
        const long long items = (long long)k.nPairs * ((k.mTiles + 1) / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2u * (unsigned)std::min<long long>(items, sms / 2));
        cfg.blockDim = dim3(TC2_THREADS); cfg.dynamicSmemBytes = TC2_SMEM; cfg.stream = st;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, k_knn2_tc_pair, mapA, k);
    }
    const long long total = (long long)k.nPairs * k.mTiles;
    const unsigned grid = (unsigned)std::min<long long>(total, sms);
    k_knn2_tc<<<grid, TC2_THREADS, TC_SMEM, st>>>(mapA, mapB, k);
    return cudaGetLastError();
}

void knn2_tc_set_cta_pair(bool on) { g_knn2Pair.store(on ? 1 : 0, std::memory_order_relaxed); }
