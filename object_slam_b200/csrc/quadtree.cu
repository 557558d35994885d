// K3 -- quadtree keypoint distribution.  Replaces ORBextractor::DistributeOctTree
// (src/ORBextractor.cc:539-763) and ExtractorNode::DivideNode (:481-537).
//
// The reference walks a std::list of nodes, splitting and push_front-ing children, and breaks
// ties between equally populated nodes by heap address (sort of pair<int, ExtractorNode*>, :684).
// The canonical execution reproduced here is the one where addresses grow with creation order
// (the reference under a bump allocator; see oracle/ref_harness.cpp).  Seen level-synchronously
// the algorithm is a sequence of passes.  In every pass the splittable nodes (more than one
// key) are exactly the children created by the previous pass, and
//   * a sweep pass (:606-665) visits them in list order and splits them all;
//   * a largest-first pass (:673-738) visits them by (key count desc, list position asc -- the
//     latest created first) and stops at the first prefix that brings the list to >= N nodes;
//   * afterwards the list is reverse(children in creation order) ++ (unsplit nodes, old order).
// With node id == list position this is counting, ranking and prefix sums over at most N+3
// nodes and one partition step over the keys per pass: one CTA runs one (image, level) problem
// with the node tables in shared memory.  Output: per final node the key with the largest
// response, the first in candidate order winning ties (:741-760), in list order.
#include "kernels.h"

namespace {

#ifndef QT_NT
#define QT_NT 512
#endif
constexpr int QT_THREADS = QT_NT;
#ifndef QT_KEYS
#define QT_KEYS 4096
#endif
constexpr int QT_SMEM_KEYS = QT_KEYS;       // problems with at most this many candidates keep keys + node ids in shared memory

struct Smem {
    // carve-up of the dynamic shared memory block for node capacity nc
    ushort4* rect[2];
    int* cnt[2];
    int* childCnt;        // 4*nc
    int* scanA;           // nc   (ne per rank -> exclusive prefix)
    int* scanB;           // nc   (flags -> exclusive prefix)
    int* neArr;           // nc
    uint16_t* rank;       // nc
    uint16_t* parentAt;   // nc
    uint16_t* selfPos;    // nc
    uint16_t* childPos;   // 4*nc
    uint32_t* keys;       // QT_SMEM_KEYS
    uint16_t* nodeOf;     // QT_SMEM_KEYS
    int* warpTmp;         // 32 + misc
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

__host__ __device__ inline size_t smem_layout(int nc, uint8_t* base, Smem* s) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o = align16(o + bytes); return r; };
    size_t oR0 = take(sizeof(ushort4) * nc), oR1 = take(sizeof(ushort4) * nc);
    size_t oC0 = take(4 * nc), oC1 = take(4 * nc);
    size_t oCC = take(16 * nc), oSA = take(4 * nc), oSB = take(4 * nc), oNE = take(4 * nc);
    size_t oRk = take(2 * nc), oPa = take(2 * nc), oSp = take(2 * nc), oCp = take(8 * nc);
    size_t oK = take(4 * QT_SMEM_KEYS), oN = take(2 * QT_SMEM_KEYS), oW = take(4 * 48);
    if (s) {
        s->rect[0] = (ushort4*)(base + oR0); s->rect[1] = (ushort4*)(base + oR1);
        s->cnt[0] = (int*)(base + oC0); s->cnt[1] = (int*)(base + oC1);
        s->childCnt = (int*)(base + oCC); s->scanA = (int*)(base + oSA); s->scanB = (int*)(base + oSB);
        s->neArr = (int*)(base + oNE);
        s->rank = (uint16_t*)(base + oRk); s->parentAt = (uint16_t*)(base + oPa);
        s->selfPos = (uint16_t*)(base + oSp); s->childPos = (uint16_t*)(base + oCp);
        s->keys = (uint32_t*)(base + oK); s->nodeOf = (uint16_t*)(base + oN); s->warpTmp = (int*)(base + oW);
    }
    return o;
}

// In-place exclusive prefix sum of a[0..n) by the whole CTA; returns the total to every thread.
__device__ int block_scan_excl(int* a, int n, int* tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = QT_THREADS / 32;
    int carry = 0;
    for (int base = 0; base < n; base += QT_THREADS) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) tmp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = lane < NW ? tmp[lane] : 0;
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            if (lane < NW) tmp[lane] = wi - w;
            if (lane == 31) tmp[32] = wi;
        }
        __syncthreads();
        if (i < n) a[i] = carry + tmp[warp] + incl - v;
        carry += tmp[32];
        __syncthreads();
    }
    return carry;
}

__device__ int block_sum(int v, int* tmp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0) tmp[warp] = v;
    __syncthreads();
    int t = 0;
    if (warp == 0) {
        t = lane < QT_THREADS / 32 ? tmp[lane] : 0;
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0) tmp[32] = t;
    }
    __syncthreads();
    t = tmp[32];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(QT_THREADS) k_quadtree(const __grid_constant__ Geom g, int nc,
                                                         const uint32_t* __restrict__ cand, const int* __restrict__ cellCount,
                                                         uint32_t* __restrict__ keyScratch, uint16_t* __restrict__ nodeScratch,
                                                         uint32_t* __restrict__ sel, int* __restrict__ selCount) {
    extern __shared__ __align__(16) uint8_t smraw[];
    Smem S;
    smem_layout(nc, smraw, &S);
    __shared__ int sMisc[8];
    pdl_entry();

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int level = g.nlevels - 1 - (int)blockIdx.x;      // any order is correct; this one just interleaves problem sizes
    const int img = blockIdx.y;
    const LevelGeom& lg = g.lv[level];
    const int N = lg.nfeat;
    uint32_t* selOut = sel + ((size_t)img * g.nlevels + level) * g.selCap;
    int* selCountOut = selCount + (size_t)img * g.nlevels + level;

    // ---- 0. gather the level's candidates in the reference's order (cell-row-major, row-major inside a cell)
    const int nCells = lg.nCols * lg.nRows;
    const int* cc = cellCount + (size_t)img * g.nCellsTotal + lg.cellBase;
    const uint32_t* slots = cand + (size_t)img * g.slotTotal + lg.slotBase;
    int part = 0;
    for (int c = tid; c < nCells; c += QT_THREADS) part += cc[c];
    const int M = block_sum(part, S.warpTmp);
    const int width = lg.w - 2 * OBS_BORDER, height = lg.h - 2 * OBS_BORDER;
    int nIni = 0;
    if (width > 0 && height > 0) nIni = (int)roundf(__fdiv_rn((float)width, (float)height));      // :543
    if (M == 0 || nIni < 1 || nIni > nc) {       // nothing to distribute (nIni < 1 is undefined behaviour in the reference)
        if (tid == 0) *selCountOut = 0;
        return;
    }
    const bool inSmem = M <= QT_SMEM_KEYS;
    uint32_t* keys = inSmem ? S.keys : keyScratch + (size_t)img * g.slotTotal + lg.slotBase;
    uint16_t* nodeOf = inSmem ? S.nodeOf : nodeScratch + (size_t)img * g.slotTotal + lg.slotBase;
    {
        int carry = 0;
        for (int base = 0; base < nCells; base += QT_THREADS) {
            const int c = base + tid;
            // chunk-local exclusive prefix of the cell counts, kept in registers + warp scan
            const int v = c < nCells ? cc[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            if (lane == 31) S.warpTmp[warp] = incl;
            __syncthreads();
            if (warp == 0) {
                const int w = lane < QT_THREADS / 32 ? S.warpTmp[lane] : 0;
                int wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
                if (lane < QT_THREADS / 32) S.warpTmp[lane] = wi - w;
                if (lane == 31) S.warpTmp[32] = wi;
            }
            __syncthreads();
            const int ofs = carry + S.warpTmp[warp] + incl - v;
            // copy: a thread moves the (few) candidates of its own cell; four loads in flight at a time
            if (v > 0) {
                const uint32_t* slot = slots + (size_t)c * lg.cellCap;
                uint32_t* dstk = keys + ofs;
                int e = 0;
                for (; e + 4 <= v; e += 4) {
                    const uint32_t a0 = slot[e], a1 = slot[e + 1], a2 = slot[e + 2], a3 = slot[e + 3];
                    dstk[e] = a0; dstk[e + 1] = a1; dstk[e + 2] = a2; dstk[e + 3] = a3;
                }
                for (; e < v; e++) dstk[e] = slot[e];
            }
            carry += S.warpTmp[32];
            __syncthreads();
        }
    }

    // ---- 1. root nodes (:545-585)
    const float hX = __fdiv_rn((float)width, (float)nIni);
    for (int i = tid; i < nIni; i += QT_THREADS) S.childCnt[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += QT_THREADS) {
        int r = (int)__fdiv_rn((float)key_x(keys[k]), hX);
        r = min(r, nIni - 1);
        nodeOf[k] = (uint16_t)r;
        atomicAdd(&S.childCnt[r], 1);
    }
    __syncthreads();
    for (int i = tid; i < nIni; i += QT_THREADS) S.scanB[i] = S.childCnt[i] > 0;
    __syncthreads();
    int L = block_scan_excl(S.scanB, nIni, S.warpTmp);
    for (int i = tid; i < nIni; i += QT_THREADS) {
        if (S.childCnt[i] > 0) {
            const int pos = S.scanB[i];
            S.rect[0][pos] = make_ushort4((unsigned short)(int)__fmul_rn(hX, (float)i), 0,
                                          (unsigned short)(int)__fmul_rn(hX, (float)(i + 1)), (unsigned short)height);
            S.cnt[0][pos] = S.childCnt[i];
            S.selfPos[i] = (uint16_t)pos;
        }
    }
    __syncthreads();
    for (int k = tid; k < M; k += QT_THREADS) nodeOf[k] = S.selfPos[nodeOf[k]];
    __syncthreads();

    // ---- 2. passes
    int cur = 0;
    bool largestFirst = false;
    while (true) {
        const ushort4* R = S.rect[cur];
        const int* C = S.cnt[cur];
        ushort4* Rn = S.rect[cur ^ 1];
        int* Cn = S.cnt[cur ^ 1];

        // 2.1 visiting rank of every splittable node
        int P;
        if (!largestFirst) {
            for (int i = tid; i < L; i += QT_THREADS) S.scanB[i] = C[i] > 1;
            __syncthreads();
            P = block_scan_excl(S.scanB, L, S.warpTmp);
            for (int i = tid; i < L; i += QT_THREADS) if (C[i] > 1) S.rank[i] = (uint16_t)S.scanB[i];
        } else {
            int mine = 0;
            for (int i = tid; i < L; i += QT_THREADS) {
                const int ci = C[i];
                if (ci > 1) {
                    mine++;
                    int r = 0;
                    for (int j = 0; j < L; j++) {
                        const int cj = C[j];
                        r += (cj > 1) & ((cj > ci) | ((cj == ci) & (j < i)));
                    }
                    S.rank[i] = (uint16_t)r;
                }
            }
            P = block_sum(mine, S.warpTmp);
        }
        if (P == 0) break;                      // no node can be split: list size unchanged (:669 / :736)
        __syncthreads();
        for (int i = tid; i < L; i += QT_THREADS) if (C[i] > 1) S.parentAt[S.rank[i]] = (uint16_t)i;
        for (int i = tid; i < 4 * L; i += QT_THREADS) S.childCnt[i] = 0;
        if (tid == 0) { sMisc[0] = P; sMisc[1] = 0; }           // [0] = nsplit, [1] = nToExpand
        __syncthreads();

        // 2.2 quadrant of every key of a splittable node (DivideNode, :481-526)
        for (int k = tid; k < M; k += QT_THREADS) {
            const int nd = nodeOf[k];
            if (C[nd] > 1) {
                const uint32_t key = keys[k];
                const ushort4 r = R[nd];
                const int mx = r.x + ((r.z - r.x + 1) >> 1);          // UL.x + ceil((UR.x-UL.x)/2)
                const int my = r.y + ((r.w - r.y + 1) >> 1);
                const int q = (key_x(key) < mx ? 0 : 1) + (key_y(key) < my ? 0 : 2);
                atomicAdd(&S.childCnt[4 * nd + q], 1);
                nodeOf[k] = (uint16_t)(nd | (q << 14));
            }
        }
        __syncthreads();

        // 2.3 non-empty children per visited parent, prefix sums, stopping point
        for (int r = tid; r < P; r += QT_THREADS) {
            const int* c4 = S.childCnt + 4 * S.parentAt[r];
            const int ne = (c4[0] > 0) + (c4[1] > 0) + (c4[2] > 0) + (c4[3] > 0);
            S.neArr[r] = ne;
            S.scanA[r] = ne;
        }
        __syncthreads();
        block_scan_excl(S.scanA, P, S.warpTmp);
        if (largestFirst) {
            for (int r = tid; r < P; r += QT_THREADS)
                if (L + S.scanA[r] + S.neArr[r] - (r + 1) >= N) atomicMin(&sMisc[0], r + 1);     // :733-734
            __syncthreads();
        }
        const int nsplit = sMisc[0];
        const int Ctot = S.scanA[nsplit - 1] + S.neArr[nsplit - 1];

        // 2.4 new positions: children in reverse creation order, then the unsplit nodes in old order
        // (in a sweep every splittable node splits and scanB still holds "splittable nodes before i" from 2.1, so the unsplit
        // nodes before i number i - scanB[i]; a largest-first pass needs a scan over the nodes it actually split)
        if (largestFirst) {
            for (int i = tid; i < L; i += QT_THREADS) S.scanB[i] = !(C[i] > 1 && S.rank[i] < nsplit);
            __syncthreads();
            block_scan_excl(S.scanB, L, S.warpTmp);
        }
        for (int i = tid; i < L; i += QT_THREADS) {
            if (!(C[i] > 1 && S.rank[i] < nsplit)) {
                const int pos = Ctot + (largestFirst ? S.scanB[i] : i - S.scanB[i]);
                S.selfPos[i] = (uint16_t)pos;
                Rn[pos] = R[i];
                Cn[pos] = C[i];
            }
        }
        int expand = 0;
        for (int r = tid; r < nsplit; r += QT_THREADS) {
            const int i = S.parentAt[r];
            const ushort4 pr = R[i];
            const int mx = pr.x + ((pr.z - pr.x + 1) >> 1);
            const int my = pr.y + ((pr.w - pr.y + 1) >> 1);
            int c = S.scanA[r];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int n = S.childCnt[4 * i + q];
                if (n > 0) {
                    const int pos = Ctot - 1 - c;
                    c++;
                    Rn[pos] = make_ushort4((q & 1) ? mx : pr.x, (q & 2) ? my : pr.y, (q & 1) ? pr.z : mx, (q & 2) ? pr.w : my);
                    Cn[pos] = n;
                    S.childPos[4 * i + q] = (uint16_t)pos;
                    expand += n > 1;
                }
            }
        }
        if (expand) atomicAdd(&sMisc[1], expand);
        __syncthreads();

        // 2.5 move the keys
        for (int k = tid; k < M; k += QT_THREADS) {
            const int v = nodeOf[k];
            const int nd = v & 0x3fff;
            const bool split = C[nd] > 1 && S.rank[nd] < nsplit;
            nodeOf[k] = split ? S.childPos[4 * nd + (v >> 14)] : S.selfPos[nd];
        }
        const int Lnew = Ctot + (L - nsplit);
        const int nToExpand = sMisc[1];
        __syncthreads();
        const bool fin = Lnew >= N || Lnew == L;          // :669, :736
        L = Lnew;
        cur ^= 1;
        if (fin) break;
        if (!largestFirst && L + 3 * nToExpand > N) largestFirst = true;     // :671
    }
    __syncthreads();

    // ---- 3. best key per node (:741-760): largest response, first in candidate order on ties
    uint32_t* best = reinterpret_cast<uint32_t*>(S.scanA);
    for (int i = tid; i < L; i += QT_THREADS) best[i] = 0;
    __syncthreads();
    for (int k = tid; k < M; k += QT_THREADS)
        atomicMax(&best[nodeOf[k] & 0x3fff], ((uint32_t)key_r(keys[k]) << 24) | (0xffffffu - (uint32_t)k));
    __syncthreads();
    for (int i = tid; i < L; i += QT_THREADS) selOut[i] = keys[0xffffffu - (best[i] & 0xffffffu)];
    if (tid == 0) *selCountOut = L;
}

}  // namespace

size_t quadtree_smem_bytes(int nodeCap) { return smem_layout(nodeCap, nullptr, nullptr); }

cudaError_t quadtree_prepare(int nodeCap) {
    cudaError_t e = OBS_ALLOW_MAX_SMEM(k_quadtree);
    if (e != cudaSuccess) return e;
    return (int)quadtree_smem_bytes(nodeCap) <= obsdetail::max_dynamic_smem((const void*)k_quadtree) ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t launch_quadtree(const Geom& g, int nodeCap, const uint32_t* cand, const int* cellCount,
                            uint32_t* keyScratch, uint16_t* nodeScratch, uint32_t* sel, int* selCount,
                            int nimg, cudaStream_t st) {
    dim3 grid(g.nlevels, nimg);
    cudaError_t le = launch_k(pdl_enabled(), k_quadtree, grid, dim3(QT_THREADS), quadtree_smem_bytes(nodeCap), st, g, nodeCap, cand, cellCount, keyScratch, nodeScratch, sel, selCount);
    if (le != cudaSuccess) return le;
    return cudaGetLastError();
}
