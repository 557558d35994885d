// Shared definitions of the sm_100a ORB front end (geometry tables, packed record formats,
// small device helpers).  Nothing here is a fallback path: every consumer is a CUDA kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define OBS_MAX_LEVELS 12
#define OBS_EDGE 19              // EDGE_THRESHOLD, src/ORBextractor.cc:63
#define OBS_BORDER 16            // EDGE_THRESHOLD-3: origin of the FAST cell grid, :773
#define OBS_HALF_PATCH 15        // HALF_PATCH_SIZE, :62
#define OBS_HDR_INTS 16          // per-image result header: [0]=n, [1..nlevels]=per-level counts

// One pyramid level of one image shape.
struct LevelGeom {
    int w, h;            // level size (cvRound(w0 * invScale), :1111-1112)
    int pitch;           // row pitch in bytes inside the pyramid / blur slabs (multiple of 128)
    unsigned off;        // byte offset of the level inside one image's slab
    int nCols, nRows;    // FAST cell grid (:781-784)
    int wCell, hCell;    // cell pitch (:785-786)
    int cellBase;        // index of the level's first cell in the per-image cell table
    int cellCap;         // capacity (entries) of one cell's candidate slot
    unsigned slotBase;   // index (u32 units) of the level's first slot in the per-image slab;
                         // the compact key array and the node-id scratch of the level start here too
    int nfeat;           // mnFeaturesPerLevel[level] (:435-446)
    int xtab, ytab;      // first entry of the level's resize coefficient tables (level >= 1)
    int fastCellsPerCta; // cells of one cell row handled by one FAST CTA
    int fastCtaBase;     // first CTA of this level in the FAST grid
    int blurTileBase;    // first CTA of this level in the blur grid
    int blurTilesX;      // tiles per row of the blur grid
    int blurEdgeBase;    // first border-strip work item of this level
    int rsPitch;         // resize of this level from the previous one: shared-memory row pitch of the staged source
    int rsRows;          //   tile (0 = taps too far apart for the tiled kernel) and its row capacity
    double rsScaleX, rsScaleY;   // resize from the previous level: source step per destination pixel, exactly the doubles the host built the
                         //   tap tables with (a CTA derives its source footprint from them without a dependent table load)
    float scale;         // mvScaleFactor[level]
    float invScale;      // mvInvScaleFactor[level]
    float patchSize;     // (float)(int)(31 * scale), :838
};

struct Geom {
    int nlevels;
    int w, h;
    int iniTh, minTh;
    int nCellsTotal;         // cells per image over all levels
    unsigned slotTotal;      // u32 entries of one image's candidate slab
    unsigned slabBytes;      // bytes of one image's pyramid slab (same layout for the blur slab)
    int selCap;              // per-level capacity of the selected-keypoint list (max nfeat + 4)
    int kpCap;               // per-image keypoint capacity of the result record
    int blurTilesTotal;
    int blurEdgeCtas;        // CTAs appended to the blur grid for the border strips
    int fastCtasTotal;
    int fastTileRows;        // rows of the FAST window tile: max hCell over the levels + 6
    int umax[OBS_HALF_PATCH + 1];
    LevelGeom lv[OBS_MAX_LEVELS];
};

// Where the levels of the current batch live.  Level 0 may be the caller's own device buffer.
struct PyrPtrs {
    const uint8_t* l0;       // level 0 of image 0
    size_t l0ImgStride;
    int l0Pitch;
    uint8_t* slab;           // levels 1.. of image 0 (level offsets from Geom)
    size_t slabStride;       // bytes between images
};

__device__ __forceinline__ const uint8_t* level_ptr(const PyrPtrs& p, const Geom& g, int img, int l, int& pitch) {
    if (l == 0) { pitch = p.l0Pitch; return p.l0 + (size_t)img * p.l0ImgStride; }
    pitch = g.lv[l].pitch;
    return p.slab + (size_t)img * p.slabStride + g.lv[l].off;
}

// ---- programmatic dependent launch (PDL).  Every kernel of the per-frame chain starts with pdl_entry(): it lets the NEXT kernel of
// the stream be scheduled right away (griddepcontrol.launch_dependents -- that kernel then runs its own prologue and parks in its
// griddepcontrol.wait) and then waits until everything the PREVIOUS kernel wrote is complete and visible.  Nothing touches global
// memory before the wait, so no kernel can overtake the data it depends on.  Without the launch attribute (host_util.h launch_k
// with pdl = false) both instructions are no-ops.
__device__ __forceinline__ void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- helpers against ptxas rematerialisation.  Under a register bound (__launch_bounds__) ptxas prefers recomputing loop
// invariants -- the lane index, shared-window bases (S2R CgaCtaId + LEA), loop bounds -- in every iteration over keeping them in
// registers; in issue-bound loops that is 10-20 % of the instructions.  A value that went through an opaque shuffle (per-lane
// values) or a volatile shared-memory load (warp-uniform values: a shuffle of those is folded away) cannot be recomputed.
__device__ __forceinline__ unsigned pin(unsigned x) {
    unsigned y;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(y) : "r"(x), "r"(threadIdx.x & 31u));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_volatile_u32(const void* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)));
    return v;
}
template <int OFS> __device__ __forceinline__ unsigned lds_u8(uint32_t a) {      // LDS.U8 [R + imm] on a shared-window address
    unsigned v;
    asm("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFS));
    return v;
}

// Candidate / key record: x (12 bits) | y (12 bits) << 12 | FAST response (8 bits) << 24,
// x and y relative to the 16-px border origin as the reference hands them to DistributeOctTree.
__device__ __forceinline__ uint32_t pack_key(int x, int y, int r) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)r << 24); }
__device__ __forceinline__ int key_x(uint32_t k) { return (int)(k & 0xfffu); }
__device__ __forceinline__ int key_y(uint32_t k) { return (int)((k >> 12) & 0xfffu); }
__device__ __forceinline__ int key_r(uint32_t k) { return (int)(k >> 24); }

struct FastCta {                // one CTA of the FAST grid: a run of cells of one cell row
    short level, ci, j0, n;
};

struct alignas(4) ResizeTap {   // one destination index of one axis
    int32_t ofs;                // source index
    int16_t a0, a1;             // 11-bit fixed-point weights of source[ofs], source[ofs+1]
};
