// C ABI of the library (include/obslam_b200.h): handle management, geometry tables, stream
// orchestration.  All image work happens in the kernels of this directory; this file holds no
// CPU implementation of any stage.
#include "../../include/obslam_b200.h"
#include "kernels.h"
#include "host_util.h"
#include "matcher_api.h"
#include "knn2_tc.h"

#include <nvtx3/nvToolsExt.h>          // header-only NVTX v3: the ranges cost nothing unless a tool is attached

#include <cmath>
#include <cstdlib>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

namespace obsdetail {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// true for cudaMallocHost / cudaHostRegister memory: such buffers are DMA-ed directly, without the staging copy
bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

bool is_device(const void* p) {
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// process-wide launch options (obs_set_option)
static std::atomic<int> g_pdl{1}, g_graphs{1};
// (inside a graph capture the kernels are launched without the attribute: graph edges between kernels are already cheaper than the
// programmatic ones -- 0.172 vs 0.176 ms per live stereo frame)
static thread_local bool tl_capturing = false;
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0 && !tl_capturing; }
void set_capturing(bool on) { tl_capturing = on; }
void set_pdl_enabled(bool on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }
bool graphs_enabled() { return g_graphs.load(std::memory_order_relaxed) != 0; }
void set_graphs_enabled(bool on) { g_graphs.store(on ? 1 : 0, std::memory_order_relaxed); }
std::atomic<unsigned long long> g_allocEpoch{0};

int max_dynamic_smem(const void* kernel) {
    int dev = 0, optin = 0;
    cudaFuncAttributes fa;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess ||
        cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); return 0; }
    return optin - (int)fa.sharedSizeBytes;
}

cudaError_t allow_max_smem(const void* kernel, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    const int lim = max_dynamic_smem(kernel);
    if (lim <= 0) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

}  // namespace obsdetail

namespace {

inline int cv_round_f(float v) { return (int)nearbyintf(v); }      // cvRound: round half to even
inline int cv_round_d(double v) { return (int)nearbyint(v); }
inline int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil_d(double v) { int i = (int)v; return i + (i < v); }
inline short sat_short(float v) { int i = cv_round_f(v); return (short)(i < -32768 ? -32768 : i > 32767 ? 32767 : i); }

constexpr int PROF_SLOTS = 128;

// NVTX range around the enqueue of one stage, only while obs_extractor_set_profiling is on (the reference's ad-hoc
// std::chrono timers around the same stages: src/Tracking.cc:270-273)
struct StageRange {
    bool on;
    StageRange(bool enabled, const char* name) : on(enabled) { if (on) nvtxRangePushA(name); }
    ~StageRange() { if (on) nvtxRangePop(); }
};

}  // namespace

struct obs_extractor {
    obs_orb_params prm;
    int device = 0;
    int maxW = 0, maxH = 0, maxBatch = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;    // the blur runs here, beside FAST + quadtree (both only need the pyramid)
    cudaEvent_t done = nullptr, fork = nullptr, join = nullptr;
    cudaEvent_t hostDone = nullptr;    // cudaEventBlockingSync: the host thread sleeps on it instead of spinning in a stream synchronise
    cudaEvent_t hostDoneSpin = nullptr; // plain event for small batches (a live frame): the wake-up of a sleeping thread costs more than the frame
    bool pendingSpin = false;
    bool pending = false;              // obs_stereo_frames_submit without its obs_stereo_frames_wait
    const int* pendingCounts[2] = {nullptr, nullptr};
    int pendingN = 0, pendingCap = 0;
    cudaStream_t cin = nullptr, cout = nullptr;       // host path: upload / download streams of the chunk pipeline
    std::vector<cudaEvent_t> chunkIn, chunkDone;

    // tables of the constructor (src/ORBextractor.cc:410-470)
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel;
    int umax[OBS_HALF_PATCH + 1];

    // geometry of the current image shape
    Geom g;
    int nodeCap = 0;
    bool geomValid = false;
    std::vector<ResizeTap> hXtab, hYtab;
    DevBuf<ResizeTap> dXtab, dYtab;
    std::vector<FastCta> hFastCtas;
    DevBuf<FastCta> dFastCtas;

    // device state of the last batch
    DevBuf<uint8_t> pyr, blur, records, rawIn;
    DevBuf<uint32_t> cand, keyScratch, sel;
    DevBuf<uint16_t> nodeScratch;
    DevBuf<int> cellCount, selCount;
    DevBuf<float> uRight, depth;
    DevBuf<int> sad, rowStart;
    DevBuf<uint2> rowIdx;
    PinBuf<uint8_t> stageIn, stageOut;
    size_t recordBytes = 0;
    PyrPtrs ptrs{};
    DescMaps blurMaps{};           // TMA descriptors of the blur slab's levels, valid for (mapsPtr, mapsB, geometry)
    DevBuf<CUtensorMap> dMaps;     // ... and their device copy, which the descriptor kernel reads
    const uint8_t* mapsPtr = nullptr;
    size_t mapsB = 0;
    int lastN = 0;                 // images in the last batch (0 = nothing extracted yet)
    cudaStream_t lastStream = nullptr;

    // CUDA graphs of obs_stereo_frames_submit (owned by the left handle): one executable graph per distinct argument set
    struct StereoGraph {
        const void* io[10];
        int n, w, h, cap;
        size_t stride;
        float mbf, minD, maxD;
        obs_extractor* right;
        unsigned long long epoch;      // g_allocEpoch at capture: any device reallocation since then invalidates the pointers inside
        bool pdl;
        cudaGraphExec_t exec;          // nullptr: arguments seen once (plain enqueue); captured when they come a second time
        unsigned long long lastUse;
        bool same_args(const StereoGraph& o) const {
            return memcmp(io, o.io, sizeof(io)) == 0 && n == o.n && w == o.w && h == o.h && cap == o.cap && stride == o.stride &&
                   mbf == o.mbf && minD == o.minD && maxD == o.maxD && right == o.right;
        }
    };
    std::vector<StereoGraph> graphs;
    // CUDA graph of the single-image host path of obs_extract (staging buffers are the handle's own, so the graph only depends on the shape)
    cudaGraphExec_t monoExec = nullptr;
    int monoW = 0, monoH = 0;
    bool monoSeen = false;
    unsigned long long monoEpoch = 0;
    unsigned long long graphClock = 0;
    cudaEvent_t gjoin[3] = {nullptr, nullptr, nullptr};   // joins of the side streams at the end of a capture

    // optional per-stage CUDA-event timing (obs_extractor_set_profiling)
    bool prof = false;
    std::vector<cudaEvent_t> pev;  // ring of PROF_SLOTS x OBS_NUM_STAGES x (start, end) events
    int profCalls = 0;
    std::vector<cudaEvent_t> sev;  // stereo: ring of PROF_SLOTS x 2 events (owned by the left handle)
    int stereoCalls = 0;
};

namespace {

void build_tables(obs_extractor* e) {
    const obs_orb_params& P = e->prm;
    const int nl = P.nlevels;
    const double sf = (double)P.scale_factor;          // the reference keeps the float argument in a double member
    e->scale.assign(nl, 1.f); e->sigma2.assign(nl, 1.f);
    e->invScale.assign(nl, 1.f); e->invSigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * sf);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nl; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    e->featPerLevel.assign(nl, 0);
    const float factor = (float)(1.0f / sf);
    float nDesired = (float)(P.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl)));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        e->featPerLevel[l] = cv_round_f(nDesired);
        sum += e->featPerLevel[l];
        nDesired *= factor;
    }
    e->featPerLevel[nl - 1] = std::max(P.nfeatures - sum, 0);
    // circular patch half-widths (:454-469)
    const int HP = OBS_HALF_PATCH;
    int v, v0;
    const int vmax = cv_floor_d(HP * sqrt(2.f) / 2 + 1);
    const int vmin = cv_ceil_d(HP * sqrt(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
}

// OpenCV resize(INTER_LINEAR, 8U) tap table of one axis: source index and 11-bit weights.
void resize_taps(int ssize, int dsize, bool isX, ResizeTap* out) {
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_d(f);
        f -= s;
        if (isX) {
            if (s < 0) { f = 0; s = 0; }
            if (s >= ssize - 1) { f = 0; s = ssize - 1; }
        }
        out[d].ofs = s;
        out[d].a0 = sat_short((1.f - f) * 2048.f);
        out[d].a1 = sat_short(f * 2048.f);
    }
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// Source footprint of the tiled resize kernel (pyramid.cu): every group of four output columns must read
// within 7 source bytes, and the staged source tile of a 128 x 64 output tile must fit shared memory.
void resize_tile_geometry(const LevelGeom& S, LevelGeom& D, const ResizeTap* xt, const ResizeTap* yt) {
    D.rsPitch = D.rsRows = 0;
    int maxBytes = 0, maxRows = 0;
    for (int x = 0; x < D.w; x += 4) {
        const int last = std::min(x + 3, D.w - 1);
        for (int i = x; i <= last; i++) if (xt[i].ofs - xt[x].ofs < 0 || xt[i].ofs - xt[x].ofs > 6) return;
    }
    for (int x = 0; x < D.w; x += RESIZE_TILE_W) {
        const int last = std::min(x + RESIZE_TILE_W, D.w) - 1;
        maxBytes = std::max(maxBytes, xt[last].ofs + 1 - (xt[x].ofs & ~15) + 1);
    }
    for (int y = 0; y < D.h; y += RESIZE_TILE_H) {
        const int last = std::min(y + RESIZE_TILE_H, D.h) - 1;
        const int y0 = std::min(std::max(yt[y].ofs, 0), S.h - 1), y1 = std::min(std::max(yt[last].ofs + 1, 0), S.h - 1);
        if (y1 < y0) return;
        maxRows = std::max(maxRows, y1 - y0 + 1);
    }
    const int pitch = round_up(maxBytes, 16) + 16;
    if ((size_t)pitch * maxRows > 96 * 1024) return;
    D.rsPitch = pitch; D.rsRows = maxRows;
}

// Level sizes, FAST cell grid, slot layout for an image shape.
void build_geometry(obs_extractor* e, int w, int h) {
    Geom& g = e->g;
    memset(&g, 0, sizeof(g));
    const int nl = e->prm.nlevels;
    g.nlevels = nl; g.w = w; g.h = h;
    g.iniTh = e->prm.ini_th_fast; g.minTh = e->prm.min_th_fast;
    memcpy(g.umax, e->umax, sizeof(g.umax));
    unsigned off = 0, slot = 0;
    int cells = 0, xt = 0, yt = 0, tiles = 0, maxFeat = 0, nodeCap = 0, fastCtas = 0, edgeItems = 0, kpWorst = 0;
    for (int l = 0; l < nl; l++) {
        LevelGeom& L = g.lv[l];
        L.w = cv_round_f((float)w * e->invScale[l]);
        L.h = cv_round_f((float)h * e->invScale[l]);
        L.pitch = round_up(std::max(L.w, 1), 128);
        L.off = off;
        off += (unsigned)L.pitch * (unsigned)std::max(L.h, 1);
        off = (off + 255u) & ~255u;
        const int width = L.w - 2 * OBS_BORDER, height = L.h - 2 * OBS_BORDER;
        L.nCols = width > 0 ? (int)((float)width / 30.f) : 0;
        L.nRows = height > 0 ? (int)((float)height / 30.f) : 0;
        if (L.nCols <= 0 || L.nRows <= 0) { L.nCols = L.nRows = 0; L.wCell = L.hCell = 1; }
        else {
            L.wCell = (int)ceilf((float)width / L.nCols);
            L.hCell = (int)ceilf((float)height / L.nRows);
        }
        L.cellBase = cells;
        L.cellCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        L.slotBase = slot;
        cells += L.nCols * L.nRows;
        slot += (unsigned)(L.nCols * L.nRows) * (unsigned)L.cellCap;
        L.fastCellsPerCta = fast_cells_per_cta_host(L.wCell, L.hCell);
        L.fastCtaBase = fastCtas;
        if (l == 0) e->hFastCtas.clear();
        for (int ci = 0; ci < L.nRows; ci++)
            for (int j0 = 0; j0 < L.nCols; j0 += L.fastCellsPerCta) {
                e->hFastCtas.push_back(FastCta{(short)l, (short)ci, (short)j0, (short)std::min(L.fastCellsPerCta, L.nCols - j0)});
                fastCtas++;
            }
        L.nfeat = e->featPerLevel[l];
        L.xtab = xt; L.ytab = yt;
        if (l > 0) { xt += L.w; yt += L.h; }
        L.blurTilesX = (L.w + 127) / 128;
        L.blurTileBase = tiles;
        tiles += L.blurTilesX * ((L.h + 127) / 128);
        {   // border strips: strip 0 and the strips whose 8-byte look-ahead crosses the right border, per 32-row band
            const int nStrips = (L.w + 3) >> 2;
            const int firstRight = std::max(1, (L.w - 8 + 4) >> 2);
            const int perBand = 1 + std::max(0, nStrips - firstRight);
            L.blurEdgeBase = edgeItems;
            edgeItems += perBand * ((L.h + 31) / 32);
        }
        L.scale = e->scale[l];
        L.invScale = e->invScale[l];
        L.patchSize = (float)(int)(31 * e->scale[l]);
        maxFeat = std::max(maxFeat, L.nfeat);
        int nIni = 0;
        if (width > 0 && height > 0) nIni = (int)roundf((float)width / (float)height);
        nodeCap = std::max(nodeCap, std::max(L.nfeat + 4, 4 * nIni + 4));
        // DistributeOctTree leaves at most N + 3 nodes, except that its first sweep always splits every root: very wide
        // levels (nIni roots of one cell height) can end with up to 4 nIni nodes whatever the quota (:606-669)
        kpWorst += std::max(L.nfeat + 3, 4 * nIni);
    }
    g.nCellsTotal = cells;
    g.slotTotal = std::max(slot, 1u);
    g.slabBytes = off;
    g.selCap = round_up(nodeCap, 4);
    g.kpCap = round_up(std::max(e->prm.nfeatures + 4 * nl, kpWorst), 32);
    g.blurTilesTotal = tiles;
    g.blurEdgeCtas = (edgeItems + 127) / 128;
    g.fastCtasTotal = fastCtas;
    g.fastTileRows = 7;
    for (int l = 0; l < nl; l++) if (g.lv[l].nCols > 0) g.fastTileRows = std::max(g.fastTileRows, g.lv[l].hCell + 6);
    e->nodeCap = g.selCap;
    e->hXtab.assign(std::max(xt, 1), ResizeTap{0, 0, 0});
    e->hYtab.assign(std::max(yt, 1), ResizeTap{0, 0, 0});
    for (int l = 1; l < nl; l++) {
        g.lv[l].rsScaleX = 1.0 / ((double)g.lv[l].w / g.lv[l - 1].w);          // as in resize_taps
        g.lv[l].rsScaleY = 1.0 / ((double)g.lv[l].h / g.lv[l - 1].h);
        resize_taps(g.lv[l - 1].w, g.lv[l].w, true, e->hXtab.data() + g.lv[l].xtab);
        resize_taps(g.lv[l - 1].h, g.lv[l].h, false, e->hYtab.data() + g.lv[l].ytab);
        resize_tile_geometry(g.lv[l - 1], g.lv[l], e->hXtab.data() + g.lv[l].xtab, e->hYtab.data() + g.lv[l].ytab);
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_blur_maps(obs_extractor* e, size_t B) {
    // resolved once per process (the initialisation of a function-local static is thread safe: handles are created and reshaped from
    // several host threads)
    static const EncodeTiledFn enc = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeTiledFn)fn;
    }();
    if (!enc) return fail(OBS_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    const Geom& g = e->g;
    for (int l = 0; l < g.nlevels; l++) {
        const LevelGeom& L = g.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)std::max(L.h, 1), (cuuint64_t)B};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)g.slabBytes};       // bytes, dimensions 1 and 2
        const cuuint32_t box[3] = {DESC_BOX_W, DESC_BOX_H, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&e->blurMaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, e->blur.p + L.off, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(OBS_ERR_CUDA, "cuTensorMapEncodeTiled(level %d) failed with %d", l, (int)r);
    }
    CU(e->dMaps.ensure(OBS_MAX_LEVELS));
    CU(cudaMemcpy(e->dMaps.p, e->blurMaps.m, sizeof(e->blurMaps.m), cudaMemcpyHostToDevice));
    e->mapsPtr = e->blur.p;
    e->mapsB = B;
    return OBS_OK;
}

int set_shape(obs_extractor* e, int w, int h, int nimg, cudaStream_t st) {
    const bool newGeom = !e->geomValid || e->g.w != w || e->g.h != h;
    if (!e->geomValid || e->g.w != w || e->g.h != h) {
        build_geometry(e, w, h);
        if (e->g.lv[e->g.nlevels - 1].w < 1 || e->g.lv[e->g.nlevels - 1].h < 1)
            return fail(OBS_ERR_INVALID, "image %dx%d too small for %d levels", w, h, e->g.nlevels);
        if (e->nodeCap > 16383 || quadtree_smem_bytes(e->nodeCap) > 200 * 1024)
            return fail(OBS_ERR_INVALID, "nfeatures too large for the quadtree kernel (node capacity %d)", e->nodeCap);
        for (int l = 0; l < e->g.nlevels; l++)
            if (e->g.lv[l].w > 4095 + 2 * OBS_BORDER || e->g.lv[l].h > 4095 + 2 * OBS_BORDER)
                return fail(OBS_ERR_INVALID, "image %dx%d exceeds the 12-bit key coordinate range", w, h);
        CU(quadtree_prepare(e->nodeCap));
        CU(fast_prepare(e->g.fastTileRows));
        CU(pyramid_prepare(e->g));
        CU(e->dXtab.ensure(e->hXtab.size()));
        CU(e->dYtab.ensure(e->hYtab.size()));
        CU(e->dFastCtas.ensure(std::max<size_t>(e->hFastCtas.size(), 1)));
        // tables are tiny; a synchronous copy keeps the host vectors free to change on the next shape
        CU(cudaStreamSynchronize(st));
        CU(cudaMemcpy(e->dXtab.p, e->hXtab.data(), e->hXtab.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(e->dYtab.p, e->hYtab.data(), e->hYtab.size() * sizeof(ResizeTap), cudaMemcpyHostToDevice));
        if (!e->hFastCtas.empty())
            CU(cudaMemcpy(e->dFastCtas.p, e->hFastCtas.data(), e->hFastCtas.size() * sizeof(FastCta), cudaMemcpyHostToDevice));
        e->geomValid = true;
    }
    const Geom& g = e->g;
    const size_t B = (size_t)nimg;
    e->recordBytes = (size_t)OBS_HDR_INTS * 4 + (size_t)g.kpCap * 60;
    e->recordBytes = (e->recordBytes + 255) & ~(size_t)255;
    CU(e->pyr.ensure(B * g.slabBytes));
    CU(e->blur.ensure(B * g.slabBytes));
    CU(e->cand.ensure(B * g.slotTotal));
    CU(e->keyScratch.ensure(B * g.slotTotal));
    CU(e->nodeScratch.ensure(B * g.slotTotal));
    CU(e->cellCount.ensure(B * std::max(g.nCellsTotal, 1)));
    CU(e->sel.ensure(B * g.nlevels * g.selCap));
    CU(e->selCount.ensure(B * g.nlevels));
    CU(e->records.ensure(B * e->recordBytes));
    if (newGeom || e->mapsPtr != e->blur.p || e->mapsB != e->blur.n / g.slabBytes) {
        const int rc = encode_blur_maps(e, e->blur.n / g.slabBytes);
        if (rc != OBS_OK) return rc;
    }
    return OBS_OK;
}

// Enqueue the whole extraction of `nimg` images on `st`.  Dependencies: pyramid -> {FAST -> quadtree} and
// pyramid -> blur, both -> describe; the blur is forked to the handle's auxiliary stream so that it fills the
// SMs the latency-bound quadtree leaves idle.
int run_pipeline(obs_extractor* e, int nimg, cudaStream_t st, int img0 = 0, bool last = true, bool forkBlur = true) {
    const Geom& g = e->g;
    const size_t o = (size_t)img0;
    PyrPtrs P = e->ptrs;
    P.l0 += o * P.l0ImgStride;
    P.slab += o * P.slabStride;
    uint8_t* blurP = e->blur.p + o * g.slabBytes;
    uint32_t* candP = e->cand.p + o * g.slotTotal;
    int* cellCountP = e->cellCount.p + o * g.nCellsTotal;
    uint32_t* selP = e->sel.p + o * g.nlevels * g.selCap;
    int* selCountP = e->selCount.p + o * g.nlevels;
    cudaEvent_t* ev = nullptr;
    if (e->prof && e->profCalls < PROF_SLOTS && img0 == 0 && last) ev = e->pev.data() + (size_t)e->profCalls * (2 * OBS_NUM_STAGES);
    StageRange whole(e->prof, "obs:extract");
    if (ev) CU(cudaEventRecord(ev[0], st));
    { StageRange r(e->prof, "obs:ComputePyramid"); CU(launch_pyramid(g, P, e->dXtab.p, e->dYtab.p, nimg, st)); }
    if (ev) CU(cudaEventRecord(ev[1], st));
    cudaStream_t bs = forkBlur ? e->aux : st;
    if (forkBlur) {
        CU(cudaEventRecord(e->fork, st));
        CU(cudaStreamWaitEvent(e->aux, e->fork, 0));
    }
    if (ev) CU(cudaEventRecord(ev[6], bs));
    { StageRange r(e->prof, "obs:GaussianBlur"); CU(launch_blur(g, P, blurP, g.slabBytes, nimg, bs)); }
    if (ev) CU(cudaEventRecord(ev[7], bs));
    if (forkBlur) CU(cudaEventRecord(e->join, e->aux));
    if (ev) CU(cudaEventRecord(ev[2], st));
    { StageRange r(e->prof, "obs:FAST"); CU(launch_fast(g, P, e->dFastCtas.p, candP, cellCountP, nimg, st)); }
    if (ev) { CU(cudaEventRecord(ev[3], st)); CU(cudaEventRecord(ev[4], st)); }
    { StageRange r(e->prof, "obs:DistributeOctTree");
      CU(launch_quadtree(g, e->nodeCap, candP, cellCountP, e->keyScratch.p + o * g.slotTotal, e->nodeScratch.p + o * g.slotTotal, selP, selCountP, nimg, st)); }
    if (ev) CU(cudaEventRecord(ev[5], st));
    if (forkBlur) CU(cudaStreamWaitEvent(st, e->join, 0));
    if (ev) CU(cudaEventRecord(ev[8], st));
    { StageRange r(e->prof, "obs:IC_Angle+computeOrbDescriptor");
      CU(launch_describe(g, P, e->dMaps.p, img0, selP, selCountP, e->records.p + o * e->recordBytes, e->recordBytes, nimg, st)); }
    if (ev) { CU(cudaEventRecord(ev[9], st)); e->profCalls++; }
    e->lastN = img0 + nimg;
    e->lastStream = st;
    return OBS_OK;
}

// Host path of a batch whose images are one page-locked block and whose outputs are page-locked too: upload of chunk c+1
// (copy-in stream), extraction of chunk c and download of chunk c-1 (copy-out stream) overlap; no staging.  Consecutive chunks
// alternate between the handle's two compute streams, so one chunk's latency-bound quadtree runs beside the next chunk's
// FAST / blur instead of serialising the pipeline.  Nothing here waits for the device: the caller synchronises on e->cout.
// counts: page-locked ints, one per image.
int enqueue_chunks(obs_extractor* e, const uint8_t* images0, int n_images, int w, int h, size_t stride,
                   obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* counts) {
    const Geom& g = e->g;
    cudaStream_t st = e->stream;
    const size_t p0 = (size_t)g.lv[0].pitch;
    const size_t imgBytes = stride * (size_t)h;
    e->ptrs.l0 = e->pyr.p;
    e->ptrs.l0ImgStride = g.slabBytes;
    e->ptrs.l0Pitch = g.lv[0].pitch;
    e->ptrs.slab = e->pyr.p;
    e->ptrs.slabStride = g.slabBytes;
    const int nChunks = n_images >= 32 ? 4 : n_images >= 8 ? 2 : 1;
    while ((int)e->chunkIn.size() < nChunks) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        e->chunkIn.push_back(a); e->chunkDone.push_back(b);
    }
    CU(e->rawIn.ensure((size_t)e->maxBatch * imgBytes + 16));
    const int m = cap < g.kpCap ? cap : g.kpCap;
    // the previous call's device work (a stereo match may still read the pyramids) precedes the first upload
    CU(cudaEventRecord(e->done, st));
    CU(cudaStreamWaitEvent(e->cin, e->done, 0));
    CU(cudaStreamWaitEvent(e->aux, e->done, 0));
    for (int c = 0; c < nChunks; c++) {
        const int c0 = (int)((long long)n_images * c / nChunks), c1 = (int)((long long)n_images * (c + 1) / nChunks);
        if (c1 <= c0) continue;
        CU(cudaMemcpyAsync(e->rawIn.p + (size_t)c0 * imgBytes, images0 + (size_t)c0 * imgBytes, (size_t)(c1 - c0) * imgBytes, cudaMemcpyHostToDevice, e->cin));
        CU(cudaEventRecord(e->chunkIn[c], e->cin));
        cudaStream_t cs = (c & 1) ? e->aux : st;
        CU(cudaStreamWaitEvent(cs, e->chunkIn[c], 0));
        CU(launch_repack(e->rawIn.p + (size_t)c0 * imgBytes, imgBytes, stride, e->pyr.p + (size_t)c0 * g.slabBytes, g.slabBytes, (int)p0, w, h, c1 - c0, cs));
        // a single chunk (a live frame) leaves the auxiliary stream free: the blur runs there beside FAST + quadtree
        int rc = run_pipeline(e, c1 - c0, cs, c0, false, nChunks == 1);
        if (rc) return rc;
        CU(cudaEventRecord(e->chunkDone[c], cs));
        CU(cudaStreamWaitEvent(e->cout, e->chunkDone[c], 0));
        const uint8_t* rec = e->records.p + (size_t)c0 * e->recordBytes;
        CU(cudaMemcpy2DAsync(counts + c0, sizeof(int), rec, e->recordBytes, sizeof(int), c1 - c0, cudaMemcpyDeviceToHost, e->cout));
        CU(cudaMemcpy2DAsync(keypoints + (size_t)c0 * cap, (size_t)cap * 28, rec + OBS_HDR_INTS * 4, e->recordBytes, (size_t)m * 28, c1 - c0, cudaMemcpyDeviceToHost, e->cout));
        CU(cudaMemcpy2DAsync(descriptors + (size_t)c0 * cap * 32, (size_t)cap * 32, rec + OBS_HDR_INTS * 4 + (size_t)g.kpCap * 28, e->recordBytes, (size_t)m * 32, c1 - c0, cudaMemcpyDeviceToHost, e->cout));
    }
    // later consumers (stereo match, frame-set build) order themselves after the handle stream
    CU(cudaEventRecord(e->join, e->aux));
    CU(cudaStreamWaitEvent(st, e->join, 0));
    e->lastN = n_images;
    e->lastStream = st;
    return OBS_OK;
}

// result records staged in page-locked memory -> the caller's arrays
int unpack_records(const obs_extractor* e, int n, obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
    int status = OBS_OK;
    for (int i = 0; i < n; i++) {
        const uint8_t* rec = e->stageOut.p + (size_t)i * e->recordBytes;
        const int cnt = *reinterpret_cast<const int*>(rec);
        n_out[i] = cnt;
        const int mm = cnt < cap ? cnt : cap;
        if (cnt > cap && (keypoints || descriptors)) status = OBS_ERR_CAPACITY;
        if (keypoints) memcpy(keypoints + (size_t)i * cap, rec + OBS_HDR_INTS * 4, (size_t)mm * 28);
        if (descriptors) memcpy(descriptors + (size_t)i * cap * 32, rec + OBS_HDR_INTS * 4 + (size_t)e->g.kpCap * 28, (size_t)mm * 32);
    }
    return status;
}

int check_handle(const obs_extractor* e) {
    if (!e) return fail(OBS_ERR_INVALID, "null extractor handle");
    cudaError_t ce = cudaSetDevice(e->device);
    if (ce != cudaSuccess) return fail(OBS_ERR_CUDA, "cudaSetDevice(%d): %s", e->device, cudaGetErrorString(ce));
    return OBS_OK;
}

}  // namespace

extern "C" {

const char* obs_last_error(void) { return obsdetail::g_err; }
const char* obs_version(void) { return "obslam_b200 0.1 sm_100a"; }

int obs_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return -fail(OBS_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

int obs_extractor_create(const obs_orb_params* params, int max_w, int max_h, int max_batch, int device, obs_extractor** out) {
    if (!params || !out) return fail(OBS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (params->nlevels < 1 || params->nlevels > OBS_MAX_LEVELS) return fail(OBS_ERR_INVALID, "nlevels must be in [1,%d]", OBS_MAX_LEVELS);
    if (params->nfeatures < 1 || params->nfeatures > 60000) return fail(OBS_ERR_INVALID, "nfeatures out of range");
    if (!(params->scale_factor > 1.0f) || params->scale_factor > 4.0f) return fail(OBS_ERR_INVALID, "scale_factor must be in (1,4]");
    if (params->ini_th_fast < 0 || params->ini_th_fast > 255 || params->min_th_fast < 0 || params->min_th_fast > 255)
        return fail(OBS_ERR_INVALID, "FAST thresholds must be in [0,255]");
    if (max_w < 1 || max_h < 1 || max_batch < 1) return fail(OBS_ERR_INVALID, "max_w, max_h, max_batch must be positive");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(OBS_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (device < 0 || device >= ndev) return fail(OBS_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(OBS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);

    obs_extractor* e = new (std::nothrow) obs_extractor;
    if (!e) return fail(OBS_ERR_INVALID, "out of host memory");
    e->prm = *params;
    e->device = device;
    e->maxW = max_w; e->maxH = max_h; e->maxBatch = max_batch;
    build_tables(e);
    cudaError_t se = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&e->aux, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->done, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->fork, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->join, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->hostDone, cudaEventDisableTiming | cudaEventBlockingSync);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&e->hostDoneSpin, cudaEventDisableTiming);
    for (cudaEvent_t& ev : e->gjoin) if (se == cudaSuccess) se = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&e->cin, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaStreamCreateWithFlags(&e->cout, cudaStreamNonBlocking);
    if (se != cudaSuccess) { delete e; return fail(OBS_ERR_CUDA, "stream/event creation: %s", cudaGetErrorString(se)); }
    int rc = set_shape(e, max_w, max_h, max_batch, e->stream);
    if (rc != OBS_OK) { obs_extractor_destroy(e); return rc; }
    *out = e;
    return OBS_OK;
}

int obs_extractor_destroy(obs_extractor* e) {
    if (!e) return OBS_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->pyr.release(); e->blur.release(); e->dMaps.release(); e->records.release(); e->rawIn.release(); e->cand.release(); e->keyScratch.release();
    e->sel.release(); e->nodeScratch.release(); e->cellCount.release(); e->selCount.release();
    e->uRight.release(); e->depth.release(); e->sad.release(); e->rowStart.release(); e->rowIdx.release(); e->dXtab.release(); e->dYtab.release(); e->dFastCtas.release();
    e->stageIn.release(); e->stageOut.release();
    for (cudaEvent_t ev : e->pev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->sev) cudaEventDestroy(ev);
    if (e->done) cudaEventDestroy(e->done);
    if (e->fork) cudaEventDestroy(e->fork);
    if (e->join) cudaEventDestroy(e->join);
    if (e->hostDone) cudaEventDestroy(e->hostDone);
    if (e->hostDoneSpin) cudaEventDestroy(e->hostDoneSpin);
    for (cudaEvent_t ev : e->gjoin) if (ev) cudaEventDestroy(ev);
    for (auto& sg : e->graphs) if (sg.exec) cudaGraphExecDestroy(sg.exec);
    if (e->monoExec) cudaGraphExecDestroy(e->monoExec);
    for (cudaEvent_t ev : e->chunkIn) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->chunkDone) cudaEventDestroy(ev);
    if (e->cin) cudaStreamDestroy(e->cin);
    if (e->cout) cudaStreamDestroy(e->cout);
    if (e->aux) cudaStreamDestroy(e->aux);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return OBS_OK;
}

int obs_extractor_levels(const obs_extractor* e) { return e ? e->prm.nlevels : -1; }
int obs_extractor_max_keypoints(const obs_extractor* e) { return e ? e->g.kpCap : -1; }

int obs_extractor_tables(const obs_extractor* e, float* sc, float* isc, float* s2, float* is2, int32_t* fpl) {
    if (!e) return fail(OBS_ERR_INVALID, "null extractor handle");
    for (int i = 0; i < e->prm.nlevels; i++) {
        if (sc) sc[i] = e->scale[i];
        if (isc) isc[i] = e->invScale[i];
        if (s2) s2[i] = e->sigma2[i];
        if (is2) is2[i] = e->invSigma2[i];
        if (fpl) fpl[i] = e->featPerLevel[i];
    }
    return OBS_OK;
}

int obs_extractor_set_profiling(obs_extractor* e, int on) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (on && e->pev.empty()) {
        e->pev.resize((size_t)PROF_SLOTS * (2 * OBS_NUM_STAGES));
        for (cudaEvent_t& ev : e->pev) CU(cudaEventCreate(&ev));
        e->sev.resize((size_t)PROF_SLOTS * 2);
        for (cudaEvent_t& ev : e->sev) CU(cudaEventCreate(&ev));
    }
    e->prof = on != 0;
    e->profCalls = 0;
    e->stereoCalls = 0;
    return OBS_OK;
}

int obs_extractor_stage_ms(obs_extractor* e, float* stage_ms, float* stereo_ms, int* n_calls, int* n_stereo_calls) {
    int rc = check_handle(e);
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    for (int s = 0; s < OBS_NUM_STAGES; s++) {
        float sum = 0.f;
        for (int c = 0; c < e->profCalls; c++) {
            float ms = 0.f;
            const cudaEvent_t* ev = e->pev.data() + (size_t)c * (2 * OBS_NUM_STAGES);
            CU(cudaEventElapsedTime(&ms, ev[2 * s], ev[2 * s + 1]));
            sum += ms;
        }
        if (stage_ms) stage_ms[s] = sum;
    }
    float ssum = 0.f;
    for (int c = 0; c < e->stereoCalls; c++) {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, e->sev[2 * c], e->sev[2 * c + 1]));
        ssum += ms;
    }
    if (stereo_ms) *stereo_ms = ssum;
    if (n_calls) *n_calls = e->profCalls;
    if (n_stereo_calls) *n_stereo_calls = e->stereoCalls;
    return OBS_OK;
}

int obs_extract_batch_device(obs_extractor* e, const uint8_t* d_images, int n_images, int w, int h,
                             size_t stride, size_t image_stride, void* stream) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!d_images || n_images < 1) return fail(OBS_ERR_INVALID, "no images");
    if (n_images > e->maxBatch) return fail(OBS_ERR_CAPACITY, "batch of %d exceeds max_batch %d", n_images, e->maxBatch);
    if (w < 1 || h < 1 || w > e->maxW || h > e->maxH) return fail(OBS_ERR_INVALID, "image %dx%d outside [1,%d]x[1,%d]", w, h, e->maxW, e->maxH);
    if (((uintptr_t)d_images & 15) || (stride & 15) || (image_stride & 15) || stride < (size_t)w)
        return fail(OBS_ERR_INVALID, "device images need a 16-byte aligned base, stride and image stride");
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    rc = set_shape(e, w, h, n_images, st);
    if (rc) return rc;
    e->ptrs.l0 = d_images;
    e->ptrs.l0ImgStride = image_stride;
    e->ptrs.l0Pitch = (int)stride;
    e->ptrs.slab = e->pyr.p;
    e->ptrs.slabStride = e->g.slabBytes;
    return run_pipeline(e, n_images, st, 0, true, !e->prof);      // profiling: stages serialised, so their brackets do not overlap
}

int obs_extract_batch(obs_extractor* e, const uint8_t* const* images, int n_images, int w, int h,
                      size_t stride, obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!images || !n_out || n_images < 1) return fail(OBS_ERR_INVALID, "null argument");
    if (n_images > e->maxBatch) return fail(OBS_ERR_CAPACITY, "batch of %d exceeds max_batch %d", n_images, e->maxBatch);
    if (w == 0 || h == 0) {                      // empty image: the reference returns without touching its outputs (:1046)
        for (int i = 0; i < n_images; i++) n_out[i] = 0;
        e->lastN = 0;
        return OBS_OK;
    }
    if (w < 1 || h < 1 || w > e->maxW || h > e->maxH) return fail(OBS_ERR_INVALID, "image %dx%d outside [1,%d]x[1,%d]", w, h, e->maxW, e->maxH);
    if (stride < (size_t)w) return fail(OBS_ERR_INVALID, "stride smaller than width");
    cudaStream_t st = e->stream;
    rc = set_shape(e, w, h, n_images, st);
    if (rc) return rc;
    const Geom& g = e->g;
    const size_t p0 = (size_t)g.lv[0].pitch, l0Bytes = p0 * h;
    for (int i = 0; i < n_images; i++)
        if (!images[i]) return fail(OBS_ERR_INVALID, "image %d is null", i);
    // page-locked images that follow each other in memory: one DMA for the whole batch, repacked to the
    // pitched level-0 slab on the device
    bool oneBlock = is_pinned(images[0]);
    const size_t imgBytes = stride * (size_t)h;
    for (int i = 1; oneBlock && i < n_images; i++) oneBlock = images[i] == images[i - 1] + imgBytes;
    e->ptrs.l0 = e->pyr.p;
    e->ptrs.l0ImgStride = g.slabBytes;
    e->ptrs.l0Pitch = g.lv[0].pitch;
    e->ptrs.slab = e->pyr.p;
    e->ptrs.slabStride = g.slabBytes;
    const bool pinnedOut = keypoints && descriptors && is_pinned(keypoints) && is_pinned(descriptors) && cap > 0;
    // One image in pageable memory, results to pageable memory -- what the C++ drop-in's operator() does for every frame: the image
    // is staged in the handle's own page-locked buffer, so upload + 9 kernels + download depend on nothing but the shape and are
    // replayed as one CUDA graph from the third call on (captured on the second).
    if (n_images == 1 && cap >= 0 && graphs_enabled() && !e->prof && !oneBlock &&
        !(keypoints && is_pinned(keypoints)) && !(descriptors && is_pinned(descriptors))) {
        CU(e->stageIn.ensure((size_t)e->maxBatch * l0Bytes));
        CU(e->stageOut.ensure((size_t)e->maxBatch * e->recordBytes));
        for (int y = 0; y < h; y++) memcpy(e->stageIn.p + (size_t)y * p0, images[0] + (size_t)y * stride, w);
        const unsigned long long epoch = g_allocEpoch.load(std::memory_order_relaxed);
        if (e->monoExec && (e->monoW != w || e->monoH != h || e->monoEpoch != epoch)) {
            cudaGraphExecDestroy(e->monoExec);
            e->monoExec = nullptr; e->monoSeen = false;
        }
        if (e->monoExec) {
            CU(cudaGraphLaunch(e->monoExec, st));
            e->lastN = 1; e->lastStream = st;
        } else {
            const bool capture = e->monoSeen && e->monoW == w && e->monoH == h && e->monoEpoch == epoch;
            if (capture) { CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed)); obsdetail::set_capturing(true); }
            cudaError_t ce = cudaMemcpyAsync(e->pyr.p, e->stageIn.p, l0Bytes, cudaMemcpyHostToDevice, st);
            rc = ce == cudaSuccess ? run_pipeline(e, 1, st, 0, true, true) : OBS_OK;
            if (ce == cudaSuccess && !rc) ce = cudaMemcpyAsync(e->stageOut.p, e->records.p, e->recordBytes, cudaMemcpyDeviceToHost, st);
            if (capture) {
                cudaGraph_t graph = nullptr;
                const cudaError_t ee = cudaStreamEndCapture(st, &graph);
                obsdetail::set_capturing(false);
                if (rc || ce != cudaSuccess || ee != cudaSuccess || !graph) {
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    if (rc) return rc;
                    return fail(OBS_ERR_CUDA, "obs_extract: graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee));
                }
                const cudaError_t ie = cudaGraphInstantiate(&e->monoExec, graph, 0);
                cudaGraphDestroy(graph);
                if (ie != cudaSuccess) { e->monoExec = nullptr; return fail(OBS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie)); }
                CU(cudaGraphLaunch(e->monoExec, st));
            } else {
                if (rc) return rc;
                CU(ce);
                e->monoSeen = true; e->monoW = w; e->monoH = h; e->monoEpoch = epoch;
            }
        }
        CU(cudaStreamSynchronize(st));
        const int status = unpack_records(e, 1, keypoints, descriptors, cap, n_out);
        if (status) return fail(status, "caller capacity %d smaller than the keypoint count", cap);
        return OBS_OK;
    }
    if (oneBlock && pinnedOut && n_images >= 8) {
        CU(e->stageOut.ensure((size_t)e->maxBatch * sizeof(int)));
        int* cnts = reinterpret_cast<int*>(e->stageOut.p);
        rc = enqueue_chunks(e, images[0], n_images, w, h, stride, keypoints, descriptors, cap, cnts);
        if (rc) return rc;
        // sleep until the last download has landed (blocking-sync event: no spinning host thread per handle)
        CU(cudaEventRecord(e->hostDone, e->cout));
        CU(cudaEventSynchronize(e->hostDone));
        int status = OBS_OK;
        for (int i = 0; i < n_images; i++) {
            n_out[i] = cnts[i];
            if (cnts[i] > cap) status = OBS_ERR_CAPACITY;
        }
        if (status) return fail(status, "caller capacity %d smaller than the keypoint count", cap);
        return OBS_OK;
    }
    if (oneBlock) {
        CU(e->rawIn.ensure((size_t)e->maxBatch * imgBytes + 16));
        CU(cudaMemcpyAsync(e->rawIn.p, images[0], (size_t)n_images * imgBytes, cudaMemcpyHostToDevice, st));
        CU(launch_repack(e->rawIn.p, imgBytes, stride, e->pyr.p, g.slabBytes, (int)p0, w, h, n_images, st));
    } else {
        for (int i = 0; i < n_images; i++) {
            if (is_pinned(images[i])) {           // page-locked caller memory: DMA straight into the pitched level-0 slab
                CU(cudaMemcpy2DAsync(e->pyr.p + (size_t)i * g.slabBytes, p0, images[i], stride, w, h, cudaMemcpyHostToDevice, st));
            } else {
                CU(e->stageIn.ensure((size_t)e->maxBatch * l0Bytes));
                uint8_t* dstp = e->stageIn.p + (size_t)i * l0Bytes;
                for (int y = 0; y < h; y++) memcpy(dstp + (size_t)y * p0, images[i] + (size_t)y * stride, w);
                CU(cudaMemcpyAsync(e->pyr.p + (size_t)i * g.slabBytes, dstp, l0Bytes, cudaMemcpyHostToDevice, st));
            }
        }
    }
    e->ptrs.l0 = e->pyr.p;
    e->ptrs.l0ImgStride = g.slabBytes;
    e->ptrs.l0Pitch = g.lv[0].pitch;
    e->ptrs.slab = e->pyr.p;
    e->ptrs.slabStride = g.slabBytes;
    rc = run_pipeline(e, n_images, st, 0, true, !e->prof);
    if (rc) return rc;
    return obs_extractor_fetch(e, keypoints, descriptors, cap, n_out);
}

int obs_extract(obs_extractor* e, const uint8_t* image, int w, int h, size_t stride,
                obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
    if (!n_out) return fail(OBS_ERR_INVALID, "null argument");
    if (!image || w == 0 || h == 0) { *n_out = 0; if (e) e->lastN = 0; return OBS_OK; }
    const uint8_t* one[1] = {image};
    return obs_extract_batch(e, one, 1, w, h, stride, keypoints, descriptors, cap, n_out);
}

int obs_extractor_fetch(obs_extractor* e, obs_keypoint* keypoints, uint8_t* descriptors, int cap, int* n_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction to fetch");
    if (!n_out || cap < 0) return fail(OBS_ERR_INVALID, "null argument");
    const int n = e->lastN;
    const Geom& g = e->g;
    const uint8_t* kpSrc = e->records.p + OBS_HDR_INTS * 4;
    const uint8_t* descSrc = kpSrc + (size_t)g.kpCap * 28;
    const int m = cap < g.kpCap ? cap : g.kpCap;
    int status = OBS_OK;
    if ((!keypoints || is_pinned(keypoints)) && (!descriptors || is_pinned(descriptors)) && m > 0) {
        // page-locked outputs: strided DMA from the result records, no staging
        CU(e->stageOut.ensure((size_t)e->maxBatch * sizeof(int)));
        int* cnts = reinterpret_cast<int*>(e->stageOut.p);
        CU(cudaMemcpy2DAsync(cnts, sizeof(int), e->records.p, e->recordBytes, sizeof(int), n, cudaMemcpyDeviceToHost, e->lastStream));
        if (keypoints) CU(cudaMemcpy2DAsync(keypoints, (size_t)cap * 28, kpSrc, e->recordBytes, (size_t)m * 28, n, cudaMemcpyDeviceToHost, e->lastStream));
        if (descriptors) CU(cudaMemcpy2DAsync(descriptors, (size_t)cap * 32, descSrc, e->recordBytes, (size_t)m * 32, n, cudaMemcpyDeviceToHost, e->lastStream));
        CU(cudaStreamSynchronize(e->lastStream));
        for (int i = 0; i < n; i++) {
            n_out[i] = cnts[i];
            if (cnts[i] > cap && (keypoints || descriptors)) status = OBS_ERR_CAPACITY;
        }
    } else {
        CU(e->stageOut.ensure((size_t)e->maxBatch * e->recordBytes));
        CU(cudaMemcpyAsync(e->stageOut.p, e->records.p, (size_t)n * e->recordBytes, cudaMemcpyDeviceToHost, e->lastStream));
        CU(cudaStreamSynchronize(e->lastStream));
        status = unpack_records(e, n, keypoints, descriptors, cap, n_out);
    }
    if (status) return fail(status, "caller capacity %d smaller than the keypoint count", cap);
    return OBS_OK;
}

int obs_extractor_fetch_counts(obs_extractor* e, int* n_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction to fetch");
    if (!n_out) return fail(OBS_ERR_INVALID, "null argument");
    CU(e->stageOut.ensure((size_t)e->lastN * e->recordBytes));
    int* tmp = reinterpret_cast<int*>(e->stageOut.p);
    CU(cudaMemcpy2DAsync(tmp, sizeof(int), e->records.p, e->recordBytes, sizeof(int), e->lastN, cudaMemcpyDeviceToHost, e->lastStream));
    CU(cudaStreamSynchronize(e->lastStream));
    for (int i = 0; i < e->lastN; i++) n_out[i] = tmp[i];
    return OBS_OK;
}

int obs_extractor_results_device(obs_extractor* e, const void** d_records, size_t* record_bytes, int* cap) {
    if (!e) return fail(OBS_ERR_INVALID, "null extractor handle");
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction yet");
    if (d_records) *d_records = e->records.p;
    if (record_bytes) *record_bytes = e->recordBytes;
    if (cap) *cap = e->g.kpCap;
    return OBS_OK;
}

int obs_extractor_get_level(obs_extractor* e, int image_index, int level, int which, uint8_t* dst, size_t dst_stride, int* w, int* h) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction yet");
    if (image_index < 0 || image_index >= e->lastN || level < 0 || level >= e->g.nlevels) return fail(OBS_ERR_INVALID, "index out of range");
    const LevelGeom& L = e->g.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (!dst) return OBS_OK;
    if (dst_stride < (size_t)L.w) return fail(OBS_ERR_INVALID, "dst_stride smaller than the level width");
    const uint8_t* src;
    size_t pitch;
    if (which == 1) { src = e->blur.p + (size_t)image_index * e->g.slabBytes + L.off; pitch = L.pitch; }
    else if (level == 0) { src = e->ptrs.l0 + (size_t)image_index * e->ptrs.l0ImgStride; pitch = e->ptrs.l0Pitch; }
    else { src = e->pyr.p + (size_t)image_index * e->g.slabBytes + L.off; pitch = L.pitch; }
    CU(cudaStreamSynchronize(e->lastStream));
    CU(cudaMemcpy2D(dst, dst_stride, src, pitch, L.w, L.h, cudaMemcpyDeviceToHost));
    return OBS_OK;
}

static int fetch_keys(obs_extractor* e, int image_index, int level, bool selected, int32_t* xyr, int cap, int* n_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction yet");
    if (!n_out) return fail(OBS_ERR_INVALID, "null argument");
    if (image_index < 0 || image_index >= e->lastN || level < 0 || level >= e->g.nlevels) return fail(OBS_ERR_INVALID, "index out of range");
    const Geom& g = e->g;
    const LevelGeom& L = g.lv[level];
    CU(cudaStreamSynchronize(e->lastStream));
    std::vector<uint32_t> keys;
    if (selected) {
        int cnt = 0;
        CU(cudaMemcpy(&cnt, e->selCount.p + (size_t)image_index * g.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
        keys.resize(cnt);
        if (cnt) CU(cudaMemcpy(keys.data(), e->sel.p + ((size_t)image_index * g.nlevels + level) * g.selCap, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    } else {
        const int nCells = L.nCols * L.nRows;
        std::vector<int> cc(std::max(nCells, 1));
        if (nCells) CU(cudaMemcpy(cc.data(), e->cellCount.p + (size_t)image_index * g.nCellsTotal + L.cellBase, (size_t)nCells * 4, cudaMemcpyDeviceToHost));
        std::vector<uint32_t> slots((size_t)nCells * L.cellCap);
        if (nCells) CU(cudaMemcpy(slots.data(), e->cand.p + (size_t)image_index * g.slotTotal + L.slotBase, slots.size() * 4, cudaMemcpyDeviceToHost));
        for (int c = 0; c < nCells; c++)
            for (int i = 0; i < cc[c]; i++) keys.push_back(slots[(size_t)c * L.cellCap + i]);
    }
    *n_out = (int)keys.size();
    if (xyr) {
        const int m = std::min((int)keys.size(), cap);
        for (int i = 0; i < m; i++) {
            xyr[3 * i] = (int)(keys[i] & 0xfffu);
            xyr[3 * i + 1] = (int)((keys[i] >> 12) & 0xfffu);
            xyr[3 * i + 2] = (int)(keys[i] >> 24);
        }
    }
    return OBS_OK;
}

int obs_extractor_get_candidates(obs_extractor* e, int image_index, int level, int32_t* xyr, int cap, int* n_out) {
    return fetch_keys(e, image_index, level, false, xyr, cap, n_out);
}
int obs_extractor_get_selected(obs_extractor* e, int image_index, int level, int32_t* xyr, int cap, int* n_out) {
    return fetch_keys(e, image_index, level, true, xyr, cap, n_out);
}

int obs_stereo_match_device(obs_extractor* L, obs_extractor* R, float mbf, float min_d, float max_d,
                            void* stream, const float** d_u_right, const float** d_depth) {
    int rc = check_handle(L);
    if (rc) return rc;
    if (!R) return fail(OBS_ERR_INVALID, "null right extractor");
    if (L->lastN < 1 || R->lastN < 1) return fail(OBS_ERR_STATE, "stereo matching needs an extraction on both handles first");
    if (L->device != R->device) return fail(OBS_ERR_INVALID, "both eyes must live on one device");
    if (L->lastN != R->lastN || L->g.w != R->g.w || L->g.h != R->g.h || L->g.nlevels != R->g.nlevels ||
        L->g.kpCap != R->g.kpCap || L->prm.scale_factor != R->prm.scale_factor)
        return fail(OBS_ERR_INVALID, "left and right extractions differ in shape, batch or parameters");
    cudaStream_t st = stream ? (cudaStream_t)stream : L->stream;
    const int n = L->lastN;
    const size_t cnt = (size_t)n * L->g.kpCap;
    CU(L->uRight.ensure(cnt));
    CU(L->depth.ensure(cnt));
    CU(L->sad.ensure(cnt));
    // row table: a right keypoint is listed under ceil(y+r) - floor(y-r) + 1 <= 4*scale + 3 rows
    const int bandMax = (int)(4.0f * L->scale[L->prm.nlevels - 1]) + 4;
    const int rowIdxCap = L->g.kpCap * bandMax;
    CU(L->rowStart.ensure((size_t)n * (L->g.h + 1)));
    CU(L->rowIdx.ensure((size_t)n * rowIdxCap));
    // order after both extractions
    if (R->lastStream != st) { CU(cudaEventRecord(R->done, R->lastStream)); CU(cudaStreamWaitEvent(st, R->done, 0)); }
    if (L->lastStream != st) { CU(cudaEventRecord(L->done, L->lastStream)); CU(cudaStreamWaitEvent(st, L->done, 0)); }
    StereoArgs a;
    a.g = L->g;
    a.left = L->ptrs; a.right = R->ptrs;
    a.recL = L->records.p; a.recR = R->records.p; a.recordBytes = L->recordBytes;
    a.mbf = mbf; a.minD = min_d; a.maxD = max_d;
    a.uRight = L->uRight.p; a.depth = L->depth.p; a.sad = L->sad.p;
    a.rowStart = L->rowStart.p; a.rowIdx = L->rowIdx.p; a.rowIdxCap = rowIdxCap;
    cudaEvent_t* sev = nullptr;
    if (L->prof && L->stereoCalls < PROF_SLOTS) sev = L->sev.data() + (size_t)L->stereoCalls * 2;
    if (sev) CU(cudaEventRecord(sev[0], st));
    { StageRange r(L->prof, "obs:ComputeStereoMatches"); CU(launch_stereo(a, n, st)); }
    if (sev) { CU(cudaEventRecord(sev[1], st)); L->stereoCalls++; }
    // later work on either handle must not overwrite the inputs while the match runs
    CU(cudaEventRecord(L->done, st));
    if (R->stream != st) CU(cudaStreamWaitEvent(R->stream, L->done, 0));
    if (R->lastStream != st && R->lastStream != R->stream) CU(cudaStreamWaitEvent(R->lastStream, L->done, 0));
    if (L->stream != st) CU(cudaStreamWaitEvent(L->stream, L->done, 0));
    if (L->lastStream != st && L->lastStream != L->stream) CU(cudaStreamWaitEvent(L->lastStream, L->done, 0));
    L->lastStream = st;
    if (d_u_right) *d_u_right = L->uRight.p;
    if (d_depth) *d_depth = L->depth.p;
    return OBS_OK;
}

int obs_stereo_match(obs_extractor* L, obs_extractor* R, float mbf, float min_d, float max_d,
                     float* u_right, float* depth, int cap) {
    if (!u_right || !depth) return fail(OBS_ERR_INVALID, "null output");
    int rc = obs_stereo_match_device(L, R, mbf, min_d, max_d, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const int n = L->lastN, kc = L->g.kpCap;
    const int m = std::min(cap, kc);
    if (is_pinned(u_right) && is_pinned(depth) && cap <= kc) {
        CU(cudaMemcpy2DAsync(u_right, (size_t)cap * 4, L->uRight.p, (size_t)kc * 4, (size_t)m * 4, n, cudaMemcpyDeviceToHost, L->stream));
        CU(cudaMemcpy2DAsync(depth, (size_t)cap * 4, L->depth.p, (size_t)kc * 4, (size_t)m * 4, n, cudaMemcpyDeviceToHost, L->stream));
        CU(cudaStreamSynchronize(L->stream));
        return OBS_OK;
    }
    std::vector<float> hu((size_t)n * kc), hd((size_t)n * kc);
    CU(cudaMemcpyAsync(hu.data(), L->uRight.p, hu.size() * 4, cudaMemcpyDeviceToHost, L->stream));
    CU(cudaMemcpyAsync(hd.data(), L->depth.p, hd.size() * 4, cudaMemcpyDeviceToHost, L->stream));
    CU(cudaStreamSynchronize(L->stream));
    for (int i = 0; i < n; i++) {
        for (int j = 0; j < cap; j++) { u_right[(size_t)i * cap + j] = -1.f; depth[(size_t)i * cap + j] = -1.f; }
        memcpy(u_right + (size_t)i * cap, hu.data() + (size_t)i * kc, (size_t)m * 4);
        memcpy(depth + (size_t)i * cap, hd.data() + (size_t)i * kc, (size_t)m * 4);
    }
    return OBS_OK;
}

int obs_frame_set_from_extractor(obs_frame_set* fs, obs_extractor* e, const float* d_u_right) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!fs) return fail(OBS_ERR_INVALID, "null frame set");
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction yet");
    if (obs_frame_set_capacity(fs) < e->g.kpCap)
        return fail(OBS_ERR_CAPACITY, "frame set keypoint capacity %d < extractor capacity %d", obs_frame_set_capacity(fs), e->g.kpCap);
    if (d_u_right && !is_device(d_u_right)) return fail(OBS_ERR_INVALID, "d_u_right must be device memory");
    if (obs_frame_set_device(fs) != e->device)
        return fail(OBS_ERR_INVALID, "the frame set lives on device %d, the extractor on device %d", obs_frame_set_device(fs), e->device);
    const uint8_t* rec = e->records.p;
    return obs_frame_set_build_device(fs, rec + OBS_HDR_INTS * 4, e->recordBytes, rec + OBS_HDR_INTS * 4 + (size_t)e->g.kpCap * 28,
                                      e->recordBytes, d_u_right, (size_t)e->g.kpCap, reinterpret_cast<const int*>(rec),
                                      e->recordBytes / 4, e->lastN, e->lastStream);
}

int obs_gray_from_color(obs_extractor* e, const uint8_t* d_src, int n_images, int w, int h, int channels, int rgb_order,
                        size_t src_stride, size_t src_image_stride, uint8_t* d_gray, size_t gray_stride, size_t gray_image_stride, void* stream) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!d_src || !d_gray || n_images < 1 || w < 1 || h < 1) return fail(OBS_ERR_INVALID, "bad argument");
    if (channels != 3 && channels != 4) return fail(OBS_ERR_INVALID, "channels must be 3 or 4 (the reference converts 3- and 4-channel images, Tracking.cc:202-227)");
    if (src_stride < (size_t)w * channels || gray_stride < (size_t)w) return fail(OBS_ERR_INVALID, "stride smaller than a row");
    if (!is_device(d_src) || !is_device(d_gray)) return fail(OBS_ERR_INVALID, "both images must be device memory");
    CU(launch_gray(d_src, src_stride, src_image_stride, d_gray, gray_stride, gray_image_stride, w, h, channels, rgb_order != 0, n_images,
                   stream ? (cudaStream_t)stream : e->stream));
    return OBS_OK;
}

int obs_depth_to_float(obs_extractor* e, const uint16_t* d_src, int n_images, int w, int h, size_t src_stride, size_t src_image_stride,
                       float factor, float* d_dst, size_t dst_stride, size_t dst_image_stride, void* stream) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!d_src || !d_dst || n_images < 1 || w < 1 || h < 1) return fail(OBS_ERR_INVALID, "bad argument");
    if (src_stride < (size_t)w * 2 || dst_stride < (size_t)w * 4 || (src_stride & 1) || (dst_stride & 3)) return fail(OBS_ERR_INVALID, "bad stride");
    if (!is_device(d_src) || !is_device(d_dst)) return fail(OBS_ERR_INVALID, "both images must be device memory");
    CU(launch_depth_scale(reinterpret_cast<const uint8_t*>(d_src), src_stride, src_image_stride, reinterpret_cast<uint8_t*>(d_dst), dst_stride,
                          dst_image_stride, w, h, factor, n_images, stream ? (cudaStream_t)stream : e->stream));
    return OBS_OK;
}

int obs_stereo_from_rgbd(obs_extractor* e, const float* d_depth, size_t depth_stride, size_t depth_image_stride, float mbf, void* stream,
                         const float** d_u_right, const float** d_depth_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!d_depth || !is_device(d_depth)) return fail(OBS_ERR_INVALID, "the depth images must be device memory");
    if (e->lastN < 1) return fail(OBS_ERR_STATE, "no extraction yet");
    if (depth_stride < (size_t)e->g.w * 4 || (depth_stride & 3)) return fail(OBS_ERR_INVALID, "bad depth stride");
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const size_t cnt = (size_t)e->lastN * e->g.kpCap;
    CU(e->uRight.ensure(cnt));
    CU(e->depth.ensure(cnt));
    if (e->lastStream != st) { CU(cudaEventRecord(e->done, e->lastStream)); CU(cudaStreamWaitEvent(st, e->done, 0)); }
    CU(launch_rgbd_stereo(e->records.p, e->recordBytes, e->g.kpCap, reinterpret_cast<const uint8_t*>(d_depth), depth_stride, depth_image_stride,
                          e->g.w, e->g.h, mbf, e->uRight.p, e->depth.p, e->lastN, st));
    e->lastStream = st;
    if (d_u_right) *d_u_right = e->uRight.p;
    if (d_depth_out) *d_depth_out = e->depth.p;
    return OBS_OK;
}

// ---- asynchronous form of obs_extract_batch for page-locked buffers
int obs_extract_batch_submit(obs_extractor* e, const uint8_t* images, int n_images, int w, int h, size_t stride,
                             obs_keypoint* keypoints, uint8_t* descriptors, int cap, int32_t* n_out) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (e->pending) return fail(OBS_ERR_STATE, "obs_extract_batch_submit: the previous submission on this handle has not been waited for");
    if (n_images < 1 || n_images > e->maxBatch) return fail(OBS_ERR_CAPACITY, "batch of %d exceeds max_batch %d", n_images, e->maxBatch);
    if (w < 1 || h < 1 || w > e->maxW || h > e->maxH) return fail(OBS_ERR_INVALID, "image %dx%d outside [1,%d]x[1,%d]", w, h, e->maxW, e->maxH);
    if (stride < (size_t)w || cap < 1) return fail(OBS_ERR_INVALID, "stride smaller than width, or cap < 1");
    const void* need[] = {images, keypoints, descriptors, n_out};
    for (const void* p : need)
        if (!p || !is_pinned(p)) return fail(OBS_ERR_INVALID, "obs_extract_batch_submit takes page-locked host buffers (obs_host_alloc)");
    if ((rc = set_shape(e, w, h, n_images, e->stream))) return rc;
    if ((rc = enqueue_chunks(e, images, n_images, w, h, stride, keypoints, descriptors, cap, n_out))) return rc;
    CU(cudaEventRecord(e->hostDone, e->cout));
    e->pending = true;
    e->pendingCounts[0] = n_out; e->pendingCounts[1] = nullptr;
    e->pendingN = n_images; e->pendingCap = cap;
    return OBS_OK;
}

int obs_extract_batch_wait(obs_extractor* e) {
    int rc = check_handle(e);
    if (rc) return rc;
    if (!e->pending || !e->pendingCounts[0] || e->pendingCounts[1]) return fail(OBS_ERR_STATE, "obs_extract_batch_wait without an obs_extract_batch_submit");
    e->pending = false;
    CU(cudaEventSynchronize(e->hostDone));
    for (int i = 0; i < e->pendingN; i++)
        if (e->pendingCounts[0][i] > e->pendingCap)
            return fail(OBS_ERR_CAPACITY, "caller capacity %d smaller than the keypoint count %d (image %d)", e->pendingCap, e->pendingCounts[0][i], i);
    return OBS_OK;
}

// ---- whole stereo frames in one call (Frame::Frame for stereo, src/Frame.cc:78-90: two ExtractORB threads + ComputeStereoMatches)
int obs_stereo_frames_submit(obs_extractor* L, obs_extractor* R, const obs_stereo_io* io, int n_frames, int w, int h, size_t stride,
                             int cap, float mbf, float min_d, float max_d) {
    int rc = check_handle(L);
    if (rc) return rc;
    if (!R || !io) return fail(OBS_ERR_INVALID, "null argument");
    if (L == R) return fail(OBS_ERR_INVALID, "the two eyes need two extractor handles");
    if (L->device != R->device) return fail(OBS_ERR_INVALID, "both eyes must live on one device");
    if (L->pending || R->pending) return fail(OBS_ERR_STATE, "obs_stereo_frames_submit: the previous submission on these handles has not been waited for");
    if (n_frames < 1 || n_frames > L->maxBatch || n_frames > R->maxBatch) return fail(OBS_ERR_CAPACITY, "batch of %d exceeds max_batch", n_frames);
    if (w < 1 || h < 1 || w > L->maxW || h > L->maxH || w > R->maxW || h > R->maxH) return fail(OBS_ERR_INVALID, "image %dx%d outside the handles' limits", w, h);
    if (stride < (size_t)w || cap < 1) return fail(OBS_ERR_INVALID, "stride smaller than width, or cap < 1");
    const void* need[] = {io->left, io->right, io->kp_left, io->desc_left, io->n_left, io->kp_right, io->desc_right, io->n_right, io->u_right, io->depth};
    for (const void* p : need)
        if (!p || !is_pinned(p)) return fail(OBS_ERR_INVALID, "obs_stereo_frames_submit takes page-locked host buffers (obs_host_alloc) for every field of obs_stereo_io");
    if ((rc = set_shape(L, w, h, n_frames, L->stream))) return rc;
    if ((rc = set_shape(R, w, h, n_frames, R->stream))) return rc;
    if (cap > L->g.kpCap) return fail(OBS_ERR_INVALID, "cap %d exceeds obs_extractor_max_keypoints (%d)", cap, L->g.kpCap);
    // Steady state of a live front end: the same pinned buffers, shape and parameters call after call.  The second call with one
    // argument set is captured into a CUDA graph (both eyes' uploads, 2 x 9 kernels, the stereo kernels, every download; ordinary
    // graph edges -- the PDL attribute is dropped inside a capture), later calls are one cudaGraphLaunch instead of ~60 runtime calls.
    obs_extractor::StereoGraph key{};
    const void* iov[10] = {io->left, io->right, io->kp_left, io->desc_left, io->n_left, io->kp_right, io->desc_right, io->n_right, io->u_right, io->depth};
    memcpy(key.io, iov, sizeof(iov));
    key.n = n_frames; key.w = w; key.h = h; key.cap = cap; key.stride = stride; key.mbf = mbf; key.minD = min_d; key.maxD = max_d; key.right = R;
    obs_extractor::StereoGraph* slot = nullptr;
    const bool useGraphs = graphs_enabled() && !L->prof && !R->prof;
    if (useGraphs) {
        for (auto& sg : L->graphs) if (sg.same_args(key)) slot = &sg;
        if (slot && slot->exec && (slot->epoch != g_allocEpoch.load(std::memory_order_relaxed) || slot->pdl != pdl_enabled())) {
            cudaGraphExecDestroy(slot->exec);
            slot->exec = nullptr;
        }
    }
    // work still queued on the handles' streams by earlier calls precedes everything below (a graph runs on L->stream only)
    {
        cudaStream_t pre[3] = {R->stream, R->lastStream, L->lastStream};
        for (int i = 0; i < 3; i++) {
            if (!pre[i] || pre[i] == L->stream || (i == 1 && pre[1] == pre[0])) continue;
            CU(cudaEventRecord(L->gjoin[i], pre[i]));
            CU(cudaStreamWaitEvent(L->stream, L->gjoin[i], 0));
        }
    }
    const bool launchGraph = slot && slot->exec;
    const bool capture = slot && !slot->exec;
    if (launchGraph) {
        CU(cudaGraphLaunch(slot->exec, L->stream));
        slot->lastUse = ++L->graphClock;
        for (obs_extractor* e : {L, R}) {              // the state enqueue_chunks / obs_stereo_match_device leave behind
            e->ptrs.l0 = e->pyr.p; e->ptrs.l0ImgStride = e->g.slabBytes; e->ptrs.l0Pitch = e->g.lv[0].pitch;
            e->ptrs.slab = e->pyr.p; e->ptrs.slabStride = e->g.slabBytes;
            e->lastN = n_frames;
        }
        L->lastStream = L->stream; R->lastStream = R->stream;
    } else {
        if (capture) {
            CU(cudaStreamBeginCapture(L->stream, cudaStreamCaptureModeRelaxed));
            obsdetail::set_capturing(true);
            cudaError_t fe = cudaEventRecord(L->fork, L->stream);
            if (fe == cudaSuccess) fe = cudaStreamWaitEvent(R->stream, L->fork, 0);
            if (fe != cudaSuccess) { obsdetail::set_capturing(false); cudaGraph_t g0 = nullptr; cudaStreamEndCapture(L->stream, &g0); if (g0) cudaGraphDestroy(g0); CU(fe); }
        }
        // the two eyes on their own streams (the reference's two extraction threads), enqueued by this one host thread
        rc = enqueue_chunks(L, io->left, n_frames, w, h, stride, io->kp_left, io->desc_left, cap, io->n_left);
        if (!rc) rc = enqueue_chunks(R, io->right, n_frames, w, h, stride, io->kp_right, io->desc_right, cap, io->n_right);
        if (!rc) rc = obs_stereo_match_device(L, R, mbf, min_d, max_d, nullptr, nullptr, nullptr);
        const int kc = L->g.kpCap;
        cudaError_t ce = cudaSuccess;
        auto step = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
        if (!rc) {
            step(cudaMemcpy2DAsync(io->u_right, (size_t)cap * 4, L->uRight.p, (size_t)kc * 4, (size_t)cap * 4, n_frames, cudaMemcpyDeviceToHost, L->stream));
            step(cudaMemcpy2DAsync(io->depth, (size_t)cap * 4, L->depth.p, (size_t)kc * 4, (size_t)cap * 4, n_frames, cudaMemcpyDeviceToHost, L->stream));
            // one stream behind everything: the downloads of both eyes and of the stereo result
            step(cudaEventRecord(L->fork, L->cout));
            step(cudaStreamWaitEvent(L->stream, L->fork, 0));
            step(cudaEventRecord(R->fork, R->cout));
            step(cudaStreamWaitEvent(L->stream, R->fork, 0));
            if (capture) {                             // every stream that joined the capture returns to its origin
                cudaStream_t side[3] = {R->stream, L->cin, R->cin};
                for (int i = 0; i < 3; i++) { step(cudaEventRecord(L->gjoin[i], side[i])); step(cudaStreamWaitEvent(L->stream, L->gjoin[i], 0)); }
            }
        }
        if (capture) {
            cudaGraph_t graph = nullptr;
            cudaError_t ee = cudaStreamEndCapture(L->stream, &graph);
            obsdetail::set_capturing(false);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess || ee != cudaSuccess || !graph) {
                if (graph) cudaGraphDestroy(graph);
                cudaGetLastError();
                return fail(OBS_ERR_CUDA, "obs_stereo_frames_submit: graph capture failed: %s", cudaGetErrorString(ce != cudaSuccess ? ce : ee));
            }
            cudaGraphExec_t exec = nullptr;
            cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return fail(OBS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(ie));
            slot->exec = exec;
            slot->epoch = g_allocEpoch.load(std::memory_order_relaxed);
            slot->pdl = pdl_enabled();
            slot->lastUse = ++L->graphClock;
            CU(cudaGraphLaunch(exec, L->stream));
        } else {
            if (rc) return rc;
            CU(ce);
            if (useGraphs && !slot) {                  // first time these arguments are seen: remember them (at most 8 sets, least recently used out)
                if (L->graphs.size() >= 8) {
                    size_t v = 0;
                    for (size_t i = 1; i < L->graphs.size(); i++) if (L->graphs[i].lastUse < L->graphs[v].lastUse) v = i;
                    if (L->graphs[v].exec) cudaGraphExecDestroy(L->graphs[v].exec);
                    L->graphs.erase(L->graphs.begin() + v);
                }
                key.exec = nullptr; key.lastUse = ++L->graphClock;
                L->graphs.push_back(key);
            }
        }
    }
    if (launchGraph || capture) {                      // later work on the right handle's stream waits for the graph
        CU(cudaEventRecord(L->done, L->stream));
        CU(cudaStreamWaitEvent(R->stream, L->done, 0));
    }
    L->pendingSpin = n_frames <= 4;
    CU(cudaEventRecord(L->pendingSpin ? L->hostDoneSpin : L->hostDone, L->stream));
    L->pending = R->pending = true;
    L->pendingCounts[0] = io->n_left; L->pendingCounts[1] = io->n_right;
    L->pendingN = n_frames; L->pendingCap = cap;
    return OBS_OK;
}

int obs_stereo_frames_wait(obs_extractor* L, obs_extractor* R) {
    int rc = check_handle(L);
    if (rc) return rc;
    if (!R) return fail(OBS_ERR_INVALID, "null argument");
    if (!L->pending || !R->pending || !L->pendingCounts[1]) return fail(OBS_ERR_STATE, "obs_stereo_frames_wait without an obs_stereo_frames_submit");
    L->pending = R->pending = false;
    if (L->pendingSpin) {                           // a live frame (<= 4 frames): poll, the frame is shorter than a thread wake-up
        cudaError_t q;
        while ((q = cudaEventQuery(L->hostDoneSpin)) == cudaErrorNotReady) {}
        CU(q);
    } else {
        CU(cudaEventSynchronize(L->hostDone));      // blocking-sync event: the thread sleeps
    }
    for (int s = 0; s < 2; s++)
        for (int i = 0; i < L->pendingN; i++)
            if (L->pendingCounts[s][i] > L->pendingCap)
                return fail(OBS_ERR_CAPACITY, "caller capacity %d smaller than the keypoint count %d (frame %d, %s eye)", L->pendingCap,
                            L->pendingCounts[s][i], i, s ? "right" : "left");
    return OBS_OK;
}

int obs_stereo_frames(obs_extractor* L, obs_extractor* R, const obs_stereo_io* io, int n_frames, int w, int h, size_t stride,
                      int cap, float mbf, float min_d, float max_d) {
    int rc = obs_stereo_frames_submit(L, R, io, n_frames, w, h, stride, cap, mbf, min_d, max_d);
    if (rc) return rc;
    return obs_stereo_frames_wait(L, R);
}

int obs_set_option(const char* name, int value) {
    if (!name) return fail(OBS_ERR_INVALID, "null option name");
    if (!strcmp(name, "pdl")) { obsdetail::set_pdl_enabled(value != 0); return OBS_OK; }
    if (!strcmp(name, "graphs")) { obsdetail::set_graphs_enabled(value != 0); return OBS_OK; }
    if (!strcmp(name, "knn2_cta_pair")) { knn2_tc_set_cta_pair(value != 0); return OBS_OK; }
    return fail(OBS_ERR_INVALID, "unknown option '%s' (known: pdl, graphs, knn2_cta_pair)", name);
}

int obs_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(OBS_ERR_INVALID, "null argument");
    *out = nullptr;
    CU(cudaMallocHost(out, bytes ? bytes : 1));
    return OBS_OK;
}

int obs_host_free(void* p) {
    if (p) CU(cudaFreeHost(p));
    return OBS_OK;
}

}  // extern "C"
