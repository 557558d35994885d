// C ABI of the matchers (include/obslam_b200.h, "Matchers"): handles, staging of host arrays,
// launches.  No matching arithmetic happens on the host.
#include "../../include/obslam_b200.h"
#include "matcher.h"
#include "matcher_api.h"
#include "knn2_tc.h"
#include "host_util.h"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

struct obs_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    // staging slots for host inputs / device copies of host outputs
    static constexpr int SLOTS = 24;
    DevBuf<uint8_t> slot[SLOTS];
    DevBuf<uint32_t> cand, pool;
    DevBuf<int> poolCursor;
    DevBuf<int> choice, rounds;
    DevBuf<uint32_t> initList;
    DevBuf<int> initCount;
    DevBuf<int> sim3Idx[2], sim3Dist[2];
    DevBuf<uint8_t> knnUsed;
    DevBuf<uint8_t> knnExpanded;         // +-1 int8 expansion of the descriptor sets of the last tensor-core knn2 call
    int knnEngine = OBS_KNN2_AUTO;
    bool hsvTables = false;
    std::vector<int> lastRounds;
    std::vector<obs_frame_set*> sets;   // frame sets created on this matcher (orphaned, not freed, when it is destroyed first)
};

struct obs_frame_set {
    obs_matcher* m = nullptr;
    obs_frame_params prm;
    int maxFrames = 0, cap = 0;
    int nFrames = 0;
    FrameSetDev d{};
    DevBuf<int> n, cellStart;
    DevBuf<float4> kp;
    DevBuf<float> angle;
    DevBuf<uint4> desc;
    DevBuf<uint16_t> cellIdx;
    // upload staging
    PinBuf<uint8_t> hStage;
    DevBuf<uint8_t> dStage;
    cudaEvent_t built = nullptr;      // recorded behind every build: searches issued from another matcher's stream wait for it
    int device = 0;
};

namespace {

int check_matcher(const obs_matcher* m) {
    if (!m) return fail(OBS_ERR_INVALID, "null matcher handle");
    cudaError_t ce = cudaSetDevice(m->device);
    if (ce != cudaSuccess) return fail(OBS_ERR_CUDA, "cudaSetDevice(%d): %s", m->device, cudaGetErrorString(ce));
    return OBS_OK;
}

// A frame set may be searched from any matcher of its device (the reference builds a Frame on the Tracking thread and searches it
// from LocalMapping / LoopClosing): a foreign matcher's stream first waits for the set's last build.  Rebuilding a set while another
// thread still searches it is the caller's race, as it would be for the reference's Frame.
int adopt_set(obs_matcher* m, obs_frame_set* fs) {
    if (fs->device != m->device) return fail(OBS_ERR_INVALID, "frame set lives on device %d, the matcher on device %d", fs->device, m->device);
    if (fs->m != m && fs->built) CU(cudaStreamWaitEvent(m->stream, fs->built, 0));
    return OBS_OK;
}

// Device view of an input array: device pointers pass through, host arrays are copied into staging slot `s`.
template <typename T> int dev_in(obs_matcher* m, int s, const T* p, size_t count, const T** out) {
    *out = nullptr;
    if (!p || count == 0) return OBS_OK;
    if (is_device(p)) { *out = p; return OBS_OK; }
    const size_t bytes = count * sizeof(T);
    CU(m->slot[s].ensure((bytes + 15) & ~(size_t)15));
    CU(cudaMemcpyAsync(m->slot[s].p, p, bytes, cudaMemcpyHostToDevice, m->stream));
    *out = reinterpret_cast<const T*>(m->slot[s].p);
    return OBS_OK;
}

// Device buffer an output is produced in: the caller's own array if it is device memory, else slot `s`.
template <typename T> int dev_out(obs_matcher* m, int s, T* p, size_t count, T** out) {
    *out = nullptr;
    if (!p || count == 0) return OBS_OK;
    if (is_device(p)) { *out = p; return OBS_OK; }
    CU(m->slot[s].ensure((count * sizeof(T) + 15) & ~(size_t)15));
    *out = reinterpret_cast<T*>(m->slot[s].p);
    return OBS_OK;
}

// Copy a staged output back to a host array (no-op for device outputs).  Returns whether a copy was queued.
template <typename T> int host_back(obs_matcher* m, T* hostp, const T* devp, size_t count, bool* queued) {
    if (!hostp || !devp || count == 0 || (const void*)hostp == (const void*)devp) return OBS_OK;
    CU(cudaMemcpyAsync(hostp, devp, count * sizeof(T), cudaMemcpyDeviceToHost, m->stream));
    *queued = true;
    return OBS_OK;
}

FrameParamsDev to_dev(const obs_frame_params& p) {
    FrameParamsDev d;
    memset(&d, 0, sizeof(d));
    d.minX = p.min_x; d.maxX = p.max_x; d.minY = p.min_y; d.maxY = p.max_y;
    d.invW = (float)OBS_GRID_COLS / (p.max_x - p.min_x);       // Frame.cc:101-102
    d.invH = (float)OBS_GRID_ROWS / (p.max_y - p.min_y);
    d.fx = p.fx; d.fy = p.fy; d.cx = p.cx; d.cy = p.cy; d.mbf = p.mbf; d.mb = p.mb;
    d.nlevels = p.nlevels;
    for (int i = 0; i < p.nlevels && i < OBS_MAX_LEVELS; i++) d.scale[i] = p.scale_factors[i];
    return d;
}

}  // namespace

int obs_frame_set_build_device(obs_frame_set* fs, const uint8_t* keys, size_t keysFrameStride, const uint8_t* desc,
                               size_t descFrameStride, const float* uRight, size_t uRightFrameStride, const int* count,
                               size_t countStrideInts, int nFrames, cudaStream_t producer) {
    int rc = check_matcher(fs ? fs->m : nullptr);
    if (rc) return rc;
    if (nFrames < 1 || nFrames > fs->maxFrames) return fail(OBS_ERR_CAPACITY, "%d frames exceed the set's capacity %d", nFrames, fs->maxFrames);
    obs_matcher* m = fs->m;
    if (producer && producer != m->stream) {
        CU(cudaEventRecord(m->ev, producer));
        CU(cudaStreamWaitEvent(m->stream, m->ev, 0));
    }
    FrameBuildArgs a;
    a.F = fs->d;
    a.keys = keys; a.keysFrameStride = keysFrameStride;
    a.desc = desc; a.descFrameStride = descFrameStride;
    a.uRight = uRight; a.uRightFrameStride = uRightFrameStride;
    a.count = count; a.countStrideInts = countStrideInts;
    CU(launch_frame_build(a, nFrames, m->stream));
    CU(cudaEventRecord(fs->built, m->stream));
    fs->nFrames = nFrames;
    return OBS_OK;
}

extern "C" {

int obs_matcher_create(int device, obs_matcher** out) {
    if (!out) return fail(OBS_ERR_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(OBS_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0");
    if (device < 0 || device >= ndev) return fail(OBS_ERR_INVALID, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(OBS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    obs_matcher* m = new (std::nothrow) obs_matcher;
    if (!m) return fail(OBS_ERR_INVALID, "out of host memory");
    m->device = device;
    cudaError_t se = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&m->ev, cudaEventDisableTiming);
    if (se != cudaSuccess) { delete m; return fail(OBS_ERR_CUDA, "stream/event creation: %s", cudaGetErrorString(se)); }
    *out = m;
    return OBS_OK;
}

int obs_matcher_destroy(obs_matcher* m) {
    if (!m) return OBS_OK;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (obs_frame_set* fs : m->sets) fs->m = nullptr;     // their buffers stay valid until obs_frame_set_destroy
    for (auto& s : m->slot) s.release();
    m->knnExpanded.release(); m->knnUsed.release(); m->cand.release(); m->pool.release(); m->poolCursor.release(); m->choice.release(); m->rounds.release(); m->initList.release(); m->initCount.release();
    if (m->ev) cudaEventDestroy(m->ev);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return OBS_OK;
}

void* obs_matcher_stream(obs_matcher* m) { return m ? (void*)m->stream : nullptr; }

int obs_matcher_sync(obs_matcher* m) {
    int rc = check_matcher(m);
    if (rc) return rc;
    CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_matcher_last_rounds(obs_matcher* m, int32_t* rounds, int cap) {
    if (!m || !rounds) return fail(OBS_ERR_INVALID, "null argument");
    const int n = (int)m->lastRounds.size();
    for (int i = 0; i < n && i < cap; i++) rounds[i] = m->lastRounds[i];
    return n;
}

int obs_frame_set_create(obs_matcher* m, const obs_frame_params* params, int max_frames, int max_keypoints, obs_frame_set** out) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!params || !out) return fail(OBS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (max_frames < 1 || max_keypoints < 1 || max_keypoints > 65535) return fail(OBS_ERR_INVALID, "max_frames >= 1 and 1 <= max_keypoints <= 65535 required");
    if (params->nlevels < 1 || params->nlevels > OBS_MAX_LEVELS) return fail(OBS_ERR_INVALID, "nlevels must be in [1,%d]", OBS_MAX_LEVELS);
    if (!(params->max_x > params->min_x) || !(params->max_y > params->min_y)) return fail(OBS_ERR_INVALID, "empty image bounds");
    obs_frame_set* fs = new (std::nothrow) obs_frame_set;
    if (!fs) return fail(OBS_ERR_INVALID, "out of host memory");
    fs->m = m;
    fs->device = m->device;
    fs->prm = *params;
    fs->maxFrames = max_frames;
    fs->cap = (max_keypoints + 31) & ~31;
    const size_t B = (size_t)max_frames, cap = (size_t)fs->cap;
    cudaError_t e = fs->n.ensure(B);
    if (e == cudaSuccess) e = fs->kp.ensure(B * cap);
    if (e == cudaSuccess) e = fs->angle.ensure(B * cap);
    if (e == cudaSuccess) e = fs->desc.ensure(B * cap * 2);
    if (e == cudaSuccess) e = fs->cellStart.ensure(B * (OBS_GRID_CELLS + 1));
    if (e == cudaSuccess) e = fs->cellIdx.ensure(B * cap);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fs->built, cudaEventDisableTiming);
    if (e != cudaSuccess) { obs_frame_set_destroy(fs); return fail(OBS_ERR_CUDA, "frame set allocation: %s", cudaGetErrorString(e)); }
    fs->d.cap = fs->cap;
    fs->d.n = fs->n.p; fs->d.kp = fs->kp.p; fs->d.angle = fs->angle.p; fs->d.desc = fs->desc.p;
    fs->d.cellStart = fs->cellStart.p; fs->d.cellIdx = fs->cellIdx.p;
    fs->d.P = to_dev(*params);
    m->sets.push_back(fs);
    *out = fs;
    return OBS_OK;
}

int obs_frame_set_destroy(obs_frame_set* fs) {
    if (!fs) return OBS_OK;
    cudaSetDevice(fs->device);
    if (fs->built) cudaEventSynchronize(fs->built);
    if (fs->m) {
        cudaStreamSynchronize(fs->m->stream);
        for (size_t i = 0; i < fs->m->sets.size(); i++)
            if (fs->m->sets[i] == fs) { fs->m->sets.erase(fs->m->sets.begin() + i); break; }
    }
    fs->n.release(); fs->kp.release(); fs->angle.release(); fs->desc.release(); fs->cellStart.release(); fs->cellIdx.release();
    fs->hStage.release(); fs->dStage.release();
    if (fs->built) cudaEventDestroy(fs->built);
    delete fs;
    return OBS_OK;
}

int obs_frame_set_count(const obs_frame_set* fs) { return fs ? fs->nFrames : -1; }
}
int obs_frame_set_capacity(const obs_frame_set* fs) { return fs ? fs->cap : -1; }
int obs_frame_set_device(const obs_frame_set* fs) { return fs ? fs->device : -1; }
extern "C" {

int obs_frame_set_upload(obs_frame_set* fs, const obs_frame_view* frames, int n_frames) {
    int rc = check_matcher(fs ? fs->m : nullptr);
    if (rc) return rc;
    if (!frames || n_frames < 1) return fail(OBS_ERR_INVALID, "no frames");
    if (n_frames > fs->maxFrames) return fail(OBS_ERR_CAPACITY, "%d frames exceed the set's capacity %d", n_frames, fs->maxFrames);
    obs_matcher* m = fs->m;
    const size_t cap = (size_t)fs->cap;
    // staging record per frame: count (16 bytes) | keys cap x 28 | desc cap x 32 | uRight cap x 4
    const size_t keysOff = 16, descOff = keysOff + cap * 28, urOff = descOff + cap * 32, frameBytes = urOff + cap * 4;
    CU(fs->dStage.ensure((size_t)fs->maxFrames * frameBytes));
    CU(cudaStreamSynchronize(m->stream));         // the previous upload may still read the host staging buffer
    CU(fs->hStage.ensure((size_t)fs->maxFrames * frameBytes));
    bool anyUr = false;
    for (int b = 0; b < n_frames; b++) anyUr |= frames[b].u_right != nullptr;
    for (int b = 0; b < n_frames; b++) {
        const obs_frame_view& f = frames[b];
        if (f.n < 0 || f.n > fs->cap) return fail(OBS_ERR_CAPACITY, "frame %d has %d keypoints, capacity %d", b, f.n, fs->cap);
        if (f.n > 0 && (!f.keys_un || !f.descriptors)) return fail(OBS_ERR_INVALID, "frame %d: null keypoints or descriptors", b);
        uint8_t* rec = fs->hStage.p + (size_t)b * frameBytes;
        *reinterpret_cast<int*>(rec) = f.n;
        if (f.n > 0 && (is_device(f.keys_un) || is_device(f.descriptors) || (f.u_right && is_device(f.u_right))))
            return fail(OBS_ERR_INVALID, "obs_frame_set_upload takes host arrays (use obs_frame_set_from_extractor for device data)");
        memcpy(rec + keysOff, f.keys_un, (size_t)f.n * 28);
        memcpy(rec + descOff, f.descriptors, (size_t)f.n * 32);
        float* ur = reinterpret_cast<float*>(rec + urOff);
        if (f.u_right) memcpy(ur, f.u_right, (size_t)f.n * 4);
        else if (anyUr) for (int i = 0; i < f.n; i++) ur[i] = -1.0f;
    }
    CU(cudaMemcpyAsync(fs->dStage.p, fs->hStage.p, (size_t)n_frames * frameBytes, cudaMemcpyHostToDevice, m->stream));
    return obs_frame_set_build_device(fs, fs->dStage.p + keysOff, frameBytes, fs->dStage.p + descOff, frameBytes,
                                      anyUr ? reinterpret_cast<const float*>(fs->dStage.p + urOff) : nullptr, frameBytes / 4,
                                      reinterpret_cast<const int*>(fs->dStage.p), frameBytes / 4, n_frames, nullptr);
}

int obs_frame_set_grid(obs_frame_set* fs, int frame, int32_t* cell_start, int32_t* cell_idx, int cap) {
    int rc = check_matcher(fs ? fs->m : nullptr);
    if (rc) return rc;
    if (frame < 0 || frame >= fs->nFrames || !cell_start) return fail(OBS_ERR_INVALID, "bad frame index or null output");
    CU(cudaStreamSynchronize(fs->m->stream));
    CU(cudaMemcpy(cell_start, fs->cellStart.p + (size_t)frame * (OBS_GRID_CELLS + 1), (OBS_GRID_CELLS + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    if (cell_idx) {
        const int total = cell_start[OBS_GRID_CELLS];
        if (total > cap) return fail(OBS_ERR_CAPACITY, "cell_idx capacity %d < %d", cap, total);
        std::vector<uint16_t> tmp((size_t)std::max(total, 1));
        if (total) CU(cudaMemcpy(tmp.data(), fs->cellIdx.p + (size_t)frame * fs->cap, (size_t)total * 2, cudaMemcpyDeviceToHost));
        for (int i = 0; i < total; i++) cell_idx[i] = tmp[i];
    }
    return OBS_OK;
}

static int run_proj(obs_matcher* m, obs_frame_set* fs, ProjSearchArgs& a, int variant, int M, const int32_t* kp_observations,
                    int32_t* kp_match, int32_t* n_matches) {
    const int B = fs->nFrames;
    const size_t kc = (size_t)B * fs->cap;
    int rc;
    if ((rc = dev_in(m, 20, kp_observations, kp_observations ? kc : 0, &a.kpObs))) return rc;
    int *dMatch = nullptr, *dN = nullptr;
    if ((rc = dev_out(m, 21, kp_match, kc, &dMatch))) return rc;
    if ((rc = dev_out(m, 22, n_matches, (size_t)B, &dN))) return rc;
    CU(m->cand.ensure((size_t)B * std::max(M, 1) * OBS_CAND_SLOTS));
    a.poolChunks = std::max(M, 64) * 2;
    CU(m->pool.ensure((size_t)B * a.poolChunks * OBS_CAND_SLOTS));
    CU(m->poolCursor.ensure((size_t)B));
    CU(m->choice.ensure((size_t)B * std::max(M, 1)));
    CU(m->rounds.ensure((size_t)B));
    // k_proj_resolve keeps four ints per keypoint slot in shared memory (about 14.5 k keypoints per frame at most)
    if ((size_t)fs->cap * 16 > 220 * 1024)
        return fail(OBS_ERR_CAPACITY, "projection searches need max_keypoints <= %d (frame set was created with %d)", 220 * 1024 / 16, fs->cap);
    a.F = fs->d;
    a.cand = m->cand.p; a.pool = m->pool.p; a.poolCursor = m->poolCursor.p; a.choice = m->choice.p; a.kpMatch = dMatch; a.nMatches = dN; a.rounds = m->rounds.p;
    CU(launch_proj_search(a, variant, B, m->stream));
    bool queued = false;
    if ((rc = host_back(m, kp_match, dMatch, kc, &queued))) return rc;
    if ((rc = host_back(m, n_matches, dN, (size_t)B, &queued))) return rc;
    if (queued) {
        m->lastRounds.resize(B);
        CU(cudaMemcpyAsync(m->lastRounds.data(), m->rounds.p, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
    }
    return OBS_OK;
}

int obs_search_by_projection(obs_matcher* m, obs_frame_set* frames, const obs_mappoint_view* pts, float th, float nnratio,
                             const int32_t* kp_observations, int32_t* kp_match, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!frames || !pts || !kp_match || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, frames))) return rc;
    if (frames->nFrames < 1) return fail(OBS_ERR_STATE, "frame set is empty");
    if (pts->n < 0) return fail(OBS_ERR_INVALID, "negative point count");
    if (pts->n > 0 && (!pts->in_view || !pts->proj_x || !pts->proj_y || !pts->proj_xr || !pts->scale_level || !pts->view_cos ||
                       !pts->descriptors || !pts->observations))
        return fail(OBS_ERR_INVALID, "null map point array");
    const int B = frames->nFrames, M = pts->n;
    const size_t cnt = (size_t)M * (pts->per_frame ? B : 1);
    ProjSearchArgs a;
    memset(&a, 0, sizeof(a));
    a.mp.n = M; a.mp.stride = pts->per_frame ? (size_t)M : 0;
    const uint8_t* dsc = nullptr;
    if ((rc = dev_in(m, 0, pts->in_view, cnt, &a.mp.inView))) return rc;
    if ((rc = dev_in(m, 1, pts->proj_x, cnt, &a.mp.projX))) return rc;
    if ((rc = dev_in(m, 2, pts->proj_y, cnt, &a.mp.projY))) return rc;
    if ((rc = dev_in(m, 3, pts->proj_xr, cnt, &a.mp.projXR))) return rc;
    if ((rc = dev_in(m, 4, pts->scale_level, cnt, &a.mp.level))) return rc;
    if ((rc = dev_in(m, 5, pts->view_cos, cnt, &a.mp.viewCos))) return rc;
    if ((rc = dev_in(m, 6, pts->descriptors, cnt * 32, &dsc))) return rc;
    if ((rc = dev_in(m, 7, pts->observations, cnt, &a.mp.obs))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    a.mp.desc = reinterpret_cast<const uint4*>(dsc);
    a.th = th; a.nnratio = nnratio;
    return run_proj(m, frames, a, 0, M, kp_observations, kp_match, n_matches);
}

int obs_search_by_projection_last(obs_matcher* m, obs_frame_set* cur, const obs_lastframe_view* last, float th, int mono,
                                  int check_orientation, const int32_t* kp_observations, int32_t* kp_match, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!cur || !last || !kp_match || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, cur))) return rc;
    if (cur->nFrames < 1) return fail(OBS_ERR_STATE, "frame set is empty");
    if (last->n < 0) return fail(OBS_ERR_INVALID, "negative point count");
    if (!last->tcw_last || !last->tcw_current) return fail(OBS_ERR_INVALID, "null pose");
    if (last->n > 0 && (!last->has_point || !last->world_pos || !last->octave || !last->angle || !last->descriptors || !last->observations))
        return fail(OBS_ERR_INVALID, "null last-frame array");
    const int B = cur->nFrames, M = last->n;
    const size_t cnt = (size_t)M * (last->per_frame ? B : 1);
    ProjSearchArgs a;
    memset(&a, 0, sizeof(a));
    a.lf.n = M; a.lf.stride = last->per_frame ? (size_t)M : 0;
    const uint8_t* dsc = nullptr;
    if ((rc = dev_in(m, 0, last->has_point, cnt, &a.lf.hasPoint))) return rc;
    if ((rc = dev_in(m, 1, last->world_pos, cnt * 3, &a.lf.pos))) return rc;
    if ((rc = dev_in(m, 2, last->octave, cnt, &a.lf.octave))) return rc;
    if ((rc = dev_in(m, 3, last->angle, cnt, &a.lf.angle))) return rc;
    if ((rc = dev_in(m, 4, last->descriptors, cnt * 32, &dsc))) return rc;
    if ((rc = dev_in(m, 5, last->observations, cnt, &a.lf.obs))) return rc;
    if ((rc = dev_in(m, 6, last->tcw_last, (size_t)B * 12, &a.lf.tcwLast))) return rc;
    if ((rc = dev_in(m, 7, last->tcw_current, (size_t)B * 12, &a.lf.tcwCur))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    a.lf.desc = reinterpret_cast<const uint4*>(dsc);
    a.lf.mono = mono != 0; a.lf.checkOri = check_orientation != 0;
    a.th = th; a.nnratio = 0.f;
    return run_proj(m, cur, a, 1, M, kp_observations, kp_match, n_matches);
}

static int run_keyframe_search(obs_matcher* m, obs_frame_set* fs, const obs_keyframe_points_view* pts, int variant, float th, int distTh,
                               int checkOri, const int32_t* kp_taken, int32_t* kp_match, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!fs || !pts || !kp_match || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, fs))) return rc;
    if (fs->nFrames < 1) return fail(OBS_ERR_STATE, "frame set is empty");
    if (pts->n < 0 || !pts->tcw) return fail(OBS_ERR_INVALID, "negative point count or null pose");
    if (pts->n > 0 && (!pts->valid || !pts->world_pos || !pts->min_distance || !pts->max_distance || !pts->max_distance_raw ||
                       !pts->descriptors || (variant == 3 && !pts->normal) || (variant == 2 && checkOri && !pts->angle)))
        return fail(OBS_ERR_INVALID, "null map point array");
    const int B = fs->nFrames, M = pts->n;
    const size_t cnt = (size_t)M * (pts->per_frame ? B : 1);
    ProjSearchArgs a;
    memset(&a, 0, sizeof(a));
    a.kf.n = M; a.kf.stride = pts->per_frame ? (size_t)M : 0;
    const uint8_t* dsc = nullptr;
    if ((rc = dev_in(m, 0, pts->valid, cnt, &a.kf.valid))) return rc;
    if ((rc = dev_in(m, 1, pts->world_pos, cnt * 3, &a.kf.pos))) return rc;
    if ((rc = dev_in(m, 2, pts->min_distance, cnt, &a.kf.minDist))) return rc;
    if ((rc = dev_in(m, 3, pts->max_distance, cnt, &a.kf.maxDist))) return rc;
    if ((rc = dev_in(m, 4, pts->max_distance_raw, cnt, &a.kf.maxDistRaw))) return rc;
    if ((rc = dev_in(m, 5, pts->normal, pts->normal ? cnt * 3 : 0, &a.kf.normal))) return rc;
    if ((rc = dev_in(m, 6, pts->angle, pts->angle ? cnt : 0, &a.kf.angle))) return rc;
    if ((rc = dev_in(m, 7, pts->descriptors, cnt * 32, &dsc))) return rc;
    if ((rc = dev_in(m, 8, pts->tcw, (size_t)B * 12, &a.kf.tcw))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    a.kf.desc = reinterpret_cast<const uint4*>(dsc);
    a.kf.logScaleFactor = logf(fs->prm.nlevels > 1 ? fs->prm.scale_factors[1] : 1.2f);      // Frame.cc:71: mfLogScaleFactor = log(mfScaleFactor)
    a.kf.distTh = distTh;
    a.kf.checkOri = checkOri != 0;
    a.th = th; a.nnratio = 0.f;
    return run_proj(m, fs, a, variant, M, kp_taken, kp_match, n_matches);
}

int obs_search_by_projection_keyframe(obs_matcher* m, obs_frame_set* current, const obs_keyframe_points_view* points, float th,
                                      int orb_dist, int check_orientation, const int32_t* kp_taken, int32_t* kp_match, int32_t* n_matches) {
    return run_keyframe_search(m, current, points, 2, th, orb_dist, check_orientation, kp_taken, kp_match, n_matches);
}

int obs_search_by_projection_sim3(obs_matcher* m, obs_frame_set* keyframes, const obs_keyframe_points_view* points, int th,
                                  const int32_t* kp_taken, int32_t* kp_match, int32_t* n_matches) {
    return run_keyframe_search(m, keyframes, points, 3, (float)th, 50 /* TH_LOW */, 0, kp_taken, kp_match, n_matches);
}

int obs_fuse_search(obs_matcher* m, obs_frame_set* fs, const obs_keyframe_points_view* pts, const float* camera_centre, float th,
                    int sim3, int32_t* best_idx, int32_t* best_dist) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!fs || !pts || !best_idx || !best_dist) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, fs))) return rc;
    if (fs->nFrames < 1) return fail(OBS_ERR_STATE, "frame set is empty");
    if (pts->n < 1 || !pts->tcw) return fail(OBS_ERR_INVALID, "point count < 1 or null pose");
    if (!pts->valid || !pts->world_pos || !pts->min_distance || !pts->max_distance || !pts->max_distance_raw || !pts->descriptors || !pts->normal)
        return fail(OBS_ERR_INVALID, "null map point array");
    if (!sim3 && !camera_centre) return fail(OBS_ERR_INVALID, "the keyframe variant needs GetCameraCenter()");
    const int B = fs->nFrames, M = pts->n;
    const size_t cnt = (size_t)M * (pts->per_frame ? B : 1);
    FuseSearchArgs a;
    memset(&a, 0, sizeof(a));
    a.F = fs->d;
    a.kf.n = M; a.kf.stride = pts->per_frame ? (size_t)M : 0;
    const uint8_t* dsc = nullptr;
    if ((rc = dev_in(m, 0, pts->valid, cnt, &a.kf.valid))) return rc;
    if ((rc = dev_in(m, 1, pts->world_pos, cnt * 3, &a.kf.pos))) return rc;
    if ((rc = dev_in(m, 2, pts->min_distance, cnt, &a.kf.minDist))) return rc;
    if ((rc = dev_in(m, 3, pts->max_distance, cnt, &a.kf.maxDist))) return rc;
    if ((rc = dev_in(m, 4, pts->max_distance_raw, cnt, &a.kf.maxDistRaw))) return rc;
    if ((rc = dev_in(m, 5, pts->normal, cnt * 3, &a.kf.normal))) return rc;
    if ((rc = dev_in(m, 7, pts->descriptors, cnt * 32, &dsc))) return rc;
    if ((rc = dev_in(m, 8, pts->tcw, (size_t)B * 12, &a.kf.tcw))) return rc;
    if ((rc = dev_in(m, 9, sim3 ? nullptr : camera_centre, sim3 ? 0 : (size_t)B * 3, &a.ow))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    a.kf.desc = reinterpret_cast<const uint4*>(dsc);
    a.kf.logScaleFactor = logf(fs->prm.nlevels > 1 ? fs->prm.scale_factors[1] : 1.2f);
    for (int l = 0; l < fs->prm.nlevels && l < OBS_MAX_LEVELS; l++) {
        const float sf = fs->prm.scale_factors[l];
        a.invSigma2[l] = 1.0f / (sf * sf);                  // mvInvLevelSigma2, src/ORBextractor.cc:422-431
    }
    a.th = th; a.sim3 = sim3 != 0;
    const size_t oc = (size_t)B * M;
    if ((rc = dev_out(m, 20, best_idx, oc, &a.bestIdx))) return rc;
    if ((rc = dev_out(m, 21, best_dist, oc, &a.bestDist))) return rc;
    CU(launch_fuse_search(a, B, m->stream));
    bool queued = false;
    if ((rc = host_back(m, best_idx, a.bestIdx, oc, &queued))) return rc;
    if ((rc = host_back(m, best_dist, a.bestDist, oc, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

static int sim3_side(obs_matcher* m, int slot0, obs_frame_set* target, const obs_keyframe_points_view* pts, const float* t2, float th,
                     int B, FuseSearchArgs* a, DevBuf<int>& idx, DevBuf<int>& dist) {
    int rc;
    if (pts->n < 1 || !pts->tcw || !pts->valid || !pts->world_pos || !pts->min_distance || !pts->max_distance || !pts->max_distance_raw || !pts->descriptors)
        return fail(OBS_ERR_INVALID, "null map point array or point count < 1");
    const int M = pts->n;
    const size_t cnt = (size_t)M * (pts->per_frame ? B : 1);
    memset(a, 0, sizeof(*a));
    a->F = target->d;
    a->kf.n = M; a->kf.stride = pts->per_frame ? (size_t)M : 0;
    const uint8_t* dsc = nullptr;
    if ((rc = dev_in(m, slot0 + 0, pts->valid, cnt, &a->kf.valid))) return rc;
    if ((rc = dev_in(m, slot0 + 1, pts->world_pos, cnt * 3, &a->kf.pos))) return rc;
    if ((rc = dev_in(m, slot0 + 2, pts->min_distance, cnt, &a->kf.minDist))) return rc;
    if ((rc = dev_in(m, slot0 + 3, pts->max_distance, cnt, &a->kf.maxDist))) return rc;
    if ((rc = dev_in(m, slot0 + 4, pts->max_distance_raw, cnt, &a->kf.maxDistRaw))) return rc;
    if ((rc = dev_in(m, slot0 + 5, pts->descriptors, cnt * 32, &dsc))) return rc;
    if ((rc = dev_in(m, slot0 + 6, pts->tcw, (size_t)B * 12, &a->kf.tcw))) return rc;
    if ((rc = dev_in(m, slot0 + 7, t2, (size_t)B * 12, &a->tcw2))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    a->kf.desc = reinterpret_cast<const uint4*>(dsc);
    a->kf.logScaleFactor = logf(target->prm.nlevels > 1 ? target->prm.scale_factors[1] : 1.2f);
    a->th = th; a->sim3 = 2;
    CU(idx.ensure((size_t)B * M));
    CU(dist.ensure((size_t)B * M));
    a->bestIdx = idx.p; a->bestDist = dist.p;
    return OBS_OK;
}

int obs_search_by_sim3(obs_matcher* m, obs_frame_set* kf1, obs_frame_set* kf2, const obs_keyframe_points_view* points1,
                       const obs_keyframe_points_view* points2, const float* t21, const float* t12, float th,
                       int32_t* match12, int32_t* n_found) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!kf1 || !kf2 || !points1 || !points2 || !t21 || !t12 || !match12 || !n_found) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, kf1)) || (rc = adopt_set(m, kf2))) return rc;
    if (kf1->nFrames < 1 || kf1->nFrames != kf2->nFrames) return fail(OBS_ERR_STATE, "frame sets are empty or differ in size");
    const int B = kf1->nFrames;
    FuseSearchArgs a1, a2;
    if ((rc = sim3_side(m, 0, kf2, points1, t21, th, B, &a1, m->sim3Idx[0], m->sim3Dist[0]))) return rc;      // KF1's points into KF2, :1153-1226
    if ((rc = sim3_side(m, 8, kf1, points2, t12, th, B, &a2, m->sim3Idx[1], m->sim3Dist[1]))) return rc;      // KF2's points into KF1, :1228-1302
    int *dMatch = nullptr, *dN = nullptr;
    const size_t oc = (size_t)B * points1->n;
    if ((rc = dev_out(m, 20, match12, oc, &dMatch))) return rc;
    if ((rc = dev_out(m, 22, n_found, (size_t)B, &dN))) return rc;
    CU(launch_fuse_search(a1, B, m->stream));
    CU(launch_fuse_search(a2, B, m->stream));
    CU(launch_sim3_agree(a1.bestIdx, a1.bestDist, points1->n, a2.bestIdx, a2.bestDist, points2->n, 100 /* TH_HIGH */, dMatch, dN, B, m->stream));
    bool queued = false;
    if ((rc = host_back(m, match12, dMatch, oc, &queued))) return rc;
    if ((rc = host_back(m, n_found, dN, (size_t)B, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_search_for_initialization(obs_matcher* m, obs_frame_set* f1, obs_frame_set* f2, float* prev_matched, int32_t* matches12,
                                  int window_size, float nnratio, int check_orientation, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!f1 || !f2 || !prev_matched || !matches12 || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if ((rc = adopt_set(m, f1)) || (rc = adopt_set(m, f2))) return rc;
    if (f1->nFrames < 1 || f1->nFrames != f2->nFrames) return fail(OBS_ERR_STATE, "both frame sets must hold the same number (>= 1) of frames");
    const int B = f1->nFrames;
    const size_t c1 = (size_t)B * f1->cap;
    // k_init_search keeps two ints per keypoint slot of either frame in shared memory; the candidate lists are cap2 entries per
    // keypoint of frame 1 (a window can hold every keypoint of frame 2)
    if (((size_t)f1->cap + (size_t)f2->cap) * 8 > 220 * 1024)
        return fail(OBS_ERR_CAPACITY, "SearchForInitialization needs max_keypoints1 + max_keypoints2 <= %d (got %d + %d)", 220 * 1024 / 8, f1->cap, f2->cap);
    if (c1 * (size_t)f2->cap * 4 > ((size_t)4 << 30))
        return fail(OBS_ERR_CAPACITY, "SearchForInitialization: %d frame pairs x %d x %d candidate slots exceed 4 GB; use smaller batches or frame sets "
                    "created with a tighter max_keypoints", B, f1->cap, f2->cap);
    InitSearchArgs a;
    memset(&a, 0, sizeof(a));
    a.F1 = f1->d; a.F2 = f2->d;
    const float* pmIn = nullptr;
    if ((rc = dev_in(m, 0, (const float*)prev_matched, c1 * 2, &pmIn))) return rc;
    a.prevMatched = const_cast<float*>(pmIn);
    if ((rc = dev_out(m, 1, matches12, c1, &a.matches12))) return rc;
    if ((rc = dev_out(m, 2, n_matches, (size_t)B, &a.nMatches))) return rc;
    a.listCap = f2->cap;
    CU(m->initList.ensure(c1 * (size_t)a.listCap));
    CU(m->initCount.ensure(c1));
    a.list = m->initList.p; a.listCount = m->initCount.p;
    a.window = window_size; a.nnratio = nnratio; a.checkOri = check_orientation != 0;
    CU(launch_init_search(a, B, m->stream));
    bool queued = false;
    if ((rc = host_back(m, prev_matched, a.prevMatched, c1 * 2, &queued))) return rc;
    if ((rc = host_back(m, matches12, a.matches12, c1, &queued))) return rc;
    if ((rc = host_back(m, n_matches, a.nMatches, (size_t)B, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_compute_three_maxima(obs_matcher* m, const int32_t* bin_sizes, int n_hist, int length, int32_t* ind) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!bin_sizes || !ind || n_hist < 1 || length < 1) return fail(OBS_ERR_INVALID, "bad argument");
    const int* dIn = nullptr; int* dOut = nullptr;
    if ((rc = dev_in(m, 0, bin_sizes, (size_t)n_hist * length, &dIn))) return rc;
    if ((rc = dev_out(m, 1, ind, (size_t)n_hist * 3, &dOut))) return rc;
    CU(launch_three_maxima(dIn, n_hist, length, dOut, m->stream));
    bool queued = false;
    if ((rc = host_back(m, ind, dOut, (size_t)n_hist * 3, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_descriptor_distance(obs_matcher* m, const uint8_t* a, const uint8_t* b, int n, int32_t* dist) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!a || !b || !dist || n < 1) return fail(OBS_ERR_INVALID, "bad argument");
    const uint8_t *da = nullptr, *db = nullptr; int* dd = nullptr;
    if ((rc = dev_in(m, 0, a, (size_t)n * 32, &da))) return rc;
    if ((rc = dev_in(m, 1, b, (size_t)n * 32, &db))) return rc;
    if (((uintptr_t)da | (uintptr_t)db) & 3) return fail(OBS_ERR_INVALID, "device descriptors must be 4-byte aligned");
    if ((rc = dev_out(m, 2, dist, (size_t)n, &dd))) return rc;
    CU(launch_descriptor_distance(da, db, n, dd, m->stream));
    bool queued = false;
    if ((rc = host_back(m, dist, dd, (size_t)n, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_matcher_set_knn2_engine(obs_matcher* m, int engine) {
    if (!m) return fail(OBS_ERR_INVALID, "null matcher handle");
    if (engine != OBS_KNN2_AUTO && engine != OBS_KNN2_POPC && engine != OBS_KNN2_TENSOR) return fail(OBS_ERR_INVALID, "unknown knn2 engine %d", engine);
    m->knnEngine = engine;
    return OBS_OK;
}

int obs_hamming_knn2(obs_matcher* m, const uint8_t* descriptors, int n_keyframes, int n_desc, const int32_t* pairs, int n_pairs,
                     int th_low, float nnratio, int32_t* best_idx, int32_t* best_dist, int32_t* second_dist) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!descriptors || !pairs || !best_idx) return fail(OBS_ERR_INVALID, "null argument");
    if (n_keyframes < 1 || n_desc < 1 || n_desc > 65535 || n_pairs < 1) return fail(OBS_ERR_INVALID, "sizes out of range (1 <= n_desc <= 65535)");
    const uint8_t* dd = nullptr; const int* dp = nullptr;
    if ((rc = dev_in(m, 0, descriptors, (size_t)n_keyframes * n_desc * 32, &dd))) return rc;
    if ((uintptr_t)dd & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    if ((rc = dev_in(m, 1, pairs, (size_t)n_pairs * 2, &dp))) return rc;
    if (!is_device(pairs))
        for (int i = 0; i < 2 * n_pairs; i++)
            if (pairs[i] < 0 || pairs[i] >= n_keyframes) return fail(OBS_ERR_INVALID, "pair %d names keyframe %d of %d", i / 2, pairs[i], n_keyframes);
    const size_t cnt = (size_t)n_pairs * n_desc;
    Knn2Args a;
    a.desc = reinterpret_cast<const uint4*>(dd);
    a.n = n_desc; a.pairs = reinterpret_cast<const int2*>(dp); a.nPairs = n_pairs;
    a.thLow = th_low; a.nnratio = nnratio;
    if ((rc = dev_out(m, 2, best_idx, cnt, &a.bestIdx))) return rc;
    if ((rc = dev_out(m, 3, best_dist, best_dist ? cnt : 0, &a.bestDist))) return rc;
    if ((rc = dev_out(m, 4, second_dist, second_dist ? cnt : 0, &a.secondDist))) return rc;
    // tensor cores (knn2_tc.cu) once a keyframe fills at least one 128 x 256 tile reasonably; the POPC kernel below that
    const bool tensor = m->knnEngine == OBS_KNN2_TENSOR || (m->knnEngine == OBS_KNN2_AUTO && n_desc >= 192);
    if (tensor) {
        CU(m->knnExpanded.ensure(knn2_tc_expanded_bytes(n_keyframes, n_desc)));
        CU(m->knnUsed.ensure(knn2_tc_scratch_bytes(n_keyframes)));
        CU(launch_knn2_tc(a, n_keyframes, m->knnExpanded.p, m->knnUsed.p, m->stream));
    } else {
        CU(launch_knn2(a, m->stream));
    }
    bool queued = false;
    if ((rc = host_back(m, best_idx, a.bestIdx, cnt, &queued))) return rc;
    if ((rc = host_back(m, best_dist, a.bestDist, cnt, &queued))) return rc;
    if ((rc = host_back(m, second_dist, a.secondDist, cnt, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

static int bow_side_dev(obs_matcher* m, int slot0, const obs_bow_side* s, int B, bool needUr, BowSideDev* d) {
    if (!s || !s->n || !s->descriptors || !s->keys_un || !s->n_nodes || !s->node_id || !s->node_start || !s->node_idx)
        return fail(OBS_ERR_INVALID, "null field in obs_bow_side");
    if (s->cap < 1 || s->cap > 65535 || s->node_cap < 1) return fail(OBS_ERR_INVALID, "obs_bow_side: 1 <= cap <= 65535, node_cap >= 1");
    int rc;
    const size_t kc = (size_t)B * s->cap;
    const uint8_t* dsc = nullptr; const obs_keypoint* keys = nullptr;
    d->cap = s->cap; d->nodeCap = s->node_cap;
    if ((rc = dev_in(m, slot0 + 0, s->n, (size_t)B, &d->n))) return rc;
    if ((rc = dev_in(m, slot0 + 1, s->descriptors, kc * 32, &dsc))) return rc;
    if ((uintptr_t)dsc & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    d->desc = reinterpret_cast<const uint4*>(dsc);
    if ((rc = dev_in(m, slot0 + 2, s->keys_un, kc, &keys))) return rc;
    d->keys = reinterpret_cast<const float*>(keys);
    if ((rc = dev_in(m, slot0 + 3, s->valid, s->valid ? kc : 0, &d->valid))) return rc;
    d->uRight = nullptr;
    if (needUr && (rc = dev_in(m, slot0 + 4, s->u_right, s->u_right ? kc : 0, &d->uRight))) return rc;
    if ((rc = dev_in(m, slot0 + 5, s->n_nodes, (size_t)B, &d->nNodes))) return rc;
    if ((rc = dev_in(m, slot0 + 6, s->node_id, (size_t)B * s->node_cap, &d->nodeId))) return rc;
    if ((rc = dev_in(m, slot0 + 7, s->node_start, (size_t)B * (s->node_cap + 1), &d->nodeStart))) return rc;
    if ((rc = dev_in(m, slot0 + 8, s->node_idx, kc, &d->nodeIdx))) return rc;
    return OBS_OK;
}

int obs_search_by_bow(obs_matcher* m, const obs_bow_side* side1, const obs_bow_side* side2, int n_pairs, int th_low,
                      int strict_low, float nnratio, int check_orientation, int32_t* match12, int32_t* match21, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!match12 || !match21 || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if (n_pairs < 1) return fail(OBS_ERR_INVALID, "n_pairs < 1");
    BowSearchArgs a;
    if ((rc = bow_side_dev(m, 0, side1, n_pairs, false, &a.A))) return rc;
    if ((rc = bow_side_dev(m, 9, side2, n_pairs, false, &a.B))) return rc;
    a.thLow = th_low; a.strictLow = strict_low != 0; a.nnratio = nnratio; a.checkOri = check_orientation != 0;
    const size_t c1 = (size_t)n_pairs * side1->cap, c2 = (size_t)n_pairs * side2->cap;
    if ((rc = dev_out(m, 20, match12, c1, &a.match12))) return rc;
    if ((rc = dev_out(m, 21, match21, c2, &a.match21))) return rc;
    if ((rc = dev_out(m, 22, n_matches, (size_t)n_pairs, &a.nMatches))) return rc;
    CU(launch_bow_search(a, n_pairs, m->stream));
    bool queued = false;
    if ((rc = host_back(m, match12, a.match12, c1, &queued))) return rc;
    if ((rc = host_back(m, match21, a.match21, c2, &queued))) return rc;
    if ((rc = host_back(m, n_matches, a.nMatches, (size_t)n_pairs, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_search_for_triangulation(obs_matcher* m, const obs_bow_side* side1, const obs_bow_side* side2, int n_pairs,
                                 const float* f12, const float* epipole, const float* level_sigma2, const float* scale_factors,
                                 int nlevels, int only_stereo, int check_orientation, int32_t* match12, int32_t* n_matches) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!f12 || !epipole || !level_sigma2 || !scale_factors || !match12 || !n_matches) return fail(OBS_ERR_INVALID, "null argument");
    if (n_pairs < 1 || nlevels < 1 || nlevels > OBS_MAX_LEVELS) return fail(OBS_ERR_INVALID, "n_pairs / nlevels out of range");
    if (is_device(level_sigma2) || is_device(scale_factors)) return fail(OBS_ERR_INVALID, "level tables are host arrays");
    TriSearchArgs a;
    memset(&a, 0, sizeof(a));
    if ((rc = bow_side_dev(m, 0, side1, n_pairs, true, &a.A))) return rc;
    if ((rc = bow_side_dev(m, 9, side2, n_pairs, true, &a.B))) return rc;
    if ((rc = dev_in(m, 18, f12, (size_t)n_pairs * 9, &a.f12))) return rc;
    if ((rc = dev_in(m, 19, epipole, (size_t)n_pairs * 2, &a.epipole))) return rc;
    for (int l = 0; l < nlevels; l++) { a.sigma2[l] = level_sigma2[l]; a.scale[l] = scale_factors[l]; }
    a.onlyStereo = only_stereo != 0; a.checkOri = check_orientation != 0;
    const size_t c1 = (size_t)n_pairs * side1->cap;
    if ((rc = dev_out(m, 20, match12, c1, &a.match12))) return rc;
    if ((rc = dev_out(m, 22, n_matches, (size_t)n_pairs, &a.nMatches))) return rc;
    CU(launch_tri_search(a, n_pairs, m->stream));
    bool queued = false;
    if ((rc = host_back(m, match12, a.match12, c1, &queued))) return rc;
    if ((rc = host_back(m, n_matches, a.nMatches, (size_t)n_pairs, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_distinctive_descriptors(obs_matcher* m, const uint8_t* descriptors, const int32_t* start, int n_points, int32_t* best) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!start || !best || n_points < 1) return fail(OBS_ERR_INVALID, "null argument or n_points < 1");
    if (is_device(start) && !is_device(descriptors)) return fail(OBS_ERR_INVALID, "start on the device needs descriptors on the device");
    const uint8_t* dd = nullptr; const int* ds = nullptr; int* db = nullptr;
    size_t total = 0;
    if (!is_device(start)) {
        for (int p = 0; p < n_points; p++) if (start[p + 1] < start[p]) return fail(OBS_ERR_INVALID, "start must not decrease");
        total = (size_t)start[n_points];
        if (total && !descriptors) return fail(OBS_ERR_INVALID, "null descriptors");
    }
    if ((rc = dev_in(m, 0, descriptors, total * 32, &dd))) return rc;
    if (is_device(descriptors)) dd = descriptors;
    if ((uintptr_t)dd & 15) return fail(OBS_ERR_INVALID, "device descriptors must be 16-byte aligned");
    if ((rc = dev_in(m, 1, start, (size_t)n_points + 1, &ds))) return rc;
    if ((rc = dev_out(m, 2, best, (size_t)n_points, &db))) return rc;
    CU(launch_distinctive(reinterpret_cast<const uint4*>(dd), ds, n_points, db, m->stream));
    bool queued = false;
    if ((rc = host_back(m, best, db, (size_t)n_points, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_assign_keypoints_to_masks(obs_matcher* m, const obs_keypoint* keys_un, const float* depth, int n, const uint8_t* masks,
                                  int n_masks, int w, int h, size_t mask_stride, size_t mask_image_stride, float th_depth,
                                  int min_keypoints, int32_t* mask_of_kp, int32_t* object_kp_indices, int32_t* object_of_mask,
                                  int32_t* n_objects) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!keys_un || !depth || !masks || !object_kp_indices || !n_objects) return fail(OBS_ERR_INVALID, "null argument");
    if (n < 1 || n_masks < 1 || w < 1 || h < 1 || mask_stride < (size_t)w || mask_image_stride < mask_stride * (size_t)h)
        return fail(OBS_ERR_INVALID, "sizes / strides out of range");
    if (mask_stride * (size_t)h > 0x7fffffffu) return fail(OBS_ERR_INVALID, "mask too large");
    MaskAssignArgs a;
    memset(&a, 0, sizeof(a));
    const obs_keypoint* dk = nullptr;
    if ((rc = dev_in(m, 0, keys_un, (size_t)n, &dk))) return rc;
    a.keys = reinterpret_cast<const float*>(dk);
    if ((rc = dev_in(m, 1, depth, (size_t)n, &a.depth))) return rc;
    if ((rc = dev_in(m, 2, masks, (size_t)n_masks * mask_image_stride, &a.masks))) return rc;
    a.n = n; a.nMasks = n_masks; a.w = w; a.h = h; a.rowStride = mask_stride; a.imageStride = mask_image_stride;
    a.thDepth = th_depth; a.minKeypoints = min_keypoints;
    CU(m->choice.ensure((size_t)n + n_masks));
    if ((rc = dev_out(m, 20, mask_of_kp, (size_t)n, &a.maskOfKp))) return rc;
    if (!a.maskOfKp) a.maskOfKp = m->choice.p;
    if ((rc = dev_out(m, 21, object_kp_indices, (size_t)n * 2, &a.objectKp))) return rc;
    if ((rc = dev_out(m, 22, object_of_mask, (size_t)n_masks, &a.objectOfMask))) return rc;
    if (!a.objectOfMask) a.objectOfMask = m->choice.p + n;
    if ((rc = dev_out(m, 23, n_objects, 1, &a.nObjects))) return rc;
    CU(launch_mask_assign(a, m->stream));
    bool queued = false;
    if ((rc = host_back(m, mask_of_kp, a.maskOfKp, (size_t)n, &queued))) return rc;
    if ((rc = host_back(m, object_kp_indices, a.objectKp, (size_t)n * 2, &queued))) return rc;
    if ((rc = host_back(m, object_of_mask, a.objectOfMask, (size_t)n_masks, &queued))) return rc;
    if ((rc = host_back(m, n_objects, a.nObjects, 1, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_hsv_histograms(obs_matcher* m, const uint8_t* bgr, size_t bgr_stride, const uint8_t* masks, int n_masks, int w, int h,
                       size_t mask_stride, size_t mask_image_stride, float* hist) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!bgr || !masks || !hist) return fail(OBS_ERR_INVALID, "null argument");
    if (n_masks < 1 || w < 1 || h < 1 || bgr_stride < (size_t)w * 3 || mask_stride < (size_t)w || mask_image_stride < mask_stride * (size_t)h)
        return fail(OBS_ERR_INVALID, "sizes / strides out of range");
    if (!m->hsvTables) { CU(hsv_tables_upload()); m->hsvTables = true; }
    const uint8_t *dImg = nullptr, *dMask = nullptr;
    float* dHist = nullptr;
    if ((rc = dev_in(m, 0, bgr, bgr_stride * (size_t)h, &dImg))) return rc;
    if ((rc = dev_in(m, 1, masks, (size_t)n_masks * mask_image_stride, &dMask))) return rc;
    if ((rc = dev_out(m, 20, hist, (size_t)n_masks * OBS_HSV_BINS, &dHist))) return rc;
    CU(m->choice.ensure((size_t)n_masks * OBS_HSV_BINS));
    CU(launch_hsv_hist(dImg, bgr_stride, dMask, mask_stride, mask_image_stride, n_masks, w, h, m->choice.p, dHist, m->stream));
    bool queued = false;
    if ((rc = host_back(m, hist, dHist, (size_t)n_masks * OBS_HSV_BINS, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

static int undistort_args(float fx, float fy, float cx, float cy, const float* dist, int n_dist, UndistortArgs* a) {
    if (n_dist < 0 || n_dist > 14 || (n_dist > 0 && !dist)) return fail(OBS_ERR_INVALID, "0..14 distortion coefficients");
    if (n_dist > 0 && is_device(dist)) return fail(OBS_ERR_INVALID, "distortion coefficients are a host array");
    memset(a, 0, sizeof(*a));
    a->fx = fx; a->fy = fy; a->cx = cx; a->cy = cy;
    for (int i = 0; i < n_dist; i++) a->k[i] = dist[i];
    return OBS_OK;
}

int obs_undistort_points(obs_matcher* m, const float* pts, int n, float fx, float fy, float cx, float cy, const float* dist_coef,
                         int n_dist, float* out) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!pts || !out || n < 1) return fail(OBS_ERR_INVALID, "null argument or n < 1");
    UndistortArgs a;
    if ((rc = undistort_args(fx, fy, cx, cy, dist_coef, n_dist, &a))) return rc;
    const float* dp = nullptr; float* dout = nullptr;
    if ((rc = dev_in(m, 0, pts, (size_t)n * 2, &dp))) return rc;
    if ((rc = dev_out(m, 20, out, (size_t)n * 2, &dout))) return rc;
    CU(launch_undistort(a, dp, 2, dout, 2, n, m->stream));
    bool queued = false;
    if ((rc = host_back(m, out, dout, (size_t)n * 2, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_undistort_keypoints(obs_matcher* m, const obs_keypoint* keys, int n, float fx, float fy, float cx, float cy,
                            const float* dist_coef, int n_dist, obs_keypoint* keys_un) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!keys || !keys_un || n < 1) return fail(OBS_ERR_INVALID, "null argument or n < 1");
    UndistortArgs a;
    if ((rc = undistort_args(fx, fy, cx, cy, dist_coef, n_dist, &a))) return rc;
    const obs_keypoint* dk = nullptr; obs_keypoint* dout = nullptr;
    if ((rc = dev_in(m, 0, keys, (size_t)n, &dk))) return rc;
    if ((rc = dev_out(m, 20, keys_un, (size_t)n, &dout))) return rc;
    if ((const void*)dk != (const void*)dout)
        CU(cudaMemcpyAsync(dout, dk, (size_t)n * sizeof(obs_keypoint), cudaMemcpyDeviceToDevice, m->stream));      // mvKeysUn[i] = mvKeys[i] with pt replaced
    if (n_dist > 0 && dist_coef[0] != 0.0f)                    // mDistCoef.at<float>(0) == 0.0: mvKeysUn = mvKeys, Frame.cc:646-650
        CU(launch_undistort(a, reinterpret_cast<const float*>(dk), 7, reinterpret_cast<float*>(dout), 7, n, m->stream));
    bool queued = false;
    if ((rc = host_back(m, keys_un, dout, (size_t)n, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

int obs_distance_transform(obs_matcher* m, const uint8_t* masks, int n_masks, int w, int h, size_t mask_stride,
                           size_t mask_image_stride, float* dist) {
    int rc = check_matcher(m);
    if (rc) return rc;
    if (!masks || !dist) return fail(OBS_ERR_INVALID, "null argument");
    if (n_masks < 1 || w < 1 || h < 1 || w > 4096 || h > 32767 || mask_stride < (size_t)w || mask_image_stride < mask_stride * (size_t)h)
        return fail(OBS_ERR_INVALID, "sizes / strides out of range (w <= 4096)");
    const uint8_t* dMask = nullptr; float* dOut = nullptr;
    const size_t cnt = (size_t)n_masks * w * h;
    if ((rc = dev_in(m, 1, masks, (size_t)n_masks * mask_image_stride, &dMask))) return rc;
    if ((rc = dev_out(m, 20, dist, cnt, &dOut))) return rc;
    CU(launch_distance_transform(dMask, mask_stride, mask_image_stride, n_masks, w, h, dOut, m->stream));
    bool queued = false;
    if ((rc = host_back(m, dist, dOut, cnt, &queued))) return rc;
    if (queued) CU(cudaStreamSynchronize(m->stream));
    return OBS_OK;
}

}  // extern "C"
