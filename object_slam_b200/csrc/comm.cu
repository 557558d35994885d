// K11 exchange step: all-gather of the keyframe descriptor shards over NCCL (NVLink / NVSwitch), the one
// collective of the path.  NCCL is bound at run time (dlopen of libnccl.so.2), so a
// host process that already carries an NCCL (e.g. through torch.distributed) shares that copy and the library
// has no link-time dependency on it.  The gather runs on the communicator's own stream in chunks, each chunk
// signalled by an event, so that matching against keyframes that are already present overlaps the transfer.
#include "../../include/obslam_b200.h"
#include "host_util.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <new>
#include <vector>

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
};

// Resolved once per process; the initialisation of a function-local static is thread safe (communicators may be created
// from several host threads).
NcclApi* nccl_api() {
    static const NcclApi api = [] {
        NcclApi api;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.h = h;
            *(void**)&api.GetUniqueId = dlsym(h, "ncclGetUniqueId");
            *(void**)&api.CommInitRank = dlsym(h, "ncclCommInitRank");
            *(void**)&api.CommDestroy = dlsym(h, "ncclCommDestroy");
            *(void**)&api.Broadcast = dlsym(h, "ncclBroadcast");
            *(void**)&api.AllGather = dlsym(h, "ncclAllGather");
            *(void**)&api.GroupStart = dlsym(h, "ncclGroupStart");
            *(void**)&api.GroupEnd = dlsym(h, "ncclGroupEnd");
            *(void**)&api.GetErrorString = dlsym(h, "ncclGetErrorString");
            *(void**)&api.GetVersion = dlsym(h, "ncclGetVersion");
            if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.Broadcast || !api.AllGather ||
                !api.GroupStart || !api.GroupEnd || !api.GetErrorString)
                api.h = nullptr;
        }
        return api;
    }();
    return api.h ? const_cast<NcclApi*>(&api) : nullptr;
}

#define NC(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return fail(OBS_ERR_CUDA, "%s: %s (%s:%d)", #call, N->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

}  // namespace

struct obs_comm {
    int device = 0, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr;            // producer of the local shard -> comm stream
    cudaEvent_t started = nullptr;          // comm stream has reached the gather -> producer stream may continue
    std::vector<cudaEvent_t> chunkDone;
    int nChunks = 0;
};

extern "C" {

int obs_comm_nccl_version(void) {
    NcclApi* N = nccl_api();
    int v = 0;
    if (!N || !N->GetVersion || N->GetVersion(&v) != ncclSuccess) return -1;
    return v;
}

int obs_comm_unique_id(uint8_t* id128) {
    if (!id128) return fail(OBS_ERR_INVALID, "null argument");
    NcclApi* N = nccl_api();
    if (!N) return fail(OBS_ERR_CUDA, "libnccl.so.2 cannot be loaded: %s", dlerror());
    ncclUniqueId id;
    NC(N->GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
    return OBS_OK;
}

int obs_comm_create(const uint8_t* id128, int rank, int n_ranks, int device, obs_comm** out) {
    if (!id128 || !out) return fail(OBS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(OBS_ERR_INVALID, "rank %d of %d", rank, n_ranks);
    NcclApi* N = nccl_api();
    if (!N) return fail(OBS_ERR_CUDA, "libnccl.so.2 cannot be loaded: %s", dlerror());
    CU(cudaSetDevice(device));
    obs_comm* c = new (std::nothrow) obs_comm;
    if (!c) return fail(OBS_ERR_INVALID, "out of host memory");
    c->device = device; c->rank = rank; c->world = n_ranks;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclResult_t r = N->CommInitRank(&c->comm, n_ranks, id, rank);
    if (r != ncclSuccess) { delete c; return fail(OBS_ERR_CUDA, "ncclCommInitRank: %s", N->GetErrorString(r)); }
    // The communicator's stream has the highest priority: NCCL's few CTAs must be scheduled ahead of the pending CTAs of a
    // matching kernel whenever an SM frees up, or a rank whose NCCL kernel is resident spins -- holding its SMs -- on a peer
    // whose NCCL kernel is queued behind a persistent matching kernel.
    int prLow = 0, prHigh = 0;
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&prLow, &prHigh);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prHigh);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->started, cudaEventDisableTiming);
    if (e != cudaSuccess) { N->CommDestroy(c->comm); delete c; return fail(OBS_ERR_CUDA, "stream/event creation: %s", cudaGetErrorString(e)); }
    *out = c;
    return OBS_OK;
}

int obs_comm_destroy(obs_comm* c) {
    if (!c) return OBS_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    NcclApi* N = nccl_api();
    if (N && c->comm) N->CommDestroy(c->comm);
    for (cudaEvent_t ev : c->chunkDone) cudaEventDestroy(ev);
    if (c->ready) cudaEventDestroy(c->ready);
    if (c->started) cudaEventDestroy(c->started);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return OBS_OK;
}

int obs_comm_allgather(obs_comm* c, const uint8_t* d_local, size_t local_bytes, uint8_t* d_all, int n_chunks, void* producer_stream) {
    if (!c || !d_local || !d_all) return fail(OBS_ERR_INVALID, "null argument");
    if (n_chunks < 1 || local_bytes == 0) return fail(OBS_ERR_INVALID, "n_chunks >= 1 and local_bytes > 0 required");
    if (!is_device(d_local) || !is_device(d_all)) return fail(OBS_ERR_INVALID, "both buffers must be device memory");
    if (local_bytes % 32) return fail(OBS_ERR_INVALID, "local_bytes must be a multiple of 32 (whole descriptors)");
    NcclApi* N = nccl_api();
    if (!N) return fail(OBS_ERR_CUDA, "libnccl.so.2 cannot be loaded");
    CU(cudaSetDevice(c->device));
    while ((int)c->chunkDone.size() < n_chunks) {
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->chunkDone.push_back(ev);
    }
    c->nChunks = n_chunks;
    // the local shard must be complete before it is sent
    CU(cudaEventRecord(c->ready, (cudaStream_t)producer_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ready, 0));
    // ... and what the producer stream enqueues next (matching against the local shard) starts once the communicator's stream
    // has arrived here, i.e. behind the launch of the first NCCL kernel on every rank rather than racing it
    CU(cudaEventRecord(c->started, c->stream));
    CU(cudaStreamWaitEvent((cudaStream_t)producer_stream, c->started, 0));
    uint8_t* mine = d_all + (size_t)c->rank * local_bytes;
    if (mine != d_local) CU(cudaMemcpyAsync(mine, d_local, local_bytes, cudaMemcpyDeviceToDevice, c->stream));
    // chunk boundaries on 32-byte descriptors
    const size_t nDesc = local_bytes / 32;
    for (int k = 0; k < n_chunks; k++) {
        const size_t lo = nDesc * k / n_chunks * 32, hi = nDesc * (k + 1) / n_chunks * 32;
        if (hi > lo && c->world > 1) {
            if (n_chunks == 1) {
                NC(N->AllGather(mine, d_all, local_bytes, ncclUint8, c->comm, c->stream));      // in place
            } else {
                NC(N->GroupStart());
                for (int r = 0; r < c->world; r++) {
                    uint8_t* p = d_all + (size_t)r * local_bytes + lo;
                    ncclResult_t rr = N->Broadcast(p, p, hi - lo, ncclUint8, r, c->comm, c->stream);
                    if (rr != ncclSuccess) { N->GroupEnd(); return fail(OBS_ERR_CUDA, "ncclBroadcast: %s", N->GetErrorString(rr)); }
                }
                NC(N->GroupEnd());
            }
        }
        CU(cudaEventRecord(c->chunkDone[k], c->stream));
    }
    return OBS_OK;
}

int obs_comm_wait(obs_comm* c, int chunk, void* consumer_stream) {
    if (!c) return fail(OBS_ERR_INVALID, "null argument");
    if (chunk < 0 || chunk >= c->nChunks) return fail(OBS_ERR_INVALID, "chunk %d of %d", chunk, c->nChunks);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamWaitEvent((cudaStream_t)consumer_stream, c->chunkDone[chunk], 0));
    return OBS_OK;
}

}  // extern "C"
