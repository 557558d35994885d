// Host-side helpers shared by the C-ABI translation units (api.cu, matcher_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstddef>

namespace obsdetail {
int fail(int code, const char* fmt, ...);     // records the thread-local message of obs_last_error(), returns code
bool is_pinned(const void* p);                // page-locked host memory
bool is_device(const void* p);                // device (or managed) memory

// The dynamic shared-memory limit of a kernel (cudaFuncAttributeMaxDynamicSharedMemorySize) is process-wide per kernel and
// device, not per handle: it is raised ONCE to the device's opt-in maximum (minus the kernel's static shared memory) and never
// lowered, so handles with different shapes or capacities cannot invalidate each other's launches.  `done` is a per-call-site
// bit mask of the devices already configured; racing threads write the same value.
cudaError_t allow_max_smem(const void* kernel, std::atomic<unsigned long long>& done);
int max_dynamic_smem(const void* kernel);      // that limit for the current device (bytes), 0 on error

// Kernel launch with or without the programmatic-stream-serialization attribute (PDL): with it, the kernel may be scheduled while
// its predecessor on the stream is still running and synchronises itself with griddepcontrol.wait (common.cuh pdl_entry).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
bool pdl_enabled();                 // process-wide switches (obs_set_option "pdl" / "graphs"), default on
void set_pdl_enabled(bool on);
void set_capturing(bool on);       // this thread is inside a stream capture: launch_k drops the PDL attribute
bool graphs_enabled();
void set_graphs_enabled(bool on);
extern std::atomic<unsigned long long> g_allocEpoch;     // bumped by every device (re)allocation: captured graphs hold raw pointers

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        g_allocEpoch.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) { cudaFree(p); g_allocEpoch.fetch_add(1, std::memory_order_relaxed); } p = nullptr; n = 0; }
};

template <typename T> struct PinBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t ensure(size_t count) {
        if (count <= n) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        g_allocEpoch.fetch_add(1, std::memory_order_relaxed);       // captured graphs hold the staging pointers too
        cudaError_t e = cudaMallocHost((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};
}  // namespace obsdetail
using obsdetail::fail;
using obsdetail::is_pinned;
using obsdetail::is_device;
using obsdetail::DevBuf;
using obsdetail::launch_k;
using obsdetail::pdl_enabled;
using obsdetail::graphs_enabled;
using obsdetail::g_allocEpoch;
using obsdetail::PinBuf;

#define OBS_ALLOW_MAX_SMEM(kernel)                                                                  \
    ([]() -> cudaError_t {                                                                          \
        static std::atomic<unsigned long long> done_{0};                                            \
        return obsdetail::allow_max_smem((const void*)(kernel), done_);                             \
    }())

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(OBS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
