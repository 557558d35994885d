// Roofline denominators measured on the device the library runs on (the
// INT/popc ceiling of the Hamming matchers has no published number for B200).
//   mode 0: register-resident 256-bit Hamming distances, the plain instruction mix of
//           ORBmatcher::DescriptorDistance on a GPU: 8 LOP3 (xor) + 8 POPC + adds per distance.
//   mode 1: the same distance through three carry-save adders (5 POPC per distance).
//   mode 3: four carry-save adders (4 POPC per distance).
//   mode 2: POPC only (8 per "distance"), the pipe ceiling.
#include "../../include/obslam_b200.h"
#include "host_util.h"
#include "knn2_tc.h"

namespace {

__device__ __forceinline__ void csa(uint32_t a, uint32_t b, uint32_t c, uint32_t& s, uint32_t& cy) {
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}

template <int MODE>
__global__ void __launch_bounds__(256) k_popc_peak(const uint32_t* seed, int iters, int* sink) {
    uint32_t q[8], b[8];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) { q[k] = seed[(t + k) & 1023]; b[k] = seed[(t * 7 + k * 13) & 1023] | 1u; }
    int acc = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            // a new database descriptor every distance: one IMAD per word (FMA pipe, where the real kernels
            // spend a shared-memory load instead), so nothing of the distance itself can be hoisted or reused
#pragma unroll
            for (int k = 0; k < 8; k++) b[k] = b[k] * 0x9E3779B1u + 0x7F4A7C15u;
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) acc += __popc(q[k] ^ b[k]);
            } else if (MODE == 1) {
                uint32_t x[8];
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = q[k] ^ b[k];
                uint32_t s1, c1, s2, c2, s3, c3;
                csa(x[0], x[1], x[2], s1, c1);
                csa(x[3], x[4], x[5], s2, c2);
                csa(s1, s2, x[6], s3, c3);
                acc += __popc(s3) + __popc(x[7]) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
            } else if (MODE == 3) {
                uint32_t x[8];
#pragma unroll
                for (int k = 0; k < 8; k++) x[k] = q[k] ^ b[k];
                uint32_t s1, c1, s2, c2, s3, c3, s4, c4;
                csa(x[0], x[1], x[2], s1, c1);
                csa(x[3], x[4], x[5], s2, c2);
                csa(s1, s2, x[6], s3, c3);
                csa(c1, c2, c3, s4, c4);
                acc += __popc(s3) + __popc(x[7]) + 2 * __popc(s4) + 4 * __popc(c4);
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++) acc += __popc(b[k]);
            }
        }
    }
    if (acc == 0x7fffffff) sink[0] = acc;
}

}  // namespace

extern "C" int obs_microbench_popc(int device, int mode, double* gdist_per_s) {
    if (!gdist_per_s || mode < 0 || mode > 3) return fail(OBS_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    uint32_t h[1024];
    uint32_t s = 0x9e3779b9u;
    for (int i = 0; i < 1024; i++) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; h[i] = s; }
    uint32_t* d = nullptr; int* sink = nullptr;
    CU(cudaMalloc(&d, sizeof(h)));
    CU(cudaMalloc(&sink, 4));
    CU(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    const int blocks = prop.multiProcessorCount * 8, iters = 4096;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(e0));
        if (mode == 0) k_popc_peak<0><<<blocks, 256>>>(d, iters, sink);
        else if (mode == 1) k_popc_peak<1><<<blocks, 256>>>(d, iters, sink);
        else if (mode == 3) k_popc_peak<3><<<blocks, 256>>>(d, iters, sink);
        else k_popc_peak<2><<<blocks, 256>>>(d, iters, sink);
        CU(cudaEventRecord(e1));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d); cudaFree(sink);
    *gdist_per_s = (double)blocks * 256 * iters * 8 / (best * 1e-3) / 1e9;
    return OBS_OK;
}

// Measured int8 tensor-core rate in the shape of k_knn2_tc (csrc/knn2_tc.cu): the denominator of the knn2 roofline.
extern "C" int obs_microbench_imma(int device, double* tops) {
    if (!tops) return fail(OBS_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(device));
    CU(knn2_tc_peak(tops));
    return OBS_OK;
}

// ---- measured pipe ceilings for the stages of the extractor: the kernels are bound by instruction issue,
// by the ALU pipe (LOP3 / SHF / PRMT / VIMNMX ...: one warp instruction per two cycles per scheduler) or by shared-memory
// wavefronts (one per cycle per SM), so those three rates are measured on the device the library runs on:
//   MODE 0  issue:  independent LOP3 (ALU pipe) and IMAD (FMA pipe) chains interleaved 1:1
//   MODE 1  ALU:    independent LOP3 chains only
//   MODE 2  LDS:    conflict-free 32-bit shared-memory loads (one wavefront each), results folded by 3-input LOP3
namespace {

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe_peak(int iters, unsigned* sink) {
    __shared__ unsigned buf[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) buf[i] = i * 2654435761u;
    __syncthreads();
    unsigned a[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = threadIdx.x * 0x9E3779B1u + k; m[k] = blockIdx.x + 3 * k + 1; }
    const unsigned x = threadIdx.x | 1u, y = blockIdx.x * 7u + 5u;
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf) + 4u * (threadIdx.x & 31u);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(x), "r"(y));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[k]) : "r"(x), "r"(y));
                }
            } else if (MODE == 1) {
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(x), "r"(y));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(m[k]) : "r"(y), "r"(x));
                }
            } else {
                unsigned v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(base + 128u * k) : "memory");
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(v[2 * k]), "r"(v[2 * k + 1]));
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) r ^= a[k] ^ m[k];
    if (r == 0x12345678u) sink[0] = r;
}

template <int MODE> cudaError_t run_pipe_peak(int blocks, int iters, unsigned* sink, double* per_s) {
    cudaEvent_t e0, e1;
    cudaError_t e;
    if ((e = cudaEventCreate(&e0)) != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) return e;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        k_pipe_peak<MODE><<<blocks, 256>>>(iters, sink);
        cudaEventRecord(e1);
        if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e != cudaSuccess) return e;
    // warp instructions (modes 0, 1: 16 per inner round) or wavefronts (mode 2: 16 per inner round) per second
    *per_s = (double)blocks * 8 /*warps*/ * (double)iters * 8 * 16 / (best * 1e-3);
    return cudaGetLastError();
}

}  // namespace

extern "C" int obs_microbench_pipes(int device, double* issue_ginst_per_s, double* alu_ginst_per_s, double* lds_gwavefronts_per_s) {
    if (!issue_ginst_per_s || !alu_ginst_per_s || !lds_gwavefronts_per_s) return fail(OBS_ERR_INVALID, "null argument");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    unsigned* sink = nullptr;
    CU(cudaMalloc(&sink, 4));
    const int blocks = prop.multiProcessorCount * 8, iters = 2048;
    double v[3] = {0, 0, 0};
    cudaError_t e = run_pipe_peak<0>(blocks, iters, sink, &v[0]);
    if (e == cudaSuccess) e = run_pipe_peak<1>(blocks, iters, sink, &v[1]);
    if (e == cudaSuccess) e = run_pipe_peak<2>(blocks, iters, sink, &v[2]);
    cudaFree(sink);
    CU(e);
    *issue_ginst_per_s = v[0] / 1e9; *alu_ginst_per_s = v[1] / 1e9; *lds_gwavefronts_per_s = v[2] / 1e9;
    return OBS_OK;
}
