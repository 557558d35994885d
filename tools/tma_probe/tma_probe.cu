// Probe: which 3-D uint8 TMA box loads does this GPU accept?  (tools only; built by hand: nvcc -arch=sm_100a tma_probe.cu -o tma_probe)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BW, int BH>
__global__ void k(const CUtensorMap* maps, int x, int y, int z, uint8_t* out) {
    __shared__ __align__(128) uint8_t buf[(BW * BH + 127) / 128 * 128];
    __shared__ __align__(8) uint64_t mbar;
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar), dst = (uint32_t)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mb), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     :: "r"(dst), "l"(reinterpret_cast<uint64_t>(maps)), "r"(x), "r"(y), "r"(z), "r"(mb) : "memory");
    }
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(mb), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}
template <int BW, int BH>
int run(EncodeTiledFn enc, uint8_t* d, const std::vector<uint8_t>& h, int pitch, int rows, int nimg, size_t imgStride, int x, int y, int z) {
    CUtensorMap m;
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)nimg};
    cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)imgStride};
    cuuint32_t box[3] = {BW, BH, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %dx%d: encode failed %d\n", BW, BH, (int)r); return 1; }
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    uint8_t* dout; cudaMalloc(&dout, BW * BH);
    k<BW, BH><<<1, 128>>>(dm, x, y, z, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("box %dx%d at (%d,%d,%d): %s\n", BW, BH, x, y, z, cudaGetErrorString(e)); return 2; }
    std::vector<uint8_t> o(BW * BH); cudaMemcpy(o.data(), dout, BW * BH, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BH; r2++) for (int c = 0; c < BW; c++) {
        const int gx = x + c, gy = y + r2;
        const uint8_t want = (gx < pitch && gy < rows) ? h[(size_t)z * imgStride + (size_t)gy * pitch + gx] : 0;
        bad += o[r2 * BW + c] != want;
    }
    printf("box %dx%d at (%d,%d,%d): ok, %d mismatching bytes\n", BW, BH, x, y, z, bad);
    return 0;
}
int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 0;
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int pitch = 1280, rows = 376, nimg = 3; const size_t imgStride = (size_t)pitch * rows + 256;
    std::vector<uint8_t> h(imgStride * nimg);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 2654435761u >> 13);
    uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    switch (which) {
        case 0: return run<48, 37>(enc, d, h, pitch, rows, nimg, imgStride, 32, 20, 1);    // aligned x
        case 1: return run<48, 37>(enc, d, h, pitch, rows, nimg, imgStride, 37, 21, 2);    // unaligned x
        case 2: return run<64, 37>(enc, d, h, pitch, rows, nimg, imgStride, 37, 21, 2);
        case 3: return run<64, 32>(enc, d, h, pitch, rows, nimg, imgStride, 32, 16, 0);
        case 4: return run<48, 37>(enc, d, h, pitch, rows, nimg, imgStride, 1250, 350, 2);  // partly outside
    }
    return 0;
}
