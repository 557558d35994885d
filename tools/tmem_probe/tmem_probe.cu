// Probe of tcgen05.st ... .unpack::16b / tcgen05.ld ... .pack::16b semantics on 32-bit TMEM columns (run on the B200):
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) k(uint32_t* out) {
    __shared__ uint32_t tb;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&tb)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t = tb + ((uint32_t)(warp * 32) << 16);
    // 1) fill columns 0..15 with 0xdeadbeef (plain 32-bit store), then overwrite columns 0..15 through 8 packed registers (unpack::16b)
    uint32_t a[16];
    for (int i = 0; i < 16; i++) a[i] = 0xdeadbeefu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(t), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t p[8];
    for (int i = 0; i < 8; i++) p[i] = (uint32_t)(0x1000 + 2 * i) | ((uint32_t)(0x9000 + 2 * i + 1) << 16);      // high halves have bit 15 set
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.unpack::16b.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(t), "r"(p[0]), "r"(p[1]), "r"(p[2]), "r"(p[3]), "r"(p[4]), "r"(p[5]), "r"(p[6]), "r"(p[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(t) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (threadIdx.x == 37) for (int i = 0; i < 16; i++) out[i] = r[i];
    // 2) 32-bit columns 16..31 hold 0x00012345 + i and negative values; read them back packed (pack::16b)
    for (int i = 0; i < 16; i++) a[i] = (i & 1) ? (uint32_t)(-64 * i) : (uint32_t)(0x00012340 + i);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(t + 16), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t q[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(t + 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (threadIdx.x == 37) for (int i = 0; i < 8; i++) out[16 + i] = q[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(64u) : "memory");
}

int main() {
    uint32_t* d; cudaMalloc(&d, 64 * 4); cudaMemset(d, 0, 64 * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    uint32_t h[24]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("unpack::16b store, plain load of columns 0..15:\n");
    for (int i = 0; i < 16; i++) printf(" %08x", h[i]);
    printf("\nplain store of columns 16..31 (even: 0x12340+i, odd: -64 i), pack::16b load (8 registers):\n");
    for (int i = 0; i < 8; i++) printf(" %08x", h[16 + i]);
    printf("\n");
    return 0;
}
