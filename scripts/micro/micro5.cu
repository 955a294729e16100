// micro5.cu — tcgen05.mma issue-to-completion time for K=16 bf16 MMAs (M=128) under different smem layouts (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int nmma, int N, int layout, int same_half) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        uint64_t lt = 0, lbo = 128, sbo = 256;
        if (layout == 1) { lt = 6; lbo = 16; sbo = 256; }        // SWIZZLE_32B: 8 rows x 32 B atoms
        if (layout == 2) { lt = 2; lbo = 16; sbo = 1024; }       // SWIZZLE_128B: 8 rows x 128 B atoms (K=16 slice of a 64-wide row)
        if (layout == 3) { lt = 4; lbo = 16; sbo = 512; }        // SWIZZLE_64B
        auto desc = [&](const void* p) { return (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4) | ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (lt << 61); };
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = desc(sm), db = desc(sm + 16 * 1024);
        long long t0 = clock64();
        for (int i = 0; i < nmma; ++i) {
            uint32_t d = tmem + (same_half ? 0 : (i & 1) * 256);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(0) : "memory");
        }
        long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t2 = clock64();
        cyc[2 * blockIdx.x] = t1 - t0; cyc[2 * blockIdx.x + 1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    long long* c; cudaMalloc(&c, 148 * 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const char* names[] = {"no swizzle (LBO 128, SBO 256)", "SWIZZLE_32B", "SWIZZLE_128B (K=16 slice)", "SWIZZLE_64B"};
    for (int layout = 0; layout < 4; ++layout)
        for (int N : {64, 128, 256})
            for (int nm : {1, 8, 64}) {
                k<<<148, 128, 64 * 1024>>>(c, nm, N, layout, 0);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[4]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
                printf("%-32s N=%3d  %2d MMAs: issue %6lld cyc, done %6lld cyc  (%.1f cyc/MMA)  %s\n", names[layout], N, nm, h[0], h[1], (double)h[1] / nm, cudaGetErrorString(e));
            }
    return 0;
}
