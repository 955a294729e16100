// micro3.cu — TMEM read (tcgen05.ld) throughput probe, alone and with a square-accumulate epilogue (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                 "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                   "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(addr));
}
__device__ __forceinline__ uint32_t consume32(const uint32_t* v) { uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) r ^= v[i];
    return r; }
__device__ __forceinline__ void tmem_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
// MODE 0: ld.x32 + wait per step.  1: two ld.x32 in flight, one wait.  2: MODE 1 + FFMA2 square-accumulate of all 64 values.
// 3: MODE 0 + FFMA2 on 32 values, software-pipelined (load next while squaring current).
template <int MODE, int THREADS> __global__ void __launch_bounds__(THREADS, 1) k(float* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp & 3) * 32u << 16);
    const int nw = blockDim.x >> 5, grp = warp >> 2, ngrp = nw >> 2;       // warps of a group share a lane quarter; groups split columns
    unsigned long long acc[4] = {0, 0, 0, 0};
    uint32_t v[64];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
            for (int c = grp * 32; c < 512; c += 32 * ngrp) { tmem_ld32(base + c, v); tmem_wait(); acc[0] ^= consume32(v); }
        } else if (MODE == 1 || MODE == 2) {
            for (int c = grp * 64; c < 512; c += 64 * ngrp) {
                tmem_ld32(base + c, v); tmem_ld32(base + c + 32, v + 32); tmem_wait();
                if (MODE == 1) { acc[0] ^= consume32(v); acc[1] ^= consume32(v + 32); }
                else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) { unsigned long long r = ((unsigned long long)v[2 * j + 1] << 32) | v[2 * j]; acc[j & 3] = ffma2(r, r, acc[j & 3]); }
                }
            }
        } else {
            uint32_t w[32];
            tmem_ld32(base + grp * 32, v); tmem_wait();
            for (int c = grp * 32; c < 512; c += 64 * ngrp) {
                tmem_ld32(base + ((c + 32 * ngrp) & 511), w);
#pragma unroll
                for (int j = 0; j < 16; ++j) { unsigned long long r = ((unsigned long long)v[2 * j + 1] << 32) | v[2 * j]; acc[j & 3] = ffma2(r, r, acc[j & 3]); }
                tmem_wait();
                tmem_ld32(base + ((c + 64 * ngrp) & 511), v);
#pragma unroll
                for (int j = 0; j < 16; ++j) { unsigned long long r = ((unsigned long long)w[2 * j + 1] << 32) | w[2 * j]; acc[j & 3] = ffma2(r, r, acc[j & 3]); }
                tmem_wait();
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    unsigned long long r = acc[0] ^ acc[1] ^ acc[2] ^ acc[3];
    if ((uint32_t)r == 0x9abcdef0u) out[0] = 1.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "n"(512) : "memory");
}
template <int MODE, int threads> void run(const char* name) {
    float* d; long long* c; cudaMalloc(&d, 4); cudaMalloc(&c, 148 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 2000;
    k<MODE, threads><<<148, threads>>>(d, c, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE, threads><<<148, threads>>>(d, c, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
    double bytes_per_sm = (double)iters * 128 * 512 * 4;      // every iteration reads the SM's whole TMEM once
    printf("%-64s warps=%2d  %8.3f ms  %7.1f B/clk/SM (clock64)  %6.1f GB/s/SM  err=%s\n", name, threads / 32, ms, bytes_per_sm / (double)h[0], bytes_per_sm / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(c);
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
#define ALL(T) run<0, T>("ld.x32 + wait"); run<1, T>("2 x ld.x32 + wait"); run<2, T>("2 x ld.x32 + wait + 32 FFMA2 (square-accumulate)"); run<3, T>("pipelined ld.x32 / 16 FFMA2");
    ALL(128) ALL(256) ALL(512) ALL(1024)
    return 0;
}
