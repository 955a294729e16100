// micro7.cu — issue rate of the packed-FMA operand shapes of the CNN convolution kernel (not part of the product).
//   A  acc = fma2( in.F32 (scalar broadcast), w.F32x2, acc.F32x2 )      the convolution kernel's shape: 5 source registers
//   B  acc = fma2( in.F32x2 (pre-duplicated),  w.F32x2, acc.F32x2 )      all-packed: 6 source registers
//   C  as A with 40 accumulators per thread (the kernel's accumulator count) instead of 16
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
template <int MODE, int NACC, int THREADS> __global__ void __launch_bounds__(THREADS) k(float* out, int iters, float s) {
    float2 acc[NACC];
    const float tw = MODE == 2 ? 0.f : 1e-12f * threadIdx.x;    // MODE 2: thread-invariant weights -> uniform registers (the constant-bank route)
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(s + i, s - i);
    float2 w[5]; for (int j = 0; j < 5; ++j) w[j] = make_float2(1.0f + 1e-7f * j + tw, 1.0f - 1e-7f * j - tw);   // thread-dependent: vector registers, like tap vectors loaded from shared memory
    float in[8]; for (int p = 0; p < 8; ++p) in[p] = 1e-7f * (p + 1) * s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            const float v = in[(i / 5) & 7];
            if (MODE == 0 || MODE == 2) acc[i] = ffma2(make_float2(v, v), w[i % 5], acc[i]);
            else { float2 vv = make_float2(in[(i / 5) & 7], in[((i / 5) + 1) & 7]); acc[i] = ffma2(vv, w[i % 5], acc[i]); }
        }
        in[it & 7] += 1e-9f;
    }
    float r = 0; for (int i = 0; i < NACC; ++i) r += acc[i].x + acc[i].y;
    if (r == 123.456f) out[0] = r;
}
template <int MODE, int NACC, int THREADS> void run(const char* name, int ctas_per_sm) {
    float* d; cudaMalloc(&d, 4); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096, blocks = 148 * ctas_per_sm;
    k<MODE, NACC, THREADS><<<blocks, THREADS>>>(d, iters, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE, NACC, THREADS><<<blocks, THREADS>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * THREADS * iters * NACC * 2.0;
    printf("%-70s %8.3f ms  %6.2f lane-FMAs/clk/SM (128 = peak)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
// scalar FFMA with the same data flow: 2*NACC scalar accumulators, 10 scalar tap weights in vector registers, scalar inputs
template <int NACC, int THREADS> __global__ void __launch_bounds__(THREADS) ks(float* out, int iters, float s) {
    float acc[2 * NACC];
    for (int i = 0; i < 2 * NACC; ++i) acc[i] = s + i;
    float w[10]; for (int j = 0; j < 10; ++j) w[j] = 1.0f + 1e-7f * j + 1e-12f * threadIdx.x;
    float in[8]; for (int p = 0; p < 8; ++p) in[p] = 1e-7f * (p + 1) * s;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * NACC; ++i) acc[i] = fmaf(in[(i / 10) & 7], w[i % 10], acc[i]);
        in[it & 7] += 1e-9f;
    }
    float r = 0; for (int i = 0; i < 2 * NACC; ++i) r += acc[i];
    if (r == 123.456f) out[0] = r;
}
template <int NACC, int THREADS> void runs(const char* name, int ctas_per_sm) {
    float* d; cudaMalloc(&d, 4); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 4096, blocks = 148 * ctas_per_sm;
    ks<NACC, THREADS><<<blocks, THREADS>>>(d, iters, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0); ks<NACC, THREADS><<<blocks, THREADS>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * THREADS * iters * NACC * 2.0;
    printf("%-70s %8.3f ms  %6.2f lane-FMAs/clk/SM (128 = peak)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    run<0, 16, 256>("A scalar-broadcast input, 16 acc, 256 thr x 4 CTAs/SM", 4);
    run<1, 16, 256>("B packed input,           16 acc, 256 thr x 4 CTAs/SM", 4);
    run<0, 40, 256>("A scalar-broadcast input, 40 acc, 256 thr x 1 CTA/SM (8 warps)", 1);
    run<1, 40, 256>("B packed input,           40 acc, 256 thr x 1 CTA/SM (8 warps)", 1);
    run<0, 40, 256>("A scalar-broadcast input, 40 acc, 256 thr x 2 CTAs/SM (16 warps)", 2);
    run<0, 50, 256>("A scalar-broadcast input, 50 acc, 256 thr x 1 CTA/SM (8 warps)", 1);
    run<2, 40, 256>("U scalar-broadcast input, uniform-register weights, 40 acc, 8 warps", 1);
    run<2, 50, 256>("U scalar-broadcast input, uniform-register weights, 50 acc, 8 warps", 1);
    runs<40, 256>("S scalar FFMA,            80 acc, 256 thr x 1 CTA/SM (8 warps)", 1);
    runs<50, 256>("S scalar FFMA,           100 acc, 256 thr x 1 CTA/SM (8 warps)", 1);
    runs<20, 256>("S scalar FFMA,            40 acc, 256 thr x 2 CTAs/SM (16 warps)", 2);
    return 0;
}
