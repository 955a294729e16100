// micro2.cu — issue-rate probes for the packed FP32 forms used by the sweep's inner loop (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
// MODE 0: scalar FFMA, 16 chains. 1: FFMA2 packed operands, 8 chains. 2: FFMA2 with scalar-broadcast multiplier+addend.
// 3: FADD2. 4: the sweep triple (ffma2 bcast, fsub2, ffma2 acc) on register data. 5: triple with the subtract as ffma2(yh,-1,y).
// 6: scalar triple (ffma, fsub, ffma).
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float2 a[8]; float f[16];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(s + i, s - i);
    for (int i = 0; i < 16; ++i) f[i] = s + i;
    float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f), neg1 = make_float2(-1.f, -1.f);
    float b0 = s * 0.5f, b1 = s * 0.25f;
    float2 x = make_float2(s, s + 1), y = make_float2(s + 2, s + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { f[2 * i] = fmaf(f[2 * i], m.x, c.x); f[2 * i + 1] = fmaf(f[2 * i + 1], m.y, c.y); }
            if (MODE == 1) a[i] = ffma2(a[i], m, c);
            if (MODE == 2) a[i] = ffma2(make_float2(b1, b1), a[i], make_float2(b0, b0));
            if (MODE == 3) a[i] = fsub2(a[i], c);
            if (MODE == 4) { float2 yh = ffma2(make_float2(b1, b1), x, make_float2(b0 + i, b0 + i)); float2 d = fsub2(y, yh); a[i] = ffma2(d, d, a[i]); }
            if (MODE == 5) { float2 yh = ffma2(make_float2(b1, b1), x, make_float2(b0 + i, b0 + i)); float2 d = ffma2(yh, neg1, y); a[i] = ffma2(d, d, a[i]); }
            if (MODE == 6) { float yh0 = fmaf(b1, x.x, b0 + i), yh1 = fmaf(b1, x.y, b0 + i); float d0 = y.x - yh0, d1 = y.y - yh1; f[2 * i] = fmaf(d0, d0, f[2 * i]); f[2 * i + 1] = fmaf(d1, d1, f[2 * i + 1]); }
        }
        if (MODE >= 4) { x.x += 1e-3f; y.y -= 1e-3f; }
    }
    float r = 0; for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y; for (int i = 0; i < 16; ++i) r += f[i];
    if (r == 123.456f) out[0] = r;
}
template <int MODE> void run(const char* name, double lane_ops_per_iter) {
    float* d; cudaMalloc(&d, 4); cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 8192, blocks = 148 * 6;
    k<MODE><<<blocks, 256>>>(d, iters, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * 256 * iters * lane_ops_per_iter;   // FP32 lane-ops (1 FMA or 1 ADD on one lane-half)
    printf("%-58s %8.3f ms  %7.2f T lane-op/s  (%.2f lane-ops/clk/SM @1.965GHz)\n", name, ms, ops / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(d);
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    run<0>("FFMA scalar x16", 16);
    run<1>("FFMA2 packed operands x8", 16);
    run<2>("FFMA2 scalar-broadcast b1,b0 x8", 16);
    run<3>("FADD2 x8", 16);
    run<4>("sweep triple: FFMA2(bcast) + FADD2 + FFMA2(acc) x8", 48);
    run<5>("sweep triple, subtract as FFMA2(yh,-1,y) x8", 48);
    run<6>("scalar triple: FFMA + FADD + FFMA x16", 48);
    return 0;
}
