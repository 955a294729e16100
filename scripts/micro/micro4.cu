// micro4.cu — numerics probe for the tcgen05 formulation of the linear-Gaussian residual (not part of the product):
// r[node][point] = y - b0 - b1*x as ONE bf16 UMMA (M=128 nodes, N=64 points, K=16) over 3-way bf16 splits of every operand.
// Prints the error of r and of the per-node sum of squares against binary64, for both readings of the descriptor's LBO/SBO fields.
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void split3(float v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
    h = __float2bfloat16_rn(v); float r1 = v - __bfloat162float(h);
    m = __float2bfloat16_rn(r1); float r2 = r1 - __bfloat162float(m);
    l = __float2bfloat16_rn(r2);
}
__device__ __forceinline__ uint32_t canon_off(int row, int k) { return (uint32_t)((row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2); }
__global__ void __launch_bounds__(128, 1) k(const float* x, const float* y, const float* nodes, float* r_out, int lbo, int sbo) {
    __shared__ __align__(1024) uint8_t sA[128 * 32];
    __shared__ __align__(1024) uint8_t sB[64 * 32];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const __nv_bfloat16 one = __float2bfloat16_rn(1.f), zero = __float2bfloat16_rn(0.f);
    {   // A: node tid
        __nv_bfloat16 h0, m0, l0, h1, m1, l1;
        split3(-nodes[3 * tid], h0, m0, l0); split3(-nodes[3 * tid + 1], h1, m1, l1);
        __nv_bfloat16 a[16] = {one, one, one, h0, m0, l0, h1, m1, h1, l1, m1, h1, l1, m1, l1, zero};
        for (int kk = 0; kk < 16; ++kk) *reinterpret_cast<__nv_bfloat16*>(sA + canon_off(tid, kk)) = a[kk];
    }
    if (tid < 64) {
        __nv_bfloat16 xh, xm, xl, yh, ym, yl;
        split3(x[tid], xh, xm, xl); split3(y[tid], yh, ym, yl);
        __nv_bfloat16 b[16] = {yh, ym, yl, one, one, one, xh, xh, xm, xh, xm, xl, xm, xl, xl, zero};
        for (int kk = 0; kk < 16; ++kk) *reinterpret_cast<__nv_bfloat16*>(sB + canon_off(tid, kk)) = b[kk];
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes → visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        auto desc = [&](const void* p) { return (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46); };
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem), "l"(desc(sA)), "l"(desc(sB)), "r"(idesc), "r"(0) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (spins > 100000000u) __trap();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    for (int half = 0; half < 2; ++half) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                     "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                       "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
                       "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + half * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) r_out[tid * 64 + half * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(-1, 1); std::normal_distribution<double> N(0, 1);
    for (int scenario = 0; scenario < 2; ++scenario) {
        std::vector<float> x(64), y(64), nodes(128 * 3);
        for (int i = 0; i < 64; ++i) { x[i] = (float)U(g); y[i] = (float)(-1 + 2 * x[i] + 0.5 * N(g)); }
        for (int p = 0; p < 128; ++p) {
            const double c[2][3] = {{1, 1, 1}, {-1, 2, 0.5}};
            for (int j = 0; j < 3; ++j) nodes[3 * p + j] = (float)(c[scenario][j] + 0.01 * N(g));
        }
        float *dx, *dy, *dn, *dr;
        cudaMalloc(&dx, 256); cudaMalloc(&dy, 256); cudaMalloc(&dn, 128 * 12); cudaMalloc(&dr, 128 * 64 * 4);
        cudaMemcpy(dx, x.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(dy, y.data(), 256, cudaMemcpyHostToDevice); cudaMemcpy(dn, nodes.data(), 128 * 12, cudaMemcpyHostToDevice);
        for (int variant = 0; variant < 2; ++variant) {
            int lbo = variant == 0 ? 128 : 256, sbo = variant == 0 ? 256 : 128;
            cudaMemset(dr, 0, 128 * 64 * 4);
            k<<<1, 128>>>(dx, dy, dn, dr, lbo, sbo);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> r(128 * 64);
            cudaMemcpy(r.data(), dr, r.size() * 4, cudaMemcpyDeviceToHost);
            double maxabs = 0, sumerr = 0, maxrel_ss = 0, max_fma = 0, sum_fma = 0, maxrel_ss_fma = 0;
            for (int p = 0; p < 128; ++p) {
                double ss = 0, ss_tc = 0, ss_f = 0;
                for (int i = 0; i < 64; ++i) {
                    double ref = (double)y[i] - (double)nodes[3 * p] - (double)nodes[3 * p + 1] * (double)x[i];
                    float rf = y[i] - fmaf(nodes[3 * p + 1], x[i], nodes[3 * p]);
                    double err = (double)r[p * 64 + i] - ref, errf = (double)rf - ref;
                    maxabs = fmax(maxabs, fabs(err)); sumerr += err; max_fma = fmax(max_fma, fabs(errf)); sum_fma += errf;
                    ss += ref * ref; ss_tc += (double)r[p * 64 + i] * r[p * 64 + i]; ss_f += (double)rf * rf;
                }
                maxrel_ss = fmax(maxrel_ss, fabs(ss_tc - ss) / ss); maxrel_ss_fma = fmax(maxrel_ss_fma, fabs(ss_f - ss) / ss);
            }
            printf("scenario %d lbo=%d sbo=%d: %s | tc: max|err r| %.3e mean err %.3e max rel err sumsq %.3e | f32 fma path: max %.3e mean %.3e sumsq %.3e\n", scenario, lbo, sbo,
                   cudaGetErrorString(e), maxabs, sumerr / (128 * 64), maxrel_ss, max_fma, sum_fma / (128 * 64), maxrel_ss_fma);
        }
    }
    return 0;
}
