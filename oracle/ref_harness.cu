// ref_harness.cu — links the REFERENCE's own log_likelihood_kernel (compiled from its source where it lies under
// /root/reference, never copied) into a shared library with a C entry point, so that tests and bench.py can run the
// reference kernel on the same B200 with fixed inputs.  TEST INFRASTRUCTURE ONLY (see oracle/pmp_oracle.c header).
//
// Build (oracle/Makefile): nvcc -DREF_SOURCE='"<path>.cu"' -DREF_VARIANT=<0 MP | 1 binary PMP> ...
// The reference program's main() is renamed and never called; its launch protocol (500_MP.cu:166-203,
// 500_PMP.cu:166-210) is restated here: zero gpu_a, upload nets (and the table, with the reference's own byte count —
// SURVEY.md quirk 1), launch <<<ceil(P/256),256>>>, download gpu_a.
#define main pmp_ref_unused_main
#include REF_SOURCE
#undef main

#include <cmath>
#include <vector>

static float* g_x = nullptr; static float* g_y = nullptr; static int g_n = 0;

extern "C" int ref_set_data(const float* x, const float* y, int n) {
    if (g_x) { cudaFree(g_x); cudaFree(g_y); }
    if (cudaMalloc(&g_x, n * sizeof(float)) != cudaSuccess) return -1;
    if (cudaMalloc(&g_y, n * sizeof(float)) != cudaSuccess) return -1;
    cudaMemcpy(g_x, x, n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(g_y, y, n * sizeof(float), cudaMemcpyHostToDevice);
    g_n = n;
    return 0;
}

// nets: [P,3] host; out_a: [P] host; kernel_ms: device time of `reps` launches / reps (CUDA events).
extern "C" int ref_loglik(const float* nets, int P, float* out_a, int reps, float* kernel_ms) {
    float *gpu_a, *gpu_nets;
    size_t net_size = P * sizeof(float), nets_size = (size_t)P * 3 * sizeof(float);
    if (cudaMalloc(&gpu_a, net_size) != cudaSuccess || cudaMalloc(&gpu_nets, nets_size) != cudaSuccess) return -1;
    cudaMemcpy(gpu_nets, nets, nets_size, cudaMemcpyHostToDevice);
    int tree_deep = (int)std::log2((double)P);
#if REF_VARIANT == 1
    int tran_table_size = P * tree_deep * 2;
    std::vector<int> tran_table(tran_table_size, -1);
    for (int deep = 0; deep < tree_deep; deep++) {          // 500_PMP.cu:170-195 (table part)
        int j = 1 << deep;
        for (int k = 0; k < j; k++) {
            tran_table[k * tree_deep * 2 + deep * 2] = k;
            tran_table[k * tree_deep * 2 + deep * 2 + 1] = k + j;
            tran_table[(k + j) * tree_deep * 2 + deep * 2] = k + j;
            tran_table[(k + j) * tree_deep * 2 + deep * 2 + 1] = k;
            if (deep - 1 > -1 && tran_table[(k + j) * tree_deep * 2 + (deep - 1) * 2] == -1)
                for (int index = 0; index < deep; index++) {
                    tran_table[(k + j) * tree_deep * 2 + index * 2] = tran_table[k * tree_deep * 2 + index * 2];
                    tran_table[(k + j) * tree_deep * 2 + index * 2 + 1] = tran_table[k * tree_deep * 2 + index * 2 + 1];
                }
        }
    }
    float* gpu_tran_table;
    cudaMalloc(&gpu_tran_table, tran_table_size * sizeof(float));
    cudaMemset(gpu_tran_table, 0, tran_table_size * sizeof(float));   // fresh cudaMalloc memory reads as zero in the reference runs
    cudaMemcpy(gpu_tran_table, tran_table.data(), tran_table_size, cudaMemcpyHostToDevice);   // byte count as shipped (500_PMP.cu:198)
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blockSize = 256, gridSize = (P + blockSize - 1) / blockSize;
    float total = 0.f;
    for (int r = 0; r < reps; ++r) {
        cudaMemset(gpu_a, 0, net_size);
        cudaEventRecord(e0);
#if REF_VARIANT == 1
        log_likelihood_kernel<<<gridSize, blockSize>>>(g_x, g_y, gpu_a, gpu_nets, gpu_tran_table, P, (int)net_size, g_n, tree_deep);
#else
        log_likelihood_kernel<<<gridSize, blockSize>>>(g_x, g_y, gpu_a, gpu_nets, P, (int)net_size, g_n, tree_deep);
#endif
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
        float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
    }
    cudaMemcpy(out_a, gpu_a, net_size, cudaMemcpyDeviceToHost);
    if (kernel_ms) *kernel_ms = total / reps;
    cudaFree(gpu_a); cudaFree(gpu_nets);
#if REF_VARIANT == 1
    cudaFree(gpu_tran_table);
#endif
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
