// ref_harness.cu — links the REFERENCE's own log_likelihood_kernel (compiled from its source where it lies under
// /root/reference, never copied) into a shared library with a C entry point, so that tests and bench.py can run the
// reference kernel on the same B200 with fixed inputs.  TEST INFRASTRUCTURE ONLY (see oracle/pmp_oracle.c header).
//
// Build (oracle/Makefile): nvcc -DREF_SOURCE='"<path>.cu"' -DREF_VARIANT=<0 MP | 1 binary PMP | 2 general PMP (conv_pmp.cu) | 3 MH (conv_mh.cu)> ...
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

// Transition table of the binary tree (500_PMP.cu:170-195, table part) or of the (N_step+1)-ary tree (conv_pmp.cu:182-221), restated:
// row [node][level] holds the (from, to) pairs of `node`'s ancestor at that level against the other members of the ancestor's
// group; a node created at level L inherits the rows of the levels below L from its parent.
static std::vector<int> build_table(int P, int depth, int n_step) {
    const int b = n_step + 1, row = depth * n_step * 2;
    std::vector<int> t((size_t)P * row, -1);
    long long s = 1;
    for (int lv = 0; lv < depth; ++lv, s *= b)
        for (long long k = 0; k < s; ++k) {
            std::vector<long long> grp(b);
            for (int j = 0; j < b; ++j) grp[j] = k + s * j;
            for (int a = 0; a < b; ++a) {
                int e = 0;
                for (int o = 0; o < b; ++o)
                    if (o != a) { t[grp[a] * row + lv * n_step * 2 + 2 * e] = (int)grp[a]; t[grp[a] * row + lv * n_step * 2 + 2 * e + 1] = (int)grp[o]; ++e; }
                if (a > 0) for (int q = 0; q < lv * n_step * 2; ++q) t[grp[a] * row + q] = t[k * row + q];
            }
        }
    return t;
}

// nets: [P,3] host; out_a: [P] host; kernel_ms: device time of `reps` launches / reps (CUDA events).
// depth / n_step describe the tree of the table variants (ignored by MP / MH); fix_table = 0 reproduces the shipped upload
// (int table behind a float*, byte count = element count: 500_PMP.cu:130-131,198, conv_pmp.cu:134-135,227), fix_table = 1 uploads
// the whole table converted to float — what the kernel source evidently expects — so the kernel's transition arithmetic is pinned too.
extern "C" int ref_loglik_ex(const float* nets, int P, int depth, int n_step, int fix_table, float* out_a, int reps, float* kernel_ms) {
    float *gpu_a, *gpu_nets;
    size_t net_size = P * sizeof(float), nets_size = (size_t)P * 3 * sizeof(float);
    if (cudaMalloc(&gpu_a, net_size) != cudaSuccess || cudaMalloc(&gpu_nets, nets_size) != cudaSuccess) return -1;
    cudaMemcpy(gpu_nets, nets, nets_size, cudaMemcpyHostToDevice);
#if REF_VARIANT == 1 || REF_VARIANT == 2
    std::vector<int> tran_table = build_table(P, depth, n_step);
    const int tran_table_size = (int)tran_table.size();
    float* gpu_tran_table;
    cudaMalloc(&gpu_tran_table, tran_table_size * sizeof(float));
    cudaMemset(gpu_tran_table, 0, tran_table_size * sizeof(float));   // fresh cudaMalloc memory reads as zero in the reference runs
    if (fix_table) {
        std::vector<float> tf(tran_table.begin(), tran_table.end());
        cudaMemcpy(gpu_tran_table, tf.data(), tran_table_size * sizeof(float), cudaMemcpyHostToDevice);
    } else cudaMemcpy(gpu_tran_table, tran_table.data(), tran_table_size, cudaMemcpyHostToDevice);   // byte count as shipped
#endif
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blockSize = 256, gridSize = (P + blockSize - 1) / blockSize;
    float total = 0.f;
    for (int r = 0; r < reps; ++r) {
        cudaMemset(gpu_a, 0, net_size);
        cudaEventRecord(e0);
#if REF_VARIANT == 1
        log_likelihood_kernel<<<gridSize, blockSize>>>(g_x, g_y, gpu_a, gpu_nets, gpu_tran_table, P, (int)net_size, g_n, depth);
#elif REF_VARIANT == 2
        log_likelihood_kernel<<<gridSize, blockSize>>>(g_x, g_y, gpu_a, gpu_nets, gpu_tran_table, P, (int)net_size, g_n, depth, n_step);
#elif REF_VARIANT == 3
        (void)gridSize;
        log_likelihood_kernel<<<1, 1>>>(g_x, g_y, gpu_a, gpu_nets, g_n);                           // conv_mh.cu:144-147; P must be 2
#else
        log_likelihood_kernel<<<gridSize, blockSize>>>(g_x, g_y, gpu_a, gpu_nets, P, (int)net_size, g_n, depth);
#endif
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -2;
        float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
    }
    cudaMemcpy(out_a, gpu_a, net_size, cudaMemcpyDeviceToHost);
    if (kernel_ms) *kernel_ms = total / reps;
    cudaFree(gpu_a); cudaFree(gpu_nets);
#if REF_VARIANT == 1 || REF_VARIANT == 2
    cudaFree(gpu_tran_table);
#endif
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return cudaGetLastError() == cudaSuccess ? 0 : -3;
}

// the time-analysis shapes: binary tree of depth log2(P), table uploaded as shipped
extern "C" int ref_loglik(const float* nets, int P, float* out_a, int reps, float* kernel_ms) {
    return ref_loglik_ex(nets, P, (int)std::log2((double)P), 1, 0, out_a, reps, kernel_ms);
}
