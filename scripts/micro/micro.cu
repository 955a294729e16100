// micro.cu — latency/throughput probes used to design the acceptance kernel (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)
__global__ void k_empty() {}
__global__ void k_dfma_chain(double* out, int n, double a, double b) { double x = threadIdx.x; for (int i = 0; i < n; ++i) x = fma(x, a, b); out[blockIdx.x * blockDim.x + threadIdx.x] = x; }
__global__ void k_ffma_chain(float* out, int n, float a, float b) { float x = threadIdx.x; for (int i = 0; i < n; ++i) x = fmaf(x, a, b); out[blockIdx.x * blockDim.x + threadIdx.x] = x; }
__global__ void k_dfma_ilp(double* out, int n, double a, double b) { double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3; for (int i = 0; i < n; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);} out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3; }
__global__ void k_explog(double* out, int n, double a) { double x = 1.0 + threadIdx.x * 1e-3; for (int i = 0; i < n; ++i) x = log(exp(x * a) + 1.0); out[blockIdx.x * blockDim.x + threadIdx.x] = x; }
__global__ void k_sync(int* out, int n) { int x = 0; for (int i = 0; i < n; ++i) { __syncthreads(); x += i; } out[threadIdx.x] = x; }
__global__ void k_atomic(unsigned long long* p, int n) { for (int i = 0; i < n; ++i) atomicAdd(p + threadIdx.x, 1ull); }
__global__ void k_fence(unsigned* p, int n) { for (int i = 0; i < n; ++i) { __threadfence(); } if (n < 0) p[0] = 1; }
__global__ void k_ldcg_chain(const int* p, int* out, int n) { int idx = 0; for (int i = 0; i < n; ++i) idx = __ldcg(p + idx); out[0] = idx; }
template <class F> float timeit(F f, int reps = 20) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaDeviceSynchronize(); cudaEventRecord(a); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps * 1e3f; }
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    double* d; CK(cudaMalloc(&d, 1 << 24)); float* f = (float*)d; int* ip = (int*)d; CK(cudaMemset(d, 0, 1 << 24));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clockRate attr %d kHz\n", clk);
    printf("empty kernel, stream launches: %.2f us each\n", timeit([&] { k_empty<<<1, 32>>>(); }, 200));
    { cudaStream_t s; cudaStreamCreate(&s); cudaGraph_t g; cudaGraphExec_t ge; cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal); for (int i = 0; i < 100; ++i) k_empty<<<1, 32, 0, s>>>(); cudaStreamEndCapture(s, &g); cudaGraphInstantiate(&ge, g, 0);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaGraphLaunch(ge, s); cudaStreamSynchronize(s); cudaEventRecord(a, s); for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s); cudaEventRecord(b, s); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); printf("empty kernel in graph (100/graph): %.2f us each\n", ms / 1000 * 1e3f);
      cudaGraph_t g2; cudaGraphExec_t ge2; cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal); for (int i = 0; i < 100; ++i) k_empty<<<444, 256, 40000, s>>>(); cudaStreamEndCapture(s, &g2); cudaGraphInstantiate(&ge2, g2, 0);
      cudaGraphLaunch(ge2, s); cudaStreamSynchronize(s); cudaEventRecord(a, s); for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge2, s); cudaEventRecord(b, s); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); printf("empty 444x256 50KB-smem kernel in graph: %.2f us each\n", ms / 1000 * 1e3f); }
    int n = 4096;
    float t = timeit([&] { k_dfma_chain<<<1, 32>>>(d, n, 1.0000001, 1e-9); }); printf("DFMA dependent chain, 1 warp: %.2f us / %d = %.1f ns per op\n", t, n, t * 1e3 / n);
    t = timeit([&] { k_ffma_chain<<<1, 32>>>(f, n, 1.0000001f, 1e-9f); }); printf("FFMA dependent chain, 1 warp: %.2f us / %d = %.1f ns per op\n", t, n, t * 1e3 / n);
    t = timeit([&] { k_dfma_chain<<<1, 256>>>(d, n, 1.0000001, 1e-9); }); printf("DFMA chain, 8 warps 1 SM: %.2f us -> %.2f ns per warp-op\n", t, t * 1e3 / n / 8);
    t = timeit([&] { k_dfma_ilp<<<148, 1024>>>(d, n, 1.0000001, 1e-9); }); printf("DFMA throughput full chip: %.2f us -> %.2f TFLOP/s\n", t, 148.0 * 1024 * n * 4 * 2 / (t * 1e-6) / 1e12);
    t = timeit([&] { k_explog<<<1, 32>>>(d, 256, 0.5); }); printf("exp+log (fp64) dependent, 1 warp: %.1f ns per pair\n", t * 1e3 / 256);
    t = timeit([&] { k_explog<<<1, 256>>>(d, 256, 0.5); }); printf("exp+log (fp64), 8 warps: %.1f ns per pair per warp-batch\n", t * 1e3 / 256);
    t = timeit([&] { k_sync<<<1, 256>>>(ip, 1000); }); printf("__syncthreads 256 thr: %.1f ns each\n", t * 1e3 / 1000);
    t = timeit([&] { k_sync<<<1, 1024>>>(ip, 1000); }); printf("__syncthreads 1024 thr: %.1f ns each\n", t * 1e3 / 1000);
    t = timeit([&] { k_atomic<<<1, 32>>>((unsigned long long*)d, 1000); }); printf("atomicAdd u64 (32 distinct addrs, 1 warp, back to back): %.1f ns each\n", t * 1e3 / 1000);
    t = timeit([&] { k_fence<<<1, 32>>>((unsigned*)d, 1000); }); printf("__threadfence idle: %.1f ns each\n", t * 1e3 / 1000);
    CK(cudaMemset(d, 0, 1 << 24));
    t = timeit([&] { k_ldcg_chain<<<1, 1>>>(ip, ip + 1024, 1000); }); printf("dependent L2 load (ldcg): %.1f ns each\n", t * 1e3 / 1000);
    CK(cudaDeviceSynchronize());
    return 0;
}
