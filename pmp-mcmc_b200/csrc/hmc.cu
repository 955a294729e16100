// hmc.cu — gradient (HMC) variants of the multi-proposal samplers.
//
// Reference: complex_nets/Cifar-10/cifar_SPhmc.py:77-137 (one leapfrog proposal, Metropolis test), cifar_MPhmc.py:77-140 (a leapfrog path of N
// nodes, weights relative to node 0), cifar_PMPhmc.py:76-171 and "Bayesian Network Training"/main.py:67-172 (binary prefetch tree whose edges
// are leapfrog steps; per-level Barker / Metropolis products over the tree).  What is on the device here:
//   * pmp_hmc_leapfrog_begin / _end — one leapfrog step along a tree edge on flat parameter vectors: momentum draw from the Philox stream
//     (or injected), half kick, drift, half kick, and the two kinetic energies |p|^2/2 the acceptance needs (cifar_PMPhmc.py:128-162);
//   * pmp_hmc_accept — the acceptance weights of all P nodes from (potential, kinetic) energies and the categorical draw, for the four rules.
// The potential's gradient is the caller's (autograd through an arbitrary torch network: LeNet+BatchNorm, torchbnn layers) and is passed as a
// device pointer — the same contract as PMP_TARGET_EXTERNAL for the likelihood callables (nets.py).  All vectors are device pointers.
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace pmp {

constexpr uint32_t STREAM_MOMENTUM = 4;
constexpr int HMC_THREADS = 256;
constexpr int HMC_MAX_P = 1024;             // nodes of an HMC tree / path (the reference runs 2 .. 32)

// p0 = scale * N(0,1) (or injected); kinetic energy of p0; p = p0 + sign * step * grad / 2; theta_child = theta_parent + sign * step * p.
// Elementwise and HBM-bound: 16-byte accesses, four elements per thread and trip (a scalar tail when dim is not a multiple of 4 or a pointer is not 16-byte aligned).
__device__ __forceinline__ void hmc_begin_elem(float th, float g, float p0, float ss, float& p, float& child, double& k) {
    k += (double)p0 * (double)p0;
    // the reference's float32 operation order: p += step * du_dx / 2 ; par += step * p   (cifar_PMPhmc.py:141-147, cifar_MPhmc.py:119-125)
    p = __fadd_rn(p0, __fdiv_rn(__fmul_rn(ss, g), 2.0f));
    child = __fadd_rn(th, __fmul_rn(ss, p));
}
__device__ __forceinline__ void hmc_block_sum(double k, double* dst) {
    __shared__ double red[HMC_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
    __syncthreads();
    if (threadIdx.x == 0) { double s = 0.0; for (int w = 0; w < HMC_THREADS / 32; ++w) s += red[w]; atomicAdd(dst, 0.5 * s); }
}
__global__ void hmc_begin_kernel(const float* __restrict__ theta_parent, const float* __restrict__ grad, float* __restrict__ theta_child, float* p_out,
                                 const float* p_init /* may alias p_out: the MP path carries one momentum buffer from node to node */, long long dim, float step, float sign,
                                 float p_scale, uint64_t seed, uint64_t iter, uint64_t idx0, double* __restrict__ ke, int vec) {
    double k = 0.0;
    const float ss = sign * step;
    const long long n4 = vec ? dim / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 th = __ldg(reinterpret_cast<const float4*>(theta_parent) + i), g = __ldg(reinterpret_cast<const float4*>(grad) + i);
        float4 p0;
        if (p_init) p0 = reinterpret_cast<const float4*>(p_init)[i];
        else if (((idx0 + 4ull * i) & 1ull) == 0) {       // four consecutive elements from an even index = the four words of TWO Philox blocks (stream_u64: block idx >> 1, word idx & 1)
            const uint64_t blk = (idx0 + 4ull * i) >> 1;
            uint32_t c0[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | ((uint32_t)STREAM_MOMENTUM << 24)};
            uint32_t c1[4] = {(uint32_t)(blk + 1), (uint32_t)((blk + 1) >> 32), c0[2], c0[3]};
            philox4x32_10(c0, (uint32_t)seed, (uint32_t)(seed >> 32));
            philox4x32_10(c1, (uint32_t)seed, (uint32_t)(seed >> 32));
            p0.x = __fmul_rn((float)det_norm_ppf(u64_to_open((uint64_t)c0[1] << 32 | c0[0])), p_scale);
            p0.y = __fmul_rn((float)det_norm_ppf(u64_to_open((uint64_t)c0[3] << 32 | c0[2])), p_scale);
            p0.z = __fmul_rn((float)det_norm_ppf(u64_to_open((uint64_t)c1[1] << 32 | c1[0])), p_scale);
            p0.w = __fmul_rn((float)det_norm_ppf(u64_to_open((uint64_t)c1[3] << 32 | c1[2])), p_scale);
        } else {
            p0.x = __fmul_rn((float)stream_normal(seed, iter, STREAM_MOMENTUM, idx0 + 4ull * i), p_scale);
            p0.y = __fmul_rn((float)stream_normal(seed, iter, STREAM_MOMENTUM, idx0 + 4ull * i + 1), p_scale);
            p0.z = __fmul_rn((float)stream_normal(seed, iter, STREAM_MOMENTUM, idx0 + 4ull * i + 2), p_scale);
            p0.w = __fmul_rn((float)stream_normal(seed, iter, STREAM_MOMENTUM, idx0 + 4ull * i + 3), p_scale);
        }
        float4 p, ch;
        hmc_begin_elem(th.x, g.x, p0.x, ss, p.x, ch.x, k); hmc_begin_elem(th.y, g.y, p0.y, ss, p.y, ch.y, k);
        hmc_begin_elem(th.z, g.z, p0.z, ss, p.z, ch.z, k); hmc_begin_elem(th.w, g.w, p0.w, ss, p.w, ch.w, k);
        reinterpret_cast<float4*>(p_out)[i] = p;
        reinterpret_cast<float4*>(theta_child)[i] = ch;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (long long)gridDim.x * blockDim.x) {
        const float p0 = p_init ? p_init[i] : __fmul_rn((float)stream_normal(seed, iter, STREAM_MOMENTUM, idx0 + (uint64_t)i), p_scale);
        float p, ch;
        hmc_begin_elem(theta_parent[i], grad[i], p0, ss, p, ch, k);
        p_out[i] = p; theta_child[i] = ch;
    }
    hmc_block_sum(k, ke);
}

// p += sign * step * grad / 2; kinetic energy of the final momentum
__global__ void hmc_end_kernel(float* __restrict__ p, const float* __restrict__ grad, long long dim, float step, float sign, double* __restrict__ ke, int vec) {
    double k = 0.0;
    const float ss = sign * step;
    const long long n4 = vec ? dim / 4 : 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 q = reinterpret_cast<float4*>(p)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(grad) + i);
        q.x = __fadd_rn(q.x, __fdiv_rn(__fmul_rn(ss, g.x), 2.0f)); q.y = __fadd_rn(q.y, __fdiv_rn(__fmul_rn(ss, g.y), 2.0f));
        q.z = __fadd_rn(q.z, __fdiv_rn(__fmul_rn(ss, g.z), 2.0f)); q.w = __fadd_rn(q.w, __fdiv_rn(__fmul_rn(ss, g.w), 2.0f));
        reinterpret_cast<float4*>(p)[i] = q;
        k += (double)q.x * (double)q.x + (double)q.y * (double)q.y + (double)q.z * (double)q.z + (double)q.w * (double)q.w;
    }
    for (long long i = 4 * n4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < dim; i += (long long)gridDim.x * blockDim.x) {
        const float q = __fadd_rn(p[i], __fdiv_rn(__fmul_rn(ss, grad[i]), 2.0f));
        p[i] = q;
        k += (double)q * (double)q;
    }
    hmc_block_sum(k, ke + 1);
}

struct HmcAcceptArgs {
    int rule, P, depth;
    const double* nl;        // [P]  nets_loss = -CrossEntropy (the negative potential)
    const double* ke_out;    // [P]  |p_s[parent(c)][c]|^2 / 2 indexed by the child c (tree rules); |p_s[j]|^2 / 2 (MP path)
    const double* ke_in;     // [P]  |p_s[c][parent(c)]|^2 / 2 indexed by the child c (tree rules)
    double u, temperature;
    float* weights;          // [P]  B as the reference hands it to torch.multinomial
    int* index;
};

// torch.min / torch.max (main.py) propagate NaN; the CIFAR script's Python min(tensor(1), x) / max(tensor(0), x) return the constant when x is NaN = fminf / fmaxf
__device__ __forceinline__ float nan_min(float a, float b) { return (isnan(a) || isnan(b)) ? NAN : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (isnan(a) || isnan(b)) ? NAN : fmaxf(a, b); }

// float32 like the reference's torch code (the tensors are float32 there); one CTA, P <= 1024
__global__ void hmc_accept_kernel(const HmcAcceptArgs a) {
    __shared__ float B[HMC_MAX_P];
    __shared__ double cdf[HMC_MAX_P];
    const int P = a.P;
    for (int node = threadIdx.x; node < P; node += blockDim.x) {
        float A;
        if (a.rule == PMP_HMC_RULE_MP) {
            // cifar_MPhmc.py:79-82: A_j = exp(min(0, (nl_j - K_j) - (nl_0 - K_0))), A_0 = N - sum_j A_j (filled below)
            A = node == 0 ? 0.f : expf(fminf(0.f, (float)a.nl[node] - (float)a.ke_out[node] - (float)a.nl[0] + (float)a.ke_out[0]));
        } else if (a.rule == PMP_HMC_RULE_SP) {
            // cifar_SPhmc.py:122-126: accept the proposal iff exp((-H_0 + H_1) * 1000) > rand, H = nl + K as the script signs them
            A = node == 0 ? 0.f : expf((float)(a.temperature) * (-((float)a.ke_out[0] + (float)a.nl[0]) + ((float)a.nl[1] + (float)a.ke_out[1])));
        } else {
            A = 1.f;
            for (int c = 0; c < a.depth; ++c) {
                const int half = 1 << c, m = node & (2 * half - 1);          // the script's `judg` loop reduces `all` modulo 2^(c+1) (cifar_PMPhmc.py:84-93)
                if (m < half) {
                    const int ch = m + half;                                 // edge (m -> ch): p_s[m][ch] is the momentum drawn at m, p_s[ch][m] the one that arrived at ch
                    const float w_new = expf((float)a.nl[m] - (float)a.ke_out[ch]);
                    const float w_old = expf((float)a.nl[ch] - (float)a.ke_in[ch]);
                    if (a.rule == PMP_HMC_RULE_TREE_CIFAR) A = A * fmaxf(0.f, 1.f - w_old / w_new);                      // cifar_PMPhmc.py:94-97
                    else { const float wo = nan_min(1.f, w_old / w_new), wn = nan_max(0.f, 1.f - wo / w_new); A = A * wn / (wn + wo); }   // main.py:84-88,95
                } else {
                    const int par = m - half;
                    const float w_new = expf((float)a.nl[m] - (float)a.ke_in[m]);     // p_s[m][m - half]: the momentum that arrived at m
                    const float w_old = expf((float)a.nl[par] - (float)a.ke_out[m]);  // p_s[m - half][m]: the momentum drawn at the parent
                    if (a.rule == PMP_HMC_RULE_TREE_CIFAR) A = A * fminf(1.f, w_new / w_old);                            // cifar_PMPhmc.py:99-102
                    else { const float wn = nan_min(1.f, w_new / w_old), wo = nan_max(0.f, 1.f - wn / w_old); A = A * wn / (wn + wo); }   // main.py:89-95
                }
            }
        }
        B[node] = A;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.rule == PMP_HMC_RULE_MP) { float s = 0.f; for (int j = 0; j < P; ++j) s += B[j]; B[0] = (float)(P - 1) - s; }
        if (a.rule == PMP_HMC_RULE_SP) {
            const int acc = B[1] > (float)a.u ? 1 : 0;
            a.weights[0] = B[0]; a.weights[1] = B[1];
            *a.index = acc;
            return;
        }
        // cifar_PMPhmc.py:104-107: NaN and Inf weights become 1; then one categorical draw (inverse CDF in place of torch.multinomial)
        double s = 0.0;
        for (int j = 0; j < P; ++j) { float b = B[j]; if (isnan(b) || isinf(b)) b = 1.f; B[j] = b; a.weights[j] = b; s += (double)b; cdf[j] = s; }
        int idx = P - 1;
        for (int j = 0; j < P; ++j) if (a.u < cdf[j] / s) { idx = j; break; }          // searchsorted(cdf / cdf[-1], u, 'right')
        *a.index = idx;
    }
}

}  // namespace pmp

using namespace pmp;

extern "C" {

static int hmc_scratch(pmp_ctx* c) {
    if (!c->d_hmc) {
        PMP_CUDA(cudaMalloc((void**)&c->d_hmc, 4096 + HMC_MAX_P * (3 * sizeof(double) + sizeof(float))));
    }
    return PMP_OK;
}

int pmp_hmc_leapfrog_begin(pmp_ctx* c, const float* theta_parent, const float* grad_parent, float* theta_child, float* p_child, const float* p_init, int64_t dim,
                           float step, float sign, float p_scale, uint64_t stream_index, double* ke_init) {
    PMP_REQUIRE(c && theta_parent && grad_parent && theta_child && p_child && dim > 0 && ke_init, "bad arguments");
    PMP_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = hmc_scratch(c))) return rc;
    double* ke = reinterpret_cast<double*>(c->d_hmc);
    PMP_CUDA(cudaMemsetAsync(ke, 0, 2 * sizeof(double), c->stream));
    long long blocks = (dim / 4 + HMC_THREADS - 1) / HMC_THREADS + 1;
    if (blocks > 16ll * c->sm_count) blocks = 16ll * c->sm_count;
    const int vec = (((uintptr_t)theta_parent | (uintptr_t)grad_parent | (uintptr_t)theta_child | (uintptr_t)p_child | (uintptr_t)p_init) & 15) == 0;
    hmc_begin_kernel<<<(unsigned)blocks, HMC_THREADS, 0, c->stream>>>(theta_parent, grad_parent, theta_child, p_child, p_init, dim, step, sign, p_scale,
                                                                       c->seed, c->host_iter, stream_index * (uint64_t)dim, ke, vec);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    PMP_CUDA(cudaMemcpyAsync(ke_init, ke, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_hmc_leapfrog_end(pmp_ctx* c, float* p_child, const float* grad_child, int64_t dim, float step, float sign, double* ke_final) {
    PMP_REQUIRE(c && p_child && grad_child && dim > 0 && ke_final, "bad arguments");
    PMP_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = hmc_scratch(c))) return rc;
    double* ke = reinterpret_cast<double*>(c->d_hmc);
    PMP_CUDA(cudaMemsetAsync(ke + 1, 0, sizeof(double), c->stream));
    long long blocks = (dim / 4 + HMC_THREADS - 1) / HMC_THREADS + 1;
    if (blocks > 16ll * c->sm_count) blocks = 16ll * c->sm_count;
    const int vec = (((uintptr_t)p_child | (uintptr_t)grad_child) & 15) == 0;
    hmc_end_kernel<<<(unsigned)blocks, HMC_THREADS, 0, c->stream>>>(p_child, grad_child, dim, step, sign, ke, vec);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    PMP_CUDA(cudaMemcpyAsync(ke_final, ke + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_hmc_accept(pmp_ctx* c, int rule, int P, const double* nets_loss, const double* ke_out, const double* ke_in, double u, double temperature,
                   float* weights_out, int32_t* index_out) {
    PMP_REQUIRE(c && nets_loss && ke_out && index_out, "bad arguments");
    PMP_REQUIRE(rule >= PMP_HMC_RULE_SP && rule <= PMP_HMC_RULE_TREE_BNN, "unknown HMC rule %d", rule);
    PMP_REQUIRE(P >= 2 && P <= HMC_MAX_P, "P=%d out of range [2, %d]", P, HMC_MAX_P);
    int depth = 0;
    if (rule == PMP_HMC_RULE_SP) PMP_REQUIRE(P == 2, "the single-proposal rule takes P == 2");
    if (rule >= PMP_HMC_RULE_TREE_CIFAR) {
        while ((1 << depth) < P) ++depth;
        PMP_REQUIRE((1 << depth) == P && ke_in, "the tree rules need P = 2^D nodes and both kinetic energies per edge");
    }
    PMP_REQUIRE(u >= 0.0 && u < 1.0, "uniform out of [0,1)");
    PMP_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = hmc_scratch(c))) return rc;
    uint8_t* base = reinterpret_cast<uint8_t*>(c->d_hmc) + 4096;
    double* d_nl = reinterpret_cast<double*>(base);
    double* d_ko = d_nl + HMC_MAX_P;
    double* d_ki = d_ko + HMC_MAX_P;
    float* d_w = reinterpret_cast<float*>(d_ki + HMC_MAX_P);
    int* d_idx = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(c->d_hmc) + 64);
    PMP_CUDA(cudaMemcpyAsync(d_nl, nets_loss, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaMemcpyAsync(d_ko, ke_out, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (ke_in) PMP_CUDA(cudaMemcpyAsync(d_ki, ke_in, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    HmcAcceptArgs a{rule, P, depth, d_nl, d_ko, d_ki, u, temperature, d_w, d_idx};
    hmc_accept_kernel<<<1, 256, 0, c->stream>>>(a);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    if (weights_out) PMP_CUDA(cudaMemcpyAsync(weights_out, d_w, P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaMemcpyAsync(index_out, d_idx, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

}  // extern "C"
