// common.cuh — context layout and small helpers shared by the translation units of libpmp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/pmp_b200.h"

namespace pmp {

void set_error(const char* fmt, ...);

#define PMP_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            pmp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PMP_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define PMP_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            pmp::set_error(__VA_ARGS__);  \
            return PMP_ERR_ARG;           \
        }                                 \
    } while (0)

// Fixed-point format of the per-node sums of squared standardised residuals: the sweep converts the float32
// partial of every CHUNK consecutive data points to an integer multiple of 2^-FX_SHIFT and from there on
// everything is integer addition — exact and associative, so the per-node sum is bit-identical for any
// block order, any grid size and any number of GPUs (shards aligned to CHUNK).
constexpr int CHUNK = 64;
constexpr int FX_SHIFT = 20;
constexpr int MAX_NODES = 8192;      // accept kernel keeps two double arrays of P entries in shared memory
constexpr int KDIM_MAX = 64;         // in-kernel closed-form MP kernel term supports dim <= KDIM_MAX

struct DeviceCounters {
    unsigned long long iteration;   // Philox iteration counter, advanced by the accept kernel
    long long trace_rows;           // rows recorded in the trace ring
    int flags;                      // bit0: fixed-point saturation seen
    int last_next;                  // accepted node of the last accept
};

struct TraceBuffers {
    long long capacity = 0;
    uint32_t what = 0;
    float* state = nullptr;
    int32_t* next = nullptr;
    int32_t* draws = nullptr;
    float* samples = nullptr;
    double* logw = nullptr;
};

}  // namespace pmp

struct pmp_ctx {
    int device = 0, world = 1, rank = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    void* nccl_comm = nullptr;         // created lazily (first all-reduce): chains that only ever use the in-kernel peer exchange never build a communicator
    unsigned char nccl_id[128] = {0};  // ncclUniqueId given to pmp_create

    bool configured = false;
    pmp_config cfg{};
    int P = 0;

    // linear-Gaussian data shard
    float* d_x = nullptr; float* d_y = nullptr;
    uint8_t* d_bimg = nullptr;         // [nchunks][2048] bf16 data operand image of the tensor-core sweep (sweep_linear_tc.cuh)
    long long n_local = 0, n_offset = 0, n_global = 0;
    size_t data_capacity = 0;          // padded float count d_x / d_y were allocated for
    bool data_borrowed = false;        // d_x / d_y / d_bimg alias another ctx's buffers (pmp_share_data): never freed here

    // chain
    uint64_t seed = 0;
    float* d_state = nullptr;          // [dim]
    float* d_props = nullptr;          // [P, dim]
    unsigned long long* d_acc = nullptr;  // [P] fixed-point partial sums (linear-Gaussian)
    double* d_lt = nullptr;            // [P] log-targets
    double* d_logw = nullptr;          // [P] log-weights of the last accept
    int32_t* d_draws = nullptr;        // [P]
    double* d_uniforms = nullptr;      // [P+1] injected uniforms
    pmp::DeviceCounters* d_cnt = nullptr;
    float* d_z = nullptr;              // [2, P*dim] prefetched standard normals (current / next iteration)
    unsigned long long* d_dbg = nullptr; // optional phase stamps
    void* d_psync = nullptr;           // PersistSync of the cooperative chain kernel
    unsigned long long* d_hs = nullptr; // flag-in-data hand-off buffers of the persistent chain kernels: sums | nodes | normals (accept_lean.cuh)
    size_t hs_words = 0; int hs_grid = 0, hs_P = 0;
    unsigned int hs_epoch = 0;         // iterations of all earlier hand-off launches on this ctx (tags are never reused)
    unsigned int* d_done = nullptr;    // CTA completion counter of the fused sweep+accept kernel
    unsigned long long host_iter = 0;  // host mirror of d_cnt->iteration
    long long z_valid_iter = -1;       // iteration whose normals are in d_z, -1: none
    bool lt_valid = false;             // d_lt holds the log-targets of the current proposals
    bool props_external = false;       // d_props were written by the caller (pmp_write_proposals), not generated about d_state by pmp_propose
    bool acc_pending = false;          // d_acc holds an un-finalised sweep

    pmp::TraceBuffers trace;

    // graph cache for pmp_run
    cudaGraphExec_t graph_exec = nullptr;
    int graph_iters = 0;
    long long graph_launches_total = 0;

    // misc
    void* d_flush = nullptr; size_t flush_bytes = 0;
    long long launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::vector<cudaEvent_t> ev_pool;

    // batched analytic chains
    long long n_chains = 0;
    float* d_chain_states = nullptr;   // [dim, n_chains] (chain fastest)
    float* d_chain_samples = nullptr;  // [iters, P, dim, n_chains]
    long long chain_samples_cap = 0;   // in floats
    long long chain_iters_recorded = 0;
    unsigned long long chain_iteration = 0;
    void* chain_scratch = nullptr;

    // FC model, GLM heads
    void* fc = nullptr;
    void* glm = nullptr;
    void* cnn = nullptr;
    void* d_hmc = nullptr;             // HMC variants: kinetic energies, acceptance inputs / outputs (hmc.cu)

    // peer-memory exchange of the per-node sums (world > 1, pmp_peer_exchange_*): own buffer + the peers' buffers mapped with CUDA IPC
    unsigned long long* d_xchg = nullptr;
    unsigned long long* peer_xchg[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool peers_attached = false;
    unsigned long long xchg_count = 0;   // exchanges done so far (identical on every rank: SPMD call sequence)

    // MP kernel term for long parameter vectors (pmp_large_dim_kernel_term)
    float* d_kt_s1 = nullptr; double* d_kt_dj2 = nullptr; double* d_kt_dot = nullptr; long long kt_dim_cap = 0;
};
