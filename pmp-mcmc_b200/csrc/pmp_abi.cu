// pmp_abi.cu — the C-ABI of include/pmp_b200.h: context, buffers, kernel launches, device-resident chain loop.
// No torch types, no CPU fallback: every compute entry point launches sm_100a kernels on the ctx stream.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdlib.h>

#include <memory>

#include "accept.cuh"
#include "accept_fast.cuh"
#include "accept_lean.cuh"
#include "chain_persistent.cuh"
#include "chain_persistent_tc.cuh"
#include "chain_persistent_multi.cuh"

extern "C" int pmp_fc_loglik(pmp_ctx* c);   // fc_sweep.cu
extern "C" int pmp_cnn_loglik(pmp_ctx* c);  // cnn_sweep.cuh (compiled with fc_sweep.cu)
extern "C" int pmp_glm_loglik(pmp_ctx* c);  // fc_sweep.cu
extern "C" int pmp_glm_destroy(pmp_ctx* c);
#include "common.cuh"
#include "sweep_linear.cuh"
#include "sweep_linear_tc.cuh"

namespace pmp {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

// ---- NCCL, resolved at run time so that libpmp_b200.so has no link-time dependency (and shares torch's libnccl
// when torch is already loaded in the process).
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.handle) return PMP_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error("NCCL not found: %s", dlerror()); return PMP_ERR_NCCL; }
    g_nccl.GetUniqueId = (decltype(&ncclGetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(&ncclCommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(&ncclCommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(&ncclAllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.GetErrorString = (decltype(&ncclGetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
        set_error("NCCL symbols missing"); return PMP_ERR_NCCL;
    }
    g_nccl.handle = h;
    return PMP_OK;
}
#define PMP_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) {                                                                         \
            pmp::set_error("%s failed: %s", #expr, pmp::g_nccl.GetErrorString ? pmp::g_nccl.GetErrorString(_r) : "?"); \
            return PMP_ERR_NCCL;                                                                         \
        }                                                                                                \
    } while (0)

static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

// The persistent sweep CTAs keep at most PERSIST_MAX_SEGS (node tile, chunk range) segments; a contiguous range of `per_cta` units
// touches at most this many tiles.  The host refuses the persistent kernels when it could exceed the cap (the kernels trap if it ever does).
static bool persist_segments_fit(long long units, int n_sweep, long long nchunks) {
    if (units == 0 || nchunks == 0) return true;
    if ((long long)n_sweep * nchunks >= units) return true;      // at least as many sweep CTAs as node tiles: one segment per CTA (sweep_partition)
    const long long per_cta = (units + n_sweep - 1) / n_sweep;
    return (per_cta + nchunks - 2) / nchunks + 1 <= PERSIST_MAX_SEGS;
}

static long long total_chunks_global(const pmp_ctx* c) { return (c->n_global + CHUNK - 1) / CHUNK + c->world; }
static double sat_limit(const pmp_ctx* c) { return 4611686018427387904.0 / (double)(total_chunks_global(c) > 0 ? total_chunks_global(c) : 1); }

template <typename T> static int dev_alloc(T** p, size_t count) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (count == 0) return PMP_OK;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return PMP_ERR_ALLOC; }
    return PMP_OK;
}

static void drop_graph(pmp_ctx* c) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; c->graph_iters = 0; }
}

// ---- launches ------------------------------------------------------------------------------------------------
static int launch_propose(pmp_ctx* c) {
    c->props_external = false;
    ProposeArgs a{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, (c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) ? 1 : 0};
    long long total = (long long)c->P * c->cfg.dim;
    if (c->cfg.tree != PMP_TREE_FLAT && total >= (1ll << 20) && c->cfg.dim >= 1024) {
        // long parameter vectors on a tree: one launch per level, one quantile per element (propose_level_kernel)
        const int b = c->cfg.tree == PMP_TREE_BINARY ? 2 : c->cfg.b;
        unsigned gx = (unsigned)((c->cfg.dim + 1023) / 1024);
        propose_level_kernel<<<dim3(gx, 1), 256, 0, c->stream>>>(a, 0);
        c->launches++;
        long long s = 1;
        for (int l = 0; l < c->cfg.depth; ++l, s *= b) {
            propose_level_kernel<<<dim3(gx, (unsigned)(s * (b - 1))), 256, 0, c->stream>>>(a, (int)s);
            c->launches++;
        }
        PMP_CUDA(cudaGetLastError());
        return PMP_OK;
    }
    long long blocks = (total + 255) / 256;
    long long cap = (long long)c->sm_count * 8;
    if (blocks > cap) blocks = cap;
    propose_kernel<<<(unsigned)blocks, 256, 0, c->stream>>>(a);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

static size_t sweep_smem_bytes() { return (size_t)2 * TILE_CHUNKS * CHUNK_STRIDE * sizeof(float); }

static AcceptArgs make_accept_args(pmp_ctx* c, int from_acc, int only_finalize, int advance, const double* d_uniforms) {
    AcceptArgs a{};
    a.cfg = c->cfg; a.P = c->P; a.n_global = c->n_global; a.state = c->d_state; a.props = c->d_props; a.acc = c->d_acc;
    a.lt = c->d_lt; a.logw = c->d_logw; a.draws = c->d_draws; a.uniforms = d_uniforms; a.cnt = c->d_cnt; a.seed = c->seed;
    a.sat_limit = sat_limit(c); a.inv_scale = 1.0 / (double)c->cfg.scale; a.dbg = c->d_dbg; a.from_acc = from_acc; a.only_finalize = only_finalize; a.advance = advance; a.trace = c->trace;
    return a;
}

// Tensor-core sweep (sweep_linear_tc.cuh) when its shared-memory plan fits: node operand + integer scratch for all of
// P, and this CTA's chunks resident.  Returns the grid size, 0 when the FMA sweep must be used.
static int tc_sweep_plan(const pmp_ctx* c, int ctas, int* max_chunks, int* max_units, size_t* smem) {
    if (!env_int("PMP_SWEEP_TC", 0) || !c->d_bimg) return 0;     // opt-in until it beats the FMA sweep
    const int ntiles = (c->P + tc::TILE_NODES - 1) / tc::TILE_NODES;
    const long long nchunks = (c->n_local + CHUNK - 1) / CHUNK;
    if (ntiles > tc::MAX_TILES || nchunks == 0) return 0;
    const long long units = nchunks * ntiles;
    long long g = ctas < units ? ctas : units;
    const long long per = (units + g - 1) / g;
    const long long mc = (per + ntiles - 1) / ntiles + 1;
    if (per > tc::MAX_UNITS) return 0;
    const size_t bytes = tc::smem_bytes(ntiles, (int)mc, (int)per);
    if (bytes > 200 * 1024) return 0;
    *max_chunks = (int)mc; *max_units = (int)per; *smem = bytes;
    return (int)g;
}

// generate: also fill the next iteration's half of the normals table as a side job.
static int launch_sweep_linear(pmp_ctx* c, int generate) {
    PMP_REQUIRE(c->d_x && c->n_local >= 0, "linear-Gaussian data not set (pmp_set_data_linear)");
    {
        int mc = 0, mu = 0; size_t smem = 0;
        const int g = tc_sweep_plan(c, c->sm_count, &mc, &mu, &smem);
        if (g > 0) {
            tc::Args ta{c->d_bimg, c->d_props, c->d_acc, c->d_cnt, (c->n_local + CHUNK - 1) / CHUNK, c->P, mc, mu, sat_limit(c), generate, c->d_z,
                        ProposeArgs{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, (c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) ? 1 : 0},
                        c->d_dbg};
            tc::sweep_linear_tc_kernel<<<(unsigned)g, tc::STANDALONE_THREADS, smem, c->stream>>>(ta);
            c->launches++;
            PMP_CUDA(cudaGetLastError());
            return PMP_OK;
        }
    }
    constexpr int R = 4;
    int need = (c->P + R - 1) / R;
    int tp_cap = env_int("PMP_SWEEP_TP", 32);
    if (tp_cap > MAX_TP) tp_cap = MAX_TP;
    int TP = 1; while (TP < need && TP < tp_cap) TP <<= 1;
    int TD = SWEEP_THREADS / TP;
    long long nchunks = (c->n_local + CHUNK - 1) / CHUNK;
    if (nchunks == 0) return PMP_OK;
    int ntiles = (c->P + TP * R - 1) / (TP * R);
    long long units = (long long)ntiles * nchunks;
    int per_sm = env_int("PMP_SWEEP_BLOCKS_PER_SM", 2);
    long long gx = (long long)c->sm_count * per_sm;
    int rounds = env_int("PMP_SWEEP_ROUNDS", 0);      // > 0: fixed chunks per data-thread per CTA (many small CTAs)
    if (rounds > 0) gx = (units + (long long)rounds * TD - 1) / ((long long)rounds * TD);
    long long want = (units + TD - 1) / TD;          // at least TD chunks per CTA so every data-thread has work
    if (gx > want) gx = want;
    if (gx < 1) gx = 1;
    SweepArgs a{c->d_x, c->d_y, c->d_props, c->d_acc, c->d_cnt, c->n_local, nchunks, c->P, TP, TD, sat_limit(c), generate, c->d_z,
                ProposeArgs{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, (c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) ? 1 : 0},
                c->d_dbg};
    if (env_int("PMP_SWEEP_SCALAR", 0)) sweep_linear_kernel<R, false><<<(unsigned)gx, SWEEP_THREADS, sweep_smem_bytes(), c->stream>>>(a);
    else sweep_linear_kernel<R, true><<<(unsigned)gx, SWEEP_THREADS, sweep_smem_bytes(), c->stream>>>(a);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// Collective: every rank of the ctx's world must reach it at the same point of its call sequence (they do: identical host code paths).
// Never called while the stream is being captured — run_impl builds the communicator before it starts a capture.
static int ensure_comm(pmp_ctx* c) {
    if (c->world <= 1 || c->nccl_comm) return PMP_OK;
    ncclUniqueId id; memcpy(&id, c->nccl_id, sizeof(id));
    ncclComm_t comm;
    PMP_NCCL(g_nccl.CommInitRank(&comm, c->world, id, c->rank));
    c->nccl_comm = comm;
    return PMP_OK;
}

static int allreduce_acc(pmp_ctx* c) {
    if (c->world <= 1) return PMP_OK;
    { int rc = ensure_comm(c); if (rc) return rc; }
    PMP_NCCL(g_nccl.AllReduce(c->d_acc, c->d_acc, (size_t)c->P, ncclUint64, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return PMP_OK;
}

static int launch_accept(pmp_ctx* c, int from_acc, int only_finalize, int advance, const double* d_uniforms) {
    AcceptArgs a = make_accept_args(c, from_acc, only_finalize, advance, d_uniforms);
    size_t smem = (size_t)c->P * 2 * sizeof(double);
    switch (c->cfg.algo) {
        case PMP_ALGO_MH: case PMP_ALGO_BARKER: accept_kernel<PMP_ALGO_MH><<<1, ACCEPT_THREADS, smem, c->stream>>>(a); break;
        case PMP_ALGO_MP: accept_kernel<PMP_ALGO_MP><<<1, ACCEPT_THREADS, smem, c->stream>>>(a); break;
        case PMP_ALGO_PSP: accept_kernel<PMP_ALGO_PSP><<<1, ACCEPT_THREADS, smem, c->stream>>>(a); break;
        case PMP_ALGO_PMP: accept_kernel<PMP_ALGO_PMP><<<1, ACCEPT_THREADS, smem, c->stream>>>(a); break;
        default: accept_kernel<PMP_ALGO_TABLE><<<1, ACCEPT_THREADS, smem, c->stream>>>(a); break;
    }
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

static bool fast_accept_ok(const pmp_ctx* c) {
    if (c->cfg.target != PMP_TARGET_LINEAR_GAUSS || env_int("PMP_ACCEPT_GENERIC", 0)) return false;
    if ((size_t)c->P * (c->cfg.algo == PMP_ALGO_PSP ? 32 : 16) > 220 * 1024) return false;
    if (c->cfg.algo == PMP_ALGO_MP || c->cfg.algo == PMP_ALGO_PSP) return true;
    return c->cfg.algo == PMP_ALGO_TABLE && ((c->cfg.flags & (PMP_FLAG_QUIRK_TABLE_CONST | PMP_FLAG_NO_KERNEL_TERM)) != 0) && !(c->cfg.flags & PMP_FLAG_STANDARDIZE);
}

static bool lean_accept_ok(const pmp_ctx* c) { return fast_accept_ok(c) && c->P <= LEAN_MAX_P && env_int("PMP_ACCEPT_LEAN", 1); }

static int launch_accept_fast(pmp_ctx* c, int make_next) {
    AcceptFastArgs fa{make_accept_args(c, 1, 0, 1, nullptr), c->d_z, make_next,
                      ProposeArgs{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, (c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) ? 1 : 0}};
    if (lean_accept_ok(c)) {     // three-phase acceptance (accept_lean.cuh); same rules, shorter critical path
        const size_t lsm = lean_smem_bytes(c->P, c->cfg.algo);
        switch (c->cfg.algo) {
            case PMP_ALGO_MP: accept_lean_kernel<PMP_ALGO_MP><<<1, ACCEPT_THREADS, lsm, c->stream>>>(fa); break;
            case PMP_ALGO_PSP: accept_lean_kernel<PMP_ALGO_PSP><<<1, ACCEPT_THREADS, lsm, c->stream>>>(fa); break;
            default: accept_lean_kernel<PMP_ALGO_TABLE><<<1, ACCEPT_THREADS, lsm, c->stream>>>(fa); break;
        }
        c->launches++;
        PMP_CUDA(cudaGetLastError());
        return PMP_OK;
    }
    size_t smem = (size_t)c->P * (c->cfg.algo == PMP_ALGO_PSP ? 4 : 2) * sizeof(double);
    switch (c->cfg.algo) {
        case PMP_ALGO_MP: accept_fast_kernel<PMP_ALGO_MP><<<1, ACCEPT_THREADS, smem, c->stream>>>(fa); break;
        case PMP_ALGO_PSP: accept_fast_kernel<PMP_ALGO_PSP><<<1, ACCEPT_THREADS, smem, c->stream>>>(fa); break;
        default: accept_fast_kernel<PMP_ALGO_TABLE><<<1, ACCEPT_THREADS, smem, c->stream>>>(fa); break;
    }
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

// one full iteration on the stream: propose → sweep → [all-reduce] → accept
static int enqueue_iteration(pmp_ctx* c, cudaEvent_t sweep_begin, cudaEvent_t sweep_end, bool first_of_graph) {
    int rc;
    if (c->cfg.target == PMP_TARGET_LINEAR_GAUSS) {
        // fused chain loop: the acceptance kernel of iteration i publishes the nodes of iteration i+1 from the normals
        // table that the sweep of iteration i filled as a side job; only the first iteration of a run (or of a captured
        // graph, which cannot know what preceded it) builds its table and nodes with the stand-alone kernels.
        // (an empty shard launches no sweep CTAs, hence nobody to fill the normals table: that rank builds its nodes with propose_kernel — same bits)
        const int fused = env_int("PMP_FUSE_PROPOSE", 1) && fast_accept_ok(c) && c->n_local > 0;
        const bool chained = fused && !first_of_graph && c->z_valid_iter == (long long)c->host_iter;
        if (!chained && (rc = launch_propose(c))) return rc;
        if (sweep_begin) PMP_CUDA(cudaEventRecord(sweep_begin, c->stream));
        if ((rc = launch_sweep_linear(c, fused))) return rc;
        if (sweep_end) PMP_CUDA(cudaEventRecord(sweep_end, c->stream));
        if ((rc = allreduce_acc(c))) return rc;
        if (fast_accept_ok(c)) { if ((rc = launch_accept_fast(c, fused))) return rc; }
        else if ((rc = launch_accept(c, 1, 0, 1, nullptr))) return rc;
        c->host_iter++;
        c->z_valid_iter = fused ? (long long)c->host_iter : -1;
        return PMP_OK;
    }
    if ((rc = launch_propose(c))) return rc;
    if (c->cfg.target == PMP_TARGET_EXTERNAL) {
        set_error("pmp_run: target %d needs host-driven log-targets (use pmp_propose / pmp_write_logtarget / pmp_accept)", c->cfg.target);
        return PMP_ERR_UNSUPPORTED;
    }
    if (c->cfg.target == PMP_TARGET_CNN && (rc = pmp_cnn_loglik(c))) return rc;   // direct convolutions + GEMM with the fused head, same contract as the FC target
    if (c->cfg.target == PMP_TARGET_FC && (rc = pmp_fc_loglik(c))) return rc;     // GEMM chain + all-reduce of the integer loss sums, all on the ctx stream
    if ((c->cfg.target == PMP_TARGET_GLM_LOGISTIC || c->cfg.target == PMP_TARGET_GLM_GAUSS) && (rc = pmp_glm_loglik(c))) return rc;
    if ((rc = launch_accept(c, 0, 0, 1, nullptr))) return rc;
    c->host_iter++;
    return PMP_OK;
}

// ---- chain diagnostics over the STATE trace (ESS/s and MSJD/s are the reference's headline comparison, README:56,
// computed offline there from the dumped samples) -------------------------------------------------------------------------------
__global__ void diag_moments_kernel(const float* __restrict__ st, long long n, int dim, double* __restrict__ mean, double* __restrict__ var) {
    __shared__ double red[32];
    const int j = blockIdx.x;
    double s = 0.0;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) s += (double)st[t * dim + j];
    s = block_sum(s, red);
    const double m = s / (double)n;
    double v = 0.0;
    for (long long t = threadIdx.x; t < n; t += blockDim.x) { const double d = (double)st[t * dim + j] - m; v = fma(d, d, v); }
    v = block_sum(v, red);
    if (threadIdx.x == 0) { mean[j] = m; var[j] = v / (double)n; }
}
// acov[k, j] = (1/n) sum_{t < n-k} (x_t - m_j)(x_{t+k} - m_j)
__global__ void diag_acov_kernel(const float* __restrict__ st, long long n, int dim, const double* __restrict__ mean, double* __restrict__ acov) {
    __shared__ double red[32];
    const int k = blockIdx.x, j = blockIdx.y;
    const double m = mean[j];
    double s = 0.0;
    for (long long t = threadIdx.x; t + k < n; t += blockDim.x) s = fma((double)st[t * dim + j] - m, (double)st[(t + k) * dim + j] - m, s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) acov[(long long)k * dim + j] = s / (double)n;
}
// out[0] = mean squared jump distance, out[1] = fraction of iterations whose accepted node is not node 0 (-1 without a NEXT trace)
__global__ void diag_jumps_kernel(const float* __restrict__ st, const int32_t* __restrict__ next, long long n, int dim, double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0, mv = 0.0;
    for (long long t = threadIdx.x; t + 1 < n; t += blockDim.x) {
        double d2 = 0.0;
        for (int j = 0; j < dim; ++j) { const double d = (double)st[(t + 1) * dim + j] - (double)st[t * dim + j]; d2 = fma(d, d, d2); }
        s += d2;
    }
    if (next) for (long long t = threadIdx.x; t < n; t += blockDim.x) mv += next[t] != 0 ? 1.0 : 0.0;
    s = block_sum(s, red);
    mv = block_sum(mv, red);
    if (threadIdx.x == 0) { out[0] = n > 1 ? s / (double)(n - 1) : 0.0; out[1] = next ? mv / (double)n : -1.0; }
}

}  // namespace pmp

using namespace pmp;

extern "C" {

const char* pmp_last_error(void) { return g_err; }
int pmp_abi_version(void) { return PMP_B200_ABI_VERSION; }

int pmp_nccl_unique_id(void* out128) {
    PMP_REQUIRE(out128, "out128 is NULL");
    int rc = load_nccl(); if (rc) return rc;
    ncclUniqueId id;
    PMP_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return PMP_OK;
}

int pmp_create(pmp_ctx** out, int device, int world_size, int rank, const void* nccl_unique_id) {
    PMP_REQUIRE(out, "out is NULL");
    PMP_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad world_size/rank %d/%d", world_size, rank);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no usable CUDA device (%s); this library has no CPU path", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return PMP_ERR_CUDA;
    }
    PMP_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
    PMP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PMP_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libpmp_b200 is built for sm_100a only", device, prop.major, prop.minor);
        return PMP_ERR_UNSUPPORTED;
    }
    pmp_ctx* c = new pmp_ctx();
    c->device = device; c->world = world_size; c->rank = rank; c->sm_count = prop.multiProcessorCount;
    PMP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PMP_CUDA(cudaEventCreate(&c->ev0)); PMP_CUDA(cudaEventCreate(&c->ev1));
    PMP_CUDA(cudaMalloc((void**)&c->d_cnt, sizeof(DeviceCounters)));
    PMP_CUDA(cudaMemset(c->d_cnt, 0, sizeof(DeviceCounters)));
    if (env_int("PMP_DEBUG_STAMPS", 0)) { PMP_CUDA(cudaMalloc((void**)&c->d_dbg, (64 + 3 * 1024) * 8)); PMP_CUDA(cudaMemset(c->d_dbg, 0, (64 + 3 * 1024) * 8)); }
    PMP_CUDA(cudaMalloc((void**)&c->d_done, sizeof(unsigned int)));
    PMP_CUDA(cudaMemset(c->d_done, 0, sizeof(unsigned int)));
    const int accept_smem = MAX_NODES * 2 * (int)sizeof(double);
    PMP_CUDA(cudaFuncSetAttribute(accept_kernel<PMP_ALGO_MH>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_kernel<PMP_ALGO_MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_kernel<PMP_ALGO_PSP>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_kernel<PMP_ALGO_PMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_kernel<PMP_ALGO_TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(tc::sweep_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PMP_CUDA(cudaFuncSetAttribute(sweep_linear_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes()));
    PMP_CUDA(cudaFuncSetAttribute(sweep_linear_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_smem_bytes()));
    // without this the driver may pick a carveout that fits a single CTA per SM
    PMP_CUDA(cudaFuncSetAttribute(sweep_linear_kernel<4, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    PMP_CUDA(cudaFuncSetAttribute(sweep_linear_kernel<4, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    PMP_CUDA(cudaFuncSetAttribute(accept_fast_kernel<PMP_ALGO_MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_fast_kernel<PMP_ALGO_PSP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * accept_smem > 220 * 1024 ? 220 * 1024 : 2 * accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_fast_kernel<PMP_ALGO_TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, accept_smem));
    PMP_CUDA(cudaFuncSetAttribute(accept_lean_kernel<PMP_ALGO_MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_smem_bytes(LEAN_MAX_P, PMP_ALGO_MP)));
    PMP_CUDA(cudaFuncSetAttribute(accept_lean_kernel<PMP_ALGO_PSP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_smem_bytes(LEAN_MAX_P, PMP_ALGO_PSP)));
    PMP_CUDA(cudaFuncSetAttribute(accept_lean_kernel<PMP_ALGO_TABLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_smem_bytes(LEAN_MAX_P, PMP_ALGO_TABLE)));
    if (world_size > 1) {
        PMP_REQUIRE(nccl_unique_id, "world_size > 1 needs an NCCL unique id");
        int rc = load_nccl(); if (rc) { delete c; return rc; }
        static_assert(sizeof(ncclUniqueId) == sizeof(c->nccl_id), "NCCL unique id size");
        memcpy(c->nccl_id, nccl_unique_id, sizeof(c->nccl_id));
        if (env_int("PMP_NCCL_EAGER", 0)) { rc = ensure_comm(c); if (rc) { delete c; return rc; } }
    }
    *out = c;
    return PMP_OK;
}

int pmp_fc_destroy(pmp_ctx* ctx);       // fc_sweep.cu
int pmp_cnn_destroy(pmp_ctx* ctx);      // cnn_sweep.cuh
int pmp_chains_destroy(pmp_ctx* ctx);   // chains.cu

int pmp_destroy(pmp_ctx* c) {
    if (!c) return PMP_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    drop_graph(c);
    pmp_fc_destroy(c);
    pmp_cnn_destroy(c);
    pmp_glm_destroy(c);
    pmp_chains_destroy(c);
    if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    if (c->data_borrowed) { c->d_x = nullptr; c->d_y = nullptr; c->d_bimg = nullptr; }     // owned by another ctx (pmp_share_data)
    for (int r = 0; r < PEER_MAX_WORLD; ++r) if (c->peer_xchg[r] && r != c->rank) cudaIpcCloseMemHandle(c->peer_xchg[r]);
    if (c->d_xchg) cudaFree(c->d_xchg);
    void* ptrs[] = {c->d_x, c->d_y, c->d_state, c->d_props, c->d_acc, c->d_lt, c->d_logw, c->d_draws, c->d_uniforms, c->d_cnt,
                    c->trace.state, c->trace.next, c->trace.draws, c->trace.samples, c->trace.logw, c->d_flush, c->d_z, c->d_done, c->d_dbg, c->d_psync, c->d_hs, c->d_bimg, c->d_kt_s1, c->d_kt_dj2, c->d_kt_dot, c->d_hmc};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (cudaEvent_t ev : c->ev_pool) cudaEventDestroy(ev);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    delete c;
    return PMP_OK;
}

int pmp_device_info(pmp_ctx* c, int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
    PMP_REQUIRE(c, "ctx is NULL");
    cudaDeviceProp prop;
    PMP_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
    return PMP_OK;
}

int pmp_configure(pmp_ctx* c, const pmp_config* cfg) {
    PMP_REQUIRE(c && cfg, "NULL argument");
    PMP_CUDA(cudaSetDevice(c->device));
    long long P;
    if (cfg->tree == PMP_TREE_FLAT) P = cfg->b;
    else if (cfg->tree == PMP_TREE_BINARY) P = 1ll << cfg->depth;
    else if (cfg->tree == PMP_TREE_BARY) P = ipow(cfg->b, cfg->depth);
    else { set_error("unknown tree kind %d", cfg->tree); return PMP_ERR_ARG; }
    PMP_REQUIRE(cfg->tree == PMP_TREE_FLAT || (cfg->depth >= 1 && cfg->depth <= 13), "depth %d out of range", cfg->depth);
    PMP_REQUIRE(cfg->tree == PMP_TREE_BINARY || cfg->b >= 1, "b=%d must be >= 1", cfg->b);
    PMP_REQUIRE(P >= 1 && P <= MAX_NODES, "P=%lld nodes out of range [1,%d]", P, MAX_NODES);
    PMP_REQUIRE(cfg->dim >= 1, "dim must be >= 1");
    PMP_REQUIRE(cfg->target >= 0 && cfg->target <= PMP_TARGET_CNN, "unknown target %d", cfg->target);
    PMP_REQUIRE(cfg->algo >= 0 && cfg->algo <= PMP_ALGO_TABLE, "unknown algo %d", cfg->algo);
    PMP_REQUIRE(cfg->draw >= 0 && cfg->draw <= PMP_DRAW_SINGLE, "unknown draw rule %d", cfg->draw);
    if (cfg->target == PMP_TARGET_LINEAR_GAUSS) PMP_REQUIRE(cfg->dim == 3, "linear-Gaussian target has dim 3 (b0,b1,sigma), got %d", cfg->dim);
    if (cfg->target == PMP_TARGET_NORMAL1D) PMP_REQUIRE(cfg->dim == 1, "NORMAL1D has dim 1");
    if (cfg->target == PMP_TARGET_BANANA) PMP_REQUIRE(cfg->dim == 2, "BANANA has dim 2");
    if (cfg->algo == PMP_ALGO_MH || cfg->algo == PMP_ALGO_BARKER) PMP_REQUIRE(P == 2, "MH/BARKER need P == 2 (FLAT, b=2), got %lld", P);
    if (cfg->algo == PMP_ALGO_PSP) PMP_REQUIRE(cfg->tree == PMP_TREE_BINARY, "PSP needs the BINARY tree");
    if (cfg->algo == PMP_ALGO_PMP) PMP_REQUIRE(cfg->tree == PMP_TREE_BARY || cfg->tree == PMP_TREE_BINARY, "PMP needs a BARY/BINARY tree");
    if (cfg->algo == PMP_ALGO_MP && !(cfg->flags & PMP_FLAG_NO_KERNEL_TERM) && cfg->target != PMP_TARGET_FC && cfg->target != PMP_TARGET_CNN && cfg->target != PMP_TARGET_EXTERNAL && cfg->target != PMP_TARGET_GLM_LOGISTIC && cfg->target != PMP_TARGET_GLM_GAUSS)
        PMP_REQUIRE(cfg->dim <= KDIM_MAX, "in-kernel MP kernel term supports dim <= %d for this target", KDIM_MAX);
    PMP_REQUIRE(cfg->scale != 0.f && cfg->kernel_sigma > 0.f, "scale must be non-zero and kernel_sigma > 0");

    PMP_CUDA(cudaStreamSynchronize(c->stream));
    drop_graph(c);
    bool realloc_state = !c->configured || c->cfg.dim != cfg->dim;
    c->cfg = *cfg; c->P = (int)P;
    int rc;
    if (realloc_state) {
        if ((rc = dev_alloc(&c->d_state, (size_t)cfg->dim))) return rc;
        PMP_CUDA(cudaMemset(c->d_state, 0, cfg->dim * sizeof(float)));
    }
    if ((rc = dev_alloc(&c->d_props, (size_t)P * cfg->dim))) return rc;
    if ((rc = dev_alloc(&c->d_acc, (size_t)P))) return rc;
    if ((rc = dev_alloc(&c->d_lt, (size_t)P))) return rc;
    if ((rc = dev_alloc(&c->d_logw, (size_t)P))) return rc;
    if ((rc = dev_alloc(&c->d_draws, (size_t)P))) return rc;
    if ((rc = dev_alloc(&c->d_uniforms, (size_t)P + 1))) return rc;
    if ((rc = dev_alloc(&c->d_z, cfg->target == PMP_TARGET_LINEAR_GAUSS ? (size_t)2 * P * cfg->dim : 0))) return rc;
    c->z_valid_iter = -1;
    PMP_CUDA(cudaMemset(c->d_acc, 0, P * sizeof(unsigned long long)));
    PMP_CUDA(cudaMemset(c->d_props, 0, (size_t)P * cfg->dim * sizeof(float)));
    PMP_CUDA(cudaMemset(c->d_lt, 0, P * sizeof(double)));
    PMP_CUDA(cudaMemset(c->d_logw, 0, P * sizeof(double)));
    c->lt_valid = false; c->acc_pending = false;
    c->configured = true;
    // trace buffers depend on P and dim
    { void* tb[] = {c->trace.state, c->trace.next, c->trace.draws, c->trace.samples, c->trace.logw}; for (void* p : tb) if (p) cudaFree(p); }
    c->trace = TraceBuffers{};
    return PMP_OK;
}

int pmp_num_nodes(pmp_ctx* c) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    return c->P;
}

int pmp_set_data_linear(pmp_ctx* c, const float* x, const float* y, int64_t n_local, int64_t n_offset, int64_t n_global) {
    PMP_REQUIRE(c, "ctx is NULL");
    PMP_REQUIRE(n_local >= 0 && n_global >= n_local && n_offset >= 0 && n_offset + n_local <= n_global, "bad shard [%lld,+%lld) of %lld",
                (long long)n_offset, (long long)n_local, (long long)n_global);
    PMP_REQUIRE(n_local == 0 || (x && y), "x/y NULL");
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    drop_graph(c);
    if (c->data_borrowed) { c->d_x = nullptr; c->d_y = nullptr; c->d_bimg = nullptr; c->data_borrowed = false; c->data_capacity = 0; }
    int rc;
    size_t padded = ((size_t)n_local + 3) / 4 * 4 + 4;
    const long long nchunks = ((long long)n_local + CHUNK - 1) / CHUNK;
    const bool reuse = c->d_x && c->d_y && c->data_capacity == padded && (c->d_bimg || nchunks == 0);   // same shard size: refill in place (contexts aliasing it stay valid)
    if (!reuse) {
        if ((rc = dev_alloc(&c->d_x, padded))) return rc;
        if ((rc = dev_alloc(&c->d_y, padded))) return rc;
        if ((rc = dev_alloc(&c->d_bimg, (size_t)nchunks * tc::CHUNK_BYTES))) return rc;
        c->data_capacity = padded;
    }
    PMP_CUDA(cudaMemsetAsync(c->d_x, 0, padded * sizeof(float), c->stream));
    PMP_CUDA(cudaMemsetAsync(c->d_y, 0, padded * sizeof(float), c->stream));
    if (n_local) {
        PMP_CUDA(cudaMemcpyAsync(c->d_x, x, n_local * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        PMP_CUDA(cudaMemcpyAsync(c->d_y, y, n_local * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    }
    // data operand of the tensor-core sweep: the shared-memory image of every 64-point chunk, written once
    if (nchunks) {
        tc::build_data_image_kernel<<<(unsigned)((nchunks * CHUNK + 255) / 256), 256, 0, c->stream>>>(c->d_x, c->d_y, n_local, nchunks, c->d_bimg);
        c->launches++;
        PMP_CUDA(cudaGetLastError());
    }
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    c->n_local = n_local; c->n_offset = n_offset; c->n_global = n_global;
    return PMP_OK;
}

int pmp_set_state(pmp_ctx* c, const float* theta, int dim) {
    PMP_REQUIRE(c && c->configured && theta, "ctx not configured or theta NULL");
    PMP_REQUIRE(dim == c->cfg.dim, "dim %d != configured %d", dim, c->cfg.dim);
    PMP_CUDA(cudaMemcpyAsync(c->d_state, theta, dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_get_state(pmp_ctx* c, float* theta, int dim) {
    PMP_REQUIRE(c && c->configured && theta, "ctx not configured or theta NULL");
    PMP_REQUIRE(dim == c->cfg.dim, "dim %d != configured %d", dim, c->cfg.dim);
    PMP_CUDA(cudaMemcpyAsync(theta, c->d_state, dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_seed(pmp_ctx* c, uint64_t seed, uint64_t iteration) {
    PMP_REQUIRE(c, "ctx is NULL");
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    if (seed != c->seed) drop_graph(c);   // the key is a baked kernel argument
    c->seed = seed;
    c->host_iter = iteration; c->z_valid_iter = -1;
    unsigned long long it = iteration;
    PMP_CUDA(cudaMemcpyAsync(&c->d_cnt->iteration, &it, sizeof(it), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_get_iteration(pmp_ctx* c, uint64_t* iteration) {
    PMP_REQUIRE(c && iteration, "NULL argument");
    unsigned long long it;
    PMP_CUDA(cudaMemcpyAsync(&it, &c->d_cnt->iteration, sizeof(it), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    *iteration = it;
    return PMP_OK;
}

int pmp_propose(pmp_ctx* c) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_CUDA(cudaSetDevice(c->device));
    c->lt_valid = false;
    return launch_propose(c);
}

int pmp_read_proposals(pmp_ctx* c, float* out, int64_t count) {
    PMP_REQUIRE(c && c->configured && out, "ctx not configured or out NULL");
    PMP_REQUIRE(count == (int64_t)c->P * c->cfg.dim, "count %lld != P*dim = %lld", (long long)count, (long long)c->P * c->cfg.dim);
    PMP_CUDA(cudaMemcpyAsync(out, c->d_props, count * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_write_proposals(pmp_ctx* c, const float* in, int64_t count) {
    PMP_REQUIRE(c && c->configured && in, "ctx not configured or in NULL");
    PMP_REQUIRE(count == (int64_t)c->P * c->cfg.dim, "count %lld != P*dim = %lld", (long long)count, (long long)c->P * c->cfg.dim);
    PMP_CUDA(cudaMemcpyAsync(c->d_props, in, count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    c->lt_valid = false;
    c->props_external = true;
    return PMP_OK;
}

int pmp_fc_loglik(pmp_ctx* c);   // fc_sweep.cu: fills d_lt for PMP_TARGET_FC
int pmp_large_dim_kernel_term(pmp_ctx* c);   // fc_sweep.cu: MP kernel term for dim > KDIM_MAX into d_logw

// internal (not in the public header): exact cross-rank sum of fixed-point partials on the ctx stream
int pmp_allreduce_u64(pmp_ctx* c, unsigned long long* buf, size_t count) {
    if (c->world <= 1) return PMP_OK;
    { int rc = ensure_comm(c); if (rc) return rc; }
    PMP_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)c->nccl_comm, c->stream));
    return PMP_OK;
}

int pmp_loglik(pmp_ctx* c, double* out_host) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_CUDA(cudaSetDevice(c->device));
    int rc;
    if (c->cfg.target == PMP_TARGET_LINEAR_GAUSS) {
        if ((rc = launch_sweep_linear(c, 0))) return rc;
        if ((rc = allreduce_acc(c))) return rc;
        if ((rc = launch_accept(c, 1, 1, 0, nullptr))) return rc;
    } else if (c->cfg.target == PMP_TARGET_FC) {
        if ((rc = pmp_fc_loglik(c))) return rc;
    } else if (c->cfg.target == PMP_TARGET_CNN) {
        if ((rc = pmp_cnn_loglik(c))) return rc;
    } else if (c->cfg.target == PMP_TARGET_GLM_LOGISTIC || c->cfg.target == PMP_TARGET_GLM_GAUSS) {
        if ((rc = pmp_glm_loglik(c))) return rc;
    } else if (c->cfg.target == PMP_TARGET_EXTERNAL) {
        PMP_REQUIRE(c->lt_valid, "EXTERNAL target: call pmp_write_logtarget first");
    } else {
        if ((rc = launch_accept(c, 0, 1, 0, nullptr))) return rc;
    }
    c->lt_valid = true;
    if (out_host) {
        PMP_CUDA(cudaMemcpyAsync(out_host, c->d_lt, c->P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PMP_OK;
}

int pmp_write_logtarget(pmp_ctx* c, const double* in, int64_t count) {
    PMP_REQUIRE(c && c->configured && in, "ctx not configured or in NULL");
    PMP_REQUIRE(count == c->P, "count %lld != P = %d", (long long)count, c->P);
    PMP_CUDA(cudaMemcpyAsync(c->d_lt, in, count * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    c->lt_valid = true;
    return PMP_OK;
}

static int uniforms_needed(const pmp_ctx* c) {
    if (c->cfg.algo == PMP_ALGO_MH || c->cfg.algo == PMP_ALGO_BARKER || c->cfg.draw == PMP_DRAW_SINGLE) return 1;
    return c->cfg.draw == PMP_DRAW_PYTHON ? c->P + 1 : c->P;
}

int pmp_accept(pmp_ctx* c, const double* uniforms, int64_t n_uniforms, int32_t* idx_out, int32_t* next_out) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_REQUIRE(c->lt_valid, "pmp_accept before pmp_loglik / pmp_write_logtarget");
    PMP_CUDA(cudaSetDevice(c->device));
    const double* d_u = nullptr;
    if (uniforms) {
        PMP_REQUIRE(n_uniforms == uniforms_needed(c), "need %d uniforms for this draw rule, got %lld", uniforms_needed(c), (long long)n_uniforms);
        PMP_CUDA(cudaMemcpyAsync(c->d_uniforms, uniforms, n_uniforms * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        d_u = c->d_uniforms;
    }
    int rc;
    if (c->cfg.target == PMP_TARGET_EXTERNAL && c->cfg.algo == PMP_ALGO_MP && !(c->cfg.flags & PMP_FLAG_NO_KERNEL_TERM) && c->cfg.dim > KDIM_MAX)
        if ((rc = pmp_large_dim_kernel_term(c))) return rc;         // kernel term of the current nodes, read by the acceptance from d_logw
    rc = launch_accept(c, 0, 0, 1, d_u);
    if (rc) return rc;
    c->lt_valid = false;
    c->host_iter++;
    if (idx_out || next_out || uniforms) {
        int nd = (c->cfg.algo == PMP_ALGO_MH || c->cfg.algo == PMP_ALGO_BARKER || c->cfg.draw == PMP_DRAW_SINGLE) ? 1 : c->P;
        if (idx_out) PMP_CUDA(cudaMemcpyAsync(idx_out, c->d_draws, nd * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        if (next_out) PMP_CUDA(cudaMemcpyAsync(next_out, &c->d_cnt->last_next, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
    }
    return PMP_OK;
}

int pmp_read_logweights(pmp_ctx* c, double* out, int64_t count) {
    PMP_REQUIRE(c && c->configured && out, "ctx not configured or out NULL");
    PMP_REQUIRE(count == c->P, "count %lld != P = %d", (long long)count, c->P);
    PMP_CUDA(cudaMemcpyAsync(out, c->d_logw, count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_trace_config(pmp_ctx* c, int64_t max_iters, uint32_t what) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_REQUIRE(max_iters >= 0, "max_iters < 0");
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    if (max_iters == c->trace.capacity && what == c->trace.what) {      // same ring: keep the buffers, rewind the cursor
        long long zero0 = 0;
        PMP_CUDA(cudaMemcpyAsync(&c->d_cnt->trace_rows, &zero0, sizeof(zero0), cudaMemcpyHostToDevice, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
        return PMP_OK;
    }
    drop_graph(c);
    int rc;
    size_t n = (size_t)max_iters, P = c->P, dim = c->cfg.dim;
    if ((rc = dev_alloc(&c->trace.state, (what & PMP_TRACE_STATE) ? n * dim : 0))) return rc;
    if ((rc = dev_alloc(&c->trace.next, (what & PMP_TRACE_NEXT) ? n : 0))) return rc;
    if ((rc = dev_alloc(&c->trace.draws, (what & PMP_TRACE_DRAWS) ? n * P : 0))) return rc;
    if ((rc = dev_alloc(&c->trace.samples, (what & PMP_TRACE_SAMPLES) ? n * P * dim : 0))) return rc;
    if ((rc = dev_alloc(&c->trace.logw, (what & PMP_TRACE_LOGW) ? n * P : 0))) return rc;
    c->trace.capacity = max_iters; c->trace.what = what;
    long long zero = 0;
    PMP_CUDA(cudaMemcpyAsync(&c->d_cnt->trace_rows, &zero, sizeof(zero), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_trace_reset(pmp_ctx* c) {
    PMP_REQUIRE(c, "ctx is NULL");
    long long zero = 0;
    PMP_CUDA(cudaMemcpyAsync(&c->d_cnt->trace_rows, &zero, sizeof(zero), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_read_trace(pmp_ctx* c, int64_t max_iters, float* state, int32_t* next, int32_t* draws, float* samples, double* logw,
                   int64_t* n_recorded) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    long long rows = 0;
    PMP_CUDA(cudaMemcpyAsync(&rows, &c->d_cnt->trace_rows, sizeof(rows), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    if (rows > max_iters) rows = max_iters;
    size_t n = (size_t)rows, P = c->P, dim = c->cfg.dim;
    if (state) { PMP_REQUIRE(c->trace.state, "STATE not traced"); PMP_CUDA(cudaMemcpyAsync(state, c->trace.state, n * dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); }
    if (next) { PMP_REQUIRE(c->trace.next, "NEXT not traced"); PMP_CUDA(cudaMemcpyAsync(next, c->trace.next, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream)); }
    if (draws) { PMP_REQUIRE(c->trace.draws, "DRAWS not traced"); PMP_CUDA(cudaMemcpyAsync(draws, c->trace.draws, n * P * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream)); }
    if (samples) { PMP_REQUIRE(c->trace.samples, "SAMPLES not traced"); PMP_CUDA(cudaMemcpyAsync(samples, c->trace.samples, n * P * dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); }
    if (logw) { PMP_REQUIRE(c->trace.logw, "LOGW not traced"); PMP_CUDA(cudaMemcpyAsync(logw, c->trace.logw, n * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream)); }
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    if (n_recorded) *n_recorded = rows;
    return PMP_OK;
}

// Hand-off buffers of the persistent kernels (zeroed once: tag 0 is never current).
static int ensure_handoff(pmp_ctx* c, Handoff* out) {
    const size_t nw = handoff_node_words(c->P), zw = handoff_z_words(c->P), sw = handoff_state_words();
    if (!c->d_hs || c->hs_P != c->P) {
        if (c->d_hs) { PMP_CUDA(cudaStreamSynchronize(c->stream)); cudaFree(c->d_hs); c->d_hs = nullptr; }
        PMP_CUDA(cudaMalloc((void**)&c->d_hs, (nw + zw + sw) * sizeof(unsigned long long)));
        PMP_CUDA(cudaMemsetAsync(c->d_hs, 0, (nw + zw + sw) * sizeof(unsigned long long), c->stream));
        c->hs_words = nw + zw + sw; c->hs_P = c->P;
    }
    out->nodes = c->d_hs; out->zt = c->d_hs + nw; out->epoch = c->hs_epoch; out->state = c->d_hs + nw + zw;
    return PMP_OK;
}

// Persistent cooperative chain loop (chain_persistent.cuh): single GPU, linear-Gaussian target, fast-acceptance rules, and a
// data slice per CTA that fits shared memory.  Returns 1 when it ran, 0 when the stepwise path must be used, < 0 on error.
static int try_run_persistent(pmp_ctx* c, int64_t iters) {
    if (!env_int("PMP_PERSISTENT", 1) || c->world != 1 || !fast_accept_ok(c) || iters < 2 || iters > 2000000000ll) return 0;
    if (c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) return 0;
    const int G = c->sm_count;                                     // one CTA per SM; the last one is the acceptance CTA
    const int n_sweep = G - 1;
    const long long nchunks = (c->n_local + CHUNK - 1) / CHUNK;
    if (nchunks == 0 || n_sweep < 1) return 0;
    const int ntiles = (c->P + PERSIST_PT - 1) / PERSIST_PT;
    const long long units = (long long)ntiles * nchunks;
    const long long max_chunks = persist_max_chunks(ntiles, nchunks, n_sweep);
    const size_t sweep_smem = (size_t)max_chunks * CHUNK_STRIDE * sizeof(float) + (size_t)PERSIST_TD * PERSIST_PT * sizeof(unsigned long long);
    if (!lean_accept_ok(c) || !persist_segments_fit(units, n_sweep, nchunks)) return 0;
    const size_t accept_smem = lean_smem_bytes(c->P, c->cfg.algo);
    const size_t smem = sweep_smem > accept_smem ? sweep_smem : accept_smem;
    if (smem > 200 * 1024) return 0;
    if (!c->d_psync) { PMP_CUDA(cudaMalloc((void**)&c->d_psync, sizeof(PersistSync))); }
    PMP_CUDA(cudaMemsetAsync(c->d_psync, 0, sizeof(PersistSync), c->stream));
    int rc;
    if ((rc = launch_propose(c))) return rc;                       // nodes of the first iteration; later ones come from the acceptance CTA
    PersistArgs pa{};
    pa.sw = SweepArgs{c->d_x, c->d_y, c->d_props, c->d_acc, c->d_cnt, c->n_local, nchunks, c->P, PERSIST_TP, PERSIST_TD, sat_limit(c), 0, c->d_z,
                      ProposeArgs{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, 0},
                      c->d_dbg};
    pa.fa = AcceptFastArgs{make_accept_args(c, 1, 0, 1, nullptr), c->d_z, 1,
                           ProposeArgs{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, 0}};
    pa.sync = reinterpret_cast<PersistSync*>(c->d_psync);
    pa.iters = (int)iters;
    pa.max_chunks = (int)max_chunks;
    const bool hs = env_int("PMP_HANDOFF", 1) != 0;   // flag-in-data hand-offs (default) or release/acquire counters
    if (hs && (rc = ensure_handoff(c, &pa.hs))) return rc;
    // flat tree: a node is state + alpha * z(node) — the acceptance publishes the accepted state and every reader derives its nodes (Handoff::state)
    pa.derive = hs && (c->cfg.tree == PMP_TREE_FLAT || c->cfg.tree == PMP_TREE_BINARY) && c->cfg.dim == 3 && n_sweep >= ntiles && ntiles <= 16 && env_int("PMP_DERIVE_NODES", 1);
    void* kargs[] = {&pa};
    const void* fn;
#define PMP_SINGLE_FN(ALGO) (hs ? (const void*)chain_persistent_kernel<ALGO, true> : (const void*)chain_persistent_kernel<ALGO, false>)
    switch (c->cfg.algo) {
        case PMP_ALGO_MP: fn = PMP_SINGLE_FN(PMP_ALGO_MP); break;
        case PMP_ALGO_PSP: fn = PMP_SINGLE_FN(PMP_ALGO_PSP); break;
        default: fn = PMP_SINGLE_FN(PMP_ALGO_TABLE); break;
    }
#undef PMP_SINGLE_FN
    PMP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PERSIST_THREADS, smem));
    if (per_sm < 1) return 0;
    PMP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PERSIST_THREADS), kargs, smem, c->stream));
    if (hs) c->hs_epoch += (unsigned int)iters + (pa.derive ? 1u : 0u);     // derive: the tag after the last iteration's is used too (state and normals of the iteration after the launch)
    c->launches++;
    c->host_iter += (unsigned long long)iters;
    c->z_valid_iter = -1;
    c->lt_valid = false;
    return 1;
}

// Persistent cooperative chain loop with the tensor-core sweep (chain_persistent_tc.cuh): P <= 1024, a data slice per CTA
// that fits shared memory, three-phase acceptance.  Opt-in (PMP_PERSISTENT_TC=1): measured 20.4 us per iteration at P=1024,
// n=100000 against 19.7 us for the FFMA2 sweep CTAs.  Returns 1 when it ran, 0 when another path must be
// used, < 0 on error.
static int try_run_persistent_tc(pmp_ctx* c, int64_t iters) {
    if (!env_int("PMP_PERSISTENT", 1) || !env_int("PMP_PERSISTENT_TC", 0) || c->world != 1 || !lean_accept_ok(c) || iters < 2 || iters > 2000000000ll) return 0;
    if ((c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) || !c->d_bimg || c->P > tc::PT_MAX_TILES * tc::TILE_NODES) return 0;
    const int G = c->sm_count, n_sweep = G - 1;
    const int ntiles = (c->P + tc::TILE_NODES - 1) / tc::TILE_NODES;
    const long long nchunks = (c->n_local + CHUNK - 1) / CHUNK;
    if (nchunks == 0 || n_sweep < 1) return 0;
    const long long units = nchunks * ntiles;
    const long long per = (units + n_sweep - 1) / n_sweep;
    const long long mc = (per + ntiles - 1) / ntiles + 1;
    if (per > tc::MAX_UNITS || !persist_segments_fit(units, n_sweep, nchunks)) return 0;
    const size_t sweep_smem = tc::pt_smem_bytes(ntiles, (int)mc, (int)per), accept_smem = lean_smem_bytes(c->P, c->cfg.algo);
    const size_t smem = sweep_smem > accept_smem ? sweep_smem : accept_smem;
    if (smem > 200 * 1024) return 0;
    if (!c->d_psync) { PMP_CUDA(cudaMalloc((void**)&c->d_psync, sizeof(PersistSync))); }
    PMP_CUDA(cudaMemsetAsync(c->d_psync, 0, sizeof(PersistSync), c->stream));
    int rc;
    if ((rc = launch_propose(c))) return rc;                       // nodes of the first iteration; later ones come from the acceptance CTA
    const ProposeArgs gen{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, 0};
    tc::PersistTcArgs pa{};
    pa.sw = tc::Args{c->d_bimg, c->d_props, c->d_acc, c->d_cnt, nchunks, c->P, (int)mc, (int)per, sat_limit(c), 0, c->d_z, gen, c->d_dbg};
    pa.fa = AcceptFastArgs{make_accept_args(c, 1, 0, 1, nullptr), c->d_z, 1, gen};
    pa.sync = reinterpret_cast<PersistSync*>(c->d_psync);
    pa.iters = (int)iters;
    void* kargs[] = {&pa};
    const void* fn;
    switch (c->cfg.algo) {
        case PMP_ALGO_MP: fn = (const void*)tc::chain_persistent_tc_kernel<PMP_ALGO_MP>; break;
        case PMP_ALGO_PSP: fn = (const void*)tc::chain_persistent_tc_kernel<PMP_ALGO_PSP>; break;
        default: fn = (const void*)tc::chain_persistent_tc_kernel<PMP_ALGO_TABLE>; break;
    }
    PMP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, tc::PT_THREADS, smem));
    if (per_sm < 1) return 0;
    PMP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(tc::PT_THREADS), kargs, smem, c->stream));
    c->launches++;
    c->host_iter += (unsigned long long)iters;
    c->z_valid_iter = -1;
    c->lt_valid = false;
    return 1;
}

// The chain loop.  A CUDA graph of GRAPH_ITERS iterations is captured once per configuration and replayed, so the
// host issues one launch per GRAPH_ITERS iterations; the iteration counter, the state and the trace cursor live on the
// device, so replay needs no parameter update.
static bool peer_fused_ok(pmp_ctx** cs, int K, int64_t iters);
static int run_multi_impl(pmp_ctx** cs, int K, int64_t iters, cudaEvent_t ev_begin, cudaEvent_t ev_end);
static thread_local bool g_in_multi_streams = false;    // set while run_multi_streams drives run_impl: that mode is the NCCL path by definition

static int run_impl(pmp_ctx* c, int64_t iters) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_REQUIRE(iters >= 0, "iters < 0");
    PMP_CUDA(cudaSetDevice(c->device));
    const int GI = env_int("PMP_GRAPH_ITERS", 32);
    int rc;
    if (c->cfg.target == PMP_TARGET_LINEAR_GAUSS) {
        if (c->world > 1 && !g_in_multi_streams) {      // a single sharded chain: the same cooperative kernel, sums exchanged over NVLink peer memory
            pmp_ctx* one[1] = {c};
            if (peer_fused_ok(one, 1, iters)) return run_multi_impl(one, 1, iters, nullptr, nullptr);
        }
        rc = try_run_persistent_tc(c, iters); if (rc != 0) return rc < 0 ? rc : PMP_OK;
        rc = try_run_persistent(c, iters); if (rc != 0) return rc < 0 ? rc : PMP_OK;
    }
    if ((rc = ensure_comm(c))) return rc;             // stepwise loop with NCCL between kernels: the communicator must exist before any stream capture
    int64_t done = 0;
    if (GI > 1 && iters >= GI && c->cfg.target != PMP_TARGET_FC && c->cfg.target != PMP_TARGET_CNN && c->cfg.target != PMP_TARGET_GLM_LOGISTIC && c->cfg.target != PMP_TARGET_GLM_GAUSS) {     // FC iterations are milliseconds of GEMMs: nothing to gain from a graph
        if (!c->graph_exec || c->graph_iters != GI) {
            drop_graph(c);
            cudaGraph_t graph;
            const long long launches_before = c->launches;
            const unsigned long long iter_before = c->host_iter;
            const long long zvalid_before = c->z_valid_iter;
            PMP_CUDA(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            rc = PMP_OK;
            for (int i = 0; i < GI && rc == PMP_OK; ++i) rc = enqueue_iteration(c, nullptr, nullptr, i == 0);
            cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
            if (rc) { if (e == cudaSuccess) cudaGraphDestroy(graph); c->host_iter = iter_before; c->z_valid_iter = zvalid_before; return rc; }
            PMP_CUDA(e);
            PMP_CUDA(cudaGraphInstantiate(&c->graph_exec, graph, 0));
            cudaGraphDestroy(graph);
            c->graph_iters = GI;
            c->graph_launches_total = c->launches - launches_before;
            c->launches = launches_before;   // enqueues during capture are not launches
            c->host_iter = iter_before; c->z_valid_iter = zvalid_before;
        }
        for (; done + GI <= iters; done += GI) {
            PMP_CUDA(cudaGraphLaunch(c->graph_exec, c->stream));
            c->launches += c->graph_launches_total;
            c->host_iter += GI;
            c->z_valid_iter = (env_int("PMP_FUSE_PROPOSE", 1) && fast_accept_ok(c) && c->n_local > 0) ? (long long)c->host_iter : -1;
        }
    }
    for (; done < iters; ++done) if ((rc = enqueue_iteration(c, nullptr, nullptr, false))) return rc;
    c->lt_valid = false;
    return PMP_OK;
}

int pmp_run(pmp_ctx* c, int64_t iters, int sync) {
    int rc = run_impl(c, iters);
    if (rc) return rc;
    if (sync) PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

// ---- co-scheduled chains (chain_persistent_multi.cuh) ---------------------------------------------------------------------
int pmp_share_data(pmp_ctx* dst, pmp_ctx* src) {
    PMP_REQUIRE(dst && src && dst != src, "bad arguments");
    PMP_REQUIRE(dst->device == src->device, "contexts live on different devices (%d, %d)", dst->device, src->device);
    PMP_REQUIRE(src->d_x && !src->data_borrowed, "src owns no linear-Gaussian data (pmp_set_data_linear it first)");
    PMP_CUDA(cudaSetDevice(dst->device));
    PMP_CUDA(cudaStreamSynchronize(dst->stream));
    PMP_CUDA(cudaStreamSynchronize(src->stream));
    drop_graph(dst);
    if (!dst->data_borrowed) { if (dst->d_x) cudaFree(dst->d_x); if (dst->d_y) cudaFree(dst->d_y); if (dst->d_bimg) cudaFree(dst->d_bimg); }
    dst->d_x = src->d_x; dst->d_y = src->d_y; dst->d_bimg = src->d_bimg;
    dst->n_local = src->n_local; dst->n_offset = src->n_offset; dst->n_global = src->n_global;
    dst->data_borrowed = true;
    dst->data_capacity = 0;
    return PMP_OK;
}

static int run_impl(pmp_ctx* c, int64_t iters);

// Sharded chains (world_size > 1): no joint kernel — every chain runs its own stepwise loop (sweep → NCCL all-reduce →
// acceptance, CUDA-graph replayed) on its own stream with its own communicator, so one chain's all-reduce and one-CTA
// acceptance overlap the other chains' sweeps.  Fork/join with events so that [ev_begin, ev_end] on ctx 0's stream
// brackets all of it.
static int run_multi_streams(pmp_ctx** cs, int K, int64_t iters, cudaEvent_t ev_begin, cudaEvent_t ev_end) {
    pmp_ctx* c0 = cs[0];
    PMP_CUDA(cudaSetDevice(c0->device));
    if (ev_begin) PMP_CUDA(cudaEventRecord(ev_begin, c0->stream));
    PMP_CUDA(cudaEventRecord(c0->ev0, c0->stream));
    for (int k = 1; k < K; ++k) PMP_CUDA(cudaStreamWaitEvent(cs[k]->stream, c0->ev0, 0));
    int rc = PMP_OK;
    g_in_multi_streams = true;
    for (int k = 0; k < K && rc == PMP_OK; ++k) rc = run_impl(cs[k], iters);
    g_in_multi_streams = false;
    if (rc) return rc;
    for (int k = 1; k < K; ++k) {
        PMP_CUDA(cudaEventRecord(cs[k]->ev1, cs[k]->stream));
        PMP_CUDA(cudaStreamWaitEvent(c0->stream, cs[k]->ev1, 0));
    }
    if (ev_end) PMP_CUDA(cudaEventRecord(ev_end, c0->stream));
    return PMP_OK;
}

// world_size > 1: can these chains run in the cooperative kernel that exchanges the sums over NVLink peer memory?
// Every rank must take the SAME decision (a rank in the cooperative kernel spins on its peers' tags while a rank on the NCCL path blocks
// in ncclAllReduce), so the shape tests use rank-independent quantities only: the largest shard any rank can hold when the rows are
// split into 64-point blocks as evenly as possible (dist.shard_bounds) — derived from n_global and world_size — never this rank's own
// n_local.  A rank whose shard is larger than that bound fails loudly (PMP_ERR_ARG) instead of silently choosing another path.
static long long max_shard_chunks(const pmp_ctx* c) {
    const long long blocks = (c->n_global + CHUNK - 1) / CHUNK;
    return (blocks + c->world - 1) / c->world;
}
static bool peer_fused_ok(pmp_ctx** cs, int K, int64_t iters) {
    pmp_ctx* c0 = cs[0];
    if (!env_int("PMP_PEER_XCHG", 1) || c0->world > PEER_MAX_WORLD || iters < 2 || iters > 2000000000ll || !env_int("PMP_PERSISTENT", 1)) return false;
    const long long nchunks = max_shard_chunks(c0);
    const int n_sweep = c0->sm_count - (K < PERSIST_MAX_ACCEPT ? K : PERSIST_MAX_ACCEPT);
    if (nchunks == 0 || n_sweep < 1) return false;
    const long long units = (long long)((c0->P + PERSIST_PT - 1) / PERSIST_PT) * nchunks;
    const size_t sweep_smem = (size_t)persist_max_chunks((c0->P + PERSIST_PT - 1) / PERSIST_PT, nchunks, n_sweep) * CHUNK_STRIDE * sizeof(float) + (size_t)PERSIST_TD * PERSIST_PT * sizeof(unsigned long long);
    if (sweep_smem > 200 * 1024 || !persist_segments_fit(units, n_sweep, nchunks)) return false;
    for (int k = 0; k < K; ++k) {
        pmp_ctx* c = cs[k];
        if (!(c->peers_attached && c->cfg.target == PMP_TARGET_LINEAR_GAUSS && lean_accept_ok(c) && !(c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) &&
              c->P == c0->P && c->cfg.algo == c0->cfg.algo && c->d_x == c0->d_x && c->n_local == c0->n_local && c->n_global == c0->n_global)) return false;
    }
    return true;
}

static int run_multi_impl(pmp_ctx** cs, int K, int64_t iters, cudaEvent_t ev_begin, cudaEvent_t ev_end) {
    PMP_REQUIRE(cs && K >= 1 && K <= PERSIST_MAX_CHAINS, "1..%d contexts", PERSIST_MAX_CHAINS);
    PMP_REQUIRE(iters >= 2 && iters <= 2000000000ll, "iters out of range");
    pmp_ctx* c0 = cs[0];
    PMP_REQUIRE(c0 && c0->configured, "ctx 0 not configured");
    if (c0->world > 1) {
        for (int k = 0; k < K; ++k) {
            PMP_REQUIRE(cs[k] && cs[k]->configured && cs[k]->device == c0->device && cs[k]->world == c0->world && cs[k]->rank == c0->rank, "ctx %d: not configured or another rank / device", k);
            for (int j = 0; j < k; ++j) PMP_REQUIRE(cs[j] != cs[k], "ctx %d given twice", k);
        }
        if (!peer_fused_ok(cs, K, iters)) return run_multi_streams(cs, K, iters, ev_begin, ev_end);
    }
    for (int k = 0; k < K; ++k) {
        pmp_ctx* c = cs[k];
        PMP_REQUIRE(c && c->configured, "ctx %d not configured", k);
        for (int j = 0; j < k; ++j) PMP_REQUIRE(cs[j] != c, "ctx %d given twice", k);
        PMP_REQUIRE(c->device == c0->device && c->world == c0->world, "co-scheduled chains: one device, one world");
        PMP_REQUIRE(c->cfg.target == PMP_TARGET_LINEAR_GAUSS && lean_accept_ok(c) && !(c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL),
                    "co-scheduled chains need the linear-Gaussian target with a device-resident acceptance rule (MP / PSP / TABLE quirk), P <= %d", LEAN_MAX_P);
        PMP_REQUIRE(c->P == c0->P && c->cfg.algo == c0->cfg.algo && c->cfg.dim == c0->cfg.dim, "ctx %d: P / algo differ from ctx 0", k);
        PMP_REQUIRE(c->d_x == c0->d_x && c->d_y == c0->d_y && c->n_local == c0->n_local && c->n_global == c0->n_global,
                    "ctx %d does not share ctx 0's data (pmp_share_data)", k);
    }
    PMP_CUDA(cudaSetDevice(c0->device));
    // Acceptance CTAs, up to one per chain: an acceptance is ~10 us of dependent latencies on one SM, so with fewer acceptance CTAs than chains
    // the acceptances — not the sweeps — bound the throughput as soon as a sweep is short (measured, n = 12 500, K = 8: 6.2 us per chain
    // iteration with 2 acceptance CTAs, 2.8 us with 8).  PMP_MULTI_ACCEPT overrides (1..min(K, PERSIST_MAX_ACCEPT)).
    const int acc_cap = K < PERSIST_MAX_ACCEPT ? K : PERSIST_MAX_ACCEPT;
    // measured (scripts/tune_multi.py, K x acceptance CTAs x groups at n = 100 000 / N): one GPU with the full dataset: K / 4 (K = 16: 10.2 us per
    // chain iteration with 4, 10.8 with 8); a shard: K / 2 (n = 12 500, K = 24: 1.73 us with 12, 1.97 with 24, 2.08 with 6)
    int n_accept = c0->world > 1 ? (K <= 2 ? K : (K / 2 < acc_cap ? K / 2 : acc_cap)) : (K < 4 ? K : (K / 4 < acc_cap ? K / 4 : acc_cap));
    { const int ov = env_int("PMP_MULTI_ACCEPT", 0); if (ov >= 1 && ov <= acc_cap) n_accept = ov; }
    const int G = c0->sm_count, n_sweep = G - n_accept;
    // world > 1: the kernel's shared-memory plan follows the LARGEST shard of any rank (rank-independent, see peer_fused_ok); this rank's own
    // shard must not exceed it.  An empty shard is fine: its sweep CTAs have no units, only arrive.
    const long long nchunks = (c0->n_local + CHUNK - 1) / CHUNK;
    const long long plan_chunks = c0->world > 1 ? max_shard_chunks(c0) : nchunks;
    PMP_REQUIRE(nchunks <= plan_chunks, "this rank's shard (%lld blocks of 64 points) exceeds the even split (%lld blocks): shard with dist.shard_bounds or set PMP_PEER_XCHG=0", nchunks, plan_chunks);
    PMP_REQUIRE(plan_chunks > 0 && n_sweep >= 1, "no data");
    const int ntiles = (c0->P + PERSIST_PT - 1) / PERSIST_PT;
    const long long max_chunks = persist_max_chunks(ntiles, plan_chunks, n_sweep);
    PMP_REQUIRE(persist_segments_fit((long long)ntiles * plan_chunks, n_sweep, plan_chunks), "co-scheduled chains: a sweep CTA would span more than %d node tiles", PERSIST_MAX_SEGS);
    const size_t sweep_smem = (size_t)max_chunks * CHUNK_STRIDE * sizeof(float) + (size_t)PERSIST_TD * PERSIST_PT * sizeof(unsigned long long);
    const size_t accept_smem = lean_smem_bytes(c0->P, c0->cfg.algo);
    const size_t smem = sweep_smem > accept_smem ? sweep_smem : accept_smem;
    if (smem > 200 * 1024) { set_error("co-scheduled chains: a sweep CTA's data slice (%zu bytes) does not fit shared memory", smem); return PMP_ERR_UNSUPPORTED; }
    const bool hs = env_int("PMP_HANDOFF", 1) != 0;
    std::unique_ptr<PersistMultiArgs> pa_owner(new PersistMultiArgs());      // per call: several host threads may drive different contexts (kernel parameters are copied at launch)
    PersistMultiArgs& pa = *pa_owner;
    int rc;
    for (int k = 0; k < K; ++k) {
        pmp_ctx* c = cs[k];
        if (!c->d_psync) { PMP_CUDA(cudaMalloc((void**)&c->d_psync, sizeof(PersistSync))); }
        PMP_CUDA(cudaMemsetAsync(c->d_psync, 0, sizeof(PersistSync), c->stream));
        if ((rc = launch_propose(c))) return rc;                   // nodes of the first iteration; later ones come from the acceptance CTA
        const ProposeArgs gen{c->d_state, c->d_props, c->d_cnt, c->seed, c->P, c->cfg.dim, c->cfg.tree, c->cfg.b, c->cfg.depth, c->cfg.alpha, 0};
        pa.ch[k].sw = SweepArgs{c->d_x, c->d_y, c->d_props, c->d_acc, c->d_cnt, c->n_local, nchunks, c->P, PERSIST_TP, PERSIST_TD, sat_limit(c), 0, c->d_z, gen, nullptr};
        AcceptArgs aa = make_accept_args(c, 1, 0, 1, nullptr);
        aa.dbg = nullptr;
        pa.ch[k].fa = AcceptFastArgs{aa, c->d_z, 1, gen};
        pa.ch[k].sync = reinterpret_cast<PersistSync*>(c->d_psync);
        pa.ch[k].xchg.world = c->world; pa.ch[k].xchg.me = c->rank; pa.ch[k].xchg.local = c->d_xchg; pa.ch[k].xchg.base = c->xchg_count;
        for (int r = 0; r < PEER_MAX_WORLD; ++r) pa.ch[k].xchg.peer[r] = c->peer_xchg[r];
        if (hs && (rc = ensure_handoff(c, &pa.ch[k].hs))) return rc;
        PMP_CUDA(cudaStreamSynchronize(c->stream));                // everything queued on this chain's own stream is done before the joint launch
    }
    pa.n_chains = K; pa.iters = (int)iters; pa.max_chunks = (int)max_chunks; pa.n_accept = n_accept;
    pa.derive = hs && K == 1 && c0->cfg.tree == PMP_TREE_FLAT && c0->cfg.dim == 3 && (G - n_accept) >= (c0->P + PERSIST_PT - 1) / PERSIST_PT && env_int("PMP_DERIVE_NODES", 1);
    void* kargs[] = {&pa};
    const void* fn;
    const int groups_dflt = K >= 8 ? 8 : (K >= 4 ? 4 : 2);                             // warp groups per sweep CTA (measured, scripts/tune_multi.py)
    int groups = env_int("PMP_MULTI_GROUPS", groups_dflt);
    if (groups != 2 && groups != 4 && groups != 8) groups = groups_dflt;
#define PMP_MULTI_FN2(ALGO, HSV) (groups == 8 ? (const void*)chain_persistent_multi_kernel<ALGO, 8, HSV> : groups == 4 ? (const void*)chain_persistent_multi_kernel<ALGO, 4, HSV> : (const void*)chain_persistent_multi_kernel<ALGO, 2, HSV>)
#define PMP_MULTI_FN(ALGO) (hs ? PMP_MULTI_FN2(ALGO, true) : PMP_MULTI_FN2(ALGO, false))
    switch (c0->cfg.algo) {
        case PMP_ALGO_MP: fn = PMP_MULTI_FN(PMP_ALGO_MP); break;
        case PMP_ALGO_PSP: fn = PMP_MULTI_FN(PMP_ALGO_PSP); break;
        default: fn = PMP_MULTI_FN(PMP_ALGO_TABLE); break;
    }
#undef PMP_MULTI_FN
#undef PMP_MULTI_FN2
    PMP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PMP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PERSIST_THREADS, smem));
    if (per_sm < 1) { set_error("co-scheduled chains: the cooperative kernel does not fit an SM"); return PMP_ERR_UNSUPPORTED; }
    if (ev_begin) PMP_CUDA(cudaEventRecord(ev_begin, c0->stream));
    PMP_CUDA(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PERSIST_THREADS), kargs, smem, c0->stream));
    if (ev_end) PMP_CUDA(cudaEventRecord(ev_end, c0->stream));
    c0->launches++;
    PMP_CUDA(cudaEventRecord(c0->ev1, c0->stream));
    for (int k = 0; k < K; ++k) {
        pmp_ctx* c = cs[k];
        if (k > 0) PMP_CUDA(cudaStreamWaitEvent(c->stream, c0->ev1, 0));   // later work on the chain's own stream is ordered after the joint kernel
        if (c->world > 1) c->xchg_count += (unsigned long long)iters;       // only once the launch is in the stream: a failed launch must not shift the tags
        if (hs) c->hs_epoch += (unsigned int)iters + (pa.derive ? 1u : 0u);
        c->host_iter += (unsigned long long)iters;
        c->z_valid_iter = -1;
        c->lt_valid = false;
    }
    return PMP_OK;
}

int pmp_run_multi(pmp_ctx** ctxs, int n_ctx, int64_t iters, int sync) {
    int rc = run_multi_impl(ctxs, n_ctx, iters, nullptr, nullptr);
    if (rc) return rc;
    if (sync) for (int k = 0; k < n_ctx; ++k) PMP_CUDA(cudaStreamSynchronize(ctxs[k]->stream));
    return PMP_OK;
}

int pmp_run_multi_timed(pmp_ctx** ctxs, int n_ctx, int64_t iters, float* total_ms) {
    PMP_REQUIRE(ctxs && n_ctx >= 1 && ctxs[0] && total_ms, "bad arguments");
    cudaEvent_t e0, e1;
    PMP_CUDA(cudaEventCreate(&e0)); PMP_CUDA(cudaEventCreate(&e1));
    int rc = run_multi_impl(ctxs, n_ctx, iters, e0, e1);
    if (rc == PMP_OK) {
        for (int k = 0; k < n_ctx; ++k) cudaStreamSynchronize(ctxs[k]->stream);
        if (cudaEventElapsedTime(total_ms, e0, e1) != cudaSuccess) { set_error("cudaEventElapsedTime failed"); rc = PMP_ERR_CUDA; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return rc;
}

int pmp_peer_exchange_handle(pmp_ctx* c, void* out64) {
    PMP_REQUIRE(c && out64, "bad arguments");
    PMP_REQUIRE(c->world > 1 && c->world <= PEER_MAX_WORLD, "peer exchange needs 2..%d ranks (world_size %d)", PEER_MAX_WORLD, c->world);
    PMP_CUDA(cudaSetDevice(c->device));
    if (!c->d_xchg) {
        PMP_CUDA(cudaMalloc((void**)&c->d_xchg, peer_xchg_words() * sizeof(unsigned long long)));
        PMP_CUDA(cudaMemset(c->d_xchg, 0, peer_xchg_words() * sizeof(unsigned long long)));
    }
    cudaIpcMemHandle_t h;
    PMP_CUDA(cudaIpcGetMemHandle(&h, c->d_xchg));
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(out64, &h, sizeof(h));
    return PMP_OK;
}

int pmp_peer_exchange_attach(pmp_ctx* c, const void* handles, int n_ranks) {
    PMP_REQUIRE(c && handles && n_ranks == c->world, "bad arguments (n_ranks %d, world_size %d)", n_ranks, c ? c->world : -1);
    PMP_REQUIRE(c->d_xchg, "pmp_peer_exchange_handle first");
    PMP_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) { c->peer_xchg[r] = c->d_xchg; continue; }
        if (c->peer_xchg[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        PMP_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_xchg[r] = (unsigned long long*)p;
    }
    c->peers_attached = true;
    return PMP_OK;
}

int pmp_trace_diagnostics(pmp_ctx* c, int max_lag, double* mean, double* var, double* acov, double* msjd, double* move_rate, int64_t* n_rows) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_REQUIRE(c->trace.state, "diagnostics need a STATE trace (pmp_trace_config with PMP_TRACE_STATE)");
    PMP_REQUIRE(max_lag >= 0 && max_lag < 65535, "max_lag out of range");
    PMP_CUDA(cudaSetDevice(c->device));
    long long rows = 0;
    PMP_CUDA(cudaMemcpyAsync(&rows, &c->d_cnt->trace_rows, sizeof(rows), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    if (rows > c->trace.capacity) rows = c->trace.capacity;
    PMP_REQUIRE(rows >= 2, "diagnostics need at least two recorded iterations (have %lld)", rows);
    if (max_lag >= rows) max_lag = (int)rows - 1;
    const int dim = c->cfg.dim;
    PMP_REQUIRE(dim <= 65535, "dim too large for the diagnostics kernels");
    double* d = nullptr;
    const size_t words = (size_t)2 * dim + (size_t)(max_lag + 1) * dim + 2;
    PMP_CUDA(cudaMalloc((void**)&d, words * sizeof(double)));
    double *d_mean = d, *d_var = d + dim, *d_acov = d + 2 * dim, *d_j = d_acov + (size_t)(max_lag + 1) * dim;
    diag_moments_kernel<<<dim, 1024, 0, c->stream>>>(c->trace.state, rows, dim, d_mean, d_var);
    diag_acov_kernel<<<dim3(max_lag + 1, dim), 1024, 0, c->stream>>>(c->trace.state, rows, dim, d_mean, d_acov);
    diag_jumps_kernel<<<1, 1024, 0, c->stream>>>(c->trace.state, (c->trace.what & PMP_TRACE_NEXT) ? c->trace.next : nullptr, rows, dim, d_j);
    c->launches += 3;
    std::vector<double> h(words);
    cudaError_t e = cudaMemcpyAsync(h.data(), d, words * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    PMP_CUDA(e);
    if (mean) memcpy(mean, h.data(), dim * sizeof(double));
    if (var) memcpy(var, h.data() + dim, dim * sizeof(double));
    if (acov) memcpy(acov, h.data() + 2 * dim, (size_t)(max_lag + 1) * dim * sizeof(double));
    if (msjd) *msjd = h[words - 2];
    if (move_rate) *move_rate = h[words - 1];
    if (n_rows) *n_rows = rows;
    return PMP_OK;
}

int pmp_sync(pmp_ctx* c) {
    PMP_REQUIRE(c, "ctx is NULL");
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_run_timed(pmp_ctx* c, int64_t iters, float* total_ms, float* sweep_ms) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    int rc;
    if (!sweep_ms) {
        PMP_CUDA(cudaEventRecord(c->ev0, c->stream));
        if ((rc = run_impl(c, iters))) return rc;
        PMP_CUDA(cudaEventRecord(c->ev1, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        // per-iteration event pairs around the sweep kernel (plain launches, no graph)
        PMP_REQUIRE(iters <= 4096, "sweep timing mode supports <= 4096 iterations per call");
        while ((int64_t)c->ev_pool.size() < 2 * iters) { cudaEvent_t ev; PMP_CUDA(cudaEventCreate(&ev)); c->ev_pool.push_back(ev); }
        PMP_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int64_t i = 0; i < iters; ++i) if ((rc = enqueue_iteration(c, c->ev_pool[2 * i], c->ev_pool[2 * i + 1], false))) return rc;
        PMP_CUDA(cudaEventRecord(c->ev1, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
        double acc = 0.0;
        if (c->cfg.target == PMP_TARGET_LINEAR_GAUSS)
            for (int64_t i = 0; i < iters; ++i) { float ms = 0.f; PMP_CUDA(cudaEventElapsedTime(&ms, c->ev_pool[2 * i], c->ev_pool[2 * i + 1])); acc += ms; }
        *sweep_ms = (float)acc;
        c->lt_valid = false;
    }
    if (total_ms) PMP_CUDA(cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
    return PMP_OK;
}

int pmp_time_sweep(pmp_ctx* c, int reps, float* ms) {
    PMP_REQUIRE(c && c->configured && ms && reps > 0, "bad arguments");
    PMP_REQUIRE(c->cfg.target == PMP_TARGET_LINEAR_GAUSS, "pmp_time_sweep: linear-Gaussian target only");
    PMP_CUDA(cudaSetDevice(c->device));
    int rc;
    for (int i = 0; i < 3; ++i) if ((rc = launch_sweep_linear(c, 0))) return rc;
    PMP_CUDA(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < reps; ++i) if ((rc = launch_sweep_linear(c, 0))) return rc;
    PMP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PMP_CUDA(cudaMemsetAsync(c->d_acc, 0, c->P * sizeof(unsigned long long), c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    PMP_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PMP_OK;
}

int pmp_launch_count(pmp_ctx* c, int64_t* launches) {
    PMP_REQUIRE(c && launches, "NULL argument");
    *launches = c->launches;
    return PMP_OK;
}

int pmp_fp32_peak(pmp_ctx* c, int packed, double* tflops) {
    PMP_REQUIRE(c && tflops, "NULL argument");
    PMP_CUDA(cudaSetDevice(c->device));
    float* d_out; PMP_CUDA(cudaMalloc((void**)&d_out, 4));
    const int iters = 16384, blocks = c->sm_count * 8;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        PMP_CUDA(cudaEventRecord(c->ev0, c->stream));
        if (packed) fp32_peak_kernel<true><<<blocks, 256, 0, c->stream>>>(d_out, iters, 1.0f);
        else fp32_peak_kernel<false><<<blocks, 256, 0, c->stream>>>(d_out, iters, 1.0f);
        c->launches++;
        PMP_CUDA(cudaEventRecord(c->ev1, c->stream));
        PMP_CUDA(cudaStreamSynchronize(c->stream));
        float ms; PMP_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        double flops = (double)blocks * 256.0 * iters * 8.0 * 2.0 * 2.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaFree(d_out);
    *tflops = best;
    return PMP_OK;
}

int pmp_debug_stamps(pmp_ctx* c, unsigned long long* out64) {
    PMP_REQUIRE(c && out64 && c->d_dbg, "debug stamps not enabled (PMP_DEBUG_STAMPS=1 before pmp_create)");
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    PMP_CUDA(cudaMemcpy(out64, c->d_dbg, (64 + 3 * 1024) * 8, cudaMemcpyDeviceToHost));
    return PMP_OK;
}

int pmp_stream_uniforms(uint64_t seed, uint64_t iteration, uint32_t stream, uint64_t idx0, int64_t count, double* out) {
    PMP_REQUIRE(out && count >= 0, "bad arguments");
    for (int64_t i = 0; i < count; ++i) out[i] = u64_to_unit(stream_u64(seed, iteration, stream, idx0 + (uint64_t)i));
    return PMP_OK;
}

int pmp_stream_normals(uint64_t seed, uint64_t iteration, uint32_t stream, uint64_t idx0, int64_t count, double* out) {
    PMP_REQUIRE(out && count >= 0, "bad arguments");
    for (int64_t i = 0; i < count; ++i) out[i] = stream_normal(seed, iteration, stream, idx0 + (uint64_t)i);
    return PMP_OK;
}

int pmp_l2_flush(pmp_ctx* c) {
    PMP_REQUIRE(c, "ctx is NULL");
    PMP_CUDA(cudaSetDevice(c->device));
    if (!c->d_flush) { c->flush_bytes = 256ull << 20; PMP_CUDA(cudaMalloc(&c->d_flush, c->flush_bytes)); }
    PMP_CUDA(cudaMemsetAsync(c->d_flush, 0x5a, c->flush_bytes, c->stream));
    return PMP_OK;
}

}  // extern "C"
