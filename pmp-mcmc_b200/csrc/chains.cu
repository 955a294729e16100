// chains.cu — batched independent chains on the analytic targets.
//
// Replaces the hop loops of simple_sampling/error/error.py (SP 17-40, MP 43-77, PSP 78-134, PMP 137-190), of
// complex_nets/correlation/com_dim.py (PMP 24-86) and the banana density of banana_data.ipynb cell 2.  Those targets have
// no data, so one chain is a few hundred flops per hop and a GPU is only useful on many chains at once: one THREAD per
// chain, state and tree nodes in chain-fastest layouts ([dim][n_chains], [P][dim][n_chains]) so that every load and store
// of a warp is one coalesced 128-byte line.  With sample recording on, the kernel is bound by the HBM write of the
// resampled points (4*dim bytes per node evaluation); everything else stays in L1/L2-backed scratch.
// Per chain and hop: Philox increments (chain id in the upper 32 bits of the element index) → tree nodes in float32 →
// log-target in binary64 → MP / binary-Barker / general-tree weights → sequential cdf (the association NumPy's cumsum
// uses) → P inverse-CDF draws and the pick of the next state.  Chain 0 uses the same stream elements as pmp_run.
#include "accept.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace pmp {

struct ChainArgs {
    pmp_config cfg;
    int P;
    long long n_chains;
    float* states;            // [dim][n_chains]
    float* nodes;             // [P][dim][n_chains] scratch
    double* work;             // [3][P][n_chains] scratch: lt, A/cdf, level scratch
    unsigned short* draws;    // [P][n_chains] scratch (P <= MAX_NODES = 8192)
    int tmp_slots;            // level scratch entries per chain: b for the general-tree rule, else 0
    float* samples;           // [iters][P][dim][n_chains] or nullptr
    unsigned long long seed, iter0;
    int iters;
};

// Scratch of a chain (log-targets, weights / cdf: binary64; level scratch of the general-tree rule: b binary64; tree nodes: float32; draws: 16-bit): P * (4 dim + 18) + 8 b
// bytes — 448 for the banana PMP (P = 16, b = 4, dim = 2).  SM = true keeps it in SHARED
// memory ([slot][thread]: conflict-free) — round 1 kept it in global memory, and with 2^20 chains it spilled out of L2: 11.8 GB of DRAM
// writes per launch against 3.2 GB of recorded samples (ncu r1c).  SM = false is the fallback for trees too large for shared memory.
__host__ __device__ inline size_t chain_scratch_bytes(int P, int dim, int tmp_slots) { return (size_t)P * (4 * (size_t)dim + 18) + 8 * (size_t)tmp_slots; }

// log K(a, b) of the Gaussian proposal kernel.  Symmetric to the last bit (the differences only change sign), which the callers use to evaluate every pair once.
// ks2 = ks * ks; a division by exactly 1.0 returns its dividend, so the default kernel_sigma = 1 skips the binary64 division.
__device__ __forceinline__ double chain_log_kernel(const float* nodes, long long nc, int dim, int a, int b, long long c, double ks2, double lnk) {
    double s = 0.0;
    for (int j = 0; j < dim; ++j) {
        double d = (double)nodes[((long long)a * dim + j) * nc + c] - (double)nodes[((long long)b * dim + j) * nc + c];
        s = fma(d, d, s);
    }
    return ks2 == 1.0 ? dim * lnk - 0.5 * s : dim * lnk - 0.5 * s / ks2;
}

// COMPACT: the warp-cooperative increments phase for long trees ((P - 1) * dim >= 64 normals per hop); short trees keep the per-element loop — with the 30 increments of the
// banana PMP the queue bookkeeping costs more than it saves (34.1 against 29.7 ms), with the 280 of N(0, I_40), binary D = 3, it is 21.9 against 34.3 ms.
template <bool SM, bool COMPACT>
__global__ void __launch_bounds__(128) chains_kernel(ChainArgs a) {
    extern __shared__ __align__(16) unsigned char chain_sm[];
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (!COMPACT && c >= a.n_chains) return;
    const bool live = COMPACT ? (c < a.n_chains) : true;   // COMPACT: threads past the last chain keep running, the increments phase is warp-cooperative
    const pmp_config& cfg = a.cfg;
    const int P = a.P, dim = cfg.dim;
    const long long nc = a.n_chains;
    // scratch addressing: element e of an array lives at [e * sn + so]
    const long long sn = SM ? (long long)blockDim.x : nc, so = SM ? (long long)threadIdx.x : c;
    double* const work = SM ? reinterpret_cast<double*>(chain_sm) : a.work;
    const size_t nd = (size_t)2 * P + a.tmp_slots;        // binary64 slots per chain
    float* const nodes = SM ? reinterpret_cast<float*>(chain_sm + nd * blockDim.x * sizeof(double)) : a.nodes;
    unsigned short* const draws = SM ? reinterpret_cast<unsigned short*>(chain_sm + nd * blockDim.x * sizeof(double) + (size_t)P * dim * blockDim.x * sizeof(float)) : a.draws;
    const int b = (cfg.tree == PMP_TREE_BINARY) ? 2 : cfg.b;
    const int D = (cfg.tree == PMP_TREE_FLAT) ? 1 : cfg.depth;
    const double ks = (double)cfg.kernel_sigma, ks2 = ks * ks;
    const double lnk = -HALF_LOG_2PI - log(ks);
    const bool use_kernel = !(cfg.flags & PMP_FLAG_NO_KERNEL_TERM);
    const int uniform = (cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) ? 1 : 0;
    const unsigned long long cbase = (unsigned long long)c << 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // tail quantiles of a warp's chains, evaluated 32 at a time (see propose_level_kernel): element, owner lane, uniform
    __shared__ int q_e[COMPACT ? 4 : 1][COMPACT ? 96 : 1], q_o[COMPACT ? 4 : 1][COMPACT ? 96 : 1];
    __shared__ double q_u[COMPACT ? 4 : 1][COMPACT ? 96 : 1];
    double* lt = work;
    double* A = work + (long long)P * sn;
    double* tmp = work + 2ll * P * sn;
#define NODE(p, j) nodes[((long long)(p) * dim + (j)) * sn + so]
#define W(arr, p) arr[(long long)(p) * sn + so]

    for (int it = 0; it < a.iters; ++it) {
        const unsigned long long iter = a.iter0 + (unsigned long long)it;
        // ---- tree nodes (parents precede children in index order for all three shapes).  Phase 1 leaves the increment fl(alpha z) of every element
        // e = p * dim + j in its node slot: one Philox block per element PAIR, the quantile's central branch inline, its tail branch (15 % of the draws) queued
        // per warp and evaluated with all lanes busy — the per-element loop made every warp run both branches for every normal (ncu r2m: the normals were
        // half of the kernel's instructions).  Phase 2 adds the parents in index order: child = fl(parent + fl(alpha z)), the same bits.
        if (live) for (int j = 0; j < dim; ++j) NODE(0, j) = a.states[(long long)j * nc + c];
        if (!COMPACT) {
            // short trees: one Philox block per element PAIR; the quantile's central branch inline, its tail branch (15 % of the draws) deferred: with 30 increments
            // nearly every warp holds a tail lane for every element, so evaluating in place makes the warp run both branches 30 times.  A lane parks its tail draws
            // (element in the draws slots, uniform in the binary64 slots — both unused until the weights) and the warp then runs the tail code max-over-lanes times.
            const int e_hi = P * dim;
            int n_tail = 0;
            const int cap = (e_hi - dim >= 16) ? P : 0;      // parked entries: draws has P slots, the binary64 scratch 2 P.  Measured: 30 increments (banana PMP) 22.4 -> 21.6 ms, 3 increments (1-D MP) 23.7 -> 24.9: in place below 16
            for (int k = dim >> 1; 2 * k < e_hi; ++k) {
                const unsigned long long blk = (cbase | (unsigned long long)(2 * k)) >> 1;
                uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (STREAM_PROPOSAL << 24)};
                philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                const uint64_t w[2] = {(uint64_t)ctr[1] << 32 | ctr[0], (uint64_t)ctr[3] << 32 | ctr[2]};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = 2 * k + h;
                    if (e < dim || e >= e_hi) continue;
                    if (uniform) { nodes[(long long)e * sn + so] = __fmul_rn(cfg.alpha, (float)PMP_FMA(2.0, u64_to_unit(w[h]), -1.0)); continue; }
                    const double u = u64_to_open(w[h]);
                    const double q = PMP_ADD(u, -0.5);
                    if (fabs(q) <= 0.425) nodes[(long long)e * sn + so] = __fmul_rn(cfg.alpha, (float)ppf_central(q));
                    else if (n_tail < cap) { W(draws, n_tail) = (unsigned short)e; W(work, n_tail) = u; ++n_tail; }
                    else nodes[(long long)e * sn + so] = __fmul_rn(cfg.alpha, (float)det_norm_ppf(u));
                }
            }
            for (int i = 0; i < n_tail; ++i) {
                const int e = W(draws, i);
                nodes[(long long)e * sn + so] = __fmul_rn(cfg.alpha, (float)det_norm_ppf(W(work, i)));
            }
        } else {
            int qn = 0;
            auto drain = [&](int count) {
                const int x = qn - count + lane;
                if (lane < count) {
                    const int e = q_e[warp][x], ol = q_o[warp][x];
                    const long long slot = SM ? (long long)(warp * 32 + ol) : (c - lane + ol);
                    nodes[(long long)e * sn + slot] = __fmul_rn(cfg.alpha, (float)det_norm_ppf(q_u[warp][x]));
                }
                qn -= count;
                __syncwarp();
            };
            const int e_hi = P * dim;
            for (int k = dim >> 1; 2 * k < e_hi; ++k) {
                double u[2]; bool tail[2] = {false, false};
                if (live) {
                    const unsigned long long blk = (cbase | (unsigned long long)(2 * k)) >> 1;
                    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (STREAM_PROPOSAL << 24)};
                    philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                    u[0] = u64_to_open((uint64_t)ctr[1] << 32 | ctr[0]);
                    u[1] = u64_to_open((uint64_t)ctr[3] << 32 | ctr[2]);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int e = 2 * k + h;
                        if (e < dim || e >= e_hi) continue;
                        const double q = PMP_ADD(u[h], -0.5);
                        if (fabs(q) <= 0.425) nodes[(long long)e * sn + so] = __fmul_rn(cfg.alpha, (float)ppf_central(q));
                        else tail[h] = true;
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned mask = __ballot_sync(0xffffffffu, tail[h]);
                    if (tail[h]) { const int x = qn + __popc(mask & ((1u << lane) - 1u)); q_e[warp][x] = 2 * k + h; q_o[warp][x] = lane; q_u[warp][x] = u[h]; }
                    qn += __popc(mask);
                }
                __syncwarp();
                while (qn >= 32) drain(32);
            }
            if (qn > 0) drain(qn);
            __syncwarp();
        }
        if (live) {
        if (cfg.tree == PMP_TREE_FLAT) {
            for (int p = 1; p < P; ++p) for (int j = 0; j < dim; ++j) NODE(p, j) = __fadd_rn(NODE(0, j), NODE(p, j));
        } else {
            for (int s = 1; s < P; s *= b) {                 // level: nodes [s, s b), node p hangs under p mod s
                int parent = 0;
                for (int p = s; p < s * b && p < P; ++p) {
                    for (int j = 0; j < dim; ++j) NODE(p, j) = __fadd_rn(NODE(parent, j), NODE(p, j));
                    if (++parent == s) parent = 0;
                }
            }
        }
        // ---- log-targets
        for (int p = 0; p < P; ++p) W(lt, p) = analytic_logtarget(cfg.target, &NODE(p, 0), (int)sn, dim, cfg.target_p0, cfg.target_p1) / (double)cfg.scale;
        // ---- log-weights
        int next;
        int n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
        if (cfg.algo == PMP_ALGO_MH || cfg.algo == PMP_ALGO_BARKER) {
            double u = u64_to_unit(stream_u64(a.seed, iter, STREAM_DRAW, cbase));
            double l0 = W(lt, 0), l1 = W(lt, 1);
            if (cfg.algo == PMP_ALGO_MH) next = u < exp((double)cfg.mh_temperature * (l1 - l0));
            else { double m = fmax(l0, l1); double w0 = exp(l0 - m), w1 = exp(l1 - m); next = (w1 / (w0 + w1)) > u; }
            W(draws, 0) = (unsigned short)next; n_draws = 1;
            W(A, 0) = l0; W(A, 1) = l1;
        } else {
            if (cfg.algo == PMP_ALGO_MP) {
                // A[j] = lt[j] + sum_{k != j, ascending} log K(j, k): every pair is evaluated once and added to both ends — node j still receives its terms in ascending k
                for (int j = 0; j < P; ++j) W(A, j) = W(lt, j);
                if (use_kernel) for (int j = 0; j < P; ++j) for (int k = j + 1; k < P; ++k) {
                    const double kv = chain_log_kernel(nodes, sn, dim, j, k, so, ks2, lnk);
                    W(A, j) += kv; W(A, k) += kv;
                }
            } else if (cfg.algo == PMP_ALGO_PSP) {
                for (int p = 0; p < P; ++p) {
                    double s = 0.0;
                    for (int l = 0; l < D; ++l) { int m = p & ((2 << l) - 1), q = m ^ (1 << l); s += logsigmoid(W(lt, m) - W(lt, q)); }
                    W(A, p) = s;
                }
            } else if (cfg.algo == PMP_ALGO_PMP) {
                for (int p = 0; p < P; ++p) W(A, p) = 0.0;
                int s = 1;
                for (int i = 0; i < D; ++i) {
                    for (int h = 0; h < s; ++h) {
                        for (int j = 0; j < b; ++j) W(tmp, j) = W(lt, h + j * s);
                        if (use_kernel) for (int j = 0; j < b; ++j) for (int k = j + 1; k < b; ++k) {      // every pair once, same order of additions per node (see MP)
                            const double kv = chain_log_kernel(nodes, sn, dim, h + j * s, h + k * s, so, ks2, lnk);
                            W(tmp, j) += kv; W(tmp, k) += kv;
                        }
                        double mx = -INFINITY;
                        for (int j = 0; j < b; ++j) mx = fmax(mx, W(tmp, j));
                        double se = 0.0;
                        for (int j = 0; j < b; ++j) se += exp(W(tmp, j) - mx);
                        double lse = mx + log(se);
                        for (int j = 0; j < b; ++j) { const int nj = h + j * s; W(A, nj) += (mx == -INFINITY) ? -INFINITY : W(tmp, j) - lse; }
                    }
                    if (i < D - 1) {
                        const int lo = s * b, hi = s * b * b;
                        const int mod = (cfg.flags & PMP_FLAG_QUIRK_LEVEL_MOD) ? b * (i + 1) : lo;
                        int r = lo % mod;
                        for (int x = lo; x < hi; ++x) { W(A, x) = W(A, r); if (++r == mod) r = 0; }
                    }
                    s *= b;
                }
            } else {   // TABLE
                for (int p = 0; p < P; ++p) {
                    double v = W(lt, p);
                    if (cfg.flags & PMP_FLAG_QUIRK_TABLE_CONST) v += (double)D * (b - 1) * dim * (-HALF_LOG_2PI);
                    else if (use_kernel) {
                        int s = 1;
                        for (int d = 0; d < D; ++d) {
                            const int m = p % (s * b), h = m % s;
                            for (int k = 0; k < b; ++k) { const int o = h + k * s; if (o != m) v += chain_log_kernel(nodes, sn, dim, m, o, so, ks2, lnk); }
                            s *= b;
                        }
                    }
                    W(A, p) = v;
                }
            }
            if (cfg.flags & PMP_FLAG_STANDARDIZE) {
                double mean = 0.0; for (int p = 0; p < P; ++p) mean += W(A, p); mean /= P;
                double var = 0.0; for (int p = 0; p < P; ++p) { double d = W(A, p) - mean; var = fma(d, d, var); }
                double sd = sqrt(var / (P - 1));
                for (int p = 0; p < P; ++p) W(A, p) = (W(A, p) - mean) / sd;
            }
            // ---- sequential cdf and draws
            double mx = -INFINITY;
            for (int p = 0; p < P; ++p) mx = fmax(mx, W(A, p));
            double run = 0.0;
            for (int p = 0; p < P; ++p) { double w = exp(W(A, p) - mx); run += (w == w) ? w : 0.0; W(A, p) = run; }
            const double total = run;
            const bool right = (cfg.draw != PMP_DRAW_CUDA);
            unsigned long long dw1 = 0ull;                 // draws 2m and 2m+1 are the two words of one Philox block
            for (int t = 0; t < n_draws; ++t) {
                unsigned long long word;
                if (!COMPACT) word = stream_u64(a.seed, iter, STREAM_DRAW, cbase | (unsigned long long)t);
                else if (t & 1) word = dw1;
                else {
                    const unsigned long long blk = (cbase | (unsigned long long)t) >> 1;
                    uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (STREAM_DRAW << 24)};
                    philox4x32_10(ctr, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
                    word = (unsigned long long)ctr[1] << 32 | ctr[0]; dw1 = (unsigned long long)ctr[3] << 32 | ctr[2];
                }
                double thr = u64_to_unit(word) * total;
                int lo = 0, hi = P;
                while (lo < hi) { int mid = (lo + hi) >> 1; double v = W(A, mid); bool go = right ? (v <= thr) : (v < thr); if (go) lo = mid + 1; else hi = mid; }
                W(draws, t) = (unsigned short)min(lo, P - 1);
            }
            if (cfg.draw == PMP_DRAW_PYTHON) {
                double up = u64_to_unit(stream_u64(a.seed, iter, STREAM_PICK, cbase));
                next = W(draws, min(P - 1, (int)(up * (double)P)));
            } else next = W(draws, 0);
        }
        // ---- record the resampled points (chain-fastest: coalesced 4-byte stores across the warp), new state
        if (a.samples) {
            float* out = a.samples + (long long)it * P * dim * nc;
            for (int t = 0; t < P; ++t) {
                int src = t < n_draws ? W(draws, t) : next;
                for (int j = 0; j < dim; ++j) out[((long long)t * dim + j) * nc + c] = NODE(src, j);
            }
        }
        for (int j = 0; j < dim; ++j) a.states[(long long)j * nc + c] = NODE(next, j);
        }   // live
        __syncwarp();
    }
#undef NODE
#undef W
}

__global__ void chains_transpose_in(const float* host_layout, float* states, long long n_chains, int dim) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_chains * dim) { long long c = i / dim; int j = (int)(i - c * dim); states[(long long)j * n_chains + c] = host_layout[i]; }
}
__global__ void chains_transpose_out(const float* states, float* host_layout, long long n_chains, int dim) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_chains * dim) { long long c = i / dim; int j = (int)(i - c * dim); host_layout[i] = states[(long long)j * n_chains + c]; }
}

struct ChainScratch { float* nodes = nullptr; double* work = nullptr; int* draws = nullptr; float* stage = nullptr; long long n_chains = 0; int P = 0, dim = 0; };

}  // namespace pmp

using namespace pmp;

static ChainScratch* scratch_of(pmp_ctx* c) { return reinterpret_cast<ChainScratch*>(c->chain_scratch); }

extern "C" {

int pmp_chains_destroy(pmp_ctx* c) {
    ChainScratch* s = scratch_of(c);
    if (s) { cudaFree(s->nodes); cudaFree(s->work); cudaFree(s->draws); cudaFree(s->stage); delete s; c->chain_scratch = nullptr; }
    if (c->d_chain_states) { cudaFree(c->d_chain_states); c->d_chain_states = nullptr; }
    if (c->d_chain_samples) { cudaFree(c->d_chain_samples); c->d_chain_samples = nullptr; c->chain_samples_cap = 0; }
    c->n_chains = 0;
    return PMP_OK;
}

int pmp_chains_create(pmp_ctx* c, int64_t n_chains, const float* init_states) {
    PMP_REQUIRE(c && c->configured, "ctx not configured");
    PMP_REQUIRE(n_chains >= 1, "n_chains must be >= 1");
    PMP_REQUIRE(c->cfg.target == PMP_TARGET_NORMAL1D || c->cfg.target == PMP_TARGET_BANANA || c->cfg.target == PMP_TARGET_STDNORMAL,
                "batched chains run the analytic targets only");
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    pmp_chains_destroy(c);
    const int P = c->P, dim = c->cfg.dim;
    ChainScratch* s = new ChainScratch();
    c->chain_scratch = s;
    s->n_chains = n_chains; s->P = P; s->dim = dim;
    PMP_CUDA(cudaMalloc((void**)&c->d_chain_states, (size_t)n_chains * dim * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->nodes, (size_t)n_chains * P * dim * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->work, (size_t)n_chains * P * 3 * sizeof(double)));
    PMP_CUDA(cudaMalloc((void**)&s->draws, (size_t)n_chains * P * sizeof(int)));
    PMP_CUDA(cudaMalloc((void**)&s->stage, (size_t)n_chains * dim * sizeof(float)));
    if (init_states) {
        PMP_CUDA(cudaMemcpyAsync(s->stage, init_states, (size_t)n_chains * dim * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        long long tot = n_chains * dim;
        chains_transpose_in<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(s->stage, c->d_chain_states, n_chains, dim);
        c->launches++;
    } else PMP_CUDA(cudaMemsetAsync(c->d_chain_states, 0, (size_t)n_chains * dim * sizeof(float), c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    c->n_chains = n_chains;
    c->chain_iteration = 0;
    c->chain_iters_recorded = 0;
    return PMP_OK;
}

static int chains_launch(pmp_ctx* c, int64_t iters, int record) {
    PMP_REQUIRE(c && c->n_chains > 0 && scratch_of(c), "pmp_chains_create first");
    PMP_REQUIRE(iters >= 0 && iters < (1ll << 31), "bad iters");
    ChainScratch* s = scratch_of(c);
    PMP_REQUIRE(s->P == c->P && s->dim == c->cfg.dim, "configuration changed since pmp_chains_create");
    PMP_CUDA(cudaSetDevice(c->device));
    if (record) {
        long long need = (long long)iters * c->P * c->cfg.dim * c->n_chains;
        if (need > c->chain_samples_cap) {
            if (c->d_chain_samples) cudaFree(c->d_chain_samples);
            c->d_chain_samples = nullptr; c->chain_samples_cap = 0;
            cudaError_t e = cudaMalloc((void**)&c->d_chain_samples, (size_t)need * sizeof(float));
            if (e != cudaSuccess) { set_error("cudaMalloc(%lld sample floats) failed: %s", need, cudaGetErrorString(e)); return PMP_ERR_ALLOC; }
            c->chain_samples_cap = need;
        }
    }
    const int tmp_slots = c->cfg.algo == PMP_ALGO_PMP ? (c->cfg.tree == PMP_TREE_BINARY ? 2 : c->cfg.b) : 0;
    ChainArgs a{c->cfg, c->P, c->n_chains, c->d_chain_states, s->nodes, s->work, reinterpret_cast<unsigned short*>(s->draws), tmp_slots, record ? c->d_chain_samples : nullptr,
                c->seed, c->chain_iteration, (int)iters};
    // scratch in shared memory when a block of >= 32 chains fits 100 KB (the largest such block), else in global memory
    const size_t per_chain = chain_scratch_bytes(c->P, c->cfg.dim, tmp_slots);
    const bool compact = !(c->cfg.flags & PMP_FLAG_UNIFORM_PROPOSAL) && (long long)(c->P - 1) * c->cfg.dim >= 64;
    static bool attr_set = false;
    if (!attr_set) {
        PMP_CUDA(cudaFuncSetAttribute(chains_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        PMP_CUDA(cudaFuncSetAttribute(chains_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    int threads = 0;
    for (int t = 128; t >= 32 && !threads; t >>= 1) if (per_chain * t <= 100 * 1024) threads = t;      // measured (banana PMP): 128-thread blocks 25.9 ms, 64 28.6, 32 27.9 — more resident warps from smaller blocks do not pay
    if (getenv("PMP_CHAINS_THREADS")) { const int t = atoi(getenv("PMP_CHAINS_THREADS")); if ((t == 32 || t == 64 || t == 128) && per_chain * t <= 100 * 1024) threads = t; }
    const bool in_smem = threads > 0 && !getenv("PMP_CHAINS_GLOBAL_SCRATCH");
    if (in_smem) {
        const size_t smem = per_chain * threads;
        const unsigned grid = (unsigned)((c->n_chains + threads - 1) / threads);
        if (compact) chains_kernel<true, true><<<grid, threads, smem, c->stream>>>(a);
        else chains_kernel<true, false><<<grid, threads, smem, c->stream>>>(a);
    } else if (compact) chains_kernel<false, true><<<(unsigned)((c->n_chains + 127) / 128), 128, 0, c->stream>>>(a);
    else chains_kernel<false, false><<<(unsigned)((c->n_chains + 127) / 128), 128, 0, c->stream>>>(a);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    c->chain_iteration += (unsigned long long)iters;
    c->chain_iters_recorded = record ? iters : 0;
    return PMP_OK;
}

int pmp_chains_run(pmp_ctx* c, int64_t iters, int record_samples) {
    int rc = chains_launch(c, iters, record_samples);
    if (rc) return rc;
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_chains_run_timed(pmp_ctx* c, int64_t iters, int record_samples, float* total_ms) {
    PMP_REQUIRE(c && total_ms, "NULL argument");
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    PMP_CUDA(cudaEventRecord(c->ev0, c->stream));
    int rc = chains_launch(c, iters, record_samples);
    if (rc) return rc;
    PMP_CUDA(cudaEventRecord(c->ev1, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    PMP_CUDA(cudaEventElapsedTime(total_ms, c->ev0, c->ev1));
    return PMP_OK;
}

int pmp_chains_read_states(pmp_ctx* c, float* out) {
    PMP_REQUIRE(c && out && c->n_chains > 0 && scratch_of(c), "pmp_chains_create first");
    ChainScratch* s = scratch_of(c);
    long long tot = c->n_chains * c->cfg.dim;
    chains_transpose_out<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(c->d_chain_states, s->stage, c->n_chains, c->cfg.dim);
    c->launches++;
    PMP_CUDA(cudaMemcpyAsync(out, s->stage, (size_t)tot * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

int pmp_chains_read_samples(pmp_ctx* c, float* out, int64_t count) {
    PMP_REQUIRE(c && out && c->d_chain_samples, "no samples recorded");
    long long have = c->chain_iters_recorded * c->P * c->cfg.dim * c->n_chains;
    PMP_REQUIRE(count == have, "count %lld != recorded %lld", (long long)count, have);
    PMP_CUDA(cudaMemcpyAsync(out, c->d_chain_samples, (size_t)have * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    return PMP_OK;
}

}  // extern "C"
