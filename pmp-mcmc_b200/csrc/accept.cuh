// accept.cuh — proposal generation, prefetch-tree expansion and multi-proposal acceptance.
//
// propose_kernel   replaces the host generators: flat (500_MP.cu:181-185, lb.py:173-176), doubling tree
//                  (500_PMP.cu:170-179, lb.py:268-272, error.py:88-91, com_dim.py:34-37, PMP_FC.py:176-182) and
//                  (N+1)-ary tree (conv_pmp.cu:182-197, lb.py:356-360, error.py:145-149).
// accept_kernel    replaces, in one CTA and without leaving the device: finalising the sweep (-n/2 log(2 pi s^2)
//                  - S/2)/scale (500_MP.cu:19), the proposal-kernel terms (500_MP.cu:22-31, 500_PMP.cu:23-30,
//                  conv_pmp.cu:22-33, lb.py:111-116), the MP / binary-Barker / general-tree weights (lb.py:144-150,
//                  216-240, 315-330), the host exp + categorical draw (500_MP.cu:207-222; pandas sample lb.py:154-156),
//                  the pick of the next state (lb.py:162-163; 500_MP.cu:240-243) and the trace rows.
// Everything after the log-targets is binary64, in the log domain with an exact max-shift (replaces adjust_A).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace pmp {

constexpr int ACCEPT_THREADS = 1024;    // one CTA; one node per thread at P = 1024
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PMP_STAMP(buf, slot) do { if ((buf) && threadIdx.x == 0) { (buf)[(slot)] = clock64(); (buf)[(slot) + 16] = pmp::globaltimer_ns(); } } while (0)

// ---------------------------------------------------------------------------------------------------------------
// tree index helpers.  Node ids are the reference's: BINARY node k+2^l is the child of k at level l;
// BARY node k + b^l (j+1) is child j of k at level l (k < b^l).
__host__ __device__ __forceinline__ long long ipow(int b, int e) { long long r = 1; for (int i = 0; i < e; ++i) r *= b; return r; }

struct ProposeArgs {
    const float* state; float* props; const DeviceCounters* cnt;
    unsigned long long seed; int P, dim, tree, b, depth; float alpha; int uniform;
};

// One thread per (node, coordinate).  The value is built along the node's ancestor chain in float32 with the
// reference's two roundings per step: child = fl(parent + fl(alpha * z))  (normal_distribution<float>(0, alpha),
// torch.normal(0, alpha)).  z for the step that created node `a` is normal number a*dim + j of the iteration.
// Ancestors of `node` in creation order (root side first): f(anc) for every level whose digit is non-zero, anc = node mod b^(l+1).
// 32-bit arithmetic, bit operations for the binary tree: this runs on the critical path of the chain (next-iteration nodes).
template <class F>
__device__ __forceinline__ void for_each_ancestor(int tree, int b, int depth, int node, F f) {
    if (tree == PMP_TREE_BINARY) {
        for (int l = 0; l < depth; ++l) if ((node >> l) & 1) f(node & ((2 << l) - 1));
    } else {
        int s = 1;
        for (int l = 0; l < depth; ++l) { const int sb = s * b, m = node % sb; if (m >= s) f(m); s = sb; }
    }
}

static __device__ __noinline__ float proposal_value(const ProposeArgs& a, unsigned long long iter, int node, int j, float v) {
    if (a.tree == PMP_TREE_FLAT) {
        if (node > 0) {
            float z = (float)stream_step(a.seed, iter, (unsigned long long)node * a.dim + j, a.uniform);
            v = __fadd_rn(v, __fmul_rn(a.alpha, z));
        }
        return v;
    }
    for_each_ancestor(a.tree, a.b, a.depth, node, [&](int anc) {
        float z = (float)stream_step(a.seed, iter, (unsigned long long)anc * a.dim + j, a.uniform);
        v = __fadd_rn(v, __fmul_rn(a.alpha, z));
    });
    return v;
}

static __global__ void __launch_bounds__(256) propose_kernel(ProposeArgs a) {
    const unsigned long long iter = a.cnt->iteration;
    long long total = (long long)a.P * a.dim;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        int node = (int)(g / a.dim), j = (int)(g - (long long)node * a.dim);
        a.props[g] = proposal_value(a, iter, node, j, a.state[j]);
    }
}

// Tree proposals for long parameter vectors (the FC model: 567 434 coordinates), one tree level per launch: the nodes created at level l
// (m in [s, s*b), s = b^l) are their parents (m mod s) plus one increment — ONE quantile evaluation per element instead of one per
// ancestor, and the same float32 operation sequence as proposal_value (child = fl(parent + fl(alpha z))), hence the same bits.
// grid.y = new node, grid.x covers the coordinates.  Level -1 (s = 0): node 0 = the current state.
// Two costs of the straightforward per-element loop are removed without changing a bit of the result (measured: 1023 x 567 434 increments 7.1 ms -> see DESIGN 4.4):
//   * one Philox4x32-10 block yields the 64-bit words of TWO consecutive elements: a thread takes element pairs and evaluates the block once;
//   * the quantile's tail branch (15 % of the draws: a hand-rolled log, a square root, another rational function) made every warp run both branches for a handful of
//     lanes.  Tail elements are pushed to a per-warp queue in shared memory instead and evaluated 32 at a time with all lanes busy.
static __device__ __forceinline__ double ppf_central(double q) {         // det_norm_ppf's |q| <= 0.425 branch, operation for operation
    double r = PMP_FMA(-q, q, 0.180625);
    double num = PMP_H8(r, 2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4, 4.5921953931549871457e+4, 1.3731693765509461125e+4,
                        1.9715909503065514427e+3, 1.3314166789178437745e+2, 3.3871328727963666080e0);
    double den = PMP_H8(r, 5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4, 2.1213794301586595867e+4, 5.3941960214247511077e+3,
                        6.8718700749205790830e+2, 4.2313330701600911252e+1, 1.0);
    return PMP_DIV(PMP_MUL(q, num), den);
}
constexpr int PLK_THREADS = 256, PLK_QCAP = 96;                          // queue capacity per warp: a push adds at most 64 entries to fewer than 32 leftovers
static __global__ void __launch_bounds__(PLK_THREADS) propose_level_kernel(ProposeArgs a, int s) {
    const unsigned long long iter = a.cnt->iteration;
    const long long dim = a.dim;
    if (s == 0) {
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x) a.props[j] = a.state[j];
        return;
    }
    const int m = s + blockIdx.y, parent = m % s;
    const float* src = a.props + (long long)parent * dim;
    float* dst = a.props + (long long)m * dim;
    const unsigned long long base = (unsigned long long)m * dim;
    if (a.uniform) {
        for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x)
            dst[j] = __fadd_rn(src[j], __fmul_rn(a.alpha, (float)stream_step(a.seed, iter, base + j, 1)));
        return;
    }
    __shared__ long long q_j[PLK_THREADS / 32][PLK_QCAP];
    __shared__ double q_u[PLK_THREADS / 32][PLK_QCAP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int qn = 0;                                                          // entries in this warp's queue (warp-uniform)
    auto drain = [&](int count) {                                        // evaluate queue entries [qn - count, qn): one per lane
        const int e = qn - count + lane;
        if (lane < count) {
            const long long j = q_j[warp][e];
            dst[j] = __fadd_rn(src[j], __fmul_rn(a.alpha, (float)det_norm_ppf(q_u[warp][e])));      // the tail branch for every active lane
        }
        qn -= count;
        __syncwarp();
    };
    // element pairs (2k, 2k+1) of the absolute stream index share a Philox block; `off` aligns the pairs when base is odd
    const long long off = (long long)(base & 1ull);
    const long long npairs = (dim + off + 1) / 2;
    for (long long k0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); k0 < npairs; k0 += (long long)gridDim.x * blockDim.x) {   // k0: the warp's first pair (warp-uniform trip count)
        const long long k = k0 + lane;
        double u[2]; long long jj[2]; bool tail[2] = {false, false};
        if (k < npairs) {
            const unsigned long long idx0 = base - (unsigned long long)off + 2ull * (unsigned long long)k;     // even absolute index
            const unsigned long long blk = idx0 >> 1;
            uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (STREAM_PROPOSAL << 24)};
            philox4x32_10(c, (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
            u[0] = u64_to_open((uint64_t)c[1] << 32 | c[0]);
            u[1] = u64_to_open((uint64_t)c[3] << 32 | c[2]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                jj[h] = 2 * k + h - off;
                if (jj[h] < 0 || jj[h] >= dim) { jj[h] = -1; continue; }
                const double q = PMP_ADD(u[h], -0.5);
                if (fabs(q) <= 0.425) dst[jj[h]] = __fadd_rn(src[jj[h]], __fmul_rn(a.alpha, (float)ppf_central(q)));
                else tail[h] = true;
            }
        } else { jj[0] = jj[1] = -1; }
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                    // warp-aggregated push of the tail elements
            const unsigned mask = __ballot_sync(0xffffffffu, tail[h]);
            if (tail[h]) { const int e = qn + __popc(mask & ((1u << lane) - 1u)); q_j[warp][e] = jj[h]; q_u[warp][e] = u[h]; }
            qn += __popc(mask);
        }
        __syncwarp();
        while (qn >= 32) drain(32);
    }
    if (qn > 0) drain(qn);
}

// ---------------------------------------------------------------------------------------------------------------
// analytic log-targets (error.py:11-14, banana_data.ipynb cell 2, com_dim.py:13-15), binary64 from float32 coordinates
__device__ __forceinline__ double analytic_logtarget(int target, const float* th, int stride, int dim, float p0, float p1) {
    if (target == PMP_TARGET_NORMAL1D) {
        double z = ((double)th[0] - (double)p0) / (double)p1;
        return -0.5 * z * z - log((double)p1) - HALF_LOG_2PI;
    }
    if (target == PMP_TARGET_BANANA) {
        double x1 = th[0], x2 = th[stride];
        double t = x2 - 2.0 * (x1 * x1 - 5.0);
        return -0.5 * x1 * x1 - 0.5 * t * t;
    }
    double s = 0.0;
    for (int j = 0; j < dim; ++j) { double v = th[(long long)j * stride]; s = fma(v, v, s); }
    return -0.5 * s - dim * HALF_LOG_2PI;
}

__device__ __forceinline__ double logsigmoid(double x) { return x >= 0.0 ? -log1p(exp(-x)) : x - log1p(exp(x)); }

struct AcceptArgs {
    pmp_config cfg;
    int P;
    long long n_global;
    float* state;                    // [dim] in/out
    const float* props;              // [P, dim]
    unsigned long long* acc;         // [P] fixed-point sums (LINEAR_GAUSS); zeroed on exit
    double* lt;                      // [P] log-targets (in when !from_acc, always out)
    double* logw;                    // [P] out
    int32_t* draws;                  // [P] out
    const double* uniforms;          // injected uniforms or nullptr
    DeviceCounters* cnt;
    unsigned long long seed;
    double sat_limit;
    int from_acc;                    // 1: finalise lt from acc; 0: lt already holds the log-targets
    int only_finalize;               // 1: stop after writing lt (pmp_loglik)
    int advance;                     // 1: update state, iteration and trace
    double inv_scale;                // 1 / cfg.scale (binary64 reciprocal of the float32 scale)
    unsigned long long* dbg;         // optional phase stamps (PMP_DEBUG_STAMPS=1): [32..63] accept kernel
    TraceBuffers trace;
};

// block-wide helpers (blockDim.x == ACCEPT_THREADS); `red` has 32 doubles
__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    if (threadIdx.x < 32) { for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o); if (threadIdx.x == 0) red[0] = t; }
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ double block_max(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : -INFINITY;
    if (threadIdx.x < 32) { for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o)); if (threadIdx.x == 0) red[0] = t; }
    __syncthreads();
    return red[0];
}

// log K(a,b) = sum_dim log N(theta_a - theta_b; 0, ks^2): unit variance in the reference (lb.py:115, 500_MP.cu:26-28)
__device__ __forceinline__ double log_kernel_pair(const float* props, int dim, int a, int b, double ks) {
    double s = 0.0;
    for (int j = 0; j < dim; ++j) { double d = (double)props[(long long)a * dim + j] - (double)props[(long long)b * dim + j]; s = fma(d, d, s); }
    return dim * (-HALF_LOG_2PI - log(ks)) - 0.5 * s / (ks * ks);
}

// Inclusive scan of w[0..P) in place, fixed association (mirrored by oracle/pmp_oracle.c: oracle_blocked_cdf):
// thread t owns items [t*ipt, (t+1)*ipt) and scans them serially; thread totals are scanned by a Kogge-Stone
// warp scan, warp totals by warp 0, offsets are added back as (warp_offset + lane_exclusive) + local.
__device__ __forceinline__ void block_inclusive_scan(double* w, int P, double* red) {
    const int ipt = (P + ACCEPT_THREADS - 1) / ACCEPT_THREADS;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int i0 = t * ipt;
    double run = 0.0;
    for (int i = 0; i < ipt; ++i) if (i0 + i < P) { run += w[i0 + i]; w[i0 + i] = run; }
    double incl = run;
    for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = n + incl; }
    __syncthreads();
    if (lane == 31) red[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double v = red[lane];
        for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = n + v; }
        red[lane] = v;   // inclusive warp totals
    }
    __syncthreads();
    const double warp_off = warp > 0 ? red[warp - 1] : 0.0;
    double lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) lane_excl = 0.0;
    const double excl = warp_off + lane_excl;
    for (int i = 0; i < ipt; ++i) if (i0 + i < P) w[i0 + i] = excl + w[i0 + i];
    __syncthreads();
}

// Runs in one CTA of ACCEPT_THREADS threads; `sm` = 2*P doubles of shared memory.  ALGO is a template parameter so that
// each instantiation carries only its own weight rule: the kernel runs once per iteration on one SM, so its code
// footprint (instruction-cache misses) and dependent-latency chain are what it costs.
template <int ALGO>
__device__ __forceinline__ void accept_device(const AcceptArgs& a, double* sm) {
    double* lt = sm;               // [P]
    double* A = sm + a.P;          // [P] log-weights → weights → cdf
    __shared__ double red[32];
    __shared__ double sred4[128];
    __shared__ double s1[KDIM_MAX];
    __shared__ int s_next;

    const int P = a.P, dim = a.cfg.dim, tid = threadIdx.x;
    unsigned long long* dbg = a.dbg ? a.dbg + 32 : nullptr;
    PMP_STAMP(dbg, 0);
    const pmp_config& cfg = a.cfg;
    const unsigned long long iter = a.cnt->iteration;
    const long long row = a.cnt->trace_rows;
    const double ks = (double)cfg.kernel_sigma;
    const double log_norm_k = (cfg.kernel_sigma == 1.0f) ? -HALF_LOG_2PI : -HALF_LOG_2PI - log(ks);   // per-coordinate log normaliser of K
    const double half_inv_ks2 = 0.5 / (ks * ks);

    // ---- 1. log-targets ------------------------------------------------------------------------------------
    for (int p = tid; p < P; p += ACCEPT_THREADS) {
        double v;
        if (a.from_acc) {
            unsigned long long q = __ldcg(a.acc + p);
            a.acc[p] = 0ull;
            double sg = (double)__ldcg(a.props + (long long)p * 3 + 2);
            double S = (double)(long long)q * (1.0 / (double)(1 << FX_SHIFT));
            v = (-0.5 * (double)a.n_global * log(6.283185307179586477 * sg * sg) - 0.5 * S) * a.inv_scale;
            if ((double)(long long)q >= a.sat_limit || !(v == v)) v = -INFINITY;
        } else if (cfg.target == PMP_TARGET_NORMAL1D || cfg.target == PMP_TARGET_BANANA || cfg.target == PMP_TARGET_STDNORMAL) {
            v = analytic_logtarget(cfg.target, a.props + (long long)p * dim, 1, dim, cfg.target_p0, cfg.target_p1) / (double)cfg.scale;
        } else {
            v = a.lt[p];
        }
        lt[p] = v;
        a.lt[p] = v;
    }
    __syncthreads();
    PMP_STAMP(dbg, 1);
    if (a.only_finalize) return;

    const bool use_kernel = !(cfg.flags & PMP_FLAG_NO_KERNEL_TERM);
    const int b = (cfg.tree == PMP_TREE_BINARY) ? 2 : cfg.b;
    const int D = (cfg.tree == PMP_TREE_FLAT) ? 1 : cfg.depth;

    // ---- 2. log-weights ------------------------------------------------------------------------------------
    if constexpr (ALGO == PMP_ALGO_MH) {
        for (int p = tid; p < P; p += ACCEPT_THREADS) A[p] = lt[p];
    } else if constexpr (ALGO == PMP_ALGO_MP) {
        // sum_{k != j} log K(j,k) in closed form about the current state: sum_k |d_j - d_k|^2 = P|d_j|^2 - 2 d_j.S1 + S2
        if (use_kernel && dim <= KDIM_MAX) {
            double my2 = 0.0;
            if (dim <= 3) {
                // one pass, one 4-value block reduction (S1 per coordinate, S2)
                double p0 = 0.0, p1 = 0.0, p2 = 0.0;
                for (int p = tid; p < P; p += ACCEPT_THREADS) {
                    double d0 = (double)__ldcg(a.props + (long long)p * dim) - (double)__ldcg(a.props);
                    double d1 = dim > 1 ? (double)__ldcg(a.props + (long long)p * dim + 1) - (double)__ldcg(a.props + 1) : 0.0;
                    double d2 = dim > 2 ? (double)__ldcg(a.props + (long long)p * dim + 2) - (double)__ldcg(a.props + 2) : 0.0;
                    p0 += d0; p1 += d1; p2 += d2;
                    my2 = fma(d0, d0, my2); my2 = fma(d1, d1, my2); my2 = fma(d2, d2, my2);
                }
                for (int o = 16; o > 0; o >>= 1) {
                    p0 += __shfl_xor_sync(0xffffffffu, p0, o); p1 += __shfl_xor_sync(0xffffffffu, p1, o);
                    p2 += __shfl_xor_sync(0xffffffffu, p2, o); my2 += __shfl_xor_sync(0xffffffffu, my2, o);
                }
                __syncthreads();
                if ((tid & 31) == 0) { int w = tid >> 5; sred4[w] = p0; sred4[32 + w] = p1; sred4[64 + w] = p2; sred4[96 + w] = my2; }
                __syncthreads();
                if (tid < 4) { double t = 0.0; for (int w = 0; w < (ACCEPT_THREADS >> 5); ++w) t += sred4[tid * 32 + w]; s1[tid] = t; }
                __syncthreads();
                my2 = s1[3];
            } else {
                for (int j = 0; j < dim; ++j) {
                    double part = 0.0;
                    for (int p = tid; p < P; p += ACCEPT_THREADS) {
                        double d = (double)a.props[(long long)p * dim + j] - (double)a.props[j];
                        part += d; my2 = fma(d, d, my2);
                    }
                    double tot = block_sum(part, red);
                    if (tid == 0) s1[j] = tot;
                }
                my2 = block_sum(my2, red);
                __syncthreads();
            }
            const double S2 = my2;
            for (int p = tid; p < P; p += ACCEPT_THREADS) {
                double dj2 = 0.0, dot = 0.0;
                for (int j = 0; j < dim; ++j) {
                    double d = (double)a.props[(long long)p * dim + j] - (double)a.props[j];
                    dj2 = fma(d, d, dj2); dot = fma(d, s1[j], dot);
                }
                double sumsq = (double)P * dj2 - 2.0 * dot + S2;
                double kt;
                if (cfg.flags & PMP_FLAG_KERNEL_MEAN)   // MP_FC.py:107-114: sum_k mean_dim(logK_jk) / P, k != j (tran[j][j] = 0)
                    kt = ((double)(P - 1) * log_norm_k - half_inv_ks2 * sumsq / (double)dim) / (double)P;
                else
                    kt = (double)(P - 1) * dim * log_norm_k - half_inv_ks2 * sumsq;
                A[p] = lt[p] + kt;
            }
        } else {
            for (int p = tid; p < P; p += ACCEPT_THREADS) A[p] = lt[p] + (use_kernel ? a.logw[p] : 0.0);  // kernel term precomputed into logw
        }
    } else if constexpr (ALGO == PMP_ALGO_PSP) {
        // binary tree Barker product (lb.py:216-240): m = a mod 2^(c+1), partner q = m xor 2^c.  K is symmetric so
        // log(w_new/(w_new+w_old)) = logsigmoid(lt[m] - lt[q]).
        for (int p = tid; p < P; p += ACCEPT_THREADS) {
            double s = 0.0;
            for (int c = 0; c < D; ++c) {
                int m = p & ((2 << c) - 1), q = m ^ (1 << c);
                s += logsigmoid(lt[m] - lt[q]);
            }
            A[p] = s;
        }
    } else if constexpr (ALGO == PMP_ALGO_PMP) {
        // general tree (lb.py:315-330): level i, stride s = b^i, group h < s = {h + j s}; A[h + j s] += log softmax_j(v),
        // v_j = lt + sum_{k != j} log K; then nodes [b^(i+1), b^(i+2)) inherit A[x mod b^(i+1)] (or the reference's
        // typo modulus b*(i+1) under PMP_FLAG_QUIRK_LEVEL_MOD).
        for (int p = tid; p < P; p += ACCEPT_THREADS) A[p] = 0.0;
        __syncthreads();
        long long s = 1;
        for (int i = 0; i < D; ++i) {
            for (long long h = tid; h < s; h += ACCEPT_THREADS) {
                double mx = -INFINITY;
                for (int j = 0; j < b; ++j) {
                    int nj = (int)(h + j * s);
                    double v = lt[nj];
                    if (use_kernel) for (int k = 0; k < b; ++k) if (k != j) v += log_kernel_pair(a.props, dim, nj, (int)(h + k * s), ks);
                    a.logw[nj] = v;        // scratch
                    mx = fmax(mx, v);
                }
                double se = 0.0;
                for (int j = 0; j < b; ++j) se += exp(a.logw[h + j * s] - mx);
                double lse = mx + log(se);
                for (int j = 0; j < b; ++j) { int nj = (int)(h + j * s); A[nj] += (mx == -INFINITY) ? -INFINITY : a.logw[nj] - lse; }
            }
            __syncthreads();
            if (i < D - 1) {
                long long lo = s * b, hi = s * b * b;
                long long mod = (cfg.flags & PMP_FLAG_QUIRK_LEVEL_MOD) ? (long long)b * (i + 1) : lo;
                for (long long x = lo + tid; x < hi; x += ACCEPT_THREADS) A[x] = A[x % mod];
                __syncthreads();
            }
            s *= b;
        }
    } else {  // PMP_ALGO_TABLE: lt + sum_d sum_{k != m in group_d(m)} log K(m,k), m = p mod b^(d+1)  (500_PMP.cu:23-30, conv_pmp.cu:22-33)
        for (int p = tid; p < P; p += ACCEPT_THREADS) {
            double v = lt[p];
            if (cfg.flags & PMP_FLAG_QUIRK_TABLE_CONST) {
                v += (double)D * (b - 1) * dim * (-HALF_LOG_2PI);       // every (from,to) reads node 0: temp = 0
            } else if (use_kernel) {
                long long s = 1;
                for (int d = 0; d < D; ++d) {
                    long long m = p % (s * b), h = m % s;
                    for (int k = 0; k < b; ++k) { long long o = h + k * s; if (o != m) v += log_kernel_pair(a.props, dim, (int)m, (int)o, ks); }
                    s *= b;
                }
            }
            A[p] = v;
        }
    }
    __syncthreads();

    if (cfg.flags & PMP_FLAG_STANDARDIZE) {   // A = (A - mean)/std, unbiased std (PMP_FC.py:138-140)
        double part = 0.0;
        for (int p = tid; p < P; p += ACCEPT_THREADS) part += A[p];
        double mean = block_sum(part, red) / P;
        part = 0.0;
        for (int p = tid; p < P; p += ACCEPT_THREADS) { double d = A[p] - mean; part = fma(d, d, part); }
        double var = block_sum(part, red) / (P - 1);
        double sd = sqrt(var);
        for (int p = tid; p < P; p += ACCEPT_THREADS) A[p] = (A[p] - mean) / sd;
        __syncthreads();
    }
    for (int p = tid; p < P; p += ACCEPT_THREADS) a.logw[p] = A[p];
    PMP_STAMP(dbg, 2);

    // ---- 3. draw -------------------------------------------------------------------------------------------
    int n_draws;
    if constexpr (ALGO == PMP_ALGO_MH) {
        n_draws = 1;
        if (tid == 0) {
            double u = a.uniforms ? a.uniforms[0] : u64_to_unit(stream_u64(a.seed, iter, STREAM_DRAW, 0));
            int acc;
            if (cfg.algo == PMP_ALGO_MH) acc = u < exp((double)cfg.mh_temperature * (lt[1] - lt[0]));
            else { double m = fmax(lt[0], lt[1]); double w0 = exp(lt[0] - m), w1 = exp(lt[1] - m); acc = (w1 / (w0 + w1)) > u; }
            a.draws[0] = acc; s_next = acc;
        }
        __syncthreads();
    } else {
        double mx = -INFINITY;
        for (int p = tid; p < P; p += ACCEPT_THREADS) mx = fmax(mx, A[p]);
        mx = block_max(mx, red);
        for (int p = tid; p < P; p += ACCEPT_THREADS) { double w = exp(A[p] - mx); A[p] = (w == w) ? w : 0.0; }
        __syncthreads();
        PMP_STAMP(dbg, 3);
        block_inclusive_scan(A, P, red);
        const double total = A[P - 1];
        PMP_STAMP(dbg, 4);
        // inverse-CDF draws on the unnormalised cdf: first k with cdf_k > u*total ('right', numpy) or >= ('left', libstdc++);
        // same index as normalising first except for u within an ulp of a boundary (oracle mirror: draw_blocked)
        n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
        const bool right = (cfg.draw != PMP_DRAW_CUDA);
        for (int t = tid; t < n_draws; t += ACCEPT_THREADS) {
            double u = a.uniforms ? a.uniforms[t] : u64_to_unit(stream_u64(a.seed, iter, STREAM_DRAW, (unsigned long long)t));
            const double thr = u * total;
            int lo = 0, hi = P;
            while (lo < hi) { int mid = (lo + hi) >> 1; bool go = right ? (A[mid] <= thr) : (A[mid] < thr); if (go) lo = mid + 1; else hi = mid; }
            a.draws[t] = min(lo, P - 1);
        }
        __syncthreads();
        if (tid == 0) {
            int nx;
            if (cfg.draw == PMP_DRAW_PYTHON) {
                double up = a.uniforms ? a.uniforms[P] : u64_to_unit(stream_u64(a.seed, iter, STREAM_PICK, 0));
                int pick = min(P - 1, (int)(up * (double)P));
                nx = a.draws[pick];
            } else nx = a.draws[0];
            s_next = nx;
        }
        __syncthreads();
    }
    const int next = s_next;
    PMP_STAMP(dbg, 5);

    // ---- 4. state, trace, counters -------------------------------------------------------------------------
    if (!a.advance) { if (tid == 0) a.cnt->last_next = next; return; }
    const bool rec = row < a.trace.capacity;
    if (rec && (a.trace.what & PMP_TRACE_DRAWS)) for (int t = tid; t < P; t += ACCEPT_THREADS) a.trace.draws[row * P + t] = t < n_draws ? a.draws[t] : -1;
    if (rec && (a.trace.what & PMP_TRACE_LOGW)) for (int t = tid; t < P; t += ACCEPT_THREADS) a.trace.logw[row * P + t] = a.logw[t];
    if (rec && (a.trace.what & PMP_TRACE_SAMPLES))
        for (long long g = tid; g < (long long)P * dim; g += ACCEPT_THREADS) {
            int t = (int)(g / dim), j = (int)(g - (long long)t * dim);
            int src = t < n_draws ? a.draws[t] : next;
            a.trace.samples[row * P * dim + g] = a.props[(long long)src * dim + j];
        }
    for (int j = tid; j < dim; j += ACCEPT_THREADS) {
        float v = a.props[(long long)next * dim + j];
        a.state[j] = v;
        if (rec && (a.trace.what & PMP_TRACE_STATE)) a.trace.state[row * dim + j] = v;
    }
    if (tid == 0) {
        if (rec && (a.trace.what & PMP_TRACE_NEXT)) a.trace.next[row] = next;
        if (rec) a.cnt->trace_rows = row + 1;
        a.cnt->iteration = iter + 1;
        a.cnt->last_next = next;
    }
    PMP_STAMP(dbg, 6);
}

template <int ALGO>
__global__ void __launch_bounds__(ACCEPT_THREADS, 1) accept_kernel(const __grid_constant__ AcceptArgs a) {
    extern __shared__ double accept_sm[];
    accept_device<ALGO>(a, accept_sm);
}

// proposal_value with the normals read from a prefetched table instead of being generated in place
__device__ __forceinline__ float proposal_value_z(const ProposeArgs& a, const float* __restrict__ z, int node, int j, float v) {
    if (a.tree == PMP_TREE_FLAT) {
        if (node > 0) v = __fadd_rn(v, __fmul_rn(a.alpha, __ldcg(z + (long long)node * a.dim + j)));   // L2 loads: the table is written by other SMs
        return v;
    }
    for_each_ancestor(a.tree, a.b, a.depth, node, [&](int anc) { v = __fadd_rn(v, __fmul_rn(a.alpha, __ldcg(z + (long long)anc * a.dim + j))); });
    return v;
}

}  // namespace pmp
