// accept_lean.cuh — the acceptance of the linear-Gaussian chain loop split by what it depends on.
//
// Same rules as accept_fast.cuh / accept.cuh (MP: GMOptimizer.step lb.py:139-164 and 500_MP.cu:207-243; binary tree:
// preMOptimizer.step lb.py:206-258; CUDA "table" rule 500_PMP.cu:23-30) and the same scan association (mirrored by
// oracle_blocked_cdf), but the work is cut into three pieces because the acceptance sits on the critical path of the
// chain — every cycle it takes is a cycle in which 147 SMs wait:
//   pre    everything that needs only the iteration counter and the nodes: Philox uniforms of the draws and of the pick,
//          the per-node constant -n/2 log(2 pi sigma^2), the proposal-kernel term of every node (MP: closed form about
//          the current state), and the standard normals of the NEXT iteration's nodes.  In the persistent kernel this
//          runs while the sweep CTAs are busy.
//   crit   what needs the sums: log-target, log-weight, exact maximum, exp, blocked scan, the ONE inverse-CDF draw that
//          decides the next state, and the next iteration's nodes written to global memory.  One L2 round trip, then
//          shared memory only.
//   post   everything nobody waits for: the other P-1 draws (the samples of the multi-proposal step: P random probes of
//          the cdf, shared-memory bank-conflict bound), log-targets / log-weights / draws to global memory, trace rows,
//          state, counters.
//          In the persistent kernel this runs after the sweep CTAs have been released.
// accept_lean_kernel runs the three pieces back to back (CUDA-graph loop, multi-GPU loop): identical arithmetic.
// P <= LEAN_MAX_P (two nodes per thread); larger trees use accept_fast.cuh.
#pragma once
#include "accept_fast.cuh"

namespace pmp {

constexpr int LEAN_MAX_P = 2 * ACCEPT_THREADS;
constexpr int LEAN_K = 2;

struct LeanSmem {
    double* lt;      // [P] log-targets
    double* A;       // [P] log-weights → weights → cdf
    double* c1;      // [P] -n/2 log(2 pi sigma^2)
    double* kt;      // [P] proposal-kernel term
    double* ls;      // [2P] PSP log-sigmoid table
    float* props;    // [3P] this iteration's nodes
    float* z;        // [3P] normals of the next iteration's nodes
    int* draw;       // [P]
};
__host__ __device__ inline size_t lean_smem_bytes(int P, int algo) { return (size_t)P * (4 * 8 + (algo == PMP_ALGO_PSP ? 16 : 0) + 12 + 12 + 4); }
__device__ __forceinline__ LeanSmem lean_carve(void* base, int P, int algo) {
    LeanSmem s; double* d = reinterpret_cast<double*>(base);
    s.lt = d; s.A = d + P; s.c1 = d + 2 * P; s.kt = d + 3 * P; d += 4 * P;
    s.ls = d; if (algo == PMP_ALGO_PSP) d += 2 * P;
    s.props = reinterpret_cast<float*>(d); s.z = s.props + 3 * P; s.draw = reinterpret_cast<int*>(s.z + 3 * P);
    return s;
}

struct LeanRegs {
    unsigned long long iter; long long row;
    double u[LEAN_K];        // uniforms of this thread's draws
    double logw[LEAN_K];     // log-weights of this thread's nodes
    double total;            // sum of the weights (last cdf entry)
    int next; float n0, n1, n2;
};

// ---- flag-in-data hand-offs of the persistent chain kernels ------------------------------------------------------------------------
// Everything that crosses between the sweep CTAs and the acceptance CTA inside an iteration travels as 8-byte words
// {32-bit payload | tag << 32}: an 8-byte access is single-copy atomic, so a word whose tag is current is complete by itself — no
// fence, no counter, no second L2 round trip (the scheme of the cross-GPU exchange in chain_persistent_multi.cuh, and of NCCL's LL
// protocol).  tag of iteration `it` of a launch = epoch + it + 1, epoch = iterations of all earlier launches on this context (never
// reused, so nothing is ever reset and a stale word can never look current).
// (The other direction — the sweep CTAs' partial sums — stays an L2-side reduction: integer RED.64 per node, fence, arrival counter.  A
//  slot-per-CTA tagged variant was built and measured: the acceptance SM has to pull ~18 slots x 16 B x 1024 nodes through its own L2
//  port every poll round, 28.3 us per iteration against 19.8.)
//   nodes  next iteration's nodes {b0 | tag}, {b1 | tag}, {sigma | tag} at [node][4]; the sweep CTAs poll their tile's nodes
//   zt     next iteration's standard normals {float | tag} at [2][3P] (half by Philox iteration parity), written by the sweep CTAs'
//          side job, read by the acceptance CTA
// One buffer of nodes suffices: the nodes of it + 1 are written after every sweep CTA has arrived for it, i.e. after it has read the
// nodes of it.  Two halves of normals: a sweep CTA writes the normals of it + 1 (half (it + 1) & 1) at the start of its sweep it, which it
// can only enter after the acceptance of it - 1 — the reader of the other half — has published its nodes.
//   state  (flat trees, chain_persistent_kernel only) the accepted state {b0 | tag}, {b1 | tag}, {sigma | tag}: with a flat tree a node is
//          state + alpha * z(node), and z depends on counters only, so every sweep CTA derives its own tile's nodes from these three words
//          and the normals it computed while it waited — the acceptance publishes 24 bytes instead of generating and storing P nodes on the
//          critical path, and the sweep CTAs save the second L2 round trip of fetch_node.
struct Handoff {
    unsigned long long* nodes;
    unsigned long long* zt;
    unsigned int epoch;
    unsigned long long* state;
};
__host__ __device__ inline size_t handoff_node_words(int P) { return (size_t)P * 4; }
__host__ __device__ inline size_t handoff_z_words(int P) { return (size_t)2 * 3 * P; }
__host__ __device__ inline size_t handoff_state_words() { return 4; }

__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_gpu_v2(unsigned long long* p, unsigned long long a, unsigned long long b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ld_relaxed_gpu_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ bool hs_tag_ok(unsigned long long w, unsigned long long tag) { return (w & 0xffffffff00000000ull) == tag; }
// a hand-off that never arrives is a protocol bug (or a dead peer CTA): trap instead of hanging the GPU
struct SpinGuard {
    unsigned spins = 0; unsigned long long t0 = 0;
    __device__ __forceinline__ void tick() { if ((++spins & 4095u) == 0) { const unsigned long long t = globaltimer_ns(); if (t0 == 0) t0 = t; else if (t - t0 > 20000000000ull) __trap(); } }
};

// proposal_value_z (accept.cuh) with the normals in shared memory
__device__ __forceinline__ float proposal_value_zs(const ProposeArgs& a, const float* z, int node, int j, float v) {
    if (a.tree == PMP_TREE_FLAT) { if (node > 0) v = __fadd_rn(v, __fmul_rn(a.alpha, z[node * a.dim + j])); return v; }
    for_each_ancestor(a.tree, a.b, a.depth, node, [&](int anc) { v = __fadd_rn(v, __fmul_rn(a.alpha, z[anc * a.dim + j])); });
    return v;
}

// Where the normals of the next iteration's nodes come from.  TABLE_PRE: the table the previous sweep KERNEL filled (stepwise
// loop: it is complete before this kernel starts).  TABLE_CRIT: the table the sweep CTAs of the same persistent kernel fill
// during the sweep — complete only once they have all arrived, so it is read at the start of the critical phase, in the
// same L2 round trip as the sums.  GENERATE: computed here (3P quantile evaluations in binary64 on one SM: slow).
// HS_PRE: the tagged table of the flag-in-data hand-offs, read in the pre phase (a tagged word is complete by itself, so the read need not wait for the
// arrivals: it spins until the sweep CTAs' side job has written it, which depends on nothing this CTA still has to do).  DERIVE: flat trees in
// chain_persistent_kernel — the critical phase publishes the accepted state only; the nodes are derived from it by their readers (the sweep CTAs, and
// lean_derive_props here, off the critical path).
enum { LEAN_Z_TABLE_PRE = 0, LEAN_Z_GENERATE = 1, LEAN_Z_TABLE_CRIT = 2, LEAN_Z_HS_PRE = 3, LEAN_Z_DERIVE = 4 };

// tagged normals -> s.z (spins until every word carries `tag`)
__device__ __forceinline__ void lean_read_tagged_z(const LeanSmem& s, const unsigned long long* zn, int P, unsigned long long tag) {
    const int tid = threadIdx.x;
    unsigned long long zw[3 * LEAN_K];
#pragma unroll
    for (int k = 0; k < 3 * LEAN_K; ++k) { const int g = tid + k * ACCEPT_THREADS; zw[k] = (g < 3 * P) ? ld_relaxed_gpu_u64(zn + g) : tag; }
#pragma unroll
    for (int k = 0; k < 3 * LEAN_K; ++k) {
        const int g = tid + k * ACCEPT_THREADS;
        if (g < 3 * P) { SpinGuard sg; while (!hs_tag_ok(zw[k], tag)) { sg.tick(); zw[k] = ld_relaxed_gpu_u64(zn + g); } s.z[g] = __uint_as_float((unsigned)zw[k]); }
    }
}

// DERIVE: the nodes of iteration `iter` from the state the previous critical phase accepted (r.n0..n2) and the tagged normals of `iter`, into shared memory and
// into the plain copy in global memory (host reads, first iteration of the next launch).
__device__ __forceinline__ void lean_derive_props(const AcceptFastArgs& fa, const LeanSmem& s, const LeanRegs& r, const Handoff& hs, unsigned long long iter, unsigned long long tag) {
    const int P = fa.base.P, tid = threadIdx.x;
    lean_read_tagged_z(s, hs.zt + (iter & 1) * (long long)(P * 3), P, tag);
    float* props_out = const_cast<float*>(fa.base.props);
    if (fa.gen.tree != PMP_TREE_FLAT) __syncthreads();      // a tree node reads its ancestors' normals, written by other threads (flat: thread g reads the s.z[g] it wrote itself)
    for (int g = tid; g < 3 * P; g += ACCEPT_THREADS) {
        const int node = g / 3, j = g - 3 * node;
        const float v = proposal_value_zs(fa.gen, s.z, node, j, j == 0 ? r.n0 : (j == 1 ? r.n1 : r.n2));
        s.props[g] = v; props_out[g] = v;
    }
}

template <int ALGO>
__device__ __forceinline__ void lean_pre(const AcceptFastArgs& fa, const LeanSmem& s, LeanRegs& r, double (*red)[32], int* s_pick, int z_mode,
                                         const Handoff* hs = nullptr, unsigned long long tag = 0, bool first = true) {
    const AcceptArgs& a = fa.base;
    const pmp_config& cfg = a.cfg;
    const int P = a.P, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* dbg = a.dbg ? a.dbg + 32 : nullptr;
    PMP_STAMP(dbg, 0);
    if (z_mode == LEAN_Z_DERIVE && !first) {      // the counters this CTA advanced itself in the previous iteration (crit / post): no L2 round trip, and no fence needed after post
        r.iter += 1;
        if (r.row < a.trace.capacity) r.row += 1;
    } else {
        r.iter = __ldcg(&a.cnt->iteration);
        r.row = __ldcg(&a.cnt->trace_rows);
    }
    const int n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
    if (tid == 0) {
        double up = 0.0;
        if (cfg.draw == PMP_DRAW_PYTHON) up = a.uniforms ? a.uniforms[P] : u64_to_unit(stream_u64(a.seed, r.iter, STREAM_PICK, 0));
        *s_pick = min(P - 1, (int)(up * (double)P));
    }
    if (z_mode == LEAN_Z_DERIVE && !first) lean_derive_props(fa, s, r, *hs, r.iter, tag);
    else for (int g = tid; g < 3 * P; g += ACCEPT_THREADS) s.props[g] = __ldcg(a.props + g);
    if (fa.make_next) {
        if (z_mode == LEAN_Z_HS_PRE) lean_read_tagged_z(s, hs->zt + ((r.iter + 1) & 1) * (long long)(P * 3), P, tag);
        else if (z_mode == LEAN_Z_GENERATE) for (int g = tid; g < 3 * P; g += ACCEPT_THREADS) s.z[g] = (float)stream_step(fa.gen.seed, r.iter + 1, (unsigned long long)g, fa.gen.uniform);
        else if (z_mode == LEAN_Z_TABLE_PRE) { const float* zn = fa.z + ((r.iter + 1) & 1) * (long long)(P * 3); for (int g = tid; g < 3 * P; g += ACCEPT_THREADS) s.z[g] = __ldcg(zn + g); }
    }
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) {
        const int t = tid + k * ACCEPT_THREADS;
        r.u[k] = (t < n_draws) ? (a.uniforms ? a.uniforms[t] : u64_to_unit(stream_u64(a.seed, r.iter, STREAM_DRAW, (unsigned long long)t))) : 0.0;
    }
    __syncthreads();
    const float s0 = s.props[0], s1v = s.props[1], s2v = s.props[2];      // node 0 = current state
    const bool use_kernel = !(cfg.flags & PMP_FLAG_NO_KERNEL_TERM);
    const int D = (cfg.tree == PMP_TREE_FLAT) ? 1 : cfg.depth;
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    for (int p = tid; p < P; p += ACCEPT_THREADS) {
        const double sg = (double)s.props[3 * p + 2];
        s.c1[p] = -0.5 * (double)a.n_global * log(6.283185307179586477 * sg * sg);
        if (ALGO == PMP_ALGO_MP) {
            const double d0 = (double)s.props[3 * p] - (double)s0, d1 = (double)s.props[3 * p + 1] - (double)s1v, d2 = sg - (double)s2v;
            p0 += d0; p1 += d1; p2 += d2;
            p3 = fma(d0, d0, p3); p3 = fma(d1, d1, p3); p3 = fma(d2, d2, p3);
        }
    }
    if (ALGO == PMP_ALGO_MP) {
        const double log_norm_k = (cfg.kernel_sigma == 1.0f) ? -HALF_LOG_2PI : -HALF_LOG_2PI - log((double)cfg.kernel_sigma);
        const double half_inv_ks2 = 0.5 / ((double)cfg.kernel_sigma * (double)cfg.kernel_sigma);
        double S1x = 0.0, S1y = 0.0, S1z = 0.0, S2 = 0.0;
        if (use_kernel) {
            p0 = warp_sum_all(p0); p1 = warp_sum_all(p1); p2 = warp_sum_all(p2); p3 = warp_sum_all(p3);
            if (lane == 0) { red[0][warp] = p0; red[1][warp] = p1; red[2][warp] = p2; red[3][warp] = p3; }
            __syncthreads();
            S1x = warp_sum_all(red[0][lane]); S1y = warp_sum_all(red[1][lane]); S1z = warp_sum_all(red[2][lane]); S2 = warp_sum_all(red[3][lane]);
        }
        for (int p = tid; p < P; p += ACCEPT_THREADS) {
            double kt = 0.0;
            if (use_kernel) {
                const double d0 = (double)s.props[3 * p] - (double)s0, d1 = (double)s.props[3 * p + 1] - (double)s1v, d2 = (double)s.props[3 * p + 2] - (double)s2v;
                const double dj2 = fma(d2, d2, fma(d1, d1, d0 * d0));
                const double dot = fma(d2, S1z, fma(d1, S1y, d0 * S1x));
                const double sumsq = (double)P * dj2 - 2.0 * dot + S2;
                if (cfg.flags & PMP_FLAG_KERNEL_MEAN) kt = ((double)(P - 1) * log_norm_k - half_inv_ks2 * sumsq / 3.0) / (double)P;
                else kt = (double)(P - 1) * 3.0 * log_norm_k - half_inv_ks2 * sumsq;
            }
            s.kt[p] = kt;
        }
    } else if (ALGO == PMP_ALGO_TABLE) {
        const double kc = (cfg.flags & PMP_FLAG_QUIRK_TABLE_CONST) ? (double)D * ((cfg.tree == PMP_TREE_BINARY ? 2 : cfg.b) - 1) * 3.0 * (-HALF_LOG_2PI) : 0.0;
        for (int p = tid; p < P; p += ACCEPT_THREADS) s.kt[p] = kc;
    }
    __syncthreads();
    PMP_STAMP(dbg, 1);
}

// qin: the per-node sums are already in registers.  hs (persistent kernels with flag-in-data hand-offs): the next iteration's normals come from the tagged table hs->zt and the next nodes
// are ALSO published as tagged words in hs->nodes; `tag` is this iteration's, `tag + 2^32` the next one's.
template <int ALGO>
__device__ __forceinline__ void lean_crit(const AcceptFastArgs& fa, const LeanSmem& s, LeanRegs& r, double (*red)[32], const int* s_pick, int z_mode,
                                          const unsigned long long* qin = nullptr, const Handoff* hs = nullptr, unsigned long long tag = 0) {
    const AcceptArgs& a = fa.base;
    const pmp_config& cfg = a.cfg;
    const int P = a.P, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* dbg = a.dbg ? a.dbg + 32 : nullptr;
    PMP_STAMP(dbg, 2);
    const int n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
    const bool right = (cfg.draw != PMP_DRAW_CUDA);
    const int D = (cfg.tree == PMP_TREE_FLAT) ? 1 : cfg.depth;

    // ---- log-targets from the integer sums (one L2 round trip), log-weights, maximum ------------------------------------
    unsigned long long q[LEAN_K];
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) { const int p = tid + k * ACCEPT_THREADS; q[k] = (p < P) ? (qin ? qin[k] : __ldcg(a.acc + p)) : 0ull; }
    if (z_mode == LEAN_Z_HS_PRE || z_mode == LEAN_Z_DERIVE) {
        // the normals were read in the pre phase / are not needed here
    } else if (hs && fa.make_next) {                          // tagged normals: written by the sweep CTAs' side job at the start of their sweep — long since there
        lean_read_tagged_z(s, hs->zt + ((r.iter + 1) & 1) * (long long)(P * 3), P, tag);
    } else if (z_mode == LEAN_Z_TABLE_CRIT && fa.make_next) {       // same L2 round trip as the sums; first read after the block barriers below
        const float* zn = fa.z + ((r.iter + 1) & 1) * (long long)(P * 3);
        float zr[3 * LEAN_K];
#pragma unroll
        for (int k = 0; k < 3 * LEAN_K; ++k) { const int g = tid + k * ACCEPT_THREADS; zr[k] = (g < 3 * P) ? __ldcg(zn + g) : 0.f; }
#pragma unroll
        for (int k = 0; k < 3 * LEAN_K; ++k) { const int g = tid + k * ACCEPT_THREADS; if (g < 3 * P) s.z[g] = zr[k]; }
    }
    double mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) {
        const int p = tid + k * ACCEPT_THREADS;
        if (p < P) {
            if (!qin) a.acc[p] = 0ull;
            const double S = (double)(long long)q[k] * (1.0 / (double)(1 << FX_SHIFT));
            double v = (s.c1[p] - 0.5 * S) * a.inv_scale;
            if ((double)(long long)q[k] >= a.sat_limit || !(v == v)) v = -INFINITY;
            s.lt[p] = v;
            if (ALGO != PMP_ALGO_PSP) { v += s.kt[p]; s.A[p] = v; r.logw[k] = v; mx = fmax(mx, v); }
        }
    }
    if (ALGO == PMP_ALGO_PSP) {
        __syncthreads();
        // ls[off(c) + m] = logsigmoid(lt[m] - lt[m ^ 2^c]), m < 2^(c+1), off(c) = 2^(c+1) - 2
        for (int e = tid; e < 2 * P - 2; e += ACCEPT_THREADS) {
            const int c = 31 - __clz(e + 2) - 1, m = e + 2 - (2 << c);
            s.ls[e] = logsigmoid(s.lt[m] - s.lt[m ^ (1 << c)]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LEAN_K; ++k) {
            const int p = tid + k * ACCEPT_THREADS;
            if (p < P) {
                double v = 0.0;
                for (int c = 0; c < D; ++c) v += s.ls[(2 << c) - 2 + (p & ((2 << c) - 1))];
                s.A[p] = v; r.logw[k] = v; mx = fmax(mx, v);
            }
        }
    }
    mx = warp_max_all(mx);
    if (lane == 0) red[0][warp] = mx;
    __syncthreads();
    mx = warp_max_all(red[0][lane]);
    PMP_STAMP(dbg, 3);

    // ---- weights and the blocked inclusive scan (association mirrored by oracle_blocked_cdf) ------------------------------
    const int ipt = (P + ACCEPT_THREADS - 1) / ACCEPT_THREADS;
    if (ipt > 1) {
        for (int p = tid; p < P; p += ACCEPT_THREADS) { double w = exp(s.A[p] - mx); s.A[p] = (w == w) ? w : 0.0; }
        __syncthreads();
    }
    const int i0 = tid * ipt;
    double run = 0.0;
    if (ipt == 1) { if (tid < P) { double w = exp(s.A[tid] - mx); run = (w == w) ? w : 0.0; } }
    else for (int i = 0; i < ipt; ++i) if (i0 + i < P) { run += s.A[i0 + i]; s.A[i0 + i] = run; }
    double incl = run;
    for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl = n + incl; }
    if (lane == 31) red[1][warp] = incl;
    __syncthreads();
    double wt = red[1][lane];                       // every warp scans the warp totals itself
    for (int o = 1; o < 32; o <<= 1) { double n = __shfl_up_sync(0xffffffffu, wt, o); if (lane >= o) wt = n + wt; }
    const double total = __shfl_sync(0xffffffffu, wt, 31);
    const double warp_off = warp > 0 ? __shfl_sync(0xffffffffu, wt, max(warp - 1, 0)) : 0.0;
    double lane_excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) lane_excl = 0.0;
    const double excl = warp_off + lane_excl;
    if (ipt == 1) { if (tid < P) s.A[tid] = excl + run; }
    else for (int i = 0; i < ipt; ++i) if (i0 + i < P) s.A[i0 + i] = excl + s.A[i0 + i];
    __syncthreads();
    PMP_STAMP(dbg, 4);

    // ---- the one inverse-CDF draw that decides the next state (the other P-1 are samples nobody waits for: lean_post) --------
    r.total = total;
    {
        // one warp probes the cdf cooperatively: 32 block pivots, then the 32 (or 64) entries of the first block that is not
        // entirely below the threshold — two shared-memory rounds instead of a ten-step dependent binary search.  The count
        // of entries that pass the test is exactly what the binary search (lean_post) returns for a monotone cdf.
        const int t_need = (cfg.draw == PMP_DRAW_PYTHON) ? *s_pick : 0;
        if (warp == ((t_need & (ACCEPT_THREADS - 1)) >> 5)) {
            const double thr = __shfl_sync(0xffffffffu, t_need >= ACCEPT_THREADS ? r.u[1] : r.u[0], t_need & 31) * total;
            const int S = (P + 31) >> 5;
            const int piv = min(P, (lane + 1) * S) - 1;
            const bool g0 = (lane * S < P) && (right ? (s.A[piv] <= thr) : (s.A[piv] < thr));
            const int base = __popc(__ballot_sync(0xffffffffu, g0)) * S;
            int cnt = base;
            if (base < P) {
                const int i1 = base + lane, i2 = base + 32 + lane;
                const bool g1 = lane < S && i1 < P && (right ? (s.A[i1] <= thr) : (s.A[i1] < thr));
                const bool g2 = 32 + lane < S && i2 < P && (right ? (s.A[i2] <= thr) : (s.A[i2] < thr));
                cnt = base + __popc(__ballot_sync(0xffffffffu, g1)) + __popc(__ballot_sync(0xffffffffu, g2));
            }
            if (lane == 0) {
                const int nx = min(cnt, P - 1);
                red[2][0] = (double)nx;
                if (z_mode == LEAN_Z_DERIVE && a.advance) {          // what the sweep CTAs wait for: 24 bytes
                    const unsigned long long tn = tag + (1ull << 32);
                    st_relaxed_gpu_v2(hs->state, (unsigned long long)__float_as_uint(s.props[3 * nx]) | tn, (unsigned long long)__float_as_uint(s.props[3 * nx + 1]) | tn);
                    st_relaxed_gpu_u64(hs->state + 2, (unsigned long long)__float_as_uint(s.props[3 * nx + 2]) | tn);
                }
            }
        }
    }
    __syncthreads();
    r.next = (int)red[2][0];
    r.n0 = s.props[3 * r.next]; r.n1 = s.props[3 * r.next + 1]; r.n2 = s.props[3 * r.next + 2];
    PMP_STAMP(dbg, 5);

    // ---- what the next sweep waits for: its nodes and the iteration counter ------------------------------------------------
    if (a.advance) {
        if (fa.make_next && z_mode != LEAN_Z_DERIVE) {
            float* props_out = const_cast<float*>(a.props);
            const unsigned long long tag_next = tag + (1ull << 32);
            for (int g = tid; g < P * 3; g += ACCEPT_THREADS) {
                const int node = g / 3, j = g - 3 * node;
                const float v = proposal_value_zs(fa.gen, s.z, node, j, j == 0 ? r.n0 : (j == 1 ? r.n1 : r.n2));
                if (hs) st_relaxed_gpu_u64(hs->nodes + 4 * node + j, (unsigned long long)__float_as_uint(v) | tag_next);   // what the sweep CTAs poll
                props_out[g] = v;
            }
        }
        if (tid == 0) a.cnt->iteration = r.iter + 1;
    }
    PMP_STAMP(dbg, 6);
}

template <int ALGO>
__device__ __forceinline__ void lean_post(const AcceptFastArgs& fa, const LeanSmem& s, const LeanRegs& r) {
    const AcceptArgs& a = fa.base;
    const pmp_config& cfg = a.cfg;
    const int P = a.P, tid = threadIdx.x;
    const int n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
    const bool right = (cfg.draw != PMP_DRAW_CUDA);
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) {       // all inverse-CDF draws (the cdf is still in shared memory)
        const int t = tid + k * ACCEPT_THREADS;
        if (t < n_draws) {
            const double thr = r.u[k] * r.total;
            int lo = 0, hi = P;
            while (lo < hi) { int mid = (lo + hi) >> 1; bool go = right ? (s.A[mid] <= thr) : (s.A[mid] < thr); if (go) lo = mid + 1; else hi = mid; }
            s.draw[t] = min(lo, P - 1);
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) {
        const int p = tid + k * ACCEPT_THREADS;
        if (p < P) { a.lt[p] = s.lt[p]; a.logw[p] = r.logw[k]; if (p < n_draws) a.draws[p] = s.draw[p]; }
    }
    if (!a.advance) { if (tid == 0) a.cnt->last_next = r.next; return; }
    const long long row = r.row;
    const bool rec = row < a.trace.capacity;
    if (rec) {
        if (a.trace.what & PMP_TRACE_DRAWS) for (int t = tid; t < P; t += ACCEPT_THREADS) a.trace.draws[row * P + t] = t < n_draws ? s.draw[t] : -1;
        if (a.trace.what & PMP_TRACE_LOGW) {
#pragma unroll
            for (int k = 0; k < LEAN_K; ++k) { const int p = tid + k * ACCEPT_THREADS; if (p < P) a.trace.logw[row * P + p] = r.logw[k]; }
        }
        if (a.trace.what & PMP_TRACE_SAMPLES)
            for (int g = tid; g < P * 3; g += ACCEPT_THREADS) { int t = g / 3, j = g - 3 * t; a.trace.samples[row * P * 3 + g] = s.props[3 * (t < n_draws ? s.draw[t] : r.next) + j]; }
    }
    if (tid == 0) {
        a.state[0] = r.n0; a.state[1] = r.n1; a.state[2] = r.n2;
        if (rec && (a.trace.what & PMP_TRACE_STATE)) { a.trace.state[row * 3] = r.n0; a.trace.state[row * 3 + 1] = r.n1; a.trace.state[row * 3 + 2] = r.n2; }
        if (rec && (a.trace.what & PMP_TRACE_NEXT)) a.trace.next[row] = r.next;
        if (rec) a.cnt->trace_rows = row + 1;
        a.cnt->last_next = r.next;
    }
    unsigned long long* dbg = a.dbg ? a.dbg + 32 : nullptr;
    PMP_STAMP(dbg, 7);
}

template <int ALGO>
__global__ void __launch_bounds__(ACCEPT_THREADS, 1) accept_lean_kernel(const __grid_constant__ AcceptFastArgs fa) {
    extern __shared__ __align__(16) unsigned char accept_lean_sm[];
    __shared__ double red[4][32];
    __shared__ int s_pick;
    const LeanSmem s = lean_carve(accept_lean_sm, fa.base.P, ALGO);
    LeanRegs r;
    lean_pre<ALGO>(fa, s, r, red, &s_pick, LEAN_Z_TABLE_PRE);
    lean_crit<ALGO>(fa, s, r, red, &s_pick, LEAN_Z_TABLE_PRE);
    __syncthreads();
    lean_post<ALGO>(fa, s, r);
}

}  // namespace pmp
