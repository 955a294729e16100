// accept_fast.cuh — latency-lean acceptance for the linear-Gaussian chain loop (dim 3; MP, binary-Barker PSP, or the
// CUDA "table" rule with its constant transition term).  Same mathematics and the same scan association as
// accept_device (accept.cuh) — tests check both against the oracle — but organised around what this kernel costs on
// a B200: it is ONE CTA that runs between two sweeps, so its time is a chain of dependent latencies, not throughput.
//   * everything that does not depend on the sweep (Philox uniforms, node coordinates, counters) is issued first,
//     in parallel, so the chain has one global round trip instead of one per phase;
//   * block reductions use redundant per-warp shuffle trees (one __syncthreads each instead of three);
//   * draws and scratch stay in shared memory; nothing written to global memory is read back;
//   * the binary tree evaluates each distinct log-sigmoid once (2P-2 of them) instead of D per node;
//   * at the end it publishes the NEXT iteration's nodes from the new state and the prefetched normals table, so the
//     next sweep starts from a plain coalesced read (replaces the separate proposal launch).
#pragma once
#include "accept.cuh"

namespace pmp {

struct AcceptFastArgs {
    AcceptArgs base;
    const float* z;        // [2, P*3] normals table; half ((iter+1) & 1) belongs to the next iteration
    int make_next;         // 1: overwrite props with the next iteration's nodes
    ProposeArgs gen;       // tree shape and alpha for make_next
};

__device__ __forceinline__ double warp_sum_all(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ double warp_max_all(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }

template <int ALGO>
__device__ __forceinline__ void accept_fast_body(const AcceptFastArgs& fa, double* sm) {
    const AcceptArgs& a = fa.base;
    const int P = a.P;
    double* lt = sm;                     // [P]   log-targets; later reused as int32 draws
    double* A = sm + P;                  // [P]   log-weights → weights → cdf
    double* ls = sm + 2 * P;             // [2P]  PSP: log-sigmoid table
    __shared__ double red[4][32];
    __shared__ int s_pick, s_next;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const pmp_config& cfg = a.cfg;
    unsigned long long* dbg = a.dbg ? a.dbg + 32 : nullptr;
    PMP_STAMP(dbg, 0);

    // ---- phase 0: everything independent of the sweep -------------------------------------------------------------
    const unsigned long long iter = __ldcg(&a.cnt->iteration);      // L2 loads throughout: in the persistent kernel these change
    const long long row = __ldcg(&a.cnt->trace_rows);               // between iterations of the same launch
    const float s0 = __ldcg(a.props), s1v = __ldcg(a.props + 1), s2v = __ldcg(a.props + 2);   // node 0 = current state
    const int n_draws = (cfg.draw == PMP_DRAW_SINGLE) ? 1 : P;
    const bool right = (cfg.draw != PMP_DRAW_CUDA);
    if (tid == 0) {
        double up = 0.0;
        if (cfg.draw == PMP_DRAW_PYTHON) up = a.uniforms ? a.uniforms[P] : u64_to_unit(stream_u64(a.seed, iter, STREAM_PICK, 0));
        s_pick = min(P - 1, (int)(up * (double)P));
    }
    const double log_norm_k = (cfg.kernel_sigma == 1.0f) ? -HALF_LOG_2PI : -HALF_LOG_2PI - log((double)cfg.kernel_sigma);
    const double half_inv_ks2 = 0.5 / ((double)cfg.kernel_sigma * (double)cfg.kernel_sigma);
    const bool use_kernel = !(cfg.flags & PMP_FLAG_NO_KERNEL_TERM);
    const int D = (cfg.tree == PMP_TREE_FLAT) ? 1 : cfg.depth;

    // ---- phase 1: log-targets from the integer sums; MP partial sums in the same pass -----------------------------
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
    for (int p = tid; p < P; p += ACCEPT_THREADS) {
        const float t0 = __ldcg(a.props + 3 * p), t1 = __ldcg(a.props + 3 * p + 1), t2 = __ldcg(a.props + 3 * p + 2);
        const unsigned long long q = __ldcg(a.acc + p);
        a.acc[p] = 0ull;
        const double sg = (double)t2;
        const double S = (double)(long long)q * (1.0 / (double)(1 << FX_SHIFT));
        double v = (-0.5 * (double)a.n_global * log(6.283185307179586477 * sg * sg) - 0.5 * S) * a.inv_scale;
        if ((double)(long long)q >= a.sat_limit || !(v == v)) v = -INFINITY;
        lt[p] = v;
        a.lt[p] = v;
        if (ALGO == PMP_ALGO_MP) {
            const double d0 = (double)t0 - (double)s0, d1 = (double)t1 - (double)s1v, d2 = (double)t2 - (double)s2v;
            p0 += d0; p1 += d1; p2 += d2;
            p3 = fma(d0, d0, p3); p3 = fma(d1, d1, p3); p3 = fma(d2, d2, p3);
        }
    }
    if (ALGO == PMP_ALGO_MP && use_kernel) {
        p0 = warp_sum_all(p0); p1 = warp_sum_all(p1); p2 = warp_sum_all(p2); p3 = warp_sum_all(p3);
        if (lane == 0) { red[0][warp] = p0; red[1][warp] = p1; red[2][warp] = p2; red[3][warp] = p3; }
    }
    __syncthreads();
    PMP_STAMP(dbg, 1);

    // ---- phase 2: log-weights -------------------------------------------------------------------------------------
    double mx = -INFINITY;
    if (ALGO == PMP_ALGO_MP) {
        double S1x = 0.0, S1y = 0.0, S1z = 0.0, S2 = 0.0;
        if (use_kernel) {   // every warp reduces the 32 per-warp partials itself: no second barrier
            S1x = warp_sum_all(red[0][lane]); S1y = warp_sum_all(red[1][lane]); S1z = warp_sum_all(red[2][lane]); S2 = warp_sum_all(red[3][lane]);
        }
        for (int p = tid; p < P; p += ACCEPT_THREADS) {
            double v = lt[p];
            if (use_kernel) {
                const double d0 = (double)__ldcg(a.props + 3 * p) - (double)s0, d1 = (double)__ldcg(a.props + 3 * p + 1) - (double)s1v,
                             d2 = (double)__ldcg(a.props + 3 * p + 2) - (double)s2v;
                const double dj2 = fma(d2, d2, fma(d1, d1, d0 * d0));
                const double dot = fma(d2, S1z, fma(d1, S1y, d0 * S1x));
                const double sumsq = (double)P * dj2 - 2.0 * dot + S2;
                if (cfg.flags & PMP_FLAG_KERNEL_MEAN) v += ((double)(P - 1) * log_norm_k - half_inv_ks2 * sumsq / 3.0) / (double)P;
                else v += (double)(P - 1) * 3.0 * log_norm_k - half_inv_ks2 * sumsq;
            }
            A[p] = v; a.logw[p] = v; mx = fmax(mx, v);
        }
    } else if (ALGO == PMP_ALGO_PSP) {
        // ls[off(c) + m] = logsigmoid(lt[m] - lt[m ^ 2^c]), m < 2^(c+1), off(c) = 2^(c+1) - 2
        for (int e = tid; e < 2 * P - 2; e += ACCEPT_THREADS) {
            const int c = 31 - __clz(e + 2) - 1, m = e + 2 - (2 << c);
            ls[e] = logsigmoid(lt[m] - lt[m ^ (1 << c)]);
        }
        __syncthreads();
        for (int p = tid; p < P; p += ACCEPT_THREADS) {
            double v = 0.0;
            for (int c = 0; c < D; ++c) v += ls[(2 << c) - 2 + (p & ((2 << c) - 1))];
            A[p] = v; a.logw[p] = v; mx = fmax(mx, v);
        }
    } else {   // TABLE with the shipped constant transition term (or none)
        const double kc = (cfg.flags & PMP_FLAG_QUIRK_TABLE_CONST) ? (double)D * ((cfg.tree == PMP_TREE_BINARY ? 2 : cfg.b) - 1) * 3.0 * (-HALF_LOG_2PI) : 0.0;
        for (int p = tid; p < P; p += ACCEPT_THREADS) { double v = lt[p] + kc; A[p] = v; a.logw[p] = v; mx = fmax(mx, v); }
    }
    PMP_STAMP(dbg, 2);
    // uniforms of this thread's draws: issued here so the integer work overlaps the reductions' latency
    double u_first = 0.0;
    if (tid < n_draws) u_first = a.uniforms ? a.uniforms[tid] : u64_to_unit(stream_u64(a.seed, iter, STREAM_DRAW, (unsigned long long)tid));

    mx = warp_max_all(mx);
    if (lane == 0) red[0][warp] = mx;
    __syncthreads();
    mx = warp_max_all(red[0][lane]);
    PMP_STAMP(dbg, 3);

    // ---- phase 3: weights and the blocked inclusive scan (association mirrored by oracle_blocked_cdf) ---------------
    const int ipt = (P + ACCEPT_THREADS - 1) / ACCEPT_THREADS;
    if (ipt > 1) {
        for (int p = tid; p < P; p += ACCEPT_THREADS) { double w = exp(A[p] - mx); A[p] = (w == w) ? w : 0.0; }
        __syncthreads();
    }
    const int i0 = tid * ipt;
    double run = 0.0;
    if (ipt == 1) { if (tid < P) { double w = exp(A[tid] - mx); run = (w == w) ? w : 0.0; } }
    else for (int i = 0; i < ipt; ++i) if (i0 + i < P) { run += A[i0 + i]; A[i0 + i] = run; }
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
    if (ipt == 1) { if (tid < P) A[tid] = excl + run; }
    else for (int i = 0; i < ipt; ++i) if (i0 + i < P) A[i0 + i] = excl + A[i0 + i];
    __syncthreads();
    PMP_STAMP(dbg, 4);

    // ---- phase 4: inverse-CDF draws ---------------------------------------------------------------------------------
    int* sdraw = reinterpret_cast<int*>(lt);
    for (int t = tid; t < n_draws; t += ACCEPT_THREADS) {
        const double u = (t == tid) ? u_first : (a.uniforms ? a.uniforms[t] : u64_to_unit(stream_u64(a.seed, iter, STREAM_DRAW, (unsigned long long)t)));
        const double thr = u * total;
        int lo = 0, hi = P;
        while (lo < hi) { int mid = (lo + hi) >> 1; bool go = right ? (A[mid] <= thr) : (A[mid] < thr); if (go) lo = mid + 1; else hi = mid; }
        lo = min(lo, P - 1);
        sdraw[t] = lo; a.draws[t] = lo;
    }
    __syncthreads();
    const int next = (cfg.draw == PMP_DRAW_PYTHON) ? sdraw[s_pick] : sdraw[0];
    PMP_STAMP(dbg, 5);

    // ---- phase 5: trace, new state, next iteration's nodes -----------------------------------------------------------
    const float n0 = __ldcg(a.props + 3 * next), n1 = __ldcg(a.props + 3 * next + 1), n2 = __ldcg(a.props + 3 * next + 2);
    if (!a.advance) { if (tid == 0) a.cnt->last_next = next; return; }
    const bool rec = row < a.trace.capacity;
    if (rec) {
        if (a.trace.what & PMP_TRACE_DRAWS) for (int t = tid; t < P; t += ACCEPT_THREADS) a.trace.draws[row * P + t] = t < n_draws ? sdraw[t] : -1;
        if (a.trace.what & PMP_TRACE_LOGW) for (int t = tid; t < P; t += ACCEPT_THREADS) a.trace.logw[row * P + t] = __ldcg(a.logw + t);
        if (a.trace.what & PMP_TRACE_SAMPLES)
            for (int g = tid; g < P * 3; g += ACCEPT_THREADS) { int t = g / 3, j = g - 3 * t; a.trace.samples[row * P * 3 + g] = __ldcg(a.props + 3 * (t < n_draws ? sdraw[t] : next) + j); }
    }
    if (tid == 0) {
        a.state[0] = n0; a.state[1] = n1; a.state[2] = n2;
        if (rec && (a.trace.what & PMP_TRACE_STATE)) { a.trace.state[row * 3] = n0; a.trace.state[row * 3 + 1] = n1; a.trace.state[row * 3 + 2] = n2; }
        if (rec && (a.trace.what & PMP_TRACE_NEXT)) a.trace.next[row] = next;
        if (rec) a.cnt->trace_rows = row + 1;
        a.cnt->iteration = iter + 1;
        a.cnt->last_next = next;
    }
    if (fa.make_next) {
        __syncthreads();   // all reads of this iteration's nodes are done
        const float* __restrict__ zn = fa.z + ((iter + 1) & 1) * (long long)(P * 3);
        float* props_out = const_cast<float*>(a.props);
        for (int g = tid; g < P * 3; g += ACCEPT_THREADS) {
            int node = g / 3, j = g - 3 * node;
            props_out[g] = proposal_value_z(fa.gen, zn, node, j, j == 0 ? n0 : (j == 1 ? n1 : n2));
        }
    }
    PMP_STAMP(dbg, 6);
}

template <int ALGO>
__global__ void __launch_bounds__(ACCEPT_THREADS, 1) accept_fast_kernel(const __grid_constant__ AcceptFastArgs fa) {
    extern __shared__ double accept_fast_sm[];
    accept_fast_body<ALGO>(fa, accept_fast_sm);
}

}  // namespace pmp
