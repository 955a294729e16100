// chain_persistent.cuh — the whole device-resident chain as ONE cooperative kernel (single GPU, linear-Gaussian target).
//
// The stepwise loop (sweep kernel → acceptance kernel, replayed from a CUDA graph) pays per iteration for two launches,
// for re-staging the same 0.8 MB of data into shared memory, for every CTA's prologue and for a cold instruction cache
// in the one-CTA acceptance kernel.  None of that is work the algorithm asks for: the data never change, the node tile
// of a CTA never changes, and the acceptance code is the same every iteration.  Here:
//   * grid = one CTA of 1024 threads per SM (cooperative launch: co-residency is guaranteed or the launch fails);
//   * CTAs 0..G-2 are sweep CTAs: each owns a fixed contiguous range of (128-node tile, 64-point chunk) units, stages
//     its data slice into shared memory ONCE, and per iteration only re-reads its tile's nodes (1.5 KB), runs the
//     packed-FMA inner loop and flushes integer partial sums (same arithmetic, same bits as sweep_linear_kernel);
//   * CTA G-1 is the acceptance CTA: it waits for all sweep CTAs of the iteration, runs accept_fast_body (hot in its
//     SM's instruction cache), publishes the next nodes and releases the next iteration;
//   * the two hand-offs per iteration are release/acquire counters in global memory (bounded spins: a protocol bug
//     traps instead of hanging the GPU).
#pragma once
#include "accept_lean.cuh"
#include "sweep_linear.cuh"

namespace pmp {

constexpr int PERSIST_THREADS = 1024;
constexpr int PERSIST_MAX_SEGS = 3;      // (node tile, chunk range) segments a sweep CTA keeps; the host checks the plan, the kernels trap beyond it
constexpr int PERSIST_TP = 32, PERSIST_TD = PERSIST_THREADS / PERSIST_TP, PERSIST_R = 4, PERSIST_PT = PERSIST_TP * PERSIST_R;

struct PersistSync {
    unsigned int arrive;     // += 1 per sweep CTA per iteration
    unsigned int version;    // iterations completed by the acceptance CTA in this launch
};

struct PersistArgs {
    SweepArgs sw;            // x, y, theta (= props), acc, cnt, n_local, nchunks, P, ... (TP/TD fields unused)
    AcceptFastArgs fa;
    PersistSync* sync;
    int iters;
    int max_chunks;          // chunks per sweep CTA that fit the dynamic shared memory
    Handoff hs;              // HS kernels: flag-in-data hand-offs (accept_lean.cuh)
    int derive;              // HS, flat or binary tree, one segment per sweep CTA: the acceptance publishes the accepted state only and every reader derives its nodes (Handoff::state)
};

// A sweep CTA's thread i < PERSIST_PT fetches node (node_base + i) of iteration `it`: plain memory for the first iteration of a launch
// (written by propose_kernel before the launch), the tagged words the acceptance CTA publishes afterwards.
__device__ __forceinline__ void fetch_node(const Handoff& hs, const float* theta, int node, int P, bool first, unsigned long long tag, float& b0, float& b1, float& sg_) {
    if (node >= P) { b0 = 0.f; b1 = 0.f; sg_ = 0.f; return; }
    if (first) { b0 = __ldcg(theta + 3ll * node); b1 = __ldcg(theta + 3ll * node + 1); sg_ = __ldcg(theta + 3ll * node + 2); return; }
    const unsigned long long* w = hs.nodes + 4ll * node;
    unsigned long long w0, w1, w2;
    SpinGuard sg;
    for (;;) { ld_relaxed_gpu_v2(w, w0, w1); w2 = ld_relaxed_gpu_u64(w + 2); if (hs_tag_ok(w0, tag) && hs_tag_ok(w1, tag) && hs_tag_ok(w2, tag)) break; sg.tick(); }
    b0 = __uint_as_float((unsigned)w0); b1 = __uint_as_float((unsigned)w1); sg_ = __uint_as_float((unsigned)w2);
}

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void spin_until_ge(const unsigned* p, unsigned target) {
    unsigned spins = 0;
    while ((int)(ld_acquire(p) - target) < 0) { if (++spins > 400000000u) __trap(); }
}

// Units of a sweep CTA.  With at least as many sweep CTAs as node tiles the CTAs are dealt to the tiles — tile t gets CTAs
// [t * n_sweep / ntiles, (t + 1) * n_sweep / ntiles) — and split the tile's chunks evenly: ONE segment per CTA.  (Contiguous ranges of the
// tile-major unit list made the 7 CTAs that straddle a tile boundary pay a second node fetch, a second flush and badly quantised chunk
// rounds, and every iteration of the chain waited ~2 us for those stragglers; measured with PMP_DEBUG_STAMPS.)  The per-node sums are
// integers, so the partition changes no bit.  Fewer CTAs than tiles: contiguous unit ranges, up to PERSIST_MAX_SEGS segments.
struct SweepRange { int nseg; int tile[PERSIST_MAX_SEGS]; long long c0[PERSIST_MAX_SEGS], c1[PERSIST_MAX_SEGS]; };
__host__ __device__ inline long long persist_max_chunks(int ntiles, long long nchunks, int n_sweep) {
    if (n_sweep >= ntiles) { const int per_tile = n_sweep / ntiles; return (nchunks + per_tile - 1) / per_tile + 1; }
    return ((long long)ntiles * nchunks + n_sweep - 1) / n_sweep + 1;
}
__device__ __forceinline__ void sweep_partition(int cta, int n_sweep, int ntiles, long long nchunks, SweepRange& r) {
    r.nseg = 0;
    if (nchunks == 0) return;
    if (n_sweep >= ntiles) {
        int t = (int)(((long long)cta * ntiles) / n_sweep);
        while ((long long)(t + 1) * n_sweep / ntiles <= cta) ++t;       // tile whose CTA range [t n / T, (t+1) n / T) holds this CTA
        while ((long long)t * n_sweep / ntiles > cta) --t;
        const int first = (int)((long long)t * n_sweep / ntiles), cnt = (int)((long long)(t + 1) * n_sweep / ntiles) - first, k = cta - first;
        const long long c0 = (long long)k * nchunks / cnt, c1 = (long long)(k + 1) * nchunks / cnt;
        if (c1 > c0) { r.tile[0] = t; r.c0[0] = c0; r.c1[0] = c1; r.nseg = 1; }
        return;
    }
    const long long units = (long long)ntiles * nchunks;
    long long u = (long long)cta * units / n_sweep;
    const long long u_end = (long long)(cta + 1) * units / n_sweep;
    while (u < u_end && r.nseg < PERSIST_MAX_SEGS) {
        const int ptile = (int)(u / nchunks);
        const long long c0 = u - (long long)ptile * nchunks, c1 = min(nchunks, c0 + (u_end - u));
        r.tile[r.nseg] = ptile; r.c0[r.nseg] = c0; r.c1[r.nseg] = c1;
        u += c1 - c0; ++r.nseg;
    }
    if (u < u_end) __trap();               // units would be dropped silently: the host-side plan (persist_segments_fit) must have refused this shape
}

// One thread of a sweep CTA (or warp group) waits for the nodes of iteration `it` with a cheap hint — the tag of the last word the
// acceptance CTA writes, polled with back-off — before the 128 fetch_node threads verify their own words: 128 threads per CTA polling
// flat out (19 000 on the chip) took ~60 % of the L2 request rate and slowed the acceptance CTA's own loads and stores.
__device__ __forceinline__ void wait_nodes_hint(const Handoff& hs, int P, unsigned long long tag) {
    const unsigned long long* w = hs.nodes + 4ll * (P - 1) + 2;
    SpinGuard sg;
    while (!hs_tag_ok(ld_relaxed_gpu_u64(w), tag)) { __nanosleep(40); sg.tick(); }
}

template <int ALGO, bool HS>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) chain_persistent_kernel(const __grid_constant__ PersistArgs pa) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int tid = threadIdx.x;
    const int n_sweep = gridDim.x - 1;

    if ((int)blockIdx.x == n_sweep) {
        // ================= acceptance CTA: pre (during the sweep) → wait → crit → release → post (during the next sweep) =====
        __shared__ double red[4][32];
        __shared__ int s_pick;
        const LeanSmem ls = lean_carve(dsm, pa.fa.base.P, ALGO);
        LeanRegs lr;
        const int zm = !HS ? LEAN_Z_TABLE_CRIT : (pa.derive ? LEAN_Z_DERIVE : LEAN_Z_HS_PRE);
        for (int it = 0; it < pa.iters; ++it) {
            lean_pre<ALGO>(pa.fa, ls, lr, red, &s_pick, zm, &pa.hs, (unsigned long long)(pa.hs.epoch + (unsigned)it + 1u) << 32, it == 0);
            if (pa.fa.base.dbg && tid == 0) { pa.fa.base.dbg[32 + 8] = clock64(); pa.fa.base.dbg[32 + 24] = globaltimer_ns(); }
            if (tid == 0) spin_until_ge(&pa.sync->arrive, (unsigned)(it + 1) * (unsigned)n_sweep);
            __syncthreads();
            if (HS) {       // the next nodes and normals travel as tagged words: no fence, no version counter
                lean_crit<ALGO>(pa.fa, ls, lr, red, &s_pick, zm, nullptr, &pa.hs, (unsigned long long)(pa.hs.epoch + (unsigned)it + 1u) << 32);
                __syncthreads();
            } else {
                lean_crit<ALGO>(pa.fa, ls, lr, red, &s_pick, LEAN_Z_TABLE_CRIT);
                __threadfence();
                __syncthreads();
                if (tid == 0) st_release(&pa.sync->version, (unsigned)(it + 1));
            }
            if (pa.fa.base.dbg && tid == 0) { pa.fa.base.dbg[32 + 9] = clock64(); pa.fa.base.dbg[32 + 25] = globaltimer_ns(); }
            lean_post<ALGO>(pa.fa, ls, lr);
            if (!(HS && pa.derive)) __threadfence();          // trace cursor, state and the plain copy of the nodes are read back by the next pre (derive: nothing is; the host reads after the kernel)
            __syncthreads();
        }
        if (HS && pa.derive && pa.iters > 0) {      // the nodes of the iteration after the last one: the plain copy the host and the next launch read
            lean_derive_props(pa.fa, ls, lr, pa.hs, lr.iter + 1, (unsigned long long)(pa.hs.epoch + (unsigned)pa.iters + 1u) << 32);
            __threadfence();
        }
        return;
    }

    // ================= sweep CTAs =================
    constexpr int R = PERSIST_R, TP = PERSIST_TP, TD = PERSIST_TD, PT = PERSIST_PT;
    const SweepArgs& a = pa.sw;
    float* tile = reinterpret_cast<float*>(dsm);                                           // [max_chunks][CHUNK_STRIDE]
    unsigned long long* sred = reinterpret_cast<unsigned long long*>(tile + (size_t)pa.max_chunks * CHUNK_STRIDE);   // [TD][PT]
    __shared__ float sprops[PT * 3];
    __shared__ double sscl[PT];
    // derive: alpha * z of the coming iteration for the tiles this CTA's nodes depend on.  Flat tree: its own tile.  Binary tree: node 128 t + i is the state plus the
    // increments of its ancestors in creation order — the low-bit prefixes of i (all in tile 0), then element i of the tiles whose index is a low-bit prefix of t, t
    // itself last (for_each_ancestor) — so the list is {0, prefixes of t .. t}: at most 1 + popcount(t) <= 5 tiles for P <= 2048.
    constexpr int DERIVE_TILES = 5;
    __shared__ float saz[DERIVE_TILES * PT * 3];
    __shared__ float s_state[4];

    const int tp = tid & (TP - 1), td = tid / TP;
    const int ntiles = (a.P + PT - 1) / PT;

    // ---- stage this CTA's data slice once: segment s covers chunks [c0, c1) of node tile ptile -----------------
    SweepRange rg;
    sweep_partition((int)blockIdx.x, n_sweep, ntiles, a.nchunks, rg);
    const int nseg = rg.nseg;
    int seg_tile[PERSIST_MAX_SEGS]; long long seg_c0[PERSIST_MAX_SEGS], seg_c1[PERSIST_MAX_SEGS]; int seg_slot[PERSIST_MAX_SEGS];
    {
        int slot = 0;
        for (int sg = 0; sg < nseg; ++sg) {
            const long long c0 = rg.c0[sg], c1 = rg.c1[sg];
            seg_tile[sg] = rg.tile[sg]; seg_c0[sg] = c0; seg_c1[sg] = c1; seg_slot[sg] = slot;
            for (long long i = tid; i < (c1 - c0) * (CHUNK / 2); i += PERSIST_THREADS) {
                int c = (int)(i / (CHUNK / 2)), k = (int)(i - (long long)c * (CHUNK / 2));
                bool isy = k >= CHUNK / 4; int kk = isy ? k - CHUNK / 4 : k;
                long long g = (c0 + c) * CHUNK + 4 * kk;
                const float* src = (isy ? a.y : a.x) + (g < a.n_local ? g : 0);
                cp_async16(tile + (size_t)(slot + c) * CHUNK_STRIDE + (isy ? CHUNK : 0) + 4 * kk, src, g < a.n_local ? 16 : 0);
            }
            slot += (int)(c1 - c0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }

    bool saturated = false;
    const unsigned long long iter0 = __ldcg(&a.cnt->iteration);      // read before the acceptance CTA can have advanced it (it waits for every sweep CTA first)
    for (int it = 0; it < pa.iters; ++it) {
        unsigned long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
        PMP_STAMP(dbg, 0);
        const unsigned long long tag = (unsigned long long)(pa.hs.epoch + (unsigned)it + 1u) << 32;      // HS: tag of this iteration's nodes, sums and normals
        if (!HS) {
            if (tid == 0 && it > 0) spin_until_ge(&pa.sync->version, (unsigned)it);
            __syncthreads();
        }
        PMP_STAMP(dbg, 1);
        if (HS && pa.derive) {
            // this tile's nodes = accepted state + alpha * z(it): the normals depend on counters only and were computed after the previous arrival (below),
            // while the acceptance was running; wait for the 24-byte state, then build the nodes from shared memory
            if (it > 0) {
                if (tid == 0) {
                    unsigned long long w0, w1, w2;
                    SpinGuard sg;
                    for (;;) { ld_relaxed_gpu_v2(pa.hs.state, w0, w1); w2 = ld_relaxed_gpu_u64(pa.hs.state + 2); if (hs_tag_ok(w0, tag) && hs_tag_ok(w1, tag) && hs_tag_ok(w2, tag)) break; sg.tick(); }
                    s_state[0] = __uint_as_float((unsigned)w0); s_state[1] = __uint_as_float((unsigned)w1); s_state[2] = __uint_as_float((unsigned)w2);
                }
                __syncthreads();
            }
        } else {   // side job: this CTA's slice of the NEXT iteration's normals (they depend on counters only); read by the acceptance CTA
            const int zcount = a.P * 3, per = (zcount + n_sweep - 1) / n_sweep;
            const unsigned long long iter = iter0 + (unsigned long long)it;
            for (int k = PERSIST_THREADS - 1 - tid; k < per; k += PERSIST_THREADS) {
                const int e = blockIdx.x * per + k;
                if (e < zcount) {
                    const float zv = (float)stream_step(a.gen.seed, iter + 1, (unsigned long long)e, a.gen.uniform);
                    if (HS) st_relaxed_gpu_u64(pa.hs.zt + ((iter + 1) & 1) * (long long)zcount + e, (unsigned long long)__float_as_uint(zv) | tag);
                    else a.z[((iter + 1) & 1) * (long long)zcount + e] = zv;
                }
            }
        }
        if (HS && !pa.derive && it > 0) {     // (a CTA without units waits too: its normals of it + 2 would overwrite a half still in use)
            if (tid == 0) wait_nodes_hint(pa.hs, a.P, tag);
            __syncthreads();
        }
        for (int s = 0; s < nseg; ++s) {
            const int node_base = seg_tile[s] * PT;
            if (HS) {
                if (tid < PT) {
                    float v0, v1, v2;
                    if (pa.derive && it > 0) {
                        const int node = node_base + tid;
                        v0 = s_state[0]; v1 = s_state[1]; v2 = s_state[2];
                        if (node >= a.P) { v0 = 0.f; v1 = 0.f; v2 = 0.f; }
                        else if (a.gen.tree == PMP_TREE_FLAT) {
                            if (node > 0) { v0 = __fadd_rn(v0, saz[3 * tid]); v1 = __fadd_rn(v1, saz[3 * tid + 1]); v2 = __fadd_rn(v2, saz[3 * tid + 2]); }
                        } else {
                            const int t = seg_tile[s];
                            for (int l = 0; l < 7; ++l) if ((tid >> l) & 1) {                 // ancestors inside tile 0: the low-bit prefixes of i
                                const float* z3 = saz + 3 * (tid & ((2 << l) - 1));
                                v0 = __fadd_rn(v0, z3[0]); v1 = __fadd_rn(v1, z3[1]); v2 = __fadd_rn(v2, z3[2]);
                            }
                            int k = 1;
                            for (int l = 0; l < 4; ++l) if ((t >> l) & 1) {                   // element i of the prefix tiles of t, t itself last
                                const float* z3 = saz + (k * PT + tid) * 3;
                                v0 = __fadd_rn(v0, z3[0]); v1 = __fadd_rn(v1, z3[1]); v2 = __fadd_rn(v2, z3[2]);
                                ++k;
                            }
                        }
                    } else
                    fetch_node(pa.hs, a.theta, node_base + tid, a.P, it == 0, tag, v0, v1, v2);
                    sprops[3 * tid] = v0; sprops[3 * tid + 1] = v1; sprops[3 * tid + 2] = v2;
                    sscl[tid] = (node_base + tid < a.P) ? (double)(1 << FX_SHIFT) / ((double)v2 * (double)v2) : 0.0;
                }
            } else {
                for (int i = tid; i < PT * 3; i += PERSIST_THREADS) {
                    int node = node_base + i / 3, j = i - (i / 3) * 3;
                    float v = (node < a.P) ? __ldcg(a.theta + (long long)node * 3 + j) : 0.f;
                    sprops[i] = v;
                    if (j == 2) sscl[i / 3] = (node < a.P) ? (double)(1 << FX_SHIFT) / ((double)v * (double)v) : 0.0;
                }
            }
            __syncthreads();
            if (s == 0) PMP_STAMP(dbg, 2);
            float b0[R], b1[R]; double scl[R]; unsigned long long accq[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { int i = tp * R + r; b0[r] = sprops[3 * i]; b1[r] = sprops[3 * i + 1]; scl[r] = sscl[i]; accq[r] = 0ull; }
            const int nct = (int)(seg_c1[s] - seg_c0[s]);
            for (int c = td; c < nct; c += TD) {
                int cnt = (int)min((long long)CHUNK, a.n_local - (seg_c0[s] + c) * CHUNK);
                float part[R];
                const float* sx = tile + (size_t)(seg_slot[s] + c) * CHUNK_STRIDE;
                chunk_sumsq<R, true>(sx, sx + CHUNK, cnt, b0, b1, part);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double dq = (double)part[r] * scl[r];
                    if (!(dq < a.sat_limit)) { dq = a.sat_limit; saturated = true; }
                    accq[r] += (unsigned long long)__double2ll_rn(dq);
                }
            }
            if (s == 0) PMP_STAMP(dbg, 3);
#pragma unroll
            for (int r = 0; r < R; ++r) sred[td * PT + tp * R + r] = accq[r];
            __syncthreads();
            for (int i = tid; i < PT; i += PERSIST_THREADS) {
                unsigned long long sum = 0ull;
                for (int k = 0; k < TD; ++k) sum += sred[k * PT + i];
                if (node_base + i < a.P && sum) atomicAdd(a.acc + node_base + i, sum);
            }
            __syncthreads();
        }
        PMP_STAMP(dbg, 4);
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(&pa.sync->arrive, 1u);
        PMP_STAMP(dbg, 5);
        if (HS && pa.derive && nseg == 1) {
            // while the acceptance runs: the normals of the NEXT iteration's nodes of this tile (and, binary tree, of the tiles they descend from).  Every CTA computes
            // them for itself; the tile's first CTA also hands its own tile's to the acceptance CTA, which needs all nodes for its own pre phase — off the critical path
            const unsigned long long iter = iter0 + (unsigned long long)it + 1;
            const bool publish = seg_c0[0] == 0;
            const int zcount = a.P * 3, t = seg_tile[0];
            int tiles[DERIVE_TILES], nt = 0;
            if (a.gen.tree == PMP_TREE_FLAT) tiles[nt++] = t;
            else { tiles[nt++] = 0; for (int l = 0; l < 4; ++l) if ((t >> l) & 1) tiles[nt++] = t & ((2 << l) - 1); }
            for (int i = tid; i < nt * PT * 3; i += PERSIST_THREADS) {
                const int k = i / (PT * 3), u = tiles[k], e = u * PT * 3 + (i - k * PT * 3);
                if (e < zcount) {
                    const float zv = (float)stream_step(a.gen.seed, iter, (unsigned long long)e, a.gen.uniform);
                    saz[i] = __fmul_rn(a.gen.alpha, zv);
                    if (publish && u == t && (k == nt - 1)) st_relaxed_gpu_u64(pa.hs.zt + (iter & 1) * (long long)zcount + e, (unsigned long long)__float_as_uint(zv) | (tag + (1ull << 32)));
                }
            }
        }
    }
    if (saturated) atomicOr(&a.cnt->flags, 1);
}

}  // namespace pmp
