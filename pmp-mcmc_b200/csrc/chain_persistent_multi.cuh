// chain_persistent_multi.cuh — K independent chains co-scheduled in ONE cooperative kernel (single GPU, linear-Gaussian target).
//
// Why.  A single chain is a strict dependency loop: sweep (all SMs) → acceptance (one SM, ~3 us critical path + two L2
// hand-offs) → next sweep.  In chain_persistent_kernel the 147 sweep SMs idle for ~8 of every ~19.7 us while the acceptance of
// THEIR OWN chain runs; nothing of that chain can be started earlier.  The reference's experiments are run as several
// independent chains (20 repeats in error.py:191-213, one process per GPU in the ESS runs, 4 chains for any R-hat); with K >= 2
// chains resident the sweep SMs simply work on chain B while chain A is being accepted:
//     sweep SMs, warps  0-15 :  | sweep A_i  ....... | wait, nodes | sweep A_i+1 ....... |
//     sweep SMs, warps 16-31 :  ..... | flush | wait, nodes | sweep B_i ....... | flush |
//     acceptance CTAs        :  one per chain: pre during the sweep, crit when the sums are in, post after the release
// Every chain keeps its own pmp_ctx (state, Philox key, nodes, integer sums, counters, trace ring), executes exactly the
// arithmetic of chain_persistent_kernel in exactly the same order, and therefore produces bit-identical traces to the same
// chain run alone (tests/test_gpu_multichain.py).  The data slice of a sweep CTA is staged into shared memory once and is
// shared by all chains (pmp_share_data makes the contexts alias one device copy of the data).
#pragma once
#include "chain_persistent.cuh"

namespace pmp {

constexpr int PERSIST_MAX_CHAINS = 32;     // chains per launch (kernel parameters: ~0.6 KB per chain, 32 KB limit)
constexpr int PERSIST_MAX_ACCEPT = 24;     // acceptance CTAs per launch (the rest of the grid sweeps)

constexpr int PEER_MAX_WORLD = 8;

// Cross-GPU exchange of one chain's per-node integer sums over NVLink peer memory (world_size > 1; data rows sharded).
// Every rank owns a buffer  slots[2][world][MAX_NODES] x 16 bytes  that its peers have mapped with CUDA IPC.  The exchange is
// flag-in-data (the scheme of NCCL's LL protocol): a 64-bit sum travels as two 8-byte words {low 32 bits | tag << 32},
// {high 32 bits | tag << 32}, tag = exchange count + 1.  8-byte stores are single-copy atomic, so a word whose tag matches is
// complete by itself: no fence, no separate flag, no second NVLink trip.  Exchange e (a launch-independent, monotone
// count): rank `me` stores its P tagged sums into slot [e & 1][me] of EVERY peer's buffer, then every thread polls its own
// nodes in slot [e & 1][r] of the local buffer until both tags read e + 1, adds the world partial vectors (integers: the
// total is bit-identical on every rank and for any world size) and goes on with the replicated acceptance.  Two slots
// suffice: a peer can only write exchange e + 2 after it has received this rank's e + 1 sums, which are sent after the e
// sums have been read; the tag of the slot's previous use (e - 1) never equals e + 1.
struct PeerXchg {
    unsigned long long* local;                 // this rank's buffer
    unsigned long long* peer[PEER_MAX_WORLD];  // peer[r]: rank r's buffer mapped here (peer[me] = local)
    int world, me;
    unsigned long long base;                   // exchange count of this chain before this launch
};
__host__ __device__ inline size_t peer_xchg_words() { return (size_t)2 * PEER_MAX_WORLD * MAX_NODES * 2; }

__device__ __forceinline__ void st_relaxed_sys_v2(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_relaxed_sys_v2(const unsigned long long* p, unsigned long long& a, unsigned long long& b) {
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

// Executed by the whole acceptance CTA once the local sweep CTAs' sums are in: q[k] (this rank's sum of node tid + k * ACCEPT_THREADS)
// := sum over ranks.  acc != nullptr (counter-based hand-off): the local sums are read from / the totals written back to acc.
__device__ __forceinline__ void peer_allreduce_acc(const PeerXchg& x, unsigned long long* acc, int P, int it, unsigned long long* qreg = nullptr) {
    const int tid = threadIdx.x;
    const unsigned long long e = x.base + (unsigned long long)it;
    const unsigned long long tag = ((e + 1ull) & 0xffffffffull) << 32;
    const size_t slot = (size_t)(e & 1ull) * PEER_MAX_WORLD * MAX_NODES * 2;
    unsigned long long q[LEAN_K];
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) { const int p = tid + k * ACCEPT_THREADS; q[k] = (p < P) ? (qreg ? qreg[k] : __ldcg(acc + p)) : 0ull; }
    for (int r = 0; r < x.world; ++r) {
        if (r == x.me) continue;
        unsigned long long* dst = x.peer[r] + slot + (size_t)x.me * MAX_NODES * 2;
#pragma unroll
        for (int k = 0; k < LEAN_K; ++k) {
            const int p = tid + k * ACCEPT_THREADS;
            if (p < P) st_relaxed_sys_v2(dst + 2 * p, (q[k] & 0xffffffffull) | tag, (q[k] >> 32) | tag);
        }
    }
    const unsigned long long t0 = globaltimer_ns();
#pragma unroll
    for (int k = 0; k < LEAN_K; ++k) {
        const int p = tid + k * ACCEPT_THREADS;
        if (p < P) {
            unsigned long long sum = q[k];
            for (int r = 0; r < x.world; ++r) {
                if (r == x.me) continue;
                const unsigned long long* src = x.local + slot + (size_t)r * MAX_NODES * 2 + 2 * p;
                unsigned long long lo, hi;
                unsigned spins = 0;
                for (;;) {
                    ld_relaxed_sys_v2(src, lo, hi);
                    if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                    if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) __trap();     // a peer that never shows up must not hang the GPU
                }
                sum += (lo & 0xffffffffull) | (hi << 32);
            }
            if (qreg) qreg[k] = sum; else acc[p] = sum;       // read back by the same thread in lean_crit
        }
    }
}

struct PersistChain {
    SweepArgs sw;
    AcceptFastArgs fa;
    PersistSync* sync;
    PeerXchg xchg;             // world == 1: unused
    Handoff hs;                // HS kernels: this chain's flag-in-data hand-off buffers (accept_lean.cuh)
};

struct PersistMultiArgs {
    PersistChain ch[PERSIST_MAX_CHAINS];
    int n_chains;
    int iters;
    int max_chunks;
    int n_accept;              // acceptance CTAs (default: one per chain)
    int derive;                // one chain, flat tree, HS: the acceptance publishes the accepted state only (Handoff::state; see chain_persistent_kernel).  Not used for
                               // co-scheduled chains: there the sweep SMs have no idle time in which the tile's normals would come for free
};

template <int NG>
__device__ __forceinline__ void group_sync(int group) { asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(PERSIST_THREADS / NG) : "memory"); }

// Roles.  CTAs [0, n_sweep) sweep; the last n_accept CTAs accept (acceptance CTA a serves chains a, a + n_accept, ...).
// A sweep CTA is split into NG (2 or 4) independent warp groups of 1024 / NG threads with private named barriers: group h serves
// chains h, h + NG, ...  While one group waits for its chain's acceptance, reads its nodes or flushes its sums — all L2 round
// trips — the other groups' packed-FMA loops have the SM's FP32 pipe to themselves, so the pipe only idles when ALL groups are
// between sweeps.  NG = 4 when at least four chains are resident: on a shard of the data (world_size > 1) the sweep itself is
// short and the per-iteration round trips dominate, so more of them must be in flight.  The per-node sums are integers, so
// splitting the chunk lanes NG ways changes no bit of the result.
template <int ALGO, int NG, bool HS>
__global__ void __launch_bounds__(PERSIST_THREADS, 1) chain_persistent_multi_kernel(const __grid_constant__ PersistMultiArgs pa) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int tid = threadIdx.x;
    const int K = pa.n_chains;
    const int n_accept = pa.n_accept;
    const int n_sweep = gridDim.x - n_accept;

    if ((int)blockIdx.x >= n_sweep) {
        // ================= acceptance CTAs: their chains in round-robin order, one three-phase acceptance each ============
        __shared__ double red[4][32];
        __shared__ int s_pick;
        const LeanSmem ls = lean_carve(dsm, pa.ch[0].fa.base.P, ALGO);
        LeanRegs lr;
        for (int it = 0; it < pa.iters; ++it) {
            for (int c = (int)blockIdx.x - n_sweep; c < K; c += n_accept) {
                const AcceptFastArgs& fa = pa.ch[c].fa;
                const int zm = (HS && pa.derive) ? LEAN_Z_DERIVE : LEAN_Z_TABLE_CRIT;
                lean_pre<ALGO>(fa, ls, lr, red, &s_pick, zm, &pa.ch[c].hs, (unsigned long long)(pa.ch[c].hs.epoch + (unsigned)it + 1u) << 32, it == 0);
                if (tid == 0) spin_until_ge(&pa.ch[c].sync->arrive, (unsigned)(it + 1) * (unsigned)n_sweep);
                __syncthreads();
                if (pa.ch[c].xchg.world > 1) peer_allreduce_acc(pa.ch[c].xchg, fa.base.acc, fa.base.P, it);
                if (HS) {
                    const Handoff& hs = pa.ch[c].hs;
                    lean_crit<ALGO>(fa, ls, lr, red, &s_pick, zm, nullptr, &hs, (unsigned long long)(hs.epoch + (unsigned)it + 1u) << 32);
                    __syncthreads();
                } else {
                    lean_crit<ALGO>(fa, ls, lr, red, &s_pick, LEAN_Z_TABLE_CRIT);
                    __threadfence();
                    __syncthreads();
                    if (tid == 0) st_release(&pa.ch[c].sync->version, (unsigned)(it + 1));
                }
                lean_post<ALGO>(fa, ls, lr);
                if (!(HS && pa.derive)) __threadfence();          // trace cursor, state and the plain copy of the nodes are read back by the next pre of this chain (derive: nothing is)
                __syncthreads();
            }
        }
        if (HS && pa.derive && pa.iters > 0 && (int)blockIdx.x == n_sweep) {      // (derive: one chain) the nodes of the iteration after the last one
            lean_derive_props(pa.ch[0].fa, ls, lr, pa.ch[0].hs, lr.iter + 1, (unsigned long long)(pa.ch[0].hs.epoch + (unsigned)pa.iters + 1u) << 32);
            __threadfence();
        }
        return;
    }

    // ================= sweep CTAs =================
    constexpr int R = PERSIST_R, TP = PERSIST_TP, PT = PERSIST_PT;
    constexpr int HT = PERSIST_THREADS / NG, TDH = HT / TP;                                 // threads / chunk lanes of one warp group
    const SweepArgs& a0 = pa.ch[0].sw;                                                      // data, shapes: the same for every chain
    float* tile = reinterpret_cast<float*>(dsm);                                           // [max_chunks][CHUNK_STRIDE], read-only after staging
    const int half = tid / HT, htid = tid - half * HT;
    unsigned long long* sred = reinterpret_cast<unsigned long long*>(tile + (size_t)pa.max_chunks * CHUNK_STRIDE) + (size_t)half * TDH * PT;   // [TDH][PT] per half
    __shared__ float sprops_all[NG][PT * 3];
    __shared__ double sscl_all[NG][PT];
    __shared__ float saz_all[NG][PT * 3];                        // derive: alpha * z of this CTA's tile for the coming iteration
    __shared__ float s_state_all[NG][4];
    __shared__ unsigned long long s_iter0[PERSIST_MAX_CHAINS];   // Philox iteration of every chain at launch, read before any acceptance can advance it
    if (tid < K) s_iter0[tid] = __ldcg(&pa.ch[tid].sw.cnt->iteration);
    float* sprops = sprops_all[half];
    double* sscl = sscl_all[half];
    float* saz = saz_all[half];
    float* s_state = s_state_all[half];

    const int tp = htid & (TP - 1), td = htid / TP;
    const int P = a0.P;
    const long long nchunks = a0.nchunks, n_local = a0.n_local;
    const int ntiles = (P + PT - 1) / PT;

    // ---- stage this CTA's data slice once (all 1024 threads): segment s covers chunks [c0, c1) of node tile ptile (sweep_partition) --
    SweepRange rg;
    sweep_partition((int)blockIdx.x, n_sweep, ntiles, nchunks, rg);
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
                const float* src = (isy ? a0.y : a0.x) + (g < n_local ? g : 0);
                cp_async16(tile + (size_t)(slot + c) * CHUNK_STRIDE + (isy ? CHUNK : 0) + 4 * kk, src, g < n_local ? 16 : 0);
            }
            slot += (int)(c1 - c0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }

    unsigned sat_mask = 0;

    for (int it = 0; it < pa.iters; ++it) {
#pragma unroll 1
        for (int c = half; c < K; c += NG) {                     // this group's chains: group, group + NG, ...
            const SweepArgs& a = pa.ch[c].sw;
            const Handoff& hs = pa.ch[c].hs;
            const unsigned long long tag = (unsigned long long)(hs.epoch + (unsigned)it + 1u) << 32;
            if (!HS) {
                if (htid == 0 && it > 0) spin_until_ge(&pa.ch[c].sync->version, (unsigned)it);
                group_sync<NG>(half);
            }
            if (HS && pa.derive) {          // as in chain_persistent_kernel: wait for the 24-byte state; the tile's normals were computed after the previous arrival
                if (it > 0) {
                    if (htid == 0) {
                        unsigned long long w0, w1, w2;
                        SpinGuard sg;
                        for (;;) { ld_relaxed_gpu_v2(hs.state, w0, w1); w2 = ld_relaxed_gpu_u64(hs.state + 2); if (hs_tag_ok(w0, tag) && hs_tag_ok(w1, tag) && hs_tag_ok(w2, tag)) break; sg.tick(); }
                        s_state[0] = __uint_as_float((unsigned)w0); s_state[1] = __uint_as_float((unsigned)w1); s_state[2] = __uint_as_float((unsigned)w2);
                    }
                    group_sync<NG>(half);
                }
            } else {   // side job: this CTA's slice of the chain's NEXT-iteration normals (they depend on counters only)
                const int zcount = P * 3, per = (zcount + n_sweep - 1) / n_sweep;
                const unsigned long long iter = s_iter0[c] + (unsigned long long)it;
                for (int k = HT - 1 - htid; k < per; k += HT) {
                    const int e = blockIdx.x * per + k;
                    if (e < zcount) {
                        const float zv = (float)stream_step(a.gen.seed, iter + 1, (unsigned long long)e, a.gen.uniform);
                        if (HS) st_relaxed_gpu_u64(hs.zt + ((iter + 1) & 1) * (long long)zcount + e, (unsigned long long)__float_as_uint(zv) | tag);
                        else a.z[((iter + 1) & 1) * (long long)zcount + e] = zv;
                    }
                }
            }
            if (HS && !pa.derive && it > 0) {     // one thread per group waits for the chain's nodes (cheap hint, back-off); CTAs without units wait too
                if (htid == 0) wait_nodes_hint(hs, P, tag);
                group_sync<NG>(half);
            }
            bool sat = false;
            for (int s = 0; s < nseg; ++s) {
                const int node_base = seg_tile[s] * PT;
                if (HS) {
                    for (int i = htid; i < PT; i += HT) {
                        float v0, v1, v2;
                        if (pa.derive && it > 0) {
                            const int node = node_base + i;
                            v0 = s_state[0]; v1 = s_state[1]; v2 = s_state[2];
                            if (node >= P) { v0 = 0.f; v1 = 0.f; v2 = 0.f; }
                            else if (node > 0) { v0 = __fadd_rn(v0, saz[3 * i]); v1 = __fadd_rn(v1, saz[3 * i + 1]); v2 = __fadd_rn(v2, saz[3 * i + 2]); }
                        } else
                        fetch_node(hs, a.theta, node_base + i, P, it == 0, tag, v0, v1, v2);
                        sprops[3 * i] = v0; sprops[3 * i + 1] = v1; sprops[3 * i + 2] = v2;
                        sscl[i] = (node_base + i < P) ? (double)(1 << FX_SHIFT) / ((double)v2 * (double)v2) : 0.0;
                    }
                } else {
                    for (int i = htid; i < PT * 3; i += HT) {
                        int node = node_base + i / 3, j = i - (i / 3) * 3;
                        float v = (node < P) ? __ldcg(a.theta + (long long)node * 3 + j) : 0.f;
                        sprops[i] = v;
                        if (j == 2) sscl[i / 3] = (node < P) ? (double)(1 << FX_SHIFT) / ((double)v * (double)v) : 0.0;
                    }
                }
                group_sync<NG>(half);
                float b0[R], b1[R]; double scl[R]; unsigned long long accq[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { int i = tp * R + r; b0[r] = sprops[3 * i]; b1[r] = sprops[3 * i + 1]; scl[r] = sscl[i]; accq[r] = 0ull; }
                const int nct = (int)(seg_c1[s] - seg_c0[s]);
                for (int cc = td; cc < nct; cc += TDH) {
                    int cnt = (int)min((long long)CHUNK, n_local - (seg_c0[s] + cc) * CHUNK);
                    float part[R];
                    const float* sx = tile + (size_t)(seg_slot[s] + cc) * CHUNK_STRIDE;
                    chunk_sumsq<R, true>(sx, sx + CHUNK, cnt, b0, b1, part);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double dq = (double)part[r] * scl[r];
                        if (!(dq < a.sat_limit)) { dq = a.sat_limit; sat = true; }
                        accq[r] += (unsigned long long)__double2ll_rn(dq);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) sred[td * PT + tp * R + r] = accq[r];
                group_sync<NG>(half);
                for (int i = htid; i < PT; i += HT) {
                    unsigned long long sum = 0ull;
                    for (int k = 0; k < TDH; ++k) sum += sred[k * PT + i];
                    if (node_base + i < P && sum) atomicAdd(a.acc + node_base + i, sum);
                }
                group_sync<NG>(half);
            }
            if (sat) sat_mask |= 1u << c;
            __threadfence();
            group_sync<NG>(half);
            if (htid == 0) atomicAdd(&pa.ch[c].sync->arrive, 1u);
            if (HS && pa.derive && nseg == 1) {      // while the acceptance (and the cross-GPU exchange) runs: the normals of this tile's nodes of the next iteration
                const unsigned long long iter = s_iter0[c] + (unsigned long long)it + 1;
                const bool publish = seg_c0[0] == 0;
                const int zcount = P * 3;
                for (int i = htid; i < PT * 3; i += HT) {
                    const int e = seg_tile[0] * PT * 3 + i;
                    if (e < zcount) {
                        const float zv = (float)stream_step(a.gen.seed, iter, (unsigned long long)e, a.gen.uniform);
                        saz[i] = __fmul_rn(a.gen.alpha, zv);
                        if (publish) st_relaxed_gpu_u64(hs.zt + (iter & 1) * (long long)zcount + e, (unsigned long long)__float_as_uint(zv) | (tag + (1ull << 32)));
                    }
                }
            }
        }
    }
    for (int c = 0; c < K; ++c) if (sat_mask & (1u << c)) atomicOr(&pa.ch[c].sw.cnt->flags, 1);
}

}  // namespace pmp
