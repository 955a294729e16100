// chain_persistent_tc.cuh — the device-resident chain as ONE cooperative kernel whose sweep runs on the tensor cores.
//
// Same loop structure and hand-offs as chain_persistent.cuh (CTAs 0..G-2 sweep, CTA G-1 runs the three-phase acceptance of
// accept_lean.cuh), same integer-exact sums, but the sweep CTAs use the formulation of sweep_linear_tc.cuh: the residual
// r[node][point] = y - b0 - b1 x comes out of ONE bf16 UMMA per (128-node tile, 64-point chunk) over exact 3-way bf16
// splits of the operands, and the CUDA cores only square and add it (1 FP32 lane-op per pair instead of 3).
//
// Structure (the variants that were measured and lost are listed in DESIGN.md 4.2):
//   * the 512 TMEM columns are EIGHT independent stages of 64; stage s owns the units v = s, s+8, ... of this CTA's walk;
//   * every stage has four dedicated reader warps, one per TMEM lane quarter (32 reader warps = the whole CTA): a reader
//     sleeps on the stage's mbarrier, reads its 32 lanes x 64 columns with two tcgen05.ld.x32, and only then folds and
//     converts — while the stage is already being refilled;
//   * there is no MMA warp: the LAST of a stage's four readers to finish its read (an acquire-release counter in shared
//     memory elects it) issues the stage's next UMMA itself, so a refill starts the moment the columns are free;
//   * the data operand slice and the unit table are staged into shared memory once per launch; per iteration a CTA only
//     rebuilds the 32-byte node operand rows (one node per thread), fills its slice of the next normals and flushes one
//     integer per node.
// Per-node integer sums are kept per CTA in shared memory as two 32-bit halves (native shared-memory atomics; the carry is
// propagated explicitly), then added to global memory once per node — all integer, so the result is bit-identical to the
// stand-alone tensor-core sweep for any CTA partition.
#pragma once
#include "accept_lean.cuh"
#include "chain_persistent.cuh"
#include "sweep_linear_tc.cuh"

namespace pmp {
namespace tc {

constexpr int PT_STAGES = 8;                       // 8 x 64 TMEM columns
constexpr int PT_THREADS = PT_STAGES * 4 * 32;     // 1024: one reader warp per (stage, lane quarter)
constexpr int PT_MAX_TILES = 8;                    // P <= 1024: one node per thread

struct PersistTcArgs {
    Args sw;                 // bimg, theta (= props), acc, cnt, nchunks, P, max_chunks, max_units, sat_limit, ...
    AcceptFastArgs fa;
    PersistSync* sync;
    int iters;
};

struct PtSmem {
    uint8_t* sA;             // [ntiles][TILE_BYTES]
    uint8_t* sB;             // [max_chunks][CHUNK_BYTES]
    double* sscl;            // [ntiles*128]
    uint32_t* slo;           // [ntiles*128] low halves of the per-node integer sums of this CTA
    uint32_t* shi;           // [ntiles*128] high halves
    uint32_t* utab;          // [max_units]
    uint64_t* full;          // [PT_STAGES]
    uint32_t* elect;         // [PT_STAGES] readers that have finished reading the stage's current unit (cumulative)
};
__host__ __device__ inline size_t pt_smem_bytes(int ntiles, int max_chunks, int max_units) {
    return (size_t)ntiles * TILE_BYTES + (size_t)max_chunks * CHUNK_BYTES + (size_t)ntiles * TILE_NODES * 16 + (size_t)((max_units + 3) & ~3) * 4 + PT_STAGES * 8 + PT_STAGES * 4;
}
__device__ __forceinline__ PtSmem pt_carve(uint8_t* base, int ntiles, int max_chunks, int max_units) {
    PtSmem s;
    s.sA = base; base += (size_t)ntiles * TILE_BYTES;
    s.sB = base; base += (size_t)max_chunks * CHUNK_BYTES;
    s.sscl = reinterpret_cast<double*>(base); base += (size_t)ntiles * TILE_NODES * 8;
    s.slo = reinterpret_cast<uint32_t*>(base); base += (size_t)ntiles * TILE_NODES * 4;
    s.shi = reinterpret_cast<uint32_t*>(base); base += (size_t)ntiles * TILE_NODES * 4;
    s.full = reinterpret_cast<uint64_t*>(base); base += PT_STAGES * 8;
    s.elect = reinterpret_cast<uint32_t*>(base); base += PT_STAGES * 4;
    s.utab = reinterpret_cast<uint32_t*>(base);
    return s;
}

__device__ __forceinline__ uint32_t atom_add_acq_rel_shared(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
// 64-bit integer add into (lo, hi) halves with native 32-bit shared-memory atomics; exact for any interleaving
__device__ __forceinline__ void smem_add64(uint32_t* lo, uint32_t* hi, unsigned long long v) {
    const uint32_t vl = (uint32_t)v, vh = (uint32_t)(v >> 32);
    const uint32_t old = atomicAdd(lo, vl);
    const uint32_t carry = (old + vl < old) ? 1u : 0u;
    if (vh + carry) atomicAdd(hi, vh + carry);
}

template <int ALGO>
__global__ void __launch_bounds__(PT_THREADS, 1) chain_persistent_tc_kernel(const __grid_constant__ PersistTcArgs pa) {
    extern __shared__ __align__(128) unsigned char dsm_pt[];
    const int tid = threadIdx.x;
    const int n_sweep = gridDim.x - 1;

    if ((int)blockIdx.x == n_sweep) {
        // ================= acceptance CTA: pre (during the sweep) → wait → crit → release → post (during the next sweep) =====
        __shared__ double red[4][32];
        __shared__ int s_pick;
        const LeanSmem ls = lean_carve(dsm_pt, pa.fa.base.P, ALGO);
        LeanRegs lr;
        for (int it = 0; it < pa.iters; ++it) {
            lean_pre<ALGO>(pa.fa, ls, lr, red, &s_pick, LEAN_Z_TABLE_CRIT);
            if (pa.fa.base.dbg && tid == 0) { pa.fa.base.dbg[32 + 8] = clock64(); pa.fa.base.dbg[32 + 24] = globaltimer_ns(); }
            if (tid == 0) spin_until_ge(&pa.sync->arrive, (unsigned)(it + 1) * (unsigned)n_sweep);
            __syncthreads();
            lean_crit<ALGO>(pa.fa, ls, lr, red, &s_pick, LEAN_Z_TABLE_CRIT);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release(&pa.sync->version, (unsigned)(it + 1));
            if (pa.fa.base.dbg && tid == 0) { pa.fa.base.dbg[32 + 9] = clock64(); pa.fa.base.dbg[32 + 25] = globaltimer_ns(); }
            lean_post<ALGO>(pa.fa, ls, lr);
            __threadfence();          // trace cursor and state are read back by the next pre / by the host
            __syncthreads();
        }
        return;
    }

    // ================= sweep CTAs =================
    const Args& a = pa.sw;
    __shared__ uint32_t tmem_slot;
    const int warp = tid >> 5, lane = tid & 31;
    const Range rg = make_range(a.nchunks, a.P, blockIdx.x, n_sweep);
    const PtSmem s = pt_carve(dsm_pt, rg.ntiles, a.max_chunks, a.max_units);
    const int nu = (int)rg.nu;
    const int P_pad = rg.ntiles * TILE_NODES;

    // ---- once per launch: data slice + unit table into shared memory, mbarriers, TMEM --------------------------------------
    if (nu > 0) {
        const long long bytes = (rg.c_hi - rg.c_lo + 1) * CHUNK_BYTES;
        const uint8_t* src = a.bimg + rg.c_lo * CHUNK_BYTES;
        for (long long o = (long long)tid * 16; o < bytes; o += (long long)PT_THREADS * 16) {
            unsigned d = smem_u32(s.sB + o);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + o) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        for (int v = tid; v < nu; v += PT_THREADS) {
            int t = 0; long long before = 0;
            for (;;) { const long long cnt = max(0ll, rg.cb(t) - rg.ca(t)); if (v < before + cnt) break; before += cnt; ++t; }
            s.utab[v] = (uint32_t)t | ((uint32_t)(rg.ca(t) + (v - before) - rg.c_lo) << 8);
        }
    }
    if (tid == 0) {
        for (int b = 0; b < PT_STAGES; ++b) { mbar_init(&s.full[b], 1); s.elect[b] = 0u; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    stage_wait();
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // this warp: stage st, TMEM lane quarter q
    const int st = warp >> 2, q = warp & 3;
    int my = q * 32 + lane, lane0 = lane == 0;
    asm volatile("" : "+r"(my), "+r"(lane0));
    const int n_mine = nu > st ? (nu - st + PT_STAGES - 1) / PT_STAGES : 0;       // units of this stage per sweep
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + st * 64;
    const uint64_t da0 = umma_desc(s.sA), db0 = umma_desc(s.sB);
    const uint32_t idesc = umma_idesc(CHUNK);
    uint32_t fills = 0;                                  // completed fills of this stage so far (mbarrier phase)
    unsigned long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;

    const unsigned long long iter0 = __ldcg(&a.cnt->iteration);      // read before the acceptance CTA can have advanced it (it waits for every sweep CTA first)
    for (int it = 0; it < pa.iters; ++it) {
        PMP_STAMP(dbg, 0);
        if (tid == 0 && it > 0) spin_until_ge(&pa.sync->version, (unsigned)it);
        __syncthreads();
        PMP_STAMP(dbg, 1);
        {   // side job: this CTA's slice of the NEXT iteration's normals (they depend on counters only); read by the acceptance CTA
            const int zcount = a.P * 3, per = (zcount + n_sweep - 1) / n_sweep;
            const unsigned long long iter = iter0 + (unsigned long long)it;
            for (int k = PT_THREADS - 1 - tid; k < per; k += PT_THREADS) {
                const int e = blockIdx.x * per + k;
                if (e < zcount) a.z[((iter + 1) & 1) * (long long)zcount + e] = (float)stream_step(a.gen.seed, iter + 1, (unsigned long long)e, a.gen.uniform);
            }
        }
        // ---- node operand row, fixed-point scale and zeroed integer sum of node `tid` ----------------------------------------
        if (tid < P_pad) {
            uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
            double scl = 0.0;
            if (tid < a.P) {
                const float b0 = __ldcg(a.theta + 3 * tid), b1 = __ldcg(a.theta + 3 * tid + 1), sg = __ldcg(a.theta + 3 * tid + 2);
                uint32_t h0, m0, l0, h1, m1, l1;
                split3(-b0, h0, m0, l0); split3(-b1, h1, m1, l1);
                // k: 0..2 one, 3 b0h, 4 b0m, 5 b0l, 6 b1h, 7 b1m | 8 b1h, 9 b1l, 10 b1m, 11 b1h, 12 b1l, 13 b1m, 14 b1l, 15 zero
                lo = make_uint4(BF16_ONE | (BF16_ONE << 16), BF16_ONE | (h0 << 16), m0 | (l0 << 16), h1 | (m1 << 16));
                hi = make_uint4(h1 | (l1 << 16), m1 | (h1 << 16), l1 | (m1 << 16), l1);
                scl = (double)(1 << FX_SHIFT) / ((double)sg * (double)sg);
            }
            uint8_t* tb = s.sA + (size_t)(tid >> 7) * TILE_BYTES;
            *reinterpret_cast<uint4*>(tb + canon_off(tid & 127, 0)) = lo;
            *reinterpret_cast<uint4*>(tb + canon_off(tid & 127, 1)) = hi;
            s.sscl[tid] = scl; s.slo[tid] = 0u; s.shi[tid] = 0u;
        }
        proxy_fence_async();          // generic-proxy writes → visible to the tensor core's async proxy
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        PMP_STAMP(dbg, 2);

        if (n_mine > 0) {
            if (q == 0 && lane0) {        // first fill of this stage in this sweep
                const uint32_t u = s.utab[st];
                umma(tmem_base + st * 64, da0 + (uint64_t)(u & 0xff) * (TILE_BYTES >> 4), db0 + (uint64_t)(u >> 8) * (CHUNK_BYTES >> 4), idesc);
                umma_commit(&s.full[st]);
            }
            int cur_tile = -1;
            unsigned long long accq = 0ull;
            double scl = 0.0;
            for (int i = 0; i < n_mine; ++i) {
                uint32_t v[32];
                unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
                mbar_wait(&s.full[st], fills & 1);
                ++fills;
                tc_fence_after();
                tmem_ld32_nowait(taddr, v); tmem_ld_wait();
                sq32(v, a0, a1, a2, a3);
                tmem_ld32_nowait(taddr + 32, v); tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane0) {              // the last of the stage's four readers refills it
                    const uint32_t old = atom_add_acq_rel_shared(&s.elect[st], 1u);
                    if ((old & 3u) == 3u && i + 1 < n_mine) {
                        tc_fence_after();
                        const uint32_t u = s.utab[PT_STAGES * (i + 1) + st];
                        umma(tmem_base + st * 64, da0 + (uint64_t)(u & 0xff) * (TILE_BYTES >> 4), db0 + (uint64_t)(u >> 8) * (CHUNK_BYTES >> 4), idesc);
                        umma_commit(&s.full[st]);
                    }
                }
                __syncwarp();
                sq32(v, a0, a1, a2, a3);
                const float part = fold(a0, a1, a2, a3);
                const int tile = (int)(s.utab[PT_STAGES * i + st] & 0xff);
                if (tile != cur_tile) {
                    if (cur_tile >= 0 && accq) smem_add64(&s.slo[cur_tile * TILE_NODES + my], &s.shi[cur_tile * TILE_NODES + my], accq);
                    cur_tile = tile; accq = 0ull;
                    scl = s.sscl[tile * TILE_NODES + my];
                }
                accq += (unsigned long long)__double2ll_rn(fmin((double)part * scl, a.sat_limit));     // NaN / hopeless nodes saturate
            }
            if (cur_tile >= 0 && accq) smem_add64(&s.slo[cur_tile * TILE_NODES + my], &s.shi[cur_tile * TILE_NODES + my], accq);
        }
        PMP_STAMP(dbg, 3);
        tc_fence_before();
        __syncthreads();
        // ---- flush: one global integer add per node this CTA touched ------------------------------------------------------
        if (tid < a.P) {
            const unsigned long long sum = ((unsigned long long)s.shi[tid] << 32) | s.slo[tid];
            if (sum) atomicAdd(a.acc + tid, sum);
        }
        PMP_STAMP(dbg, 4);
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(&pa.sync->arrive, 1u);
        PMP_STAMP(dbg, 5);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

}  // namespace tc
}  // namespace pmp
