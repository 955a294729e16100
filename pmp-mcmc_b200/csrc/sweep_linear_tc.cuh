// sweep_linear_tc.cuh — the proposals x data sweep of the linear-Gaussian model on the 5th-generation tensor cores.
//
// Same job as sweep_linear.cuh (replaces log_likelihood_kernel's data loop, 500_MP.cu:16-21, and BayesNet.loglik,
// lb.py:103-108): per node p the sum over this rank's points of ((y_i - b0_p - b1_p x_i)/sigma_p)^2 as a 2^-FX_SHIFT
// fixed-point integer.  The residual is a contraction over K = 3 — r[p][i] = [1, x_i, y_i] . [-b0_p, -b1_p, 1] — so the
// CUDA cores only have to SQUARE AND ADD it (1 FMA lane-op per (node, point) instead of 3) if the tensor pipe produces r:
//
//   * every float32 operand is split exactly into three bf16 pieces (8+8+8 mantissa bits: v = h + m + l, no rounding
//     left), and the products are laid out along K = 16 of ONE bf16 UMMA (kind::f16, fp32 accumulation in TMEM):
//         k    0..2         3..5            6..14                              15
//         A    1 1 1        -b0{h,m,l}      -b1{h,m,h,l,m,h,l,m,l}             0        (node operand, M = 128 nodes)
//         B    y{h,m,l}     1 1 1           x {h,h,m,h,m,l,m,l,l}              0        (data operand, N = 64..256 points)
//     all nine cross terms of b1*x are present, so the only error left is the tensor core's fp32 accumulation
//     (measured: |err r| <= 4e-7, relative error of a 64-point sum of squares 1e-7 — scripts/micro/micro4.cu — the same
//     size as the float32 FMA path's own rounding);
//   * the data operand never changes: pmp_set_data_linear writes it once, chunk by chunk, as the exact shared-memory
//     image of the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices, K-adjacent cores 128 B apart,
//     row groups 256 B apart), so staging is a plain 16-byte-vector copy and consecutive chunks form one N <= 256 operand;
//   * the node operand (32 B per node) is rebuilt from the published nodes at the start of every sweep by all threads;
//   * the CTA runs four independent pipelines ("sets"): set j owns TMEM columns [128j, 128j+128) as two 64-column stages,
//     the units v = j, j+4, ... of the CTA's walk, one issuer warp (one N = 64 UMMA per unit, mbarrier full/empty hand-off)
//     and four epilogue warps, one per TMEM lane quarter, that read the residuals with tcgen05.ld and run
//     acc = fma(r, r, acc) as packed fma.rn.f32x2; one float32 partial per (node, chunk) in a fixed order, then the same
//     binary64 scale + round-to-integer + integer adds as the FMA path, so the sum is still bit-identical for any CTA
//     order, grid size, loop structure and number of GPUs.
// Units of work are (chunk, 128-node tile) pairs in chunk-major order cut into gridDim.x equal contiguous ranges; a CTA
// keeps its chunks resident in shared memory and walks them tile by tile.
// Bounds and what was measured: the epilogue's ceiling is TMEM read + FP32 issue (440 B/clk/SM with the FMA,
// scripts/micro/micro3.cu); in practice the per-unit instruction overhead and UMMA latency x TMEM capacity bind first.
#pragma once
#include <cuda_bf16.h>
#ifndef PMP_TC_ABL
#define PMP_TC_ABL 0
#endif

#include "accept.cuh"
#include "common.cuh"

namespace pmp {
namespace tc {

constexpr int TILE_NODES = 128;                   // UMMA M
constexpr int CHUNK_BYTES = CHUNK * 32;           // 64 points x 16 bf16
constexpr int TILE_BYTES = TILE_NODES * 32;       // 128 nodes x 16 bf16
constexpr int SETS = 4;                           // independent MMA → epilogue pipelines per CTA
constexpr int EPI_WARPS = 4 * SETS;               // one epilogue warp per (set, TMEM lane quarter)
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int STANDALONE_THREADS = (EPI_WARPS + SETS) * 32;
constexpr int NODES_PER_THREAD = 4;               // epilogue thread (j, q, lane) owns row q*32+lane of tiles j, j+4, ...
constexpr int TMEM_COLS = 512;
constexpr int MAX_TILES = SETS * 4;                // P <= 2048 (node operand 64 KB + integer scratch 64 KB)

struct Args {
    const uint8_t* __restrict__ bimg;     // [nchunks][CHUNK_BYTES] data operand image
    const float* __restrict__ theta;      // [P,3] nodes (b0, b1, sigma)
    unsigned long long* __restrict__ acc; // [P]
    DeviceCounters* cnt;
    long long nchunks;
    int P;
    int max_chunks;                       // chunks per CTA that fit the shared-memory data area
    int max_units;                        // (tile, chunk) units per CTA (unit table size)
    double sat_limit;
    int generate;                         // 1: also fill the next iteration's half of the normals table (side job)
    float* z;
    ProposeArgs gen;
    unsigned long long* dbg;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint expires)
// instead of spinning — spinning waiters would take issue slots from the epilogue warps that share their scheduler
#ifndef PMP_TC_WAIT_HINT_NS
#define PMP_TC_WAIT_HINT_NS 100000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(PMP_TC_WAIT_HINT_NS) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > 50000u) __trap();           // a protocol bug must fail loudly, never hang the GPU (50000 x 100 us)
}
__device__ __forceinline__ uint64_t umma_desc(const void* smem) {
    // K-major, no swizzle: LBO (K-adjacent core matrices) 128 B, SBO (8-row groups) 256 B, descriptor version 1 (sm_100)
    return (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t umma_idesc(int n_cols) {   // D f32, A/B bf16, both K-major, M = 128
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(TILE_NODES >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(0) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                 "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ unsigned long long sq_acc2(uint32_t lo, uint32_t hi, unsigned long long acc) {
    unsigned long long r = ((unsigned long long)hi << 32) | lo, d;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(d) : "l"(r), "l"(acc));
    return d;
}
__device__ __forceinline__ float pair_sum(unsigned long long a) { return __fadd_rn(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)(a >> 32))); }

// exact 3-way bf16 split of a float32 (round-to-nearest pieces; the remainders are exact float32 differences)
__device__ __forceinline__ void split3(float v, uint32_t& h, uint32_t& m, uint32_t& l) {
    __nv_bfloat16 bh = __float2bfloat16_rn(v); float r1 = __fsub_rn(v, __bfloat162float(bh));
    __nv_bfloat16 bm = __float2bfloat16_rn(r1); float r2 = __fsub_rn(r1, __bfloat162float(bm));
    __nv_bfloat16 bl = __float2bfloat16_rn(r2);
    h = __bfloat16_as_ushort(bh); m = __bfloat16_as_ushort(bm); l = __bfloat16_as_ushort(bl);
}
// byte offset of (row, k = 0 | 8) inside a canonical K-major no-swizzle operand: 8-row groups 256 B apart, the two K cores 128 B apart
__device__ __forceinline__ uint32_t canon_off(int row, int khalf) { return (uint32_t)((row >> 3) * 256 + khalf * 128 + (row & 7) * 16); }

constexpr uint32_t BF16_ONE = 0x3F80u;

// Data operand image, written once per pmp_set_data_linear: one thread per point.
__global__ void build_data_image_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n_local, long long nchunks, uint8_t* __restrict__ out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nchunks * CHUNK) return;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
    if (i < n_local) {
        uint32_t xh, xm, xl, yh, ym, yl;
        split3(x[i], xh, xm, xl); split3(y[i], yh, ym, yl);
        // k: 0 yh, 1 ym, 2 yl, 3..5 one, 6 xh, 7 xh | 8 xm, 9 xh, 10 xm, 11 xl, 12 xm, 13 xl, 14 xl, 15 zero
        lo = make_uint4(yh | (ym << 16), yl | (BF16_ONE << 16), BF16_ONE | (BF16_ONE << 16), xh | (xh << 16));
        hi = make_uint4(xm | (xh << 16), xm | (xl << 16), xm | (xl << 16), xl);
    }
    uint8_t* base = out + (i / CHUNK) * CHUNK_BYTES;
    const int row = (int)(i % CHUNK);
    *reinterpret_cast<uint4*>(base + canon_off(row, 0)) = lo;
    *reinterpret_cast<uint4*>(base + canon_off(row, 1)) = hi;
}

// A CTA's share of the (chunk, tile) units and the cursor that walks it tile by tile.
struct Range {
    int ntiles, t_lo, t_hi;          // t_lo: first tile of the first chunk; t_hi: end tile (exclusive) of the last chunk
    long long c_lo, c_hi;            // first / last chunk touched (inclusive); c_hi < c_lo: empty
    long long nu;                    // units
    __device__ __forceinline__ long long ca(int t) const { return c_lo + (t < t_lo ? 1 : 0); }
    __device__ __forceinline__ long long cb(int t) const { return c_hi + 1 - (t >= t_hi ? 1 : 0); }
};
__device__ __forceinline__ Range make_range(long long nchunks, int P, int cta, int nctas) {
    Range r;
    r.ntiles = (P + TILE_NODES - 1) / TILE_NODES;
    const long long units = nchunks * r.ntiles;
    const long long g0 = (long long)cta * units / nctas, g1 = (long long)(cta + 1) * units / nctas;
    r.nu = g1 - g0;
    if (r.nu <= 0) { r.c_lo = 0; r.c_hi = -1; r.t_lo = 0; r.t_hi = 0; r.nu = 0; return r; }
    r.c_lo = g0 / r.ntiles; r.t_lo = (int)(g0 - r.c_lo * r.ntiles);
    r.c_hi = (g1 - 1) / r.ntiles; r.t_hi = (int)(g1 - 1 - r.c_hi * r.ntiles) + 1;
    return r;
}
constexpr int MAX_UNITS = 4096;       // (tile, chunk) units per CTA that the unit table holds

struct Smem {                  // carved out of the dynamic shared memory of the hosting kernel
    uint8_t* sA;               // [ntiles][TILE_BYTES]
    uint8_t* sB;               // [max_chunks][CHUNK_BYTES]
    double* sscl;              // [ntiles*128]
    unsigned long long* sacc;  // [SETS][ntiles*128] integer sums private to (warp set, node)
    uint32_t* utab;            // [max_units] unit v of this CTA's walk: tile | local chunk << 8
    uint64_t* full;            // [SETS][2]
    uint64_t* empty;           // [SETS][2]
};
__host__ __device__ inline size_t smem_bytes(int ntiles, int max_chunks, int max_units) {
    return (size_t)ntiles * TILE_BYTES + (size_t)max_chunks * CHUNK_BYTES + (size_t)ntiles * TILE_NODES * 8 + (size_t)SETS * ntiles * TILE_NODES * 8 +
           (size_t)((max_units + 3) & ~3) * 4 + 2 * SETS * 2 * 8;
}
__device__ __forceinline__ Smem carve(uint8_t* base, int ntiles, int max_chunks, int max_units) {
    Smem s;
    s.sA = base; base += (size_t)ntiles * TILE_BYTES;
    s.sB = base; base += (size_t)max_chunks * CHUNK_BYTES;
    s.sscl = reinterpret_cast<double*>(base); base += (size_t)ntiles * TILE_NODES * 8;
    s.sacc = reinterpret_cast<unsigned long long*>(base); base += (size_t)SETS * ntiles * TILE_NODES * 8;
    s.full = reinterpret_cast<uint64_t*>(base); s.empty = s.full + SETS * 2; base += 2 * SETS * 2 * 8;
    s.utab = reinterpret_cast<uint32_t*>(base);
    return s;
}

// mbarrier phase bookkeeping across sweeps (persistent kernel): how often this thread's set has filled each of its two stages
struct PipeState { uint32_t n0, n1; };

// Stage this CTA's chunks [c_lo, c_hi] into shared memory (16-byte LDGSTS by all threads) and build the unit table.  Caller syncs.
__device__ __forceinline__ void stage_data(const Args& a, const Smem& s, const Range& rg, int nthreads) {
    const long long bytes = (rg.c_hi - rg.c_lo + 1) * CHUNK_BYTES;
    const uint8_t* src = a.bimg + rg.c_lo * CHUNK_BYTES;
    for (long long o = (long long)threadIdx.x * 16; o < bytes; o += (long long)nthreads * 16) {
        unsigned d = smem_u32(s.sB + o);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + o) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // unit table: the walk is tile by tile, chunks ascending inside a tile
    for (int v = threadIdx.x; v < (int)rg.nu; v += nthreads) {
        int t = 0; long long before = 0;
        for (;;) { const long long cnt = max(0ll, rg.cb(t) - rg.ca(t)); if (v < before + cnt) break; before += cnt; ++t; }
        s.utab[v] = (uint32_t)t | ((uint32_t)(rg.ca(t) + (v - before) - rg.c_lo) << 8);
    }
}
__device__ __forceinline__ void stage_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 32 residuals → packed square-accumulate into four accumulator pairs
__device__ __forceinline__ void sq32(const uint32_t (&v)[32], unsigned long long& a0, unsigned long long& a1, unsigned long long& a2, unsigned long long& a3) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { a0 = sq_acc2(v[8 * i], v[8 * i + 1], a0); a1 = sq_acc2(v[8 * i + 2], v[8 * i + 3], a1); a2 = sq_acc2(v[8 * i + 4], v[8 * i + 5], a2); a3 = sq_acc2(v[8 * i + 6], v[8 * i + 7], a3); }
}
__device__ __forceinline__ float fold(unsigned long long a0, unsigned long long a1, unsigned long long a2, unsigned long long a3) {
    return __fadd_rn(__fadd_rn(pair_sum(a0), pair_sum(a1)), __fadd_rn(pair_sum(a2), pair_sum(a3)));
}

// One sweep of this CTA's range: node operand → MMA/epilogue pipelines → integer flush.  Called by every thread of the CTA
// (warps 0..15 = epilogue, warps 16..19 = MMA issuers, further warps idle); uses named barriers 1 and 2; the caller provides
// the CTA-wide synchronisation before (nodes visible, data + unit table staged, mbarriers initialised) and after.
//
// The CTA runs SETS = 4 independent pipelines.  Set j owns TMEM columns [128j, 128j+128) as two 64-column stages, the units
// v = j, j+4, ... of the walk, one issuer warp and four epilogue warps (one per TMEM lane quarter).  Decoupled sets matter:
// one UMMA has ~400 cycles issue-to-visible latency and an epilogue warp's per-unit chain is ~400 cycles too, so the only
// way to keep the FP32 pipe fed is many units in flight that do not wait for each other.
__device__ __forceinline__ void sweep_range(const Args& a, const Smem& s, const Range& rg, uint32_t tmem_base, PipeState& ps, unsigned long long* dbg) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int P_pad = rg.ntiles * TILE_NODES;
    const int nu = (int)rg.nu;

    if (warp >= EPI_WARPS && warp < EPI_WARPS + SETS) {
        // ===== MMA issuer of set j: one N = 64 UMMA per unit into the set's stage (i & 1)
        const int j = warp - EPI_WARPS;
        const int n_mine = nu > j ? (nu - j + SETS - 1) / SETS : 0;
        proxy_fence_async();
        tc_fence_before();
        named_bar_sync(1, EPI_THREADS + SETS * 32);          // node operand written by the epilogue warps
        tc_fence_after();
        if (lane == 0) {
            const uint64_t da0 = umma_desc(s.sA), db0 = umma_desc(s.sB);
            const uint32_t idesc = umma_idesc(CHUNK);
            uint32_t par[2] = {(ps.n0 - 1) & 1, (ps.n1 - 1) & 1};          // parity of the previous use of each stage
            uint32_t used[2] = {ps.n0, ps.n1};
            for (int i = 0; i < n_mine; ++i) {
                const int st = i & 1;
                const uint32_t u = s.utab[SETS * i + j];
                if (used[st] > 0) { mbar_wait(&s.empty[2 * j + st], par[st]); tc_fence_after(); }
                par[st] ^= 1; used[st] = 1;
                umma(tmem_base + j * 128 + st * 64, da0 + (uint64_t)(u & 0xff) * (TILE_BYTES >> 4), db0 + (uint64_t)(u >> 8) * (CHUNK_BYTES >> 4), idesc);
                umma_commit(&s.full[2 * j + st]);
#ifdef PMP_TC_STAMPS
                if (dbg && j == 0 && i < 16) dbg[832 + i] = clock64();
#endif
            }
        }
        __syncwarp();
        ps.n0 += (uint32_t)((n_mine + 1) >> 1); ps.n1 += (uint32_t)(n_mine >> 1);
    } else if (warp < EPI_WARPS) {
        // ===== epilogue warp (q, j): TMEM lanes [32q, 32q+32) of set j's stages; rows `my` of tiles j, j+4, ... of the node operand
        const int q = warp & 3, j = warp >> 2;
        int my = q * 32 + lane;
        int lane0 = lane == 0;
        asm volatile("" : "+r"(my), "+r"(lane0));        // keep them in registers: re-reading %tid inside the unit loop costs a scoreboard wait each time
        const int n_mine = nu > j ? (nu - j + SETS - 1) / SETS : 0;
        float sg[NODES_PER_THREAD];
        {   // ---- node operand: all loads first, then split + store
            float b0[NODES_PER_THREAD], b1[NODES_PER_THREAD];
#pragma unroll
            for (int u = 0; u < NODES_PER_THREAD; ++u) {
                const int node = tid + EPI_THREADS * u;
                const bool in = node < a.P;
                b0[u] = in ? __ldcg(a.theta + 3ll * node) : 0.f; b1[u] = in ? __ldcg(a.theta + 3ll * node + 1) : 0.f; sg[u] = in ? __ldcg(a.theta + 3ll * node + 2) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < NODES_PER_THREAD; ++u) {
                const int node = tid + EPI_THREADS * u, t = j + SETS * u;
                if (t < rg.ntiles) {
                    uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
                    if (node < a.P) {
                        uint32_t h0, m0, l0, h1, m1, l1;
                        split3(-b0[u], h0, m0, l0); split3(-b1[u], h1, m1, l1);
                        // k: 0..2 one, 3 b0h, 4 b0m, 5 b0l, 6 b1h, 7 b1m | 8 b1h, 9 b1l, 10 b1m, 11 b1h, 12 b1l, 13 b1m, 14 b1l, 15 zero
                        lo = make_uint4(BF16_ONE | (BF16_ONE << 16), BF16_ONE | (h0 << 16), m0 | (l0 << 16), h1 | (m1 << 16));
                        hi = make_uint4(h1 | (l1 << 16), m1 | (h1 << 16), l1 | (m1 << 16), l1);
                    }
                    uint8_t* tb = s.sA + (size_t)t * TILE_BYTES;
                    *reinterpret_cast<uint4*>(tb + canon_off(my, 0)) = lo;
                    *reinterpret_cast<uint4*>(tb + canon_off(my, 1)) = hi;
                }
            }
        }
        proxy_fence_async();          // generic-proxy writes (node operand, staged data) → visible to the tensor core's async proxy
        tc_fence_before();
        named_bar_sync(1, EPI_THREADS + SETS * 32);
        tc_fence_after();
        if (dbg && tid == 0) { dbg[2] = clock64(); dbg[18] = globaltimer_ns(); }

        const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + j * 128;       // stage 0 of this set, this warp's lanes
        uint32_t par0 = ps.n0 & 1, par1 = ps.n1 & 1;
        uint32_t va[32], vb[32];
        if (n_mine > 0) {             // first TMEM read in flight while the fixed-point scales are computed
            mbar_wait(&s.full[2 * j], par0);
            tc_fence_after();
            tmem_ld32_nowait(tbase, va); tmem_ld32_nowait(tbase + 32, vb);
        }
        // ---- fixed-point scales 2^FX_SHIFT / sigma^2 of this thread's nodes, zeroed integer scratch of its own slots
#pragma unroll
        for (int u = 0; u < NODES_PER_THREAD; ++u) {
            const int node = tid + EPI_THREADS * u;
            if (j + SETS * u < rg.ntiles) s.sscl[node] = node < a.P ? (double)(1 << FX_SHIFT) / ((double)sg[u] * (double)sg[u]) : 0.0;
        }
        const int slot_base = j * P_pad + my;
        for (int t = 0; t < rg.ntiles; ++t) s.sacc[slot_base + t * TILE_NODES] = 0ull;
        named_bar_sync(2, EPI_THREADS);

        int cur_tile = -1;
        unsigned long long accq = 0ull;
        double scl = 0.0;
        // Units are taken four at a time: the reads, the barrier traffic and the packed FMAs of the four run back to back, and
        // the four float32 → fixed-point conversions (independent binary64 chains) are interleaved afterwards, so their latency
        // is paid once per group instead of once per unit.
        for (int i0 = 0; i0 < n_mine; i0 += 4) {
            float part[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int i = i0 + g;
                part[g] = 0.f;
                if (i < n_mine) {
                    const int st = g & 1;                                      // i0 is a multiple of 4: stage parity = g & 1
                    tmem_ld_wait();                                            // unit i: 64 residuals per lane in registers
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive(&s.empty[2 * j + st]);              // the stage is free once the set's four warps have read it
                    if (st) par1 ^= 1; else par0 ^= 1;
                    unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
                    sq32(va, a0, a1, a2, a3);
                    sq32(vb, a0, a1, a2, a3);
                    part[g] = fold(a0, a1, a2, a3);
                    if (i + 1 < n_mine) {                                      // next unit's read in flight
                        mbar_wait(&s.full[2 * j + (st ^ 1)], st ? par0 : par1);
                        tc_fence_after();
                        const uint32_t tn = tbase + (st ^ 1) * 64;
                        tmem_ld32_nowait(tn, va); tmem_ld32_nowait(tn + 32, vb);
                    }
                }
            }
            // float32 partials → 2^-FX_SHIFT fixed point, added to the running sum of the unit's tile; when the walk moves on to
            // another tile the finished sum goes to this thread's private scratch slot
            int tile[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) tile[g] = (i0 + g < n_mine) ? (int)(s.utab[SETS * (i0 + g) + j] & 0xff) : -1;
            if (tile[0] == cur_tile && tile[3] == cur_tile) {                  // common case: the whole group belongs to the running tile
                long long q4[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) q4[g] = __double2ll_rn(fmin((double)part[g] * scl, a.sat_limit));
                accq += (unsigned long long)((q4[0] + q4[1]) + (q4[2] + q4[3]));
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (tile[g] < 0) continue;
                    if (tile[g] != cur_tile) {
                        if (cur_tile >= 0) s.sacc[slot_base + cur_tile * TILE_NODES] = accq;
                        cur_tile = tile[g]; accq = 0ull;
                        scl = s.sscl[cur_tile * TILE_NODES + my];
                    }
                    accq += (unsigned long long)__double2ll_rn(fmin((double)part[g] * scl, a.sat_limit));   // NaN / hopeless nodes saturate
                }
            }
        }
        if (cur_tile >= 0) s.sacc[slot_base + cur_tile * TILE_NODES] = accq;
        ps.n0 += (uint32_t)((n_mine + 1) >> 1); ps.n1 += (uint32_t)(n_mine >> 1);
        if (dbg && tid == 0) { dbg[3] = clock64(); dbg[19] = globaltimer_ns(); }
        tc_fence_before();
        named_bar_sync(2, EPI_THREADS);
        // ---- flush: one global integer add per node this CTA touched
#pragma unroll
        for (int u = 0; u < NODES_PER_THREAD; ++u) {
            const int node = tid + EPI_THREADS * u;
            if (j + SETS * u < rg.ntiles && node < a.P) {
                unsigned long long sum = 0ull;
#pragma unroll
                for (int g = 0; g < SETS; ++g) sum += s.sacc[g * P_pad + node];
                if (sum) atomicAdd(a.acc + node, sum);
            }
        }
        if (dbg && tid == 0) { dbg[4] = clock64(); dbg[20] = globaltimer_ns(); }
    }
}

// One sweep as one launch (pmp_loglik, the CUDA-graph chain loop, the sharded multi-GPU loop): grid <= SM count, one CTA
// per SM (all 512 TMEM columns), 20 warps.
__global__ void __launch_bounds__(STANDALONE_THREADS, 1) sweep_linear_tc_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t dsm_tc[];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const Range rg = make_range(a.nchunks, a.P, blockIdx.x, gridDim.x);
    const Smem s = carve(dsm_tc, rg.ntiles, a.max_chunks, a.max_units);
    unsigned long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
    PMP_STAMP(dbg, 0);
    unsigned long long cta_t0 = 0; unsigned smid = 0;
    if (a.dbg && tid == 0) { cta_t0 = globaltimer_ns(); asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); }
    if (rg.nu > 0) stage_data(a, s, rg, STANDALONE_THREADS);
    if (tid == 0) {
        for (int b = 0; b < 2 * SETS; ++b) { mbar_init(&s.full[b], 1); mbar_init(&s.empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.generate) {
        // side job (as in sweep_linear_kernel): this CTA's slice of the NEXT iteration's normals — they depend on counters only
        const unsigned long long iter = a.gen.cnt->iteration;
        const int zcount = a.P * 3, per = (zcount + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int k = STANDALONE_THREADS - 1 - tid; k < per; k += STANDALONE_THREADS) {
            const int e = blockIdx.x * per + k;
            if (e < zcount) a.z[((iter + 1) & 1) * (long long)zcount + e] = (float)stream_step(a.gen.seed, iter + 1, (unsigned long long)e, a.gen.uniform);
        }
    }
    stage_wait();
    PMP_STAMP(dbg, 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    PipeState ps{0u, 0u};
    sweep_range(a, s, rg, tmem_base, ps, dbg);
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    PMP_STAMP(dbg, 5);
    if (a.dbg && tid == 0 && blockIdx.x < 1024) { a.dbg[64 + 3 * blockIdx.x] = cta_t0; a.dbg[64 + 3 * blockIdx.x + 1] = globaltimer_ns(); a.dbg[64 + 3 * blockIdx.x + 2] = smid; }
}

}  // namespace tc
}  // namespace pmp
