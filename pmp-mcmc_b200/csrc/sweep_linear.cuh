// sweep_linear.cuh — the proposals x data likelihood sweep of the 3-parameter linear-Gaussian model.
//
// Replaces log_likelihood_kernel's data loop (500_MP.cu:16-21, 500_PMP.cu:16-21, conv_pmp.cu:16-20,
// conv_mh.cu:10-26) and BayesNet.loglik (lb.py:103-108).  The reference runs ONE thread per proposal with a
// serial loop over all n points and a global-memory read-modify-write per point.  Here:
//   * a thread owns R proposals in registers (b0, b1) and walks data that the CTA staged in shared memory;
//     every lane of a warp reads the same (x,y) words, so each LDS.128 is a broadcast feeding 4 points x R nodes;
//   * per (node, point) the work is  yhat = fma(b1,x,b0);  d = y - yhat;  acc = fma(d,d,acc)  — the reference's
//     own per-point arithmetic (500_MP.cu:17-18) — issued as packed FFMA2/FADD2 (fma.rn.f32x2), two points per op;
//   * the data axis is split over the grid; each CHUNK of points yields one float32 partial per node that is
//     converted to 2^-FX_SHIFT fixed point (already divided by sigma^2) and from then on only integer adds happen
//     (registers → shared atomics → one global RED per node per CTA), so the result does not depend on the order
//     in which CTAs, or GPUs, contribute.
// Roofline: FP32 issue.  3 FP32 lane-ops per (node, point) = 6 flop; nothing else scales with P*n.
#pragma once
#include "accept.cuh"
#include "common.cuh"

namespace pmp {

struct SweepArgs {
    const float* __restrict__ x;
    const float* __restrict__ y;
    const float* __restrict__ theta;      // [P,3] (b0, b1, sigma)
    unsigned long long* __restrict__ acc; // [P]
    DeviceCounters* cnt;
    long long n_local;
    long long nchunks;                    // ceil(n_local / CHUNK)
    int P;
    int TP;                               // threads along the node axis (power of two, <= 256)
    int TD;                               // 256 / TP: threads along the chunk axis
    double sat_limit;                     // per-partial saturation bound in fixed-point units
    int generate;                         // 1: also fill the next iteration's half of the normals table z (side job)
    float* z;                             // [2, P*3] normals table: half (iter & 1) is the current iteration's, the other is filled here
    ProposeArgs gen;
    unsigned long long* dbg;              // optional phase stamps of CTA 0: [0..15] clock64, [16..31] globaltimer
};

constexpr int SWEEP_THREADS = 256;
#ifndef SWEEP_MIN_CTAS
#define SWEEP_MIN_CTAS 4
#endif
constexpr int TILE_CHUNKS = 32;                    // chunks staged per shared-memory buffer (3072 points, 24.75 KB)
constexpr int MAX_TP = 64;                         // node tile <= 256 nodes
constexpr int CHUNK_STRIDE = 2 * CHUNK + 4;        // floats per staged chunk: x[CHUNK], y[CHUNK], 16 B pad (bank skew)

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

// Sum of squared residuals of one chunk for R nodes, float32, fixed evaluation order:
// lane .x takes even points, lane .y odd points, then (.x + .y), then the scalar tail in order.
template <int R, bool PACKED>
__device__ __forceinline__ void chunk_sumsq(const float* __restrict__ sx, const float* __restrict__ sy, int cnt,
                                            const float (&b0)[R], const float (&b1)[R], float (&out)[R]) {
    if (PACKED) {
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
        const float4* x4 = reinterpret_cast<const float4*>(sx);
        const float4* y4 = reinterpret_cast<const float4*>(sy);
        if (cnt == CHUNK) {
#pragma unroll 4
            for (int g = 0; g < CHUNK / 4; ++g) {
                float4 xv = x4[g], yv = y4[g];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float2 yh = ffma2(make_float2(b1[r], b1[r]), make_float2(xv.x, xv.y), make_float2(b0[r], b0[r]));
                    float2 d = fsub2(make_float2(yv.x, yv.y), yh);
                    acc[r] = ffma2(d, d, acc[r]);
                    yh = ffma2(make_float2(b1[r], b1[r]), make_float2(xv.z, xv.w), make_float2(b0[r], b0[r]));
                    d = fsub2(make_float2(yv.z, yv.w), yh);
                    acc[r] = ffma2(d, d, acc[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) out[r] = __fadd_rn(acc[r].x, acc[r].y);
            return;
        }
        int ng = cnt >> 2;
        for (int g = 0; g < ng; ++g) {
            float4 xv = x4[g], yv = y4[g];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float2 yh = ffma2(make_float2(b1[r], b1[r]), make_float2(xv.x, xv.y), make_float2(b0[r], b0[r]));
                float2 d = fsub2(make_float2(yv.x, yv.y), yh);
                acc[r] = ffma2(d, d, acc[r]);
                yh = ffma2(make_float2(b1[r], b1[r]), make_float2(xv.z, xv.w), make_float2(b0[r], b0[r]));
                d = fsub2(make_float2(yv.z, yv.w), yh);
                acc[r] = ffma2(d, d, acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = __fadd_rn(acc[r].x, acc[r].y);
        for (int i = ng << 2; i < cnt; ++i) {
            float xv = sx[i], yv = sy[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float d = __fsub_rn(yv, __fmaf_rn(b1[r], xv, b0[r]));
                out[r] = __fmaf_rn(d, d, out[r]);
            }
        }
    } else {
        // scalar FFMA variant with the packed variant's evaluation order: even points of the full groups of 4
        // into one accumulator, odd points into the other, (even + odd), then the ragged tail in order
        float ae[R], ao[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { ae[r] = 0.f; ao[r] = 0.f; }
        const int full = (cnt >> 2) << 2;
#pragma unroll 8
        for (int g = 0; g < (full >> 1); ++g) {
            float x0 = sx[2 * g], x1 = sx[2 * g + 1], y0 = sy[2 * g], y1 = sy[2 * g + 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float d0 = __fsub_rn(y0, __fmaf_rn(b1[r], x0, b0[r]));
                float d1 = __fsub_rn(y1, __fmaf_rn(b1[r], x1, b0[r]));
                ae[r] = __fmaf_rn(d0, d0, ae[r]);
                ao[r] = __fmaf_rn(d1, d1, ao[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) out[r] = __fadd_rn(ae[r], ao[r]);
        for (int i = full; i < cnt; ++i) {
            float xv = sx[i], yv = sy[i];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float d = __fsub_rn(yv, __fmaf_rn(b1[r], xv, b0[r]));
                out[r] = __fmaf_rn(d, d, out[r]);
            }
        }
    }
}

// cp.async (LDGSTS) 16-byte copy; src_bytes = 0 zero-fills the destination without touching global memory
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Work decomposition.  A unit is (node tile, chunk): PT = TP*R nodes x CHUNK points.  Units are ordered tile-major and
// cut into gridDim.x equal contiguous ranges, so every CTA (and therefore every SM) gets the same amount of FP32 work
// to within one unit, for any P and n.  A range touches one tile, or two when it straddles a tile boundary; each
// (tile, chunk range) segment ends with ONE integer atomic per node of the tile — P*gridDim.x/ntiles atomics per sweep
// instead of P per CTA.  The chunks of a segment are staged with cp.async into a double-buffered shared-memory tile;
// the first stage is issued before anything that depends on the chain state, so its latency hides behind the
// construction of the tile's nodes.
template <int R, bool PACKED>
__global__ void __launch_bounds__(SWEEP_THREADS, SWEEP_MIN_CTAS) sweep_linear_kernel(const __grid_constant__ SweepArgs a) {
    extern __shared__ __align__(16) float tile[];                 // 2 x TILE_CHUNKS x CHUNK_STRIDE floats
    __shared__ float sprops[MAX_TP * R * 3];                      // the tile's nodes (b0, b1, sigma)
    __shared__ double sscl[MAX_TP * R];                           // 2^FX_SHIFT / sigma^2 per node

    const int tid = threadIdx.x;
    const int tp = tid & (a.TP - 1);
    const int td = tid / a.TP;
    const int PT = a.TP * R;
    const int ntiles = (a.P + PT - 1) / PT;
    const long long units = (long long)ntiles * a.nchunks;
    long long u = (long long)blockIdx.x * units / gridDim.x;
    const long long u_end = (long long)(blockIdx.x + 1) * units / gridDim.x;
    const int zcount = a.P * 3;
    bool saturated = false;
    bool first_segment = true;
    unsigned long long* dbg = (a.dbg && blockIdx.x == 0) ? a.dbg : nullptr;
    PMP_STAMP(dbg, 0);
    unsigned long long cta_t0 = 0; unsigned smid = 0;
    if (a.dbg && tid == 0) { cta_t0 = globaltimer_ns(); asm volatile("mov.u32 %0, %smid;" : "=r"(smid)); }

    auto stage = [&](int buf, long long t0, int nct) {
        float* base = tile + buf * (TILE_CHUNKS * CHUNK_STRIDE);
        for (int i = tid; i < nct * (CHUNK / 2); i += SWEEP_THREADS) {      // CHUNK/4 x-vectors + CHUNK/4 y-vectors per chunk
            int c = i / (CHUNK / 2), k = i - c * (CHUNK / 2);
            bool isy = k >= CHUNK / 4;
            int kk = isy ? k - CHUNK / 4 : k;
            long long g = (t0 + c) * CHUNK + 4 * kk;
            const float* src = (isy ? a.y : a.x) + (g < a.n_local ? g : 0);
            cp_async16(base + c * CHUNK_STRIDE + (isy ? CHUNK : 0) + 4 * kk, src, g < a.n_local ? 16 : 0);
        }
        cp_async_commit();
    };

    while (u < u_end) {
        const int ptile = (int)(u / a.nchunks);
        const long long c_begin = u - (long long)ptile * a.nchunks;
        const long long c_end = min(a.nchunks, c_begin + (u_end - u));
        const int node_base = ptile * PT;
        const int ntl = (int)((c_end - c_begin + TILE_CHUNKS - 1) / TILE_CHUNKS);

        __syncthreads();                       // previous segment is done with the tile buffers and sprops
        stage(0, c_begin, (int)min((long long)TILE_CHUNKS, c_end - c_begin));

        PMP_STAMP(dbg, 1);

        // ---- the tile's nodes (published by the acceptance kernel / pmp_propose) and their fixed-point scale factors
        for (int i = tid; i < PT * 3; i += SWEEP_THREADS) {
            int node = node_base + i / 3, j = i - (i / 3) * 3;
            float v = (node < a.P) ? __ldcg(a.theta + (long long)node * 3 + j) : 0.f;
            sprops[i] = v;
            if (j == 2) sscl[i / 3] = (node < a.P) ? (double)(1 << FX_SHIFT) / ((double)v * (double)v) : 0.0;
        }
        __syncthreads();
        PMP_STAMP(dbg, 2);
        float b0[R], b1[R];
        double scl[R];
        unsigned long long accq[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            int i = tp * R + r;
            b0[r] = sprops[3 * i]; b1[r] = sprops[3 * i + 1];
            scl[r] = sscl[i];
            accq[r] = 0ull;
        }

        // ---- stream the segment's chunks through the double-buffered tile
        for (int t = 0; t < ntl; ++t) {
            const long long t0 = c_begin + (long long)t * TILE_CHUNKS;
            const int nct = (int)min((long long)TILE_CHUNKS, c_end - t0);
            if (t + 1 < ntl) {
                stage((t + 1) & 1, t0 + TILE_CHUNKS, (int)min((long long)TILE_CHUNKS, c_end - t0 - TILE_CHUNKS));
                cp_async_wait<1>();
            } else cp_async_wait<0>();
            __syncthreads();
            if (t == 0) PMP_STAMP(dbg, 3);
            const float* base = tile + (t & 1) * (TILE_CHUNKS * CHUNK_STRIDE);
            for (int c = td; c < nct; c += a.TD) {
                int cnt = (int)min((long long)CHUNK, a.n_local - (t0 + c) * CHUNK);
                float part[R];
                const float* sx = base + c * CHUNK_STRIDE;
                chunk_sumsq<R, PACKED>(sx, sx + CHUNK, cnt, b0, b1, part);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    double dq = (double)part[r] * scl[r];
                    if (!(dq < a.sat_limit)) { dq = a.sat_limit; saturated = true; }
                    accq[r] += (unsigned long long)__double2ll_rn(dq);
                }
            }
            if (t + 1 < ntl) __syncthreads();
        }

        PMP_STAMP(dbg, 4);
        if (first_segment && a.generate) {
            // side job of the last warp (it owns the fewest chunks) before it joins the flush: this CTA's slice of the
            // normals of the NEXT iteration (they depend on counters only, never on the chain state), so that no launch
            // ever waits for a binary64 quantile evaluation
            const unsigned long long iter = a.gen.cnt->iteration;
            const int per = (zcount + gridDim.x - 1) / gridDim.x;
            for (int k = SWEEP_THREADS - 1 - tid; k < per; k += SWEEP_THREADS) {      // per > SWEEP_THREADS when the grid is small (few units)
                const int e = blockIdx.x * per + k;
                if (e < zcount) a.z[((iter + 1) & 1) * (long long)zcount + e] = (float)stream_step(a.gen.seed, iter + 1, (unsigned long long)e, a.gen.uniform);
            }
        }
        first_segment = false;
        // ---- segment flush, integer adds only: lanes that share a node (warp shuffles) → one row per warp in shared
        // memory → one global atomic per node.  (64-bit shared atomics compile to CAS spin loops; not used.)
        if (a.TP < 32) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                for (int o = 16; o >= a.TP; o >>= 1) accq[r] += __shfl_xor_sync(0xffffffffu, accq[r], o);
        }
        __syncthreads();                       // everyone is done reading the data tile: reuse it as integer scratch
        unsigned long long* sred = reinterpret_cast<unsigned long long*>(tile);
        const int lane = tid & 31, warp = tid >> 5;
        const int rows = a.TP < 32 ? SWEEP_THREADS / 32 : a.TD;       // partial rows per node
        const int row = a.TP < 32 ? warp : td;
        if (a.TP >= 32 || lane < a.TP) {
#pragma unroll
            for (int r = 0; r < R; ++r) sred[row * PT + tp * R + r] = accq[r];
        }
        __syncthreads();
        for (int i = tid; i < PT; i += SWEEP_THREADS) {
            unsigned long long s = 0ull;
            for (int k = 0; k < rows; ++k) s += sred[k * PT + i];
            if (node_base + i < a.P && s) atomicAdd(a.acc + node_base + i, s);
        }
        u += c_end - c_begin;
        PMP_STAMP(dbg, 5);
    }
    if (saturated) atomicOr(&a.cnt->flags, 1);
    if (a.dbg && tid == 0 && blockIdx.x < 1024) { a.dbg[64 + 3 * blockIdx.x] = cta_t0; a.dbg[64 + 3 * blockIdx.x + 1] = globaltimer_ns(); a.dbg[64 + 3 * blockIdx.x + 2] = smid; }
}

// FP32 issue-rate microbenchmark: the denominator of the sweep's roofline (MEASURED_PEAKS.json has no FP32 number).
template <bool PACKED>
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* out, int iters, float seedv) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seedv + i, seedv - i);
    float2 m = make_float2(1.0000001f, 0.9999999f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) a[i] = ffma2(a[i], m, c);
            else { a[i].x = __fmaf_rn(a[i].x, m.x, c.x); a[i].y = __fmaf_rn(a[i].y, m.y, c.y); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;
}

}  // namespace pmp
