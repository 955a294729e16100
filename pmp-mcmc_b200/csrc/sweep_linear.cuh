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
};

constexpr int SWEEP_THREADS = 256;
constexpr int TILE_CHUNKS = 32;                    // chunks staged per shared-memory tile (2048 points)
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

template <int R, bool PACKED>
__global__ void __launch_bounds__(SWEEP_THREADS) sweep_linear_kernel(SweepArgs a) {
    __shared__ __align__(16) float tile[TILE_CHUNKS * CHUNK_STRIDE];
    __shared__ unsigned long long sacc[SWEEP_THREADS * R];   // used when TD > 1

    const int tid = threadIdx.x;
    const int tp = tid & (a.TP - 1);
    const int td = tid / a.TP;
    const int node0 = (blockIdx.y * a.TP + tp) * R;

    float b0[R], b1[R];
    double scl[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int p = node0 + r;
        if (p < a.P) {
            b0[r] = a.theta[3 * p]; b1[r] = a.theta[3 * p + 1];
            double s = (double)a.theta[3 * p + 2];
            scl[r] = (double)(1 << FX_SHIFT) / (s * s);
        } else { b0[r] = 0.f; b1[r] = 0.f; scl[r] = 0.0; }
    }
    unsigned long long accq[R];
#pragma unroll
    for (int r = 0; r < R; ++r) accq[r] = 0ull;
    bool saturated = false;

    const long long c_begin = (long long)blockIdx.x * a.nchunks / gridDim.x;
    const long long c_end = (long long)(blockIdx.x + 1) * a.nchunks / gridDim.x;

    for (long long t0 = c_begin; t0 < c_end; t0 += TILE_CHUNKS) {
        const int nct = (int)min((long long)TILE_CHUNKS, c_end - t0);
        // stage nct chunks: coalesced 16-byte loads, x then y of each chunk
        for (int i = tid; i < nct * (CHUNK / 4); i += SWEEP_THREADS) {
            int c = i / (CHUNK / 4), k = i - c * (CHUNK / 4);
            long long g = (t0 + c) * CHUNK + 4 * k;
            float4 xv, yv;
            if (g + 3 < a.n_local) {
                xv = __ldg(reinterpret_cast<const float4*>(a.x + g));
                yv = __ldg(reinterpret_cast<const float4*>(a.y + g));
            } else {
                float tx[4], ty[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bool ok = g + j < a.n_local;
                    tx[j] = ok ? a.x[g + j] : 0.f; ty[j] = ok ? a.y[g + j] : 0.f;
                }
                xv = make_float4(tx[0], tx[1], tx[2], tx[3]); yv = make_float4(ty[0], ty[1], ty[2], ty[3]);
            }
            float* dst = tile + c * CHUNK_STRIDE;
            reinterpret_cast<float4*>(dst)[k] = xv;
            reinterpret_cast<float4*>(dst + CHUNK)[k] = yv;
        }
        __syncthreads();
        for (int c = td; c < nct; c += a.TD) {
            long long first = (t0 + c) * CHUNK;
            int cnt = (int)min((long long)CHUNK, a.n_local - first);
            float part[R];
            const float* sx = tile + c * CHUNK_STRIDE;
            chunk_sumsq<R, PACKED>(sx, sx + CHUNK, cnt, b0, b1, part);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double dq = (double)part[r] * scl[r];
                if (!(dq < a.sat_limit)) { dq = a.sat_limit; saturated = true; }
                accq[r] += (unsigned long long)__double2ll_rn(dq);
            }
        }
        __syncthreads();
    }

    if (a.TD > 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) sacc[tid * R + r] = 0ull;   // only slots [0, TP*R) are accumulated into
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (accq[r]) atomicAdd(&sacc[tp * R + r], accq[r]);
        __syncthreads();
        if (td == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (node0 + r < a.P && sacc[tp * R + r]) atomicAdd(a.acc + node0 + r, sacc[tp * R + r]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (node0 + r < a.P && accq[r]) atomicAdd(a.acc + node0 + r, accq[r]);
    }
    if (saturated) atomicOr(&a.cnt->flags, 1);
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
