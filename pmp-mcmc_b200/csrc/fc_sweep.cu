// fc_sweep.cu — log-target sweep of the Bayesian FC model: for P candidate weight vectors, -CrossEntropy(MLP_p(X), y).
//
// Replaces the loop `for all in range(N+1): weights[all] = exp(-loss(proposal_nets[all]))` (PMP_FC.py:117-118,
// MP_FC.py:112-114, MH_FC.py:98) with Model = 784-512-256-128-10 ReLU MLP (PMP_FC.py:21-36) and
// loss = CrossEntropyLoss(mean)(net(X), y) / 10 (PMP_FC.py:40-44).
//
// Numerics.  Proposals differ from the current state by alpha = 1e-4 per weight (PMP_FC.py:15) and the acceptance
// standardises the log-weights (PMP_FC.py:138-140), so what matters is the DIFFERENCE of losses between nodes: a plain
// bf16 contraction (8 mantissa bits, ulp(0.03) = 1.2e-4) cannot even represent theta + delta.  The contraction is therefore
// run as a 3-term split ("bf16x3"): every operand is h + l with h = bf16(v), l = bf16(v - h); A.B ~= Ah.Bh + Ah.Bl + Al.Bh
// keeps ~16 mantissa bits.  The three terms are one GEMM with K tripled: A' = [Ah | Ah | Al], B' = [Bh | Bl | Bh], so the
// tensor-core kernel is an ordinary K-major bf16 GEMM with fp32 accumulation in TMEM; the roofline counts the ALGORITHMIC
// flops (2.566528e5.n per node), the hardware executes 3x that.
//
// Kernel (sm_100a): one 128 x BN output tile per CTA; warp 0 = TMA producer (cp.async.bulk.tensor, SWIZZLE_128B boxes of
// 64 bf16 x rows, 4-stage mbarrier ring), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16,
// cta_group::1, accumulator in BN TMEM columns), warps 2-5 = epilogue (tcgen05.ld 32x32b, one TMEM lane quarter each):
//   RELU_SPLIT  bias + ReLU, re-split to (h,l) and written straight into the next layer's A' = [h | h | l];
//   NLL         bias, log-softmax over the 10 classes, pick the label, sum over rows into a 2^-32 fixed-point integer
//               (exact, order-free: identical bits for any CTA order and any data sharding).
// Every mbarrier wait is bounded (trap instead of hang).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace pmp {
namespace fc {

constexpr int BM = 128, BK = 64, UMMA_K = 16, STAGES = 4;
constexpr int GEMM_THREADS = 192;
constexpr int D_IN = 784, D_IN_PAD = 832, H1 = 512, H2 = 256, H3 = 128, NCLS = 10, NCLS_PAD = 16;
constexpr long long THETA_DIM = 567434;
constexpr long long OFF_W1 = 0, OFF_B1 = OFF_W1 + (long long)H1 * D_IN, OFF_W2 = OFF_B1 + H1, OFF_B2 = OFF_W2 + (long long)H2 * H1,
                    OFF_W3 = OFF_B2 + H2, OFF_B3 = OFF_W3 + (long long)H3 * H2, OFF_W4 = OFF_B3 + H3, OFF_B4 = OFF_W4 + (long long)NCLS * H3;
static_assert(OFF_B4 + NCLS == THETA_DIM, "theta layout");
constexpr int LOSS_FX_SHIFT = 32;

enum { EPI_RELU_SPLIT = 0, EPI_NLL = 1 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > 200000000u) __trap();       // a protocol bug must fail loudly, never hang the GPU
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_tile) {
    // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row atoms 1024 B apart (SBO), LBO unused (1), version 1 (sm_100)
    return (uint64_t)((smem_u32(smem_tile) & 0x3FFFF) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                 "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) { return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16); }

struct GemmArgs {
    int M;                  // rows (data points of this shard)
    int K3;                 // 3 * padded K: length of the concatenated contraction axis
    int a_shared;           // 1: A has no batch axis (layer 1: X' is the same for every node)
    const float* bias;      // [batch, bias_stride]
    int bias_stride;
    __nv_bfloat16* out;     // RELU_SPLIT: next A' [batch, M, 3*N_total]
    int n_total;            // N of the whole layer (row length of `out` is 3*n_total)
    const int* labels;      // NLL: [M]
    unsigned long long* loss;   // NLL: [batch] fixed-point sums
};

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1) fc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * A_BYTES;
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ float s_red[4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // rasterisation: (n-block, node) fastest, so the CTAs that share an A tile (the data rows) run together and the tile is
    // read from HBM once and then from L2; the per-node weights (a few MB) stay L2-resident across the whole launch
    const int nblk_n = g.n_total / BN;
    const int m_blk = blockIdx.y, n_blk = blockIdx.x % nblk_n, batch = blockIdx.x / nblk_n;
    const int num_k = g.K3 / BK;
    constexpr int TMEM_COLS = BN < 32 ? 32 : BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {                                         // ===== TMA producer =====
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                if (kb >= STAGES) mbar_wait(&empty_bar[s], ((kb / STAGES) - 1) & 1);
                mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
                tma_load_3d(sA + s * A_BYTES, &tmA, &full_bar[s], kb * BK, m_blk * BM, g.a_shared ? 0 : batch);
                tma_load_3d(sB + s * B_BYTES, &tmB, &full_bar[s], kb * BK, n_blk * BN, batch);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                         // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                mbar_wait(&full_bar[s], (kb / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = umma_desc_sw128(sA + s * A_BYTES), db = umma_desc_sw128(sB + s * B_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)                // +32 B along K inside the 128-byte swizzle atom = +2 in the address field
                    umma_bf16(tmem_base, da + 2ull * k, db + 2ull * k, idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);                            // frees the smem slot when these MMAs retire
            }
            umma_commit(&tmem_full_bar);                               // accumulator complete
        }
    } else {                                                     // ===== epilogue warps 2..5 =====
        const int quarter = warp & 3;                            // TMEM lanes [32q, 32q+32) are the only ones this warp may read
        const int row = m_blk * BM + quarter * 32 + lane;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (EPI == EPI_RELU_SPLIT) {
            const float* bias = g.bias + (long long)batch * g.bias_stride + n_blk * BN;
            const long long ldo = 3ll * g.n_total;
            __nv_bfloat16* orow = g.out + ((long long)batch * g.M + row) * ldo + n_blk * BN;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; ++cc) {
                uint32_t v[32];
                tmem_ld32(taddr + cc * 32, v);
                uint32_t hp[16], lp[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float a0 = fmaxf(__uint_as_float(v[2 * i]) + __ldg(bias + cc * 32 + 2 * i), 0.f);
                    float a1 = fmaxf(__uint_as_float(v[2 * i + 1]) + __ldg(bias + cc * 32 + 2 * i + 1), 0.f);
                    __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
                    __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
                    hp[i] = pack_bf16x2(h0, h1); lp[i] = pack_bf16x2(l0, l1);
                }
                if (row < g.M) {
                    uint4* d0 = reinterpret_cast<uint4*>(orow + cc * 32);
                    uint4* d1 = reinterpret_cast<uint4*>(orow + g.n_total + cc * 32);
                    uint4* d2 = reinterpret_cast<uint4*>(orow + 2 * g.n_total + cc * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 hv = make_uint4(hp[4 * q], hp[4 * q + 1], hp[4 * q + 2], hp[4 * q + 3]);
                        d0[q] = hv; d1[q] = hv;
                        d2[q] = make_uint4(lp[4 * q], lp[4 * q + 1], lp[4 * q + 2], lp[4 * q + 3]);
                    }
                }
            }
        } else {
            uint32_t v[16];
            tmem_ld16(taddr, v);
            const float* bias = g.bias + (long long)batch * g.bias_stride;
            float z[NCLS], mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < NCLS; ++i) { z[i] = __uint_as_float(v[i]) + __ldg(bias + i); mx = fmaxf(mx, z[i]); }
            float se = 0.f;
#pragma unroll
            for (int i = 0; i < NCLS; ++i) se += expf(z[i] - mx);
            float nll = 0.f;
            if (row < g.M) {
                const int lab = g.labels[row];
                float zl = z[0];
#pragma unroll
                for (int i = 1; i < NCLS; ++i) zl = (lab == i) ? z[i] : zl;
                nll = (mx + logf(se)) - zl;
            }
            // per-row NLL → 2^-32 fixed point; everything after this line is integer (exact, order-free)
            long long q = (row < g.M) ? __double2ll_rn((double)nll * 4294967296.0) : 0ll;
            for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            if (lane == 0 && q != 0) atomicAdd(g.loss + batch, (unsigned long long)q);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    (void)s_red;
}

// X f32 [n, 784] → X' bf16 [n, 3*832] = [h | h | l], zero padded
__global__ void split_x_kernel(const float* __restrict__ X, __nv_bfloat16* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * D_IN_PAD) return;
    long long r = i / D_IN_PAD; int c = (int)(i - r * D_IN_PAD);
    float v = c < D_IN ? X[r * D_IN + c] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    __nv_bfloat16* o = out + r * (3ll * D_IN_PAD);
    o[c] = h; o[D_IN_PAD + c] = h; o[2 * D_IN_PAD + c] = l;
}

// theta_p f32 → W' bf16 [rows_pad, 3*Kpad] = [h | l | h] (zero padded rows/cols) for one layer of `nb` nodes
__global__ void split_w_kernel(const float* __restrict__ theta, long long theta_stride, long long w_off, int rows, int K, int rows_pad, int Kpad,
                               __nv_bfloat16* __restrict__ out, int nb) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)rows_pad * Kpad;
    if (i >= per * nb) return;
    int b = (int)(i / per); long long rem = i - b * per;
    int r = (int)(rem / Kpad), c = (int)(rem - (long long)r * Kpad);
    float v = (r < rows && c < K) ? theta[b * theta_stride + w_off + (long long)r * K + c] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    __nv_bfloat16* o = out + ((long long)b * rows_pad + r) * (3ll * Kpad);
    o[c] = h; o[Kpad + c] = l; o[2 * Kpad + c] = h;
}

__global__ void gather_bias_kernel(const float* __restrict__ theta, long long theta_stride, float* __restrict__ out, int nb) {
    // out [nb, 512 + 256 + 128 + 16]
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = H1 + H2 + H3 + NCLS_PAD;
    if (i >= per * nb) return;
    int b = i / per, c = i - b * per;
    const float* t = theta + b * theta_stride;
    float v;
    if (c < H1) v = t[OFF_B1 + c];
    else if (c < H1 + H2) v = t[OFF_B2 + c - H1];
    else if (c < H1 + H2 + H3) v = t[OFF_B3 + c - H1 - H2];
    else { int k = c - H1 - H2 - H3; v = k < NCLS ? t[OFF_B4 + k] : 0.f; }
    out[i] = v;
}

__global__ void finalize_loss_kernel(const unsigned long long* loss_fx, double* lt, int P, double n_global, double inv_scale) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) lt[p] = -((double)(long long)loss_fx[p] * (1.0 / 4294967296.0) / n_global) * inv_scale;   // -(mean CE)/loss_div
}

// sum_k mean_dim (theta_j - theta_k)^2 pieces for the MP-FC kernel term (MP_FC.py:107-114), closed form about node 0
__global__ void fc_s1_kernel(const float* __restrict__ theta, long long dim, int P, float* __restrict__ s1) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= dim) return;
    float t0 = theta[j]; double s = 0.0;
    for (int p = 0; p < P; ++p) s += (double)theta[(long long)p * dim + j] - (double)t0;
    s1[j] = (float)s;
}
__global__ void fc_dots_kernel(const float* __restrict__ theta, const float* __restrict__ s1, long long dim, double* __restrict__ dj2, double* __restrict__ dot) {
    const int p = blockIdx.y;
    double a = 0.0, b = 0.0;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < dim; j += (long long)gridDim.x * blockDim.x) {
        double d = (double)theta[(long long)p * dim + j] - (double)theta[j];
        a = fma(d, d, a); b = fma(d, (double)s1[j], b);
    }
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(dj2 + p, a); atomicAdd(dot + p, b); }
}
__global__ void fc_kterm_kernel(const double* dj2, const double* dot, int P, double dim, double ks, double* logw) {
    __shared__ double sS2;
    if (threadIdx.x == 0) { double s = 0.0; for (int p = 0; p < P; ++p) s += dj2[p]; sS2 = s; }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        double sumsq = (double)P * dj2[p] - 2.0 * dot[p] + sS2;
        logw[p] = ((double)(P - 1) * (-0.91893853320467274178 - log(ks)) - 0.5 * sumsq / (ks * ks) / dim) / (double)P;
    }
}

struct FcState {
    long long n_local = 0, n_global = 0;
    __nv_bfloat16* xs = nullptr;       // X' [n, 3*832]
    int* labels = nullptr;
    int nb = 0;                        // nodes per batch
    __nv_bfloat16 *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w4 = nullptr;   // W' per batch
    float* bias = nullptr;             // [nb, 912]
    __nv_bfloat16 *a2 = nullptr, *a3 = nullptr, *a4 = nullptr;                  // A' of layers 2..4 per batch
    unsigned long long* loss = nullptr;   // [P]
    float* s1 = nullptr; double* dj2 = nullptr; double* dot = nullptr;
    CUtensorMap tmX, tmW1, tmA2, tmW2, tmA3, tmW3, tmA4, tmW4;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    }
    return fn;
}
// bf16 tensor [batch, rows, K3] row-major; box = 64 x box_rows x 1, 128-byte swizzle
static int make_map(CUtensorMap* m, void* base, long long K3, long long rows, long long batch, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return PMP_ERR_CUDA; }
    cuuint64_t dims[3] = {(cuuint64_t)K3, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)K3 * 2, (cuuint64_t)K3 * 2 * (cuuint64_t)rows};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for [%lld,%lld,%lld]", (int)r, batch, rows, K3); return PMP_ERR_CUDA; }
    return PMP_OK;
}

template <int BN, int EPI> static size_t gemm_smem() { return (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024; }

template <int BN, int EPI>
static int launch_gemm(pmp_ctx* c, const CUtensorMap& a, const CUtensorMap& b, const GemmArgs& g, int n_total, int nb) {
    static bool attr = false;
    if (!attr) { PMP_CUDA(cudaFuncSetAttribute(fc_gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem<BN, EPI>())); attr = true; }
    dim3 grid((n_total / BN) * nb, (g.M + BM - 1) / BM, 1);
    fc_gemm_kernel<BN, EPI><<<grid, GEMM_THREADS, gemm_smem<BN, EPI>(), c->stream>>>(a, b, g);
    c->launches++;
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

static void free_state(FcState* s) {
    void* ptrs[] = {s->xs, s->labels, s->w1, s->w2, s->w3, s->w4, s->bias, s->a2, s->a3, s->a4, s->loss, s->s1, s->dj2, s->dot};
    for (void* p : ptrs) if (p) cudaFree(p);
}

}  // namespace fc
}  // namespace pmp

using namespace pmp;
using namespace pmp::fc;

extern "C" {

int pmp_allreduce_u64(pmp_ctx* c, unsigned long long* buf, size_t count);   // pmp_abi.cu

int pmp_fc_destroy(pmp_ctx* c) {
    if (c->fc) { free_state(reinterpret_cast<FcState*>(c->fc)); delete reinterpret_cast<FcState*>(c->fc); c->fc = nullptr; }
    return PMP_OK;
}

int pmp_set_data_fc(pmp_ctx* c, const float* X, const int64_t* labels, int64_t n_local, int64_t n_offset, int64_t n_global) {
    PMP_REQUIRE(c && X && labels && n_local > 0 && n_global >= n_local, "bad arguments");
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    pmp_fc_destroy(c);
    FcState* s = new FcState();
    c->fc = s;
    s->n_local = n_local; s->n_global = n_global;
    s->nb = getenv("PMP_FC_BATCH") ? atoi(getenv("PMP_FC_BATCH")) : 8;
    if (s->nb < 1) s->nb = 1;
    float* d_x32 = nullptr;
    std::vector<int> lab32((size_t)n_local);
    for (int64_t i = 0; i < n_local; ++i) { PMP_REQUIRE(labels[i] >= 0 && labels[i] < NCLS, "label %lld out of range at row %lld", (long long)labels[i], (long long)i); lab32[i] = (int)labels[i]; }
    PMP_CUDA(cudaMalloc((void**)&d_x32, (size_t)n_local * D_IN * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->xs, (size_t)n_local * 3 * D_IN_PAD * sizeof(__nv_bfloat16)));
    PMP_CUDA(cudaMalloc((void**)&s->labels, (size_t)n_local * sizeof(int)));
    PMP_CUDA(cudaMemcpyAsync(d_x32, X, (size_t)n_local * D_IN * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaMemcpyAsync(s->labels, lab32.data(), (size_t)n_local * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    long long tot = n_local * D_IN_PAD;
    split_x_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, c->stream>>>(d_x32, s->xs, n_local);
    c->launches++;
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_x32);
    const int nb = s->nb;
    PMP_CUDA(cudaMalloc((void**)&s->w1, (size_t)nb * H1 * 3 * D_IN_PAD * 2));
    PMP_CUDA(cudaMalloc((void**)&s->w2, (size_t)nb * H2 * 3 * H1 * 2));
    PMP_CUDA(cudaMalloc((void**)&s->w3, (size_t)nb * H3 * 3 * H2 * 2));
    PMP_CUDA(cudaMalloc((void**)&s->w4, (size_t)nb * NCLS_PAD * 3 * H3 * 2));
    PMP_CUDA(cudaMalloc((void**)&s->bias, (size_t)nb * (H1 + H2 + H3 + NCLS_PAD) * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->a2, (size_t)nb * n_local * 3 * H1 * 2));
    PMP_CUDA(cudaMalloc((void**)&s->a3, (size_t)nb * n_local * 3 * H2 * 2));
    PMP_CUDA(cudaMalloc((void**)&s->a4, (size_t)nb * n_local * 3 * H3 * 2));
    int rc;
    if ((rc = make_map(&s->tmX, s->xs, 3 * D_IN_PAD, n_local, 1, BM))) return rc;
    if ((rc = make_map(&s->tmW1, s->w1, 3 * D_IN_PAD, H1, nb, 256))) return rc;
    if ((rc = make_map(&s->tmA2, s->a2, 3 * H1, n_local, nb, BM))) return rc;
    if ((rc = make_map(&s->tmW2, s->w2, 3 * H1, H2, nb, 256))) return rc;
    if ((rc = make_map(&s->tmA3, s->a3, 3 * H2, n_local, nb, BM))) return rc;
    if ((rc = make_map(&s->tmW3, s->w3, 3 * H2, H3, nb, 128))) return rc;
    if ((rc = make_map(&s->tmA4, s->a4, 3 * H3, n_local, nb, BM))) return rc;
    if ((rc = make_map(&s->tmW4, s->w4, 3 * H3, NCLS_PAD, nb, NCLS_PAD))) return rc;
    return PMP_OK;
}

// fills d_lt[p] = -(sum NLL over all shards / n_global) / scale for every node; d_logw gets the MP-FC kernel term
int pmp_fc_loglik(pmp_ctx* c) {
    PMP_REQUIRE(c->fc, "FC data not set (pmp_set_data_fc)");
    PMP_REQUIRE(c->cfg.dim == THETA_DIM, "FC target needs dim = %lld (784-512-256-128-10 MLP), got %d", THETA_DIM, c->cfg.dim);
    FcState* s = reinterpret_cast<FcState*>(c->fc);
    const int P = c->P, M = (int)s->n_local;
    int rc;
    if (!s->loss) PMP_CUDA(cudaMalloc((void**)&s->loss, (size_t)MAX_NODES * sizeof(unsigned long long)));
    PMP_CUDA(cudaMemsetAsync(s->loss, 0, (size_t)P * sizeof(unsigned long long), c->stream));
    const int bias_stride = H1 + H2 + H3 + NCLS_PAD;
    for (int p0 = 0; p0 < P; p0 += s->nb) {
        const int nb = (P - p0) < s->nb ? (P - p0) : s->nb;
        const float* th = c->d_props + (long long)p0 * THETA_DIM;
        long long t;
        t = (long long)nb * H1 * D_IN_PAD;   split_w_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, THETA_DIM, OFF_W1, H1, D_IN, H1, D_IN_PAD, s->w1, nb);
        t = (long long)nb * H2 * H1;         split_w_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, THETA_DIM, OFF_W2, H2, H1, H2, H1, s->w2, nb);
        t = (long long)nb * H3 * H2;         split_w_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, THETA_DIM, OFF_W3, H3, H2, H3, H2, s->w3, nb);
        t = (long long)nb * NCLS_PAD * H3;   split_w_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, THETA_DIM, OFF_W4, NCLS, H3, NCLS_PAD, H3, s->w4, nb);
        t = (long long)nb * bias_stride;     gather_bias_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, THETA_DIM, s->bias, nb);
        c->launches += 5;
        PMP_CUDA(cudaGetLastError());
        GemmArgs g1{M, 3 * D_IN_PAD, 1, s->bias, bias_stride, s->a2, H1, nullptr, nullptr};
        if ((rc = launch_gemm<256, EPI_RELU_SPLIT>(c, s->tmX, s->tmW1, g1, H1, nb))) return rc;
        GemmArgs g2{M, 3 * H1, 0, s->bias + H1, bias_stride, s->a3, H2, nullptr, nullptr};
        if ((rc = launch_gemm<256, EPI_RELU_SPLIT>(c, s->tmA2, s->tmW2, g2, H2, nb))) return rc;
        GemmArgs g3{M, 3 * H2, 0, s->bias + H1 + H2, bias_stride, s->a4, H3, nullptr, nullptr};
        if ((rc = launch_gemm<128, EPI_RELU_SPLIT>(c, s->tmA3, s->tmW3, g3, H3, nb))) return rc;
        GemmArgs g4{M, 3 * H3, 0, s->bias + H1 + H2 + H3, bias_stride, nullptr, NCLS_PAD, s->labels, s->loss + p0};
        if ((rc = launch_gemm<NCLS_PAD, EPI_NLL>(c, s->tmA4, s->tmW4, g4, NCLS_PAD, nb))) return rc;
    }
    if ((rc = pmp_allreduce_u64(c, s->loss, (size_t)P))) return rc;
    finalize_loss_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(s->loss, c->d_lt, P, (double)s->n_global, 1.0 / (double)c->cfg.scale);
    c->launches++;
    if (c->cfg.algo == PMP_ALGO_MP && !(c->cfg.flags & PMP_FLAG_NO_KERNEL_TERM)) {
        if (!s->s1) { PMP_CUDA(cudaMalloc((void**)&s->s1, (size_t)THETA_DIM * sizeof(float))); PMP_CUDA(cudaMalloc((void**)&s->dj2, MAX_NODES * sizeof(double))); PMP_CUDA(cudaMalloc((void**)&s->dot, MAX_NODES * sizeof(double))); }
        PMP_CUDA(cudaMemsetAsync(s->dj2, 0, P * sizeof(double), c->stream));
        PMP_CUDA(cudaMemsetAsync(s->dot, 0, P * sizeof(double), c->stream));
        fc_s1_kernel<<<(unsigned)((THETA_DIM + 255) / 256), 256, 0, c->stream>>>(c->d_props, THETA_DIM, P, s->s1);
        fc_dots_kernel<<<dim3(64, P), 256, 0, c->stream>>>(c->d_props, s->s1, THETA_DIM, s->dj2, s->dot);
        fc_kterm_kernel<<<1, 256, 0, c->stream>>>(s->dj2, s->dot, P, (double)THETA_DIM, (double)c->cfg.kernel_sigma, c->d_logw);
        c->launches += 3;
    }
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

}  // extern "C"
