// cnn_sweep.cuh — log-target sweep of the Bayesian CNN: for P candidate weight vectors, -CrossEntropy(CNN_p(X), y).
// Included at the end of fc_sweep.cu (same translation unit: it reuses fc_gemm2_kernel, the tensor maps and the split kernels).
//
// Replaces `weights[all] = exp(-loss(proposal_nets[all]))` of complex_nets/Mnist/CNN/PMP_CNN.py:119-120 (MP_CNN.py:120, MH_CNN.py:105) with
// Model = conv1(1->10, 5x5) - ReLU - maxpool 2x2 - conv2(10->20, 3x3) - ReLU - fc1(2000->500) - ReLU - fc2(500->10) - log_softmax
// (PMP_CNN.py:22-44) and loss = CrossEntropyLoss(mean)(net(X), y) / 10 (PMP_CNN.py:48-52; the cross-entropy applies a second log-softmax to
// the model's log-probabilities, mirrored here).  Parameter layout = torch order: conv1.weight[10,1,5,5], conv1.bias[10], conv2.weight[20,10,3,3],
// conv2.bias[20], fc1.weight[500,2000], fc1.bias[500], fc2.weight[10,500], fc2.bias[10] = 1 007 590 floats (CNN_model.pkl).
//
// Work per (node, image): conv1 144 000 MACs, conv2 180 000, fc1 1 000 000, fc2 5 000.
//   * fc1 (75 % of the flops) is a dense [n, 2000] x [2000, 500] contraction per node: fc_gemm2_kernel (TMA-fed tcgen05 CTA-pair GEMM, bf16x3,
//     fp32 accumulation in TMEM) with the 500 -> 10 layer folded into its epilogue (EPI2_HEAD) — the hidden activations never leave the SM;
//   * the two convolutions run on the CUDA cores in float32 (cnn_conv_kernel, packed fma.rn.f32x2).  As GEMMs they are K = 25 / 90 deep and
//     N = 10 / 20 wide per node: one output element per 25-90 MACs.  On the tensor pipe every output element still costs ~10 CUDA-core
//     instructions of epilogue / im2col staging (bias, ReLU, 2x2 pool across TMEM lanes, bf16 re-split, layout) against 12-45 packed FMAs for
//     computing it directly, the A tile of an N <= 32 UMMA is shared-memory-bandwidth bound (4 KB per 128x32x16 MMA), and conv1's pooling needs
//     four M tiles per window in TMEM — measured arithmetic in DESIGN.md 4.9.  The direct kernel keeps the image tile, the pooled conv1
//     activations and the node's filters in shared memory and writes conv2's output straight into the tile-major [h | l] A operand of fc1.
#pragma once

namespace pmp {
namespace fc {

constexpr long long CNN_DIM = 1007590;
constexpr long long C_OFF_W1 = 0, C_OFF_B1 = 250, C_OFF_W2 = 260, C_OFF_B2 = 2060, C_OFF_FW1 = 2080, C_OFF_FB1 = 1002080, C_OFF_FW2 = 1002580, C_OFF_FB2 = 1007580;
static_assert(C_OFF_FB2 + 10 == CNN_DIM, "CNN theta layout");
constexpr int CNN_PIX = 784, CNN_C1 = 10, CNN_C2 = 20, CNN_FLAT = 2000, CNN_FLAT_PAD = 2048, CNN_HID = 500, CNN_HID_PAD = 512;
constexpr int CNN_G = 8;                    // images per CTA pass
constexpr int CNN_THREADS = 256;
constexpr int CNN_SMEM = CNN_G * CNN_PIX * 4 + CNN_G * 1440 * 4 + 25 * 12 * 4 + 16 * 4 + 90 * 20 * 8 + 32 * 4;

struct CnnConvArgs {
    const float* X;            // [n, 784] float32
    int n;
    const float* theta;        // first node of the batch
    long long theta_stride;
    int nb;
    __nv_bfloat16* out;        // fc1's A operand, tile-major [nb][mb128][64 k-tiles: 32 h, 32 l][128][64]
    int mb128;
};

// One CTA = one node of the batch (its filters stay in shared memory) x a strided range of CNN_G-image groups.
__global__ void __launch_bounds__(CNN_THREADS, 2) cnn_conv_kernel(const CnnConvArgs g) {
    extern __shared__ __align__(16) uint8_t cnn_smem[];
    float* s_img = reinterpret_cast<float*>(cnn_smem);                       // [G][784]
    float* s_p1 = s_img + CNN_G * CNN_PIX;                                  // [G][10][12][12]
    float* s_w1 = s_p1 + CNN_G * 1440;                                      // [25 taps][12]: 10 channels + 2 zero pads
    float* s_b1 = s_w1 + 25 * 12;                                           // [16]
    float2* s_w2 = reinterpret_cast<float2*>(s_b1 + 16);                    // [90 taps = c*9 + ky*3 + kx][20 channels] as (w, w)
    float* s_b2 = reinterpret_cast<float*>(s_w2 + 90 * 20);                 // [32]

    const int tid = threadIdx.x;
    const int b = blockIdx.x % g.nb, slot = blockIdx.x / g.nb, nslots = gridDim.x / g.nb;
    const float* th = g.theta + (long long)b * g.theta_stride;
    for (int i = tid; i < 25 * 12; i += CNN_THREADS) { const int tap = i / 12, c = i - tap * 12; s_w1[i] = c < CNN_C1 ? __ldg(th + C_OFF_W1 + c * 25 + tap) : 0.f; }
    if (tid < 16) s_b1[tid] = tid < CNN_C1 ? __ldg(th + C_OFF_B1 + tid) : 0.f;
    for (int i = tid; i < 90 * 20; i += CNN_THREADS) { const int tap = i / 20, o = i - tap * 20; const float w = __ldg(th + C_OFF_W2 + o * 90 + tap); s_w2[i] = make_float2(w, w); }
    if (tid < 32) s_b2[tid] = tid < CNN_C2 ? __ldg(th + C_OFF_B2 + tid) : 0.f;

    const int ngroups = (g.n + CNN_G - 1) / CNN_G;
    for (int grp = slot; grp < ngroups; grp += nslots) {
        const int img0 = grp * CNN_G;
        const int cnt = (g.n - img0) < CNN_G ? (g.n - img0) : CNN_G;
        __syncthreads();                                                     // the previous pass is done with s_img / s_p1 (and the filters are in)
        {
            const float4* src = reinterpret_cast<const float4*>(g.X + (long long)img0 * CNN_PIX);
            float4* dst = reinterpret_cast<float4*>(s_img);
            for (int i = tid; i < cnt * (CNN_PIX / 4); i += CNN_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        // ---- conv1 (5x5, 1 -> 10) + bias + ReLU + 2x2 max pool: one task = one pooled position, all 10 channels; FFMA2 over channel pairs ----
        for (int t = tid; t < cnt * 144; t += CNN_THREADS) {
            const int gi = t / 144, w = t - gi * 144, py = w / 12, px = w - py * 12;
            const float* im = s_img + gi * CNN_PIX + (2 * py) * 28 + 2 * px;
            float patch[6][6];
#pragma unroll
            for (int r = 0; r < 6; ++r) {
#pragma unroll
                for (int q = 0; q < 3; ++q) { const float2 v = *reinterpret_cast<const float2*>(im + r * 28 + 2 * q); patch[r][2 * q] = v.x; patch[r][2 * q + 1] = v.y; }
            }
            float2 acc[4][5];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[p][j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) {
                    const float4* wp = reinterpret_cast<const float4*>(s_w1 + (ky * 5 + kx) * 12);
                    const float4 wa = wp[0], wb = wp[1], wc = wp[2];
                    const float2 wj[5] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w), make_float2(wc.x, wc.y)};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float v = patch[ky + (p >> 1)][kx + (p & 1)];
                        const float2 vv = make_float2(v, v);
#pragma unroll
                        for (int j = 0; j < 5; ++j) acc[p][j] = ffma2f(vv, wj[j], acc[p][j]);
                    }
                }
            }
            // max over the window first, then bias + ReLU: x -> relu(x + b) is monotone, so this equals pool(relu(conv + b)) bit for bit
            float* o = s_p1 + gi * 1440 + py * 12 + px;
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const float m0 = fmaxf(fmaxf(acc[0][j].x, acc[1][j].x), fmaxf(acc[2][j].x, acc[3][j].x));
                const float m1 = fmaxf(fmaxf(acc[0][j].y, acc[1][j].y), fmaxf(acc[2][j].y, acc[3][j].y));
                o[(2 * j) * 144] = fmaxf(m0 + s_b1[2 * j], 0.f);
                o[(2 * j + 1) * 144] = fmaxf(m1 + s_b1[2 * j + 1], 0.f);
            }
        }
        __syncthreads();
        // ---- conv2 (3x3, 10 -> 20) + bias + ReLU: one task = a 2x2 block of output positions x 10 channels; FFMA2 over horizontal position pairs ----
        for (int t = tid; t < cnt * 50; t += CNN_THREADS) {
            const int gi = t / 50, r = t - gi * 50, hh = r / 25, win = r - hh * 25, wy = win / 5, wx = win - wy * 5;
            const float* p1 = s_p1 + gi * 1440 + (2 * wy) * 12 + 2 * wx;
            float2 acc[2][10];
#pragma unroll
            for (int oy = 0; oy < 2; ++oy)
#pragma unroll
                for (int j = 0; j < 10; ++j) acc[oy][j] = make_float2(0.f, 0.f);
#pragma unroll 2
            for (int c = 0; c < CNN_C1; ++c) {
                float v[4][4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const float2 a = *reinterpret_cast<const float2*>(p1 + c * 144 + rr * 12), bq = *reinterpret_cast<const float2*>(p1 + c * 144 + rr * 12 + 2);
                    v[rr][0] = a.x; v[rr][1] = a.y; v[rr][2] = bq.x; v[rr][3] = bq.y;
                }
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float4* wp = reinterpret_cast<const float4*>(s_w2 + (c * 9 + ky * 3 + kx) * 20 + hh * 10);
                        float2 wd[10];
#pragma unroll
                        for (int j = 0; j < 5; ++j) { const float4 q = wp[j]; wd[2 * j] = make_float2(q.x, q.y); wd[2 * j + 1] = make_float2(q.z, q.w); }
#pragma unroll
                        for (int oy = 0; oy < 2; ++oy) {
                            const float2 in = make_float2(v[oy + ky][kx], v[oy + ky][kx + 1]);
#pragma unroll
                            for (int j = 0; j < 10; ++j) acc[oy][j] = ffma2f(in, wd[j], acc[oy][j]);
                        }
                    }
                }
            }
            // bias + ReLU, split to [h | l] bf16 and store into fc1's tile-major A operand: k = o*100 + y*10 + x (torch's view(in_size, -1))
            const int row = img0 + gi;
            __nv_bfloat16* oblk = g.out + ((long long)b * g.mb128 + (row >> 7)) * 64ll * 8192ll + (long long)(row & 127) * 64;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
                const int o = hh * 10 + j;
                const float bo = s_b2[o];
#pragma unroll
                for (int oy = 0; oy < 2; ++oy) {
                    const float a0 = fmaxf(acc[oy][j].x + bo, 0.f), a1 = fmaxf(acc[oy][j].y + bo, 0.f);
                    const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
                    const __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
                    const int k = o * 100 + (2 * wy + oy) * 10 + 2 * wx;           // even: the pair stays inside one 64-wide k-tile
                    __nv_bfloat16* d = oblk + (long long)(k >> 6) * 8192 + (k & 63);
                    *reinterpret_cast<uint32_t*>(d) = pack_bf16x2(h0, h1);
                    *reinterpret_cast<uint32_t*>(d + 32ll * 8192) = pack_bf16x2(l0, l1);
                }
            }
        }
    }
}

__global__ void cnn_gather_bias_kernel(const float* __restrict__ theta, long long theta_stride, float* __restrict__ out, int nb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb * CNN_HID_PAD) return;
    const int b = i / CNN_HID_PAD, j = i - b * CNN_HID_PAD;
    out[i] = j < CNN_HID ? theta[b * theta_stride + C_OFF_FB1 + j] : 0.f;
}

// logits = partial(column block 0) + partial(column block 1) + bias, fixed order; log_softmax (the model's, PMP_CNN.py:43) and the cross-entropy's own
// log-softmax on top of it (PMP_CNN.py:50-51); per-row NLL -> 2^-32 fixed point, integer sums from there on (exact, order-free, shard-invariant)
__global__ void cnn_head_nll_kernel(const float* __restrict__ part, int nparts, int M, const float* __restrict__ theta, long long theta_stride,
                                    const int* __restrict__ labels, unsigned long long* __restrict__ loss) {
    const int b = blockIdx.y;
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    long long q = 0;
    if (row < M) {
        float z[NCLS];
        const float* b4 = theta + (long long)b * theta_stride + C_OFF_FB2;
#pragma unroll
        for (int c = 0; c < NCLS; ++c) z[c] = 0.f;
        for (int p = 0; p < nparts; ++p) {
            const float4* src = reinterpret_cast<const float4*>(part + (((long long)b * nparts + p) * M + row) * HEAD_PITCH);
            const float4 a = __ldg(src), bb = __ldg(src + 1), cc = __ldg(src + 2);
            z[0] += a.x; z[1] += a.y; z[2] += a.z; z[3] += a.w; z[4] += bb.x; z[5] += bb.y; z[6] += bb.z; z[7] += bb.w; z[8] += cc.x; z[9] += cc.y;
        }
#pragma unroll
        for (int c = 0; c < NCLS; ++c) z[c] += __ldg(b4 + c);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            float mx = z[0];
#pragma unroll
            for (int c = 1; c < NCLS; ++c) mx = fmaxf(mx, z[c]);
            float se = 0.f;
#pragma unroll
            for (int c = 0; c < NCLS; ++c) se += expf(z[c] - mx);
            const float lse = mx + logf(se);
#pragma unroll
            for (int c = 0; c < NCLS; ++c) z[c] -= lse;
        }
        const int lab = labels[row];
        float zl = z[0];
#pragma unroll
        for (int c = 1; c < NCLS; ++c) zl = (lab == c) ? z[c] : zl;
        q = __double2ll_rn((double)(-zl) * 4294967296.0);
    }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if ((threadIdx.x & 31) == 0 && q != 0) atomicAdd(loss + b, (unsigned long long)q);
}

struct CnnState {
    long long n_local = 0, n_global = 0;
    float* x32 = nullptr;              // [n, 784]
    int* labels = nullptr;
    int nb = 0;
    __nv_bfloat16* a2 = nullptr;       // conv2 output = fc1's A operand, [nb][mb128][64][128][64]
    __nv_bfloat16* w = nullptr;        // fc1 weights [nb][512][2 * 2048] = [h | l]
    float* bias = nullptr;             // [nb][512]
    float* part = nullptr;             // [nb][2][n][12]
    unsigned long long* loss = nullptr;
    CUtensorMap tmA, tmW;
};

static void free_cnn(CnnState* s) {
    void* ptrs[] = {s->x32, s->labels, s->a2, s->w, s->bias, s->part, s->loss};
    for (void* p : ptrs) if (p) cudaFree(p);
}

}  // namespace fc
}  // namespace pmp

extern "C" {

int pmp_cnn_destroy(pmp_ctx* c) {
    if (c->cnn) { free_cnn(reinterpret_cast<CnnState*>(c->cnn)); delete reinterpret_cast<CnnState*>(c->cnn); c->cnn = nullptr; }
    return PMP_OK;
}

int pmp_set_data_cnn(pmp_ctx* c, const float* X, const int64_t* labels, int64_t n_local, int64_t n_offset, int64_t n_global) {
    PMP_REQUIRE(c && X && labels && n_local > 0 && n_global >= n_local, "bad arguments");
    (void)n_offset;
    PMP_CUDA(cudaSetDevice(c->device));
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    pmp_cnn_destroy(c);
    CnnState* s = new CnnState();
    c->cnn = s;
    s->n_local = n_local; s->n_global = n_global;
    s->nb = getenv("PMP_CNN_BATCH") ? atoi(getenv("PMP_CNN_BATCH")) : 8;
    if (s->nb < 1) s->nb = 1;
    if (s->nb > 2 * c->sm_count) s->nb = 2 * c->sm_count;
    std::vector<int> lab32((size_t)n_local);
    for (int64_t i = 0; i < n_local; ++i) { PMP_REQUIRE(labels[i] >= 0 && labels[i] < NCLS, "label %lld out of range at row %lld", (long long)labels[i], (long long)i); lab32[i] = (int)labels[i]; }
    const long long mb128 = (n_local + 127) / 128;
    const size_t a2_bytes = (size_t)s->nb * mb128 * 64 * 8192 * sizeof(__nv_bfloat16);
    PMP_CUDA(cudaMalloc((void**)&s->x32, (size_t)n_local * CNN_PIX * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->labels, (size_t)n_local * sizeof(int)));
    PMP_CUDA(cudaMalloc((void**)&s->a2, a2_bytes));
    PMP_CUDA(cudaMalloc((void**)&s->w, (size_t)s->nb * CNN_HID_PAD * 2 * CNN_FLAT_PAD * sizeof(__nv_bfloat16)));
    PMP_CUDA(cudaMalloc((void**)&s->bias, (size_t)s->nb * CNN_HID_PAD * sizeof(float)));
    PMP_CUDA(cudaMalloc((void**)&s->part, (size_t)s->nb * 2 * n_local * HEAD_PITCH * sizeof(float)));
    PMP_CUDA(cudaMemcpyAsync(s->x32, X, (size_t)n_local * CNN_PIX * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaMemcpyAsync(s->labels, lab32.data(), (size_t)n_local * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    PMP_CUDA(cudaMemsetAsync(s->a2, 0, a2_bytes, c->stream));              // k in [2000, 2048) and rows >= n of the last block stay zero for ever
    PMP_CUDA(cudaStreamSynchronize(c->stream));
    int rc;
    if ((rc = make_map_tiled(&s->tmA, s->a2, 2 * CNN_FLAT_PAD / 64, mb128, s->nb))) return rc;
    if ((rc = make_map(&s->tmW, s->w, 2 * CNN_FLAT_PAD, CNN_HID_PAD, s->nb, 128))) return rc;
    return PMP_OK;
}

// fills d_lt[p] = -(sum NLL over all shards / n_global) / scale for every node; d_logw gets the MP kernel term
int pmp_cnn_loglik(pmp_ctx* c) {
    PMP_REQUIRE(c->cnn, "CNN data not set (pmp_set_data_cnn)");
    PMP_REQUIRE(c->cfg.dim == CNN_DIM, "CNN target needs dim = %lld (PMP_CNN.py:22-44), got %d", CNN_DIM, c->cfg.dim);
    CnnState* s = reinterpret_cast<CnnState*>(c->cnn);
    const int P = c->P, M = (int)s->n_local;
    const int mb128 = (M + 127) / 128;
    int rc;
    if (!s->loss) PMP_CUDA(cudaMalloc((void**)&s->loss, (size_t)MAX_NODES * sizeof(unsigned long long)));
    PMP_CUDA(cudaMemsetAsync(s->loss, 0, (size_t)P * sizeof(unsigned long long), c->stream));
    static bool attr = false;
    if (!attr) { PMP_CUDA(cudaFuncSetAttribute(cnn_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CNN_SMEM)); attr = true; }
    for (int p0 = 0; p0 < P; p0 += s->nb) {
        const int nb = (P - p0) < s->nb ? (P - p0) : s->nb;
        const float* th = c->d_props + (long long)p0 * CNN_DIM;
        long long t = (long long)nb * CNN_HID_PAD * CNN_FLAT_PAD;
        split2_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, CNN_DIM, C_OFF_FW1, CNN_HID, CNN_FLAT, CNN_HID_PAD, CNN_FLAT_PAD, s->w, nb);
        cnn_gather_bias_kernel<<<(nb * CNN_HID_PAD + 255) / 256, 256, 0, c->stream>>>(th, CNN_DIM, s->bias, nb);
        CnnConvArgs ca{s->x32, M, th, CNN_DIM, nb, s->a2, mb128};
        int per_node = (2 * c->sm_count) / nb;
        const int ngroups = (M + CNN_G - 1) / CNN_G;
        if (per_node > ngroups) per_node = ngroups;
        if (per_node < 1) per_node = 1;
        cnn_conv_kernel<<<dim3((unsigned)(per_node * nb)), CNN_THREADS, CNN_SMEM, c->stream>>>(ca);
        c->launches += 3;
        PMP_CUDA(cudaGetLastError());
        Gemm2Args g{};
        g.M = M; g.Kpad = CNN_FLAT_PAD; g.a_shared = 0; g.n_total = CNN_HID_PAD; g.nb = nb; g.bias = s->bias; g.bias_stride = CNN_HID_PAD; g.a_tiled = 1; g.mb128 = mb128;
        g.theta_stride = CNN_DIM; g.head_w = th + C_OFF_FW2; g.head_k = CNN_HID; g.head_part = s->part;
        if ((rc = launch_gemm2<256, EPI2_HEAD, 3>(c, s->tmA, s->tmW, g))) return rc;
        cnn_head_nll_kernel<<<dim3((unsigned)((M + 255) / 256), (unsigned)nb), 256, 0, c->stream>>>(s->part, CNN_HID_PAD / 256, M, th, CNN_DIM, s->labels, s->loss + p0);
        c->launches++;
        PMP_CUDA(cudaGetLastError());
    }
    if ((rc = pmp_allreduce_u64(c, s->loss, (size_t)P))) return rc;
    finalize_loss_kernel<<<(P + 255) / 256, 256, 0, c->stream>>>(s->loss, c->d_lt, P, (double)s->n_global, 1.0 / (double)c->cfg.scale);
    c->launches++;
    if (c->cfg.algo == PMP_ALGO_MP && !(c->cfg.flags & PMP_FLAG_NO_KERNEL_TERM)) { if ((rc = pmp_large_dim_kernel_term(c))) return rc; }
    PMP_CUDA(cudaGetLastError());
    return PMP_OK;
}

}  // extern "C"
