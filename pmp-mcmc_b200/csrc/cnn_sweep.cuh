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
constexpr int CNN_G = 8;                    // images per pass
#ifndef CNN_THREADS_DEF
#define CNN_THREADS_DEF 256
#endif
constexpr int CNN_THREADS = CNN_THREADS_DEF;            // 8 warps, two per SM sub-partition; one CTA per SM (12 warps measured slower: the 23 chunks of a step leave a longer tail)
constexpr int CNN_P1_ROW = 13, CNN_P1_CH = 12 * CNN_P1_ROW, CNN_P1_STRIDE = 10 * CNN_P1_CH + 10;   // pooled conv1 activations [image][channel][12 rows][13]: rows of 13 floats and an image stride = 2 (mod 32)
                                            // put the ten 2x5 position blocks of an image, and those of the next image in the same warp, on distinct banks (conv2's input loads were 2-way conflicted: ncu r3g)
constexpr int CNN_IMG_BYTES = CNN_G * CNN_PIX * 4, CNN_P1_BYTES = CNN_G * CNN_P1_STRIDE * 4, CNN_STAGE_BYTES = CNN_G * 2 * CNN_FLAT_PAD * 2;
constexpr int CNN_SMEM = 2 * CNN_IMG_BYTES + 2 * CNN_P1_BYTES + CNN_STAGE_BYTES + 25 * 12 * 4 + 16 * 4 + 90 * 2 * 12 * 4 + 32 * 4;
static_assert(CNN_SMEM <= 226 * 1024, "shared-memory plan of cnn_conv_kernel");
constexpr int CNN_CHUNKS1 = CNN_G * 72 / 32, CNN_CHUNKS2 = CNN_G * 20 / 32;      // warp-sized task chunks per full pass: 18 (conv1) and 5 (conv2)

struct CnnConvArgs {
    const float* X;            // [n, 784] float32
    int n;
    const float* theta;        // first node of the batch
    long long theta_stride;
    int nb;
    __nv_bfloat16* out;        // fc1's A operand, tile-major [nb][mb128][64 k-tiles: 32 h, 32 l][128][64]; k' = (y*10 + x)*20 + o (cnn_split_w_kernel permutes fc1.weight to match)
    int mb128;
};

__device__ __forceinline__ float2 dup2(float v) { return make_float2(v, v); }

// conv1 (5x5, 1 -> 10) + bias + ReLU + 2x2 max pool: one task = two horizontally adjacent pool windows (2 x 4 conv positions), all 10 channels
__device__ __forceinline__ void cnn_conv1_task(int t, const float* s_img, float* s_p1, const float* s_w1, const float* s_b1) {
    const int gi = t / 72, w = t - gi * 72, py = w / 6, pxp = w - py * 6;
    const float* im = s_img + gi * CNN_PIX + (2 * py) * 28 + 4 * pxp;
    float patch[6][8];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(im + r * 28), c4 = *reinterpret_cast<const float4*>(im + r * 28 + 4);
        patch[r][0] = a.x; patch[r][1] = a.y; patch[r][2] = a.z; patch[r][3] = a.w; patch[r][4] = c4.x; patch[r][5] = c4.y; patch[r][6] = c4.z; patch[r][7] = c4.w;
    }
    float2 acc[8][5];
#pragma unroll
    for (int p = 0; p < 8; ++p)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[p][j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 5; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
            const float* wp = s_w1 + (ky * 5 + kx) * 12;
            const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
            const float2 wc = *reinterpret_cast<const float2*>(wp + 8);
            const float2 wj[5] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w), wc};
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const float2 vv = dup2(patch[ky + (p >> 2)][kx + (p & 3)]);
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[p][j] = ffma2f(vv, wj[j], acc[p][j]);
            }
        }
    }
    // max over the window first, then bias + ReLU: x -> relu(x + b) is monotone, so this equals pool(relu(conv + b)) bit for bit
    float* o = s_p1 + gi * CNN_P1_STRIDE + py * CNN_P1_ROW + 2 * pxp;
#pragma unroll
    for (int wdw = 0; wdw < 2; ++wdw) {
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float2 a0 = acc[2 * wdw][j], a1 = acc[2 * wdw + 1][j], a2 = acc[4 + 2 * wdw][j], a3 = acc[4 + 2 * wdw + 1][j];
            o[(2 * j) * CNN_P1_CH + wdw] = fmaxf(fmaxf(fmaxf(a0.x, a1.x), fmaxf(a2.x, a3.x)) + s_b1[2 * j], 0.f);
            o[(2 * j + 1) * CNN_P1_CH + wdw] = fmaxf(fmaxf(fmaxf(a0.y, a1.y), fmaxf(a2.y, a3.y)) + s_b1[2 * j + 1], 0.f);
        }
    }
}

// conv2 (3x3, 10 -> 20) + bias + ReLU: one task = 2 rows x 5 columns of output positions x 10 channels; the result goes to the staging area as
// [h | l] bf16 with k' = (y*10 + x)*20 + o (the ten channels of one position are 20 contiguous bytes)
__device__ __forceinline__ void cnn_conv2_task(int t, const float* s_p1, __nv_bfloat16* s_out, const float* s_w2, const float* s_b2) {
    const int gi = t / 20, r20 = t - gi * 20, hh = r20 / 10, blk = r20 - hh * 10, ry = blk >> 1, cx = blk & 1;
    float2 acc[10][5];
#pragma unroll
    for (int p = 0; p < 10; ++p)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[p][j] = make_float2(0.f, 0.f);
    const float* p1 = s_p1 + gi * CNN_P1_STRIDE + (2 * ry) * CNN_P1_ROW + 5 * cx;
#pragma unroll 1
    for (int c = 0; c < CNN_C1; ++c) {
        float v[4][7];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
            for (int q = 0; q < 7; ++q) v[rr][q] = p1[c * CNN_P1_CH + rr * CNN_P1_ROW + q];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* wp = s_w2 + ((c * 9 + ky * 3 + kx) * 2 + hh) * 12;
                const float4 wa = *reinterpret_cast<const float4*>(wp), wb = *reinterpret_cast<const float4*>(wp + 4);
                const float2 wc = *reinterpret_cast<const float2*>(wp + 8);
                const float2 wj[5] = {make_float2(wa.x, wa.y), make_float2(wa.z, wa.w), make_float2(wb.x, wb.y), make_float2(wb.z, wb.w), wc};
#pragma unroll
                for (int p = 0; p < 10; ++p) {
                    const float2 vv = dup2(v[(p / 5) + ky][(p % 5) + kx]);
#pragma unroll
                    for (int j = 0; j < 5; ++j) acc[p][j] = ffma2f(vv, wj[j], acc[p][j]);
                }
            }
        }
    }
    __nv_bfloat16* so = s_out + gi * (2 * CNN_FLAT_PAD);
#pragma unroll
    for (int p = 0; p < 10; ++p) {
        const int pos = (2 * ry + p / 5) * 10 + 5 * cx + (p % 5);
        uint32_t* dh = reinterpret_cast<uint32_t*>(so + pos * 20 + hh * 10);
        uint32_t* dl = reinterpret_cast<uint32_t*>(so + CNN_FLAT_PAD + pos * 20 + hh * 10);
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const float a0 = fmaxf(acc[p][j].x + s_b2[hh * 10 + 2 * j], 0.f), a1 = fmaxf(acc[p][j].y + s_b2[hh * 10 + 2 * j + 1], 0.f);
            const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
            dh[j] = pack_bf16x2(h0, h1);
            dl[j] = pack_bf16x2(__float2bfloat16_rn(a0 - __bfloat162float(h0)), __float2bfloat16_rn(a1 - __bfloat162float(h1)));
        }
    }
}

// One CTA per SM = one node of the batch (its filters stay in shared memory) x a strided range of CNN_G-image groups, software-pipelined over the groups:
// in one compute phase the CTA's 8 warps pull warp-sized task chunks from a shared counter — first the 5 conv2 chunks of group i (reading its pooled
// activations, writing the staging area), then the 18 conv1 chunks of group i+1 (reading its image tile, writing ITS pooled activations) — so every SM
// sub-partition stays busy whatever the chunk sizes (a static split leaves one of the four idle for half of each phase: ncu r3d, FMA pipe 54 %);
// then the staging area is copied out with full 128-byte rows while the image tile of group i+2 is loaded.
//
// Shared-memory bandwidth, not the FMA pipe, is what a direct convolution runs out of first (ncu r3c on the first version: shared wavefronts 71 %, FMA
// pipe 51 %): a warp-wide load delivers 128 B per clock to the register file even when every lane reads the same filter tap.  Both convolutions therefore
// pack two OUTPUT CHANNELS per fma.rn.f32x2 (the filter taps are read as natural (w_c, w_c+1) pairs), every thread owns 8 (conv1) / 10 (conv2) output
// positions so that one 10-channel tap vector (10 registers from shared memory) feeds 40 / 50 packed FMAs.
__global__ void __launch_bounds__(CNN_THREADS, 1) cnn_conv_kernel(const CnnConvArgs g) {
    extern __shared__ __align__(16) uint8_t cnn_smem[];
    float* s_img0 = reinterpret_cast<float*>(cnn_smem);                                       // 2 x [G][784]
    float* s_p10 = reinterpret_cast<float*>(cnn_smem + 2 * CNN_IMG_BYTES);                    // 2 x [G][10][12][13] (+10 pad per image)
    __nv_bfloat16* s_out = reinterpret_cast<__nv_bfloat16*>(cnn_smem + 2 * CNN_IMG_BYTES + 2 * CNN_P1_BYTES);   // staging [G][2 planes][2048]
    float* s_w1 = reinterpret_cast<float*>(cnn_smem + 2 * CNN_IMG_BYTES + 2 * CNN_P1_BYTES + CNN_STAGE_BYTES);  // [25 taps][12]: 10 channels + 2 pads
    float* s_b1 = s_w1 + 25 * 12;                                                             // [16]
    float* s_w2 = s_b1 + 16;                                                                  // [90 taps = c*9 + ky*3 + kx][2 channel halves][12]
    float* s_b2 = s_w2 + 90 * 2 * 12;                                                         // [32]
    __shared__ int s_ctr;

    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t ctr_addr = smem_u32(&s_ctr);
    const int b = blockIdx.x % g.nb, slot = blockIdx.x / g.nb, nslots = ((int)gridDim.x - b + g.nb - 1) / g.nb;   // CTAs b, b + nb, ... serve node b (every SM is used even when nb does not divide the SM count)
    const float* th = g.theta + (long long)b * g.theta_stride;
    for (int i = tid; i < 25 * 12; i += CNN_THREADS) { const int tap = i / 12, c = i - tap * 12; s_w1[i] = c < CNN_C1 ? __ldg(th + C_OFF_W1 + c * 25 + tap) : 0.f; }
    if (tid < 16) s_b1[tid] = tid < CNN_C1 ? __ldg(th + C_OFF_B1 + tid) : 0.f;
    for (int i = tid; i < 90 * 24; i += CNN_THREADS) { const int tap = i / 24, r = i - tap * 24, hh = r / 12, j = r - hh * 12; s_w2[i] = j < 10 ? __ldg(th + C_OFF_W2 + (hh * 10 + j) * 90 + tap) : 0.f; }
    if (tid < 32) s_b2[tid] = tid < CNN_C2 ? __ldg(th + C_OFF_B2 + tid) : 0.f;

    const int ngroups = (g.n + CNN_G - 1) / CNN_G;
    auto group_cnt = [&](int grp) { const int rem = g.n - grp * CNN_G; return rem < CNN_G ? rem : CNN_G; };
    // asynchronous (cp.async, 16 bytes per request): the tile of pass s+1 lands while pass s computes
    auto load_images = [&](int grp, float* dst_img) {
        const float* src = g.X + (long long)grp * CNN_G * CNN_PIX;
        const uint32_t dst = smem_u32(dst_img);
        const int nv = group_cnt(grp) * (CNN_PIX / 4);
        for (int i = tid; i < nv; i += CNN_THREADS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)i), "l"(src + 4 * i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // pipeline step s = 0 .. npass: conv2 of my pass s-1 and conv1 of my pass s share one compute phase
    const int npass = slot < ngroups ? (ngroups - slot + nslots - 1) / nslots : 0;
    if (npass > 0) load_images(slot, s_img0);
    if (tid == 0) s_ctr = 0;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int s = 0; s <= npass; ++s) {
        const int cur = s & 1, prv = cur ^ 1;
        if (s + 1 < npass) load_images(slot + (s + 1) * nslots, s_img0 + prv * (CNN_G * CNN_PIX));   // the image tile conv1 read in the previous step is free
        const int grp1 = slot + s * nslots, grp2 = slot + (s - 1) * nslots;            // conv1's group (if s < npass), conv2's group (if s > 0)
        const int n2 = s > 0 ? (group_cnt(grp2) * 20 + 31) / 32 : 0, t2 = s > 0 ? group_cnt(grp2) * 20 : 0;
        const int n1 = s < npass ? (group_cnt(grp1) * 72 + 31) / 32 : 0, t1 = s < npass ? group_cnt(grp1) * 72 : 0;
        const float* img = s_img0 + cur * (CNN_G * CNN_PIX);
        float* p1w = s_p10 + cur * (CNN_G * CNN_P1_STRIDE);
        const float* p1r = s_p10 + prv * (CNN_G * CNN_P1_STRIDE);
        for (;;) {
            int id = 0;
            if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(id) : "r"(ctr_addr) : "memory");   // (atomicAdd on the generic address cost an S2R SR_CgaCtaId per chunk: 7 % of the stall samples, ncu r3g)
            id = __shfl_sync(0xffffffffu, id, 0);
            if (id >= n2 + n1) break;
            if (id < n2) { const int t = id * 32 + lane; if (t < t2) cnn_conv2_task(t, p1r, s_out, s_w2, s_b2); }
            else { const int t = (id - n2) * 32 + lane; if (t < t1) cnn_conv1_task(t, img, p1w, s_w1, s_b1); }
        }
        __syncthreads();                                                     // staging complete, pooled activations of the next group complete
        if (tid == 0) s_ctr = 0;
        if (s > 0) {
            // coalesced copy-out: per image and plane 250 16-byte vectors (k' < 2000; the pad columns stay zero from the allocation), 8 vectors = one 128-byte row of a k-tile
            const int img0 = grp2 * CNN_G, nv = group_cnt(grp2) * 500;
            for (int i = tid; i < nv; i += CNN_THREADS) {
                const int im = i / 500, r = i - im * 500, plane = r / 250, vq = r - plane * 250;
                const int row = img0 + im, kt = vq >> 3, kin = (vq & 7) * 8;
                const uint4 val = *reinterpret_cast<const uint4*>(s_out + im * (2 * CNN_FLAT_PAD) + plane * CNN_FLAT_PAD + vq * 8);
                __nv_bfloat16* d = g.out + (((long long)b * g.mb128 + (row >> 7)) * 64ll + plane * 32 + kt) * 8192ll + (long long)(row & 127) * 64 + kin;
                *reinterpret_cast<uint4*>(d) = val;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
}

// fc1.weight [500, 2000] (torch column k = o*100 + y*10 + x) -> bf16 [nb][512][2 * 2048] = [h | l] with the columns permuted to the conv kernel's
// k' = (y*10 + x)*20 + o, zero padded
__global__ void cnn_split_w_kernel(const float* __restrict__ theta, long long theta_stride, __nv_bfloat16* __restrict__ out, int nb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)CNN_HID_PAD * CNN_FLAT_PAD;
    if (i >= per * nb) return;
    const int b = (int)(i / per); const long long rem = i - b * per;
    const int r = (int)(rem / CNN_FLAT_PAD), c = (int)(rem - (long long)r * CNN_FLAT_PAD);
    float v = 0.f;
    if (r < CNN_HID && c < CNN_FLAT) { const int o = c % 20, pos = c / 20; v = theta[b * theta_stride + C_OFF_FW1 + (long long)r * CNN_FLAT + o * 100 + pos]; }
    const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    __nv_bfloat16* o2 = out + ((long long)b * CNN_HID_PAD + r) * (2ll * CNN_FLAT_PAD);
    o2[c] = h; o2[CNN_FLAT_PAD + c] = l;
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
    if (s->nb > c->sm_count) s->nb = c->sm_count;
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
        cnn_split_w_kernel<<<(unsigned)((t + 255) / 256), 256, 0, c->stream>>>(th, CNN_DIM, s->w, nb);
        cnn_gather_bias_kernel<<<(nb * CNN_HID_PAD + 255) / 256, 256, 0, c->stream>>>(th, CNN_DIM, s->bias, nb);
        CnnConvArgs ca{s->x32, M, th, CNN_DIM, nb, s->a2, mb128};
        const int ngroups = (M + CNN_G - 1) / CNN_G;
        long long ctas = c->sm_count;                                      // one CTA per SM; CTA i serves node i % nb
        if (ctas > (long long)ngroups * nb) ctas = (long long)ngroups * nb;
        if (ctas < nb) ctas = nb;
        cnn_conv_kernel<<<dim3((unsigned)ctas), CNN_THREADS, CNN_SMEM, c->stream>>>(ca);
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
