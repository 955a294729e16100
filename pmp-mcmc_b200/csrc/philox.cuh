// philox.cuh — counter-based random streams for the proposal generator and the acceptance draws.
//
// Replaces the reference's host-side serial generators (std::mt19937 + normal_distribution in
// 500_MP.cu:110-125,177-185; torch.normal in lb.py:131-136; np.random.normal in error.py:53,91,149).
// Those are unseeded, so "the same stream" can only mean: this stream is a pure function of
// (seed, iteration, stream id, element index) and can be regenerated anywhere.  To make that true across
// GPU and CPU *bit for bit*, the uniform→normal map uses only IEEE-exact operations (fma/mul/add/div/sqrt in
// binary64, explicit intrinsics so nvcc cannot contract or reassociate): Wichura's AS241 PPND16 rational
// approximations with a hand-rolled log for the tails.  tests/ checks it against an independent C restatement.
#pragma once
#include <stdint.h>
#include <string.h>

namespace pmp {

enum : uint32_t { STREAM_PROPOSAL = 0, STREAM_DRAW = 1, STREAM_PICK = 2, STREAM_CHAIN_INIT = 3 };

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), hi1 = __umulhi(M1, c[2]);
#else
    uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c[0]) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c[2]) >> 32);
#endif
    uint32_t lo0 = M0 * c[0], lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10 (Salmon et al., SC'11).
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// One 64-bit word of the stream: counter = (idx>>1 lo, idx>>1 hi, iter lo, iter hi[23:0] | stream<<24);
// word = (idx & 1) ? (r3:r2) : (r1:r0).
__host__ __device__ __forceinline__ uint64_t stream_u64(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx) {
    uint64_t blk = idx >> 1;
    uint32_t c[4] = { (uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter,
                      ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (stream << 24) };
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (idx & 1) ? ((uint64_t)c[3] << 32 | c[2]) : ((uint64_t)c[1] << 32 | c[0]);
}

// [0,1) with 53 bits, the layout numpy's random_sample uses for a 64-bit word.
__host__ __device__ __forceinline__ double u64_to_unit(uint64_t w) { return (double)(w >> 11) * (1.0 / 9007199254740992.0); }
// (0,1) open on both sides, symmetric about 1/2: (k + 1/2) * 2^-52, k < 2^52 — exact in binary64.
__host__ __device__ __forceinline__ double u64_to_open(uint64_t w) { return ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0); }

#ifdef __CUDA_ARCH__
#define PMP_FMA(a, b, c) __fma_rn((a), (b), (c))
#define PMP_MUL(a, b) __dmul_rn((a), (b))
#define PMP_ADD(a, b) __dadd_rn((a), (b))
#define PMP_DIV(a, b) __ddiv_rn((a), (b))
#define PMP_SQRT(a) __dsqrt_rn((a))
#else
#include <math.h>
#define PMP_FMA(a, b, c) fma((a), (b), (c))
#define PMP_MUL(a, b) ((a) * (b))
#define PMP_ADD(a, b) ((a) + (b))
#define PMP_DIV(a, b) ((a) / (b))
#define PMP_SQRT(a) sqrt((a))
#endif

// log(p) for a normal (non-subnormal) positive double from exact operations only:
// p = m * 2^e, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m-1)/(m+1); 12-term odd series.
__host__ __device__ __forceinline__ double det_log(double p) {
#ifdef __CUDA_ARCH__
    uint64_t bits = (uint64_t)__double_as_longlong(p);
#else
    uint64_t bits; memcpy(&bits, &p, 8);
#endif
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    uint64_t mb = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
#ifdef __CUDA_ARCH__
    double m = __longlong_as_double((long long)mb);
#else
    double m; memcpy(&m, &mb, 8);
#endif
    if (m > 1.4142135623730951) { m = PMP_MUL(m, 0.5); e += 1; }
    double s = PMP_DIV(PMP_ADD(m, -1.0), PMP_ADD(m, 1.0));
    double s2 = PMP_MUL(s, s);
    double t = 1.0 / 23.0;
    t = PMP_FMA(t, s2, 1.0 / 21.0);
    t = PMP_FMA(t, s2, 1.0 / 19.0);
    t = PMP_FMA(t, s2, 1.0 / 17.0);
    t = PMP_FMA(t, s2, 1.0 / 15.0);
    t = PMP_FMA(t, s2, 1.0 / 13.0);
    t = PMP_FMA(t, s2, 1.0 / 11.0);
    t = PMP_FMA(t, s2, 1.0 / 9.0);
    t = PMP_FMA(t, s2, 1.0 / 7.0);
    t = PMP_FMA(t, s2, 1.0 / 5.0);
    t = PMP_FMA(t, s2, 1.0 / 3.0);
    t = PMP_FMA(t, s2, 1.0);
    double logm = PMP_MUL(PMP_ADD(s, s), t);
    return PMP_FMA((double)e, 0.6931471805599453, logm);
}

#define PMP_H8(r, c7, c6, c5, c4, c3, c2, c1, c0)                                                              \
    PMP_FMA(PMP_FMA(PMP_FMA(PMP_FMA(PMP_FMA(PMP_FMA(PMP_FMA((c7), (r), (c6)), (r), (c5)), (r), (c4)), (r), (c3)), \
                            (r), (c2)), (r), (c1)), (r), (c0))

// Standard normal quantile, AS241 PPND16 (Wichura 1988), |rel err| ~ 1e-16.
__host__ __device__ __forceinline__ double det_norm_ppf(double u) {
    double q = PMP_ADD(u, -0.5);
    if (fabs(q) <= 0.425) {
        double r = PMP_FMA(-q, q, 0.180625);
        double num = PMP_H8(r, 2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4,
                            4.5921953931549871457e+4, 1.3731693765509461125e+4, 1.9715909503065514427e+3,
                            1.3314166789178437745e+2, 3.3871328727963666080e0);
        double den = PMP_H8(r, 5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4,
                            2.1213794301586595867e+4, 5.3941960214247511077e+3, 6.8718700749205790830e+2,
                            4.2313330701600911252e+1, 1.0);
        return PMP_DIV(PMP_MUL(q, num), den);
    }
    double p = q < 0.0 ? u : PMP_ADD(1.0, -u);
    double r = PMP_SQRT(-det_log(p));
    double z;
    if (r <= 5.0) {
        r = PMP_ADD(r, -1.6);
        double num = PMP_H8(r, 7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1,
                            1.27045825245236838258e0, 3.64784832476320460504e0, 5.76949722146069140550e0,
                            4.63033784615654529590e0, 1.42343711074968357734e0);
        double den = PMP_H8(r, 1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2,
                            1.48103976427480074590e-1, 6.89767334985100004550e-1, 1.67638483018380384940e0,
                            2.05319162663775882187e0, 1.0);
        z = PMP_DIV(num, den);
    } else {
        r = PMP_ADD(r, -5.0);
        double num = PMP_H8(r, 2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3,
                            2.65321895265761230930e-2, 2.96560571828504891230e-1, 1.78482653991729133580e0,
                            5.46378491116411436990e0, 6.65790464350110377720e0);
        double den = PMP_H8(r, 2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5,
                            7.86869131145613259100e-4, 1.48753612908506148525e-2, 1.36929880922735805310e-1,
                            5.99832206555887937690e-1, 1.0);
        z = PMP_DIV(num, den);
    }
    return q < 0.0 ? -z : z;
}

// Standard normal number `idx` of (seed, iter, stream), binary64.
__host__ __device__ __forceinline__ double stream_normal(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx) {
    return det_norm_ppf(u64_to_open(stream_u64(seed, iter, stream, idx)));
}

// Proposal increment number `idx`: a standard normal, or 2u-1 (uniform on [-1,1)) for PMP_FLAG_UNIFORM_PROPOSAL.
__host__ __device__ __forceinline__ double stream_step(uint64_t seed, uint64_t iter, uint64_t idx, int uniform) {
    if (uniform) return PMP_FMA(2.0, u64_to_unit(stream_u64(seed, iter, STREAM_PROPOSAL, idx)), -1.0);
    return stream_normal(seed, iter, STREAM_PROPOSAL, idx);
}

}  // namespace pmp
