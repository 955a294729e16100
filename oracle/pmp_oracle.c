/* pmp_oracle.c — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the
 * product (pmp-mcmc_b200/) never does.  Every function cites the reference lines it restates (paths relative to the
 * reference root).  Pinning: the reference ships no tests or golden vectors; this restatement is
 * pinned against (a) outputs of the reference's own Python code imported in the build container
 * (oracle/make_golden.py → tests/golden/*.npz), (b) the reference CUDA kernel compiled from its own source
 * (oracle/Makefile → oracle/_ref/) on the GPU box, (c) the fp64 closed form from sufficient statistics, and
 * (d) Random123's published Philox4x32-10 known-answer vectors.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no fused multiply-add unless written as fma()).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------------------
 * 1. Counter-based stream (independent restatement of the product's definition in csrc/philox.cuh):
 *    Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11), counter (idx>>1 lo, idx>>1 hi, iter lo, iter hi24 | stream<<24),
 *    key = seed; word = idx odd ? (r3:r2) : (r1:r0).  Normals by AS241 PPND16 from exactly rounded operations. */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof(c));
}

uint64_t oracle_stream_u64(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx) {
    uint64_t blk = idx >> 1;
    uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)iter, ((uint32_t)(iter >> 32) & 0x00FFFFFFu) | (stream << 24)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (idx & 1) ? ((uint64_t)c[3] << 32 | c[2]) : ((uint64_t)c[1] << 32 | c[0]);
}

double oracle_u64_to_unit(uint64_t w) { return (double)(w >> 11) * (1.0 / 9007199254740992.0); }
static double u64_to_open(uint64_t w) { return ((double)(w >> 12) + 0.5) * (1.0 / 4503599627370496.0); }

/* log of a positive normal double: p = m 2^e, m in [sqrt(.5), sqrt 2), log m = 2 atanh((m-1)/(m+1)) by its odd series */
double oracle_det_log(double p) {
    uint64_t bits; memcpy(&bits, &p, 8);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    uint64_t mb = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m; memcpy(&m, &mb, 8);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m + -1.0) / (m + 1.0);
    double s2 = s * s;
    static const double c[12] = {1.0, 1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0, 1.0 / 13.0, 1.0 / 15.0,
                                 1.0 / 17.0, 1.0 / 19.0, 1.0 / 21.0, 1.0 / 23.0};
    double t = c[11];
    for (int k = 10; k >= 0; --k) t = fma(t, s2, c[k]);
    double logm = (s + s) * t;
    return fma((double)e, 0.6931471805599453, logm);
}

static double horner8(double r, const double* c) { /* c[0] is the highest-order coefficient */
    double t = c[0];
    for (int k = 1; k < 8; ++k) t = fma(t, r, c[k]);
    return t;
}

/* Wichura, "Algorithm AS 241: The percentage points of the normal distribution", Appl. Statist. 37 (1988), PPND16 */
double oracle_norm_ppf(double u) {
    static const double A[8] = {2.5090809287301226727e+3, 3.3430575583588128105e+4, 6.7265770927008700853e+4, 4.5921953931549871457e+4,
                                1.3731693765509461125e+4, 1.9715909503065514427e+3, 1.3314166789178437745e+2, 3.3871328727963666080e0};
    static const double B[8] = {5.2264952788528545610e+3, 2.8729085735721942674e+4, 3.9307895800092710610e+4, 2.1213794301586595867e+4,
                                5.3941960214247511077e+3, 6.8718700749205790830e+2, 4.2313330701600911252e+1, 1.0};
    static const double C[8] = {7.74545014278341407640e-4, 2.27238449892691845833e-2, 2.41780725177450611770e-1, 1.27045825245236838258e0,
                                3.64784832476320460504e0, 5.76949722146069140550e0, 4.63033784615654529590e0, 1.42343711074968357734e0};
    static const double D[8] = {1.05075007164441684324e-9, 5.47593808499534494600e-4, 1.51986665636164571966e-2, 1.48103976427480074590e-1,
                                6.89767334985100004550e-1, 1.67638483018380384940e0, 2.05319162663775882187e0, 1.0};
    static const double E[8] = {2.01033439929228813265e-7, 2.71155556874348757815e-5, 1.24266094738807843860e-3, 2.65321895265761230930e-2,
                                2.96560571828504891230e-1, 1.78482653991729133580e0, 5.46378491116411436990e0, 6.65790464350110377720e0};
    static const double F[8] = {2.04426310338993978564e-15, 1.42151175831644588870e-7, 1.84631831751005468180e-5, 7.86869131145613259100e-4,
                                1.48753612908506148525e-2, 1.36929880922735805310e-1, 5.99832206555887937690e-1, 1.0};
    double q = u + -0.5;
    if (fabs(q) <= 0.425) {
        double r = fma(-q, q, 0.180625);
        return (q * horner8(r, A)) / horner8(r, B);
    }
    double p = q < 0.0 ? u : 1.0 + -u;
    double r = sqrt(-oracle_det_log(p));
    double z;
    if (r <= 5.0) { r = r + -1.6; z = horner8(r, C) / horner8(r, D); }
    else { r = r + -5.0; z = horner8(r, E) / horner8(r, F); }
    return q < 0.0 ? -z : z;
}

double oracle_stream_normal(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx) {
    return oracle_norm_ppf(u64_to_open(oracle_stream_u64(seed, iter, stream, idx)));
}

void oracle_stream_normals(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx0, int64_t count, double* out) {
    for (int64_t i = 0; i < count; ++i) out[i] = oracle_stream_normal(seed, iter, stream, idx0 + (uint64_t)i);
}
void oracle_stream_uniforms(uint64_t seed, uint64_t iter, uint32_t stream, uint64_t idx0, int64_t count, double* out) {
    for (int64_t i = 0; i < count; ++i) out[i] = oracle_u64_to_unit(oracle_stream_u64(seed, iter, stream, idx0 + (uint64_t)i));
}

/* ------------------------------------------------------------------------------------------------------------
 * 2. Proposal generators, in the reference's own loop order and float32 arithmetic
 *    (child = parent + normal(0, alpha), i.e. fl(parent + fl(alpha*z))).  z of the step that creates node c is stream
 *    normal number c*dim + j — the one convention this repo adds, since the reference is unseeded.
 *    tree 0: flat        500_MP.cu:181-185, lb.py:173-176
 *    tree 1: doubling    500_PMP.cu:170-179, lb.py:268-272, error.py:88-91, com_dim.py:34-37, PMP_FC.py:176-182
 *    tree 2: (N+1)-ary   conv_pmp.cu:182-197, lb.py:356-360, error.py:145-149 */
static float step32(float parent, float alpha, double z) { volatile float inc = alpha * (float)z; return parent + inc; }

static uint64_t g_chain = 0;       /* chain id: upper 32 bits of the element index (csrc/chains.cu) */
void oracle_set_chain(uint64_t chain) { g_chain = chain; }
static int g_uniform_steps = 0;   /* 1: increments alpha*(2u-1) — random.uniform(-alpha, alpha), error.py:27 */
void oracle_set_uniform_steps(int on) { g_uniform_steps = on; }
static double step_value(uint64_t seed, uint64_t iter, uint64_t idx) {
    idx |= g_chain << 32;
    if (g_uniform_steps) return fma(2.0, oracle_u64_to_unit(oracle_stream_u64(seed, iter, 0, idx)), -1.0);
    return oracle_stream_normal(seed, iter, 0, idx);
}

void oracle_propose(int tree, int b, int depth, int dim, float alpha, const float* state, uint64_t seed, uint64_t iter, float* props) {
    for (int j = 0; j < dim; ++j) props[j] = state[j];
    if (tree == 0) {
        for (int i = 1; i < b; ++i)
            for (int j = 0; j < dim; ++j)
                props[(size_t)i * dim + j] = step32(props[j], alpha, step_value(seed, iter, (uint64_t)i * dim + j));
    } else if (tree == 1) {
        for (int l = 0; l < depth; ++l) {
            long jj = 1L << l;
            for (long k = 0; k < jj; ++k)
                for (int j = 0; j < dim; ++j)
                    props[(size_t)(k + jj) * dim + j] = step32(props[(size_t)k * dim + j], alpha, step_value(seed, iter, (uint64_t)(k + jj) * dim + j));
        }
    } else {
        long temp = 1;
        for (int l = 0; l < depth; ++l) {
            for (int jn = 0; jn < b - 1; ++jn)
                for (long k = 0; k < temp; ++k) {
                    long to = k + temp * (jn + 1);
                    for (int j = 0; j < dim; ++j)
                        props[(size_t)to * dim + j] = step32(props[(size_t)k * dim + j], alpha, step_value(seed, iter, (uint64_t)to * dim + j));
                }
            temp *= b;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * 3. Linear-Gaussian log-likelihood sweep. */

/* (a) The reference CUDA kernel's data loop, operation for operation (500_MP.cu:12,16-20; identical in 500_PMP.cu,
 *     100000_*.cu, conv_*.cu, ess_per_s_*.cu up to the SCALE literal): float32 fma for y_hat, float32 subtract and
 *     divide, widened to double, logf of a float argument, divide by SCALE in double, running sum rounded to float32
 *     after every point (global-memory += on a float). */
void oracle_loglik_linear_refcuda(const float* x, const float* y, int64_t n, const float* nets, int P, double scale, float* gpu_a) {
    const float MPI = 3.14159265359f;
    for (int idx = 0; idx < P; ++idx) {
        float acc = gpu_a[idx];
        const float b0 = nets[idx * 3], b1 = nets[idx * 3 + 1], sg = nets[idx * 3 + 2];
        for (int64_t i = 0; i < n; ++i) {
            float y_hat = fmaf(b1, x[i], b0);
            volatile float tf = (y[i] - y_hat) / sg;
            double temp = tf;
            volatile float arg = 2 * MPI * sg * sg;
            double term = (-0.5 * (double)logf(arg) - 0.5 * temp * temp) / scale;
            acc = (float)((double)acc + term);
        }
        gpu_a[idx] = acc;
    }
}

/* (b) Same per-point arithmetic (float32 y_hat and residual, as the kernel and torch compute them), summed exactly
 *     enough to serve as ground truth: long double accumulation of double terms.  Returns loglik/scale. */
void oracle_loglik_linear_f64(const float* x, const float* y, int64_t n, const float* nets, int P, double scale, double* out) {
    for (int idx = 0; idx < P; ++idx) {
        const float b0 = nets[idx * 3], b1 = nets[idx * 3 + 1];
        const double sg = nets[idx * 3 + 2];
        long double s = 0.0L;
        for (int64_t i = 0; i < n; ++i) {
            float y_hat = fmaf(b1, x[i], b0);
            volatile float d = y[i] - y_hat;
            s += (long double)((double)d * (double)d);
        }
        out[idx] = (-0.5 * (double)n * log(6.283185307179586477 * sg * sg) - 0.5 * (double)s / (sg * sg)) / scale;
    }
}

/* (c) Closed form from sufficient statistics in binary64 (SURVEY.md §4 "free known-answer test"): no per-point float32
 *     rounding at all, so it differs from (b) by the float32 residual rounding only. */
void oracle_loglik_linear_suffstat(const float* x, const float* y, int64_t n, const float* nets, int P, double scale, double* out) {
    long double sx = 0, sy = 0, sxx = 0, sxy = 0, syy = 0;
    for (int64_t i = 0; i < n; ++i) {
        long double xi = x[i], yi = y[i];
        sx += xi; sy += yi; sxx += xi * xi; sxy += xi * yi; syy += yi * yi;
    }
    for (int idx = 0; idx < P; ++idx) {
        long double b0 = nets[idx * 3], b1 = nets[idx * 3 + 1];
        double sg = nets[idx * 3 + 2];
        long double ss = syy - 2 * b0 * sy - 2 * b1 * sxy + (long double)n * b0 * b0 + 2 * b0 * b1 * sx + b1 * b1 * sxx;
        out[idx] = (-0.5 * (double)n * log(6.283185307179586477 * sg * sg) - 0.5 * (double)ss / (sg * sg)) / scale;
    }
}

/* (d) Bit-level mirror of the product's chunked evaluation order (csrc/sweep_linear.cuh): per 64-point chunk (aligned
 *     to the GLOBAL point index), float32 even/odd accumulators over the full groups of four, (even+odd), ragged tail,
 *     then round(partial/sigma^2 * 2^20) summed as integers.  Used to check that the device sums are bit-exact and
 *     independent of grid size and GPU count. */
void oracle_sumsq_fixed_mirror(const float* x, const float* y, int64_t n, const float* nets, int P, int64_t total_chunks_for_limit, uint64_t* out) {
    const int CH = 64;
    double limit = 4611686018427387904.0 / (double)(total_chunks_for_limit > 0 ? total_chunks_for_limit : 1);
    for (int idx = 0; idx < P; ++idx) {
        const float b0 = nets[idx * 3], b1 = nets[idx * 3 + 1];
        const double sg = nets[idx * 3 + 2];
        const double scl = 1048576.0 / (sg * sg);
        uint64_t acc = 0;
        for (int64_t c0 = 0; c0 < n; c0 += CH) {
            int cnt = (int)((n - c0) < CH ? (n - c0) : CH);
            int full = (cnt >> 2) << 2;
            float ae = 0.f, ao = 0.f;
            for (int i = 0; i < full; i += 2) {
                volatile float d0 = y[c0 + i] - fmaf(b1, x[c0 + i], b0);
                volatile float d1 = y[c0 + i + 1] - fmaf(b1, x[c0 + i + 1], b0);
                ae = fmaf(d0, d0, ae); ao = fmaf(d1, d1, ao);
            }
            volatile float part = ae + ao;
            float pp = part;
            for (int i = full; i < cnt; ++i) { volatile float d = y[c0 + i] - fmaf(b1, x[c0 + i], b0); pp = fmaf(d, d, pp); }
            double dq = (double)pp * scl;
            if (!(dq < limit)) dq = limit;
            acc += (uint64_t)llrint(dq);
        }
        out[idx] = acc;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * 4. Mirror of the device's blocked inclusive scan (csrc/accept.cuh: block_inclusive_scan) for exact cdf equality. */
void oracle_blocked_cdf(const double* w, int P, double* cdf) {
    const int T = 1024;   /* ACCEPT_THREADS */
    int ipt = (P + T - 1) / T;
    double* run = (double*)calloc(T, sizeof(double));
    double* incl = (double*)calloc(T, sizeof(double));
    for (int t = 0; t < T; ++t) {
        double r = 0.0;
        for (int i = 0; i < ipt; ++i) { int k = t * ipt + i; if (k < P) { r += w[k]; cdf[k] = r; } }
        run[t] = r; incl[t] = r;
    }
    for (int wp = 0; wp < T / 32; ++wp)           /* Kogge-Stone inside each warp */
        for (int o = 1; o < 32; o <<= 1) {
            double tmp[32];
            for (int l = 0; l < 32; ++l) tmp[l] = incl[wp * 32 + l];
            for (int l = o; l < 32; ++l) incl[wp * 32 + l] = tmp[l - o] + tmp[l];
        }
    double wt[32];
    for (int wp = 0; wp < 32; ++wp) wt[wp] = wp < T / 32 ? incl[wp * 32 + 31] : 0.0;
    for (int o = 1; o < 32; o <<= 1) {
        double tmp[32]; memcpy(tmp, wt, sizeof(tmp));
        for (int l = o; l < 32; ++l) wt[l] = tmp[l - o] + tmp[l];
    }
    for (int t = 0; t < T; ++t) {
        int wp = t >> 5, lane = t & 31;
        double woff = wp > 0 ? wt[wp - 1] : 0.0;
        double lex = lane > 0 ? incl[t - 1] : 0.0;
        double excl = woff + lex;
        for (int i = 0; i < ipt; ++i) { int k = t * ipt + i; if (k < P) cdf[k] = excl + cdf[k]; }
    }
    free(run); free(incl);   /* unnormalised: the device compares cdf_k with u * cdf[P-1] */
}
