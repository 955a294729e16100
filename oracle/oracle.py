"""oracle.py — Python face of the CPU restatement.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
nothing under pmp-mcmc_b200/ does.  It wraps oracle/liboracle.so (pmp_oracle.c) and restates, in plain numpy loops that
follow the reference line by line, the acceptance rules and the categorical draws:

    mp_logweights        GMOptimizer.step         simple_net/lb.py:139-150   (500_MP.cu:22-31; error.py:58-64)
    psp_logweights       preMOptimizer.step       simple_net/lb.py:206-240   (error.py:96-121, com_dim.py:44-68, PMP_FC.py:117-136)
    pmp_logweights       GMpreOptimizerV2.step    simple_net/lb.py:304-330   (error.py:151-173)
    table_logweights     CUDA PMP kernels         500_PMP.cu:23-30, conv_pmp.cu:22-33 (+ host tables 500_PMP.cu:180-193, conv_pmp.cu:198-218)
    draw_numpy           pandas sample → RandomState.choice: cdf=cumsum(p); cdf/=cdf[-1]; searchsorted(cdf,u,'right')
    draw_libstdcxx       std::discrete_distribution: p=w/sum; partial_sum; lower_bound   (500_MP.cu:218-222)
    loglik_lb_torch      BayesNet.loglik          simple_net/lb.py:103-108 (torch float32, as the reference runs it)

Pinning status: see the header of pmp_oracle.c and tests/golden/README.md.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "pmp_oracle.c")):
            subprocess.run(["make", "-C", _HERE, "--no-print-directory", "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(path)
        u64, u32, i64, dbl = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, ctypes.c_double
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        L.oracle_stream_u64.restype = u64
        L.oracle_stream_u64.argtypes = [u64, u64, u32, u64]
        L.oracle_stream_normal.restype = dbl
        L.oracle_stream_normal.argtypes = [u64, u64, u32, u64]
        L.oracle_norm_ppf.restype = dbl
        L.oracle_norm_ppf.argtypes = [dbl]
        L.oracle_det_log.restype = dbl
        L.oracle_det_log.argtypes = [dbl]
        L.oracle_stream_normals.argtypes = [u64, u64, u32, u64, i64, dp]
        L.oracle_stream_uniforms.argtypes = [u64, u64, u32, u64, i64, dp]
        L.oracle_philox4x32_10.argtypes = [ctypes.POINTER(u32), ctypes.POINTER(u32), ctypes.POINTER(u32)]
        L.oracle_propose.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, fp, u64, u64, fp]
        L.oracle_loglik_linear_refcuda.argtypes = [fp, fp, i64, fp, ctypes.c_int, dbl, fp]
        L.oracle_loglik_linear_f64.argtypes = [fp, fp, i64, fp, ctypes.c_int, dbl, dp]
        L.oracle_loglik_linear_suffstat.argtypes = [fp, fp, i64, fp, ctypes.c_int, dbl, dp]
        L.oracle_sumsq_fixed_mirror.argtypes = [fp, fp, i64, fp, ctypes.c_int, i64, ctypes.POINTER(u64)]
        L.oracle_blocked_cdf.argtypes = [dp, ctypes.c_int, dp]
        L.oracle_set_chain.argtypes = [u64]
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


STREAM_PROPOSAL, STREAM_DRAW, STREAM_PICK, STREAM_CHAIN_INIT = 0, 1, 2, 3
TREE_FLAT, TREE_BINARY, TREE_BARY = 0, 1, 2


# ---------------------------------------------------------------------------------------------------------------
# streams
def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(v) for v in o]


def stream_normals(seed, it, stream, idx0, count):
    out = np.empty(count, dtype=np.float64)
    lib().oracle_stream_normals(seed, it, stream, idx0, count, _dptr(out))
    return out


def stream_uniforms(seed, it, stream, idx0, count):
    out = np.empty(count, dtype=np.float64)
    lib().oracle_stream_uniforms(seed, it, stream, idx0, count, _dptr(out))
    return out


def num_nodes(tree, b, depth):
    return b if tree == TREE_FLAT else (2 ** depth if tree == TREE_BINARY else b ** depth)


def propose(tree, b, depth, dim, alpha, state, seed, it, uniform=False, chain=0):
    P = num_nodes(tree, b, depth)
    st = _f32(state)
    out = np.empty((P, dim), dtype=np.float32)
    lib().oracle_set_uniform_steps(1 if uniform else 0)
    lib().oracle_set_chain(ctypes.c_uint64(chain))
    lib().oracle_propose(tree, b, depth, dim, ctypes.c_float(alpha), _fptr(st), seed, it, _fptr(out))
    return out


# ---------------------------------------------------------------------------------------------------------------
# linear-Gaussian sweep
def loglik_linear_refcuda(x, y, nets, scale):
    x, y, nets = _f32(x), _f32(y), _f32(nets)
    out = np.zeros(nets.shape[0], dtype=np.float32)
    lib().oracle_loglik_linear_refcuda(_fptr(x), _fptr(y), x.size, _fptr(nets), nets.shape[0], scale, _fptr(out))
    return out


def loglik_linear_f64(x, y, nets, scale):
    x, y, nets = _f32(x), _f32(y), _f32(nets)
    out = np.zeros(nets.shape[0], dtype=np.float64)
    lib().oracle_loglik_linear_f64(_fptr(x), _fptr(y), x.size, _fptr(nets), nets.shape[0], scale, _dptr(out))
    return out


def loglik_linear_suffstat(x, y, nets, scale):
    x, y, nets = _f32(x), _f32(y), _f32(nets)
    out = np.zeros(nets.shape[0], dtype=np.float64)
    lib().oracle_loglik_linear_suffstat(_fptr(x), _fptr(y), x.size, _fptr(nets), nets.shape[0], scale, _dptr(out))
    return out


def sumsq_fixed_mirror(x, y, nets, total_chunks_for_limit):
    x, y, nets = _f32(x), _f32(y), _f32(nets)
    out = np.zeros(nets.shape[0], dtype=np.uint64)
    lib().oracle_sumsq_fixed_mirror(_fptr(x), _fptr(y), x.size, _fptr(nets), nets.shape[0], total_chunks_for_limit,
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    return out


def loglik_linear_from_fixed(acc, nets, n_global, scale):
    """log-target from the integer sums, as csrc/accept.cuh finalises them."""
    sg = _f32(nets)[:, 2].astype(np.float64)
    S = acc.astype(np.int64).astype(np.float64) / float(1 << 20)
    return (-0.5 * n_global * np.log(6.283185307179586477 * sg * sg) - 0.5 * S) / scale


def loglik_lb_torch(x, y, nets, threads=None):
    """BayesNet.loglik (lb.py:103-108) for every row of nets, the way the reference's Python loop evaluates it:
    torch float32 on the CPU, one vectorised pass per proposal, Normal(yhat, |sigma|).log_prob(y).sum() / n * 50."""
    import torch
    import torch.distributions as dist
    if threads:
        torch.set_num_threads(threads)
    xt = torch.as_tensor(np.asarray(x, dtype=np.float32))
    yt = torch.as_tensor(np.asarray(y, dtype=np.float32))
    out = np.empty(len(nets), dtype=np.float64)
    with torch.no_grad():
        for i, (b0, b1, sg) in enumerate(np.asarray(nets, dtype=np.float32)):
            yhat = torch.tensor([b0]) + torch.tensor([b1]) * xt
            logprob = dist.Normal(yhat, torch.tensor([sg]).abs()).log_prob(yt)
            out[i] = (logprob.sum() / yt.shape[0] * 50).item()
    return out


# ---------------------------------------------------------------------------------------------------------------
# proposal kernel and acceptance rules (binary64, log domain; loops follow the reference)
def log_kernel(a, b, ks=1.0):
    """sum over parameters of log N(a_j; b_j, ks) — log_trans_prob lb.py:111-116 (ks = 1 there), 500_MP.cu:26-28."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sum(-0.5 * math.log(2 * math.pi) - math.log(ks) - 0.5 * ((a - b) / ks) ** 2))


def mp_logweights(lt, props, ks=1.0, use_kernel=True):
    P = len(lt)
    if P > 64:                              # same sums as the loop below, one row of pairs at a time (the loop is O(P^2) Python)
        th = np.asarray(props, dtype=np.float64)
        A = np.array(lt, dtype=np.float64)
        if use_kernel:
            c = th.shape[1] * (-0.5 * math.log(2 * math.pi) - math.log(ks))
            for j in range(P):
                d2 = np.sum(((th[j] - th) / ks) ** 2, axis=1)
                A[j] += (P - 1) * c - 0.5 * (np.sum(d2) - d2[j])
        return A
    A = np.empty(P)
    for j in range(P):                      # lb.py:144-150
        temp = 0.0
        for k in range(P):
            if j != k and use_kernel:
                temp += log_kernel(props[j], props[k], ks)
        A[j] = temp + lt[j]
    return A


def _psp_partner(node, c):
    """The `judg` loop of lb.py:218-238: reduce node modulo 2^(c+1), partner differs in bit c."""
    judg = node
    j = 2 ** (c + 1)
    half_j = j // 2
    deep = int(math.log2(judg)) if judg > 0 else 0
    while judg > j - 1:
        if judg >= 2 ** deep:
            judg -= 2 ** deep
        deep -= 1
    return (judg, judg + half_j) if judg < half_j else (judg, judg - half_j)


def psp_logweights(lt, props, depth, ks=1.0, use_kernel=True):
    P = 2 ** depth
    A = np.zeros(P)
    for a in range(P):
        for c in range(depth):
            m, q = _psp_partner(a, c)
            lw_new = lt[m] + (log_kernel(props[m], props[q], ks) if use_kernel else 0.0)
            lw_old = lt[q] + (log_kernel(props[q], props[m], ks) if use_kernel else 0.0)
            # log(w_new/(w_new+w_old)), lb.py:240, evaluated stably
            with np.errstate(invalid="ignore"):
                A[a] += -np.logaddexp(0.0, lw_old - lw_new)
    return A


def pmp_logweights(lt, props, b, depth, ks=1.0, use_kernel=True, quirk_level_mod=False):
    P = b ** depth
    A = np.zeros(P)                          # log of lb.py:308's ones
    for i in range(depth):                   # lb.py:315-330
        temp = b ** i
        for h in range(temp):
            v = np.empty(b)
            for j in range(b):
                v[j] = lt[h + j * temp]
                for k in range(b):
                    if j != k and use_kernel:
                        v[j] += log_kernel(props[h + j * temp], props[h + k * temp], ks)
            mx = np.max(v)
            lse = mx + math.log(np.sum(np.exp(v - mx))) if np.isfinite(mx) else -np.inf
            for j in range(b):
                A[h + j * temp] += (v[j] - lse) if np.isfinite(mx) else -np.inf
        if i < depth - 1:
            mod = b * (i + 1) if quirk_level_mod else b ** (i + 1)
            for l in range(b ** (i + 2) - b ** (i + 1)):
                A[l + b ** (i + 1)] = A[(l + b ** (i + 1)) % mod]
    return A


def table_logweights(lt, props, b, depth, ks=1.0, quirk_const=False):
    """lt + sum_d sum_{k != m in group} log K(m,k), m = p mod b^(d+1)  (500_PMP.cu:23-30, conv_pmp.cu:22-33)."""
    P = b ** depth
    dim = np.asarray(props).shape[1]
    A = np.empty(P)
    for p in range(P):
        v = lt[p]
        for d in range(depth):
            s = b ** d
            m = p % (s * b)
            h = m % s
            for k in range(b):
                o = h + k * s
                if o != m:
                    v += dim * (-0.5 * math.log(2 * math.pi)) if quirk_const else log_kernel(props[m], props[o], ks)
        A[p] = v
    return A


def standardize(A):
    """(A - mean)/std with the unbiased std of torch.std (PMP_FC.py:138-140)."""
    A = np.asarray(A, dtype=np.float64)
    return (A - A.mean()) / A.std(ddof=1)


def weights_from_log(A):
    """exp(A - max A) with the maximum taken over the non-NaN entries; NaN log-weights (0/0 in the reference's linear
    domain, e.g. two dead partner nodes in the Barker tree — pandas would raise there) get weight 0, as on the device."""
    A = np.asarray(A, dtype=np.float64)
    m = np.nanmax(A) if np.any(~np.isnan(A)) else 0.0
    with np.errstate(invalid="ignore"):
        w = np.exp(A - m)
    return np.where(np.isnan(w), 0.0, w)


def draw_numpy(w, u):
    """pandas DataFrame.sample(weights=) → np.random.RandomState.choice(p=): verified against both in tests."""
    p = np.asarray(w, dtype=np.float64)
    p = p / p.sum()
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    return np.minimum(cdf.searchsorted(np.asarray(u), side="right"), len(p) - 1).astype(np.int32)


def draw_libstdcxx(w, u):
    """std::discrete_distribution (libstdc++ bits/random.tcc): normalise, partial_sum, last = 1, lower_bound."""
    p = np.asarray(w, dtype=np.float64)
    p = p / np.sum(p)
    cp = np.cumsum(p)
    cp[-1] = 1.0
    return np.minimum(cp.searchsorted(np.asarray(u), side="left"), len(p) - 1).astype(np.int32)


def blocked_cdf(w):
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty_like(w)
    lib().oracle_blocked_cdf(_dptr(w), len(w), _dptr(out))
    return out


def draw_blocked(w, u, side="right"):
    """The device's draw, association for association: unnormalised blocked cdf compared with u * total."""
    cdf = blocked_cdf(w)
    thr = np.asarray(u, dtype=np.float64) * cdf[-1]
    return np.minimum(cdf.searchsorted(thr, side=side), len(w) - 1).astype(np.int32)


def pick_index(u_pick, P):
    """np.random.choice(arange(P), 1) (lb.py:162) driven by one injected uniform."""
    return min(P - 1, int(u_pick * P))


# analytic log-targets ------------------------------------------------------------------------------------------
def log_normal1d(x, mu, sigma):          # error.py:11-14 (log of)
    return -0.5 * ((x - mu) / sigma) ** 2 - math.log(sigma) - 0.5 * math.log(2 * math.pi)


def log_banana(x):                       # banana_data.ipynb cell 2 L1-5 (log of)
    x1, x2 = float(x[0]), float(x[1])
    return -(x1 ** 2) / 2 - ((x2 - 2 * (x1 ** 2 - 5)) ** 2) / 2


def log_stdnormal(x):                    # com_dim.py:13-15 with mu=0, cov=I (log of)
    x = np.asarray(x, dtype=np.float64)
    return float(-0.5 * np.sum(x * x) - 0.5 * len(x) * math.log(2 * math.pi))


# ---------------------------------------------------------------------------------------------------------------
# analytic chains
def analytic_logtarget(target, x, p0=0.0, p1=1.0):
    """target: 1 normal1d (error.py:11-14), 2 banana (banana_data.ipynb cell 2), 3 N(0, I_d) (com_dim.py:13-15)."""
    x = np.asarray(x, dtype=np.float64)
    if target == 1:
        return log_normal1d(float(x[0]), p0, p1)
    if target == 2:
        return log_banana(x)
    return log_stdnormal(x)


def logweights(algo, lt, props, b, depth, ks=1.0, use_kernel=True, quirk_level_mod=False, quirk_const=False):
    """algo ids of include/pmp_b200.h: 2 MP, 3 PSP, 4 PMP, 5 TABLE."""
    if algo == 2:
        return mp_logweights(lt, props, ks, use_kernel)
    if algo == 3:
        return psp_logweights(lt, props, depth, ks, use_kernel)
    if algo == 4:
        return pmp_logweights(lt, props, b, depth, ks, use_kernel, quirk_level_mod)
    if not use_kernel and not quirk_const:
        return np.array(lt, dtype=np.float64)
    return table_logweights(lt, props, b, depth, ks, quirk_const)


def draw_sequential(w, u, side="right"):
    """csrc/chains.cu: plain sequential cumsum (NumPy's association), unnormalised compare with u*total."""
    cdf = np.cumsum(np.asarray(w, dtype=np.float64))
    return np.minimum(cdf.searchsorted(np.asarray(u, dtype=np.float64) * cdf[-1], side=side), len(w) - 1).astype(np.int32)


def analytic_chain(tree, b, depth, dim, target, algo, draw, alpha, seed, iters, state0, chain=0, scale=1.0, p0=0.0, p1=1.0,
                   ks=1.0, use_kernel=True, quirk_level_mod=False, uniform=False, it0=0):
    """One chain of csrc/chains.cu (and of pmp_run on an analytic target for chain=0), step by step.
    algo 0 MH / 1 Barker need b=2 flat.  draw: 0 python rule, 1 cuda rule, 2 single."""
    P = num_nodes(tree, b, depth)
    bb = 2 if tree == TREE_BINARY else b
    state = _f32(state0).copy()
    base = chain << 32
    samples = np.empty((iters, P, dim), dtype=np.float32)
    states = np.empty((iters, dim), dtype=np.float32)
    nexts = np.empty(iters, dtype=np.int32)
    for k in range(iters):
        it = it0 + k
        props = propose(tree, b, depth, dim, alpha, state, seed, it, uniform, chain)
        lt = np.array([analytic_logtarget(target, props[p], p0, p1) / scale for p in range(P)])
        if algo in (0, 1):
            u = stream_uniforms(seed, it, STREAM_DRAW, base, 1)[0]
            if algo == 0:
                nxt = int(u < math.exp(min(700.0, lt[1] - lt[0])))
            else:
                m = max(lt[0], lt[1]); w0, w1 = math.exp(lt[0] - m), math.exp(lt[1] - m)
                nxt = int(w1 / (w0 + w1) > u)
            d = np.full(P, nxt, dtype=np.int32)
            d[0] = nxt
        else:
            A = logweights(algo, lt, props.astype(np.float64), bb, depth, ks, use_kernel, quirk_level_mod)
            w = weights_from_log(A)
            nd = 1 if draw == 2 else P
            u = stream_uniforms(seed, it, STREAM_DRAW, base, nd)
            dd = draw_sequential(w, u, "left" if draw == 1 else "right")
            nxt = int(dd[pick_index(stream_uniforms(seed, it, STREAM_PICK, base, 1)[0], P)]) if draw == 0 else int(dd[0])
            d = np.full(P, nxt, dtype=np.int32)
            d[:nd] = dd
        samples[k] = props[d]
        state = props[nxt].copy()
        states[k] = state
        nexts[k] = nxt
    return {"samples": samples, "states": states, "next": nexts}


def error_py_replay(kind, hops, mu, sigma, N, deep, x0, normals, us, picks):
    """simple_sampling/error/error.py MP (43-77), PSP (78-134), PMP (137-190) restated with the oracle's weight rules,
    driven by recorded streams: `normals` in the order the script draws them, `us[h]` the P uniforms pandas' sample
    consumed at hop h, `picks[h]` the index np.random.choice returned.  Returns the full X (before the burn-in cut)."""
    b = N + 1
    if kind == "MP":
        P, tree, algo, depth = b, TREE_FLAT, 2, 1
    elif kind == "PSP":
        depth = int(math.log2(N + 1)); P, tree, algo = N + 1, TREE_BINARY, 3
    else:
        depth = deep; P, tree, algo = b ** deep, TREE_BARY, 4
    nz = iter(normals)
    X = np.empty(hops * P)
    Y = np.empty(P)

    def regenerate(root):
        Y[0] = root
        if kind == "MP":                                   # error.py:51-53 / 73-75
            for i in range(N):
                Y[i + 1] = root + next(nz)
        elif kind == "PSP":                                # error.py:88-91 / 129-132
            for i in range(depth):
                j = 2 ** i
                for k in range(j):
                    Y[k + j] = Y[k] + next(nz)
        else:                                              # error.py:145-149 / 184-188
            for dee in range(deep):
                temp = b ** dee
                for j in range(N):
                    for k in range(temp):
                        Y[k + temp * (j + 1)] = Y[k] + next(nz)

    regenerate(x0)
    for h in range(hops):
        lt = np.array([log_normal1d(Y[p], mu, sigma) for p in range(P)])
        A = logweights(algo, lt, Y.reshape(-1, 1), 2 if kind == "PSP" else b, depth, 1.0, True, quirk_level_mod=True)
        d = draw_numpy(weights_from_log(A), us[h])
        X[h * P:(h + 1) * P] = Y[d]
        regenerate(X[h * P + picks[h]])
    return X


# ---------------------------------------------------------------------------------------------------------------
# FC model (complex_nets/Mnist/FC/PMP_FC.py:21-44): 784-512-256-128-10 ReLU MLP, loss = CrossEntropy(mean) / 10
FC_SHAPES = [(512, 784), (512,), (256, 512), (256,), (128, 256), (128,), (10, 128), (10,)]   # torch parameter order
FC_DIM = sum(int(np.prod(s)) for s in FC_SHAPES)


def fc_unpack(theta):
    out, off = [], 0
    for s in FC_SHAPES:
        k = int(np.prod(s))
        out.append(np.asarray(theta[off:off + k]).reshape(s))
        off += k
    return out


def fc_init_theta(seed=0):
    """nn.Linear's default init (uniform +-1/sqrt(fan_in)) from a seeded generator: a deterministic stand-in for FC_model.pkl."""
    rng = np.random.default_rng(seed)
    parts = []
    for s in FC_SHAPES:
        fan_in = s[1] if len(s) == 2 else {512: 784, 256: 512, 128: 256, 10: 128}[s[0]]
        parts.append(rng.uniform(-1, 1, size=int(np.prod(s))) / math.sqrt(fan_in))
    return np.concatenate(parts).astype(np.float32)


def fc_mean_ce_f64(X, y, theta):
    """Ground truth: the forward pass of PMP_FC.py:30-36 and CrossEntropyLoss(mean) in binary64 from the float32 inputs."""
    W1, b1, W2, b2, W3, b3, W4, b4 = [a.astype(np.float64) for a in fc_unpack(theta)]
    h = np.maximum(np.asarray(X, dtype=np.float64).reshape(len(y), -1) @ W1.T + b1, 0)
    h = np.maximum(h @ W2.T + b2, 0)
    h = np.maximum(h @ W3.T + b3, 0)
    z = h @ W4.T + b4
    m = z.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(z - m).sum(axis=1))
    return float(np.mean(lse - z[np.arange(len(y)), np.asarray(y)]))


def fc_loss_torch32(X, y, theta, div=10.0):
    """loss(net) exactly as the reference evaluates it (PMP_FC.py:40-44): torch float32, CrossEntropyLoss()(net(X), y) / 10."""
    import torch
    import torch.nn.functional as F
    W1, b1, W2, b2, W3, b3, W4, b4 = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in fc_unpack(theta)]
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).view(-1, 28 * 28)
        x = F.relu(F.linear(x, W1, b1)); x = F.relu(F.linear(x, W2, b2)); x = F.relu(F.linear(x, W3, b3))
        x = F.linear(x, W4, b4)
        return float(torch.nn.CrossEntropyLoss()(x, torch.from_numpy(np.asarray(y, dtype=np.int64))) / div)


# ---------------------------------------------------------------------------------------------------------------
# CNN model (complex_nets/Mnist/CNN/PMP_CNN.py:22-52): conv1(1->10, 5x5) ReLU maxpool2 conv2(10->20, 3x3) ReLU fc1(2000->500) ReLU fc2(500->10)
# log_softmax; loss = CrossEntropy(mean)(log-probabilities, y) / 10 (the cross-entropy applies its own log-softmax on top: a no-op in exact arithmetic)
CNN_SHAPES = [(10, 1, 5, 5), (10,), (20, 10, 3, 3), (20,), (500, 2000), (500,), (10, 500), (10,)]   # torch parameter order
CNN_DIM = sum(int(np.prod(s)) for s in CNN_SHAPES)


def cnn_unpack(theta):
    out, off = [], 0
    for s in CNN_SHAPES:
        k = int(np.prod(s))
        out.append(np.asarray(theta[off:off + k]).reshape(s))
        off += k
    return out


def cnn_init_theta(seed=0):
    """torch's default init scale (uniform +-1/sqrt(fan_in)) from a seeded generator: a deterministic stand-in for CNN_model.pkl."""
    rng = np.random.default_rng(seed)
    fan = [25, 25, 90, 90, 2000, 2000, 500, 500]
    return np.concatenate([rng.uniform(-1, 1, size=int(np.prod(s))) / math.sqrt(f) for s, f in zip(CNN_SHAPES, fan)]).astype(np.float32)


def cnn_logits_numpy(X, theta, dtype=np.float64):
    """The forward pass PMP_CNN.py:32-42 written out with explicit loops over the filter taps (small n only)."""
    W1, b1, W2, b2, F1, c1, F2, c2 = [a.astype(dtype) for a in cnn_unpack(theta)]
    x = np.asarray(X, dtype=dtype).reshape(-1, 28, 28)
    n = len(x)
    t = np.zeros((n, 10, 24, 24), dtype)
    for ky in range(5):
        for kx in range(5):
            t += W1[None, :, 0, ky, kx, None, None] * x[:, None, ky:ky + 24, kx:kx + 24]
    t = np.maximum(t + b1[None, :, None, None], 0)
    p = t.reshape(n, 10, 12, 2, 12, 2).max(axis=(3, 5))
    u = np.zeros((n, 20, 10, 10), dtype)
    for c in range(10):
        for ky in range(3):
            for kx in range(3):
                u += W2[None, :, c, ky, kx, None, None] * p[:, None, c, ky:ky + 10, kx:kx + 10]
    u = np.maximum(u + b2[None, :, None, None], 0).reshape(n, 2000)
    h = np.maximum(u @ F1.T + c1, 0)
    return h @ F2.T + c2


def cnn_mean_ce_f64(X, y, theta, chunk=4096):
    """Ground truth: forward pass + CrossEntropyLoss(mean) in binary64 from the float32 inputs (torch float64 convolutions, chunked over rows;
    pinned against cnn_logits_numpy in tests/test_cpu_cnn.py)."""
    import torch
    import torch.nn.functional as F
    W1, b1, W2, b2, F1, c1, F2, c2 = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)) for a in cnn_unpack(theta)]
    y = np.asarray(y)
    tot = 0.0
    with torch.no_grad():
        for i in range(0, len(y), chunk):
            x = torch.from_numpy(np.ascontiguousarray(X[i:i + chunk], dtype=np.float64)).view(-1, 1, 28, 28)
            o = F.max_pool2d(F.relu(F.conv2d(x, W1, b1)), 2, 2)
            o = F.relu(F.conv2d(o, W2, b2)).reshape(len(x), -1)
            z = F.linear(F.relu(F.linear(o, F1, c1)), F2, c2)
            ls = F.log_softmax(z, dim=1)
            tot += float(-ls[torch.arange(len(x)), torch.from_numpy(y[i:i + chunk].astype(np.int64))].sum())
    return tot / len(y)


def cnn_loss_torch32(X, y, theta, div=10.0):
    """loss(net) exactly as the reference evaluates it (PMP_CNN.py:48-52): torch float32, CrossEntropyLoss()(log_softmax(net(X)), y) / 10."""
    import torch
    import torch.nn.functional as F
    W1, b1, W2, b2, F1, c1, F2, c2 = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in cnn_unpack(theta)]
    with torch.no_grad():
        x = torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32)).view(-1, 1, 28, 28)
        o = F.max_pool2d(F.relu(F.conv2d(x, W1, b1)), 2, 2)
        o = F.relu(F.conv2d(o, W2, b2)).view(len(x), -1)
        z = F.log_softmax(F.linear(F.relu(F.linear(o, F1, c1)), F2, c2), dim=1)
        return float(torch.nn.CrossEntropyLoss()(z, torch.from_numpy(np.asarray(y, dtype=np.int64))) / div)


# ---------------------------------------------------------------------------------------------------------------
# HMC variants (complex_nets/Cifar-10/cifar_{SP,MP,PMP}hmc.py, "Bayesian Network Training"/main.py): acceptance weights and the fit loops
STREAM_MOMENTUM = 4
HMC_SP, HMC_MP, HMC_TREE_CIFAR, HMC_TREE_BNN = range(4)


def hmc_weights(rule, nl, ke_out, ke_in=None):
    """B as step() hands it to torch.multinomial, float32 like the reference's tensors.  Tree rules: cifar_PMPhmc.py:77-108 / main.py:67-103 with
    ke_out[c] = |p_s[parent(c)][c]|^2/2 and ke_in[c] = |p_s[c][parent(c)]|^2/2 (indexed by the child c); MP: cifar_MPhmc.py:79-86 with ke_out[j] = |p_s[j]|^2/2."""
    f = np.float32
    nl = np.asarray(nl, dtype=f); ko = np.asarray(ke_out, dtype=f)
    P = len(nl)
    with np.errstate(all="ignore"):
        if rule == HMC_MP:
            A = np.zeros(P, dtype=f)
            for j in range(1, P):
                A[j] = np.exp(min(f(0), f(f(f(nl[j] - ko[j]) - nl[0]) + ko[0])))
            A[0] = f(P - 1) - A.sum(dtype=f)
        else:
            ki = np.asarray(ke_in, dtype=f)
            depth = int(round(math.log2(P)))
            A = np.ones(P, dtype=f)
            for a in range(P):
                for c in range(depth):
                    j = 2 ** (c + 1); half = j // 2
                    m = a % j                                                  # the `judg` reduction loop, cifar_PMPhmc.py:84-93
                    if m < half:
                        w_new = np.exp(f(nl[m] - ko[m + half])); w_old = np.exp(f(nl[m + half] - ki[m + half]))
                        if rule == HMC_TREE_CIFAR:
                            A[a] = A[a] * max(f(0), f(f(1) - f(w_old / w_new)))
                        else:
                            w_old = np.minimum(f(1), f(w_old / w_new)); w_new = np.maximum(f(0), f(f(1) - f(w_old / w_new)))   # torch.min / torch.max propagate NaN; the CIFAR script's Python min / max do not
                            A[a] = f(A[a] * w_new) / f(w_new + w_old)
                    else:
                        w_new = np.exp(f(nl[m] - ki[m])); w_old = np.exp(f(nl[m - half] - ko[m]))
                        if rule == HMC_TREE_CIFAR:
                            A[a] = A[a] * min(f(1), f(w_new / w_old))
                        else:
                            w_new = np.minimum(f(1), f(w_new / w_old)); w_old = np.maximum(f(0), f(f(1) - f(w_new / w_old)))
                            A[a] = f(A[a] * w_new) / f(w_new + w_old)
    B = A.copy()
    B[np.isnan(B)] = 1; B[np.isinf(B)] = 1
    return B


def hmc_momentum(seed, it, stream_index, dim):
    """what hmc.cu draws for torch.randn(d): float32(N(0,1)) from the momentum stream; the caller multiplies by 0.0005 like the scripts"""
    return stream_normals(seed, it, STREAM_MOMENTUM, stream_index * dim, dim).astype(np.float32)


def _flat_grad(net):
    import torch
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()])


def hmc_fit_restated(kind, net, X, y, num_steps, seed, N=3, step_size=0.1, device="cpu"):
    """The fit loops of the four HMC scripts restated on top of torch autograd (same module copies, the same accumulating .grad buffers, the same
    float32 leapfrog arithmetic), with the scripts' unseeded generators replaced by this repo's streams: torch.randn(d) -> hmc_momentum, torch.multinomial /
    torch.rand -> inverse CDF / comparison with stream_uniforms(.., STREAM_DRAW), random.uniform -> STREAM_PICK.  Returns (loss list, accepted indices, final net).
    kind: 'SP' cifar_SPhmc.py:77-137, 'MP' cifar_MPhmc.py:91-153, 'PMP' cifar_PMPhmc.py:114-172, 'BNN' main.py:104-172."""
    import copy
    import torch
    loss_fn = torch.nn.CrossEntropyLoss().to(device)
    d = sum(p.numel() for p in net.parameters())
    losses, picks = [], []

    def randn(it, si):
        return torch.from_numpy(hmc_momentum(seed, it, si, d)).to(device)

    def kick_drift(child, parent_for_grad, p, sign):
        du = _flat_grad(parent_for_grad).reshape(d)
        p += sign * step_size * du / 2
        off = 0
        for par in child.parameters():
            k = par.numel()
            par.data += sign * step_size * p[off:off + k].reshape(par.data.shape)
            off += k

    for s in range(num_steps):
        u = float(stream_uniforms(seed, s, STREAM_DRAW, 0, 1)[0])
        if kind == "SP":
            prop = copy.deepcopy(net)
            p_old = randn(s, 0) * 0.0005
            p_new = copy.deepcopy(p_old).to(device)
            x0 = -loss_fn(net(X), y)
            H0 = (p_old * p_old).sum() / 2 + x0
            x0.backward()
            kick_drift(prop, net, p_new, 1.0)
            x1 = -loss_fn(prop(X), y)
            x1.backward()
            p_new += step_size * _flat_grad(prop).reshape(d) / 2
            p1 = (p_new * p_new).sum() / 2
            x1 = -loss_fn(prop(X), y)
            H1 = x1 + p1
            acc = bool(torch.exp((-H0 + H1) * 1000) > u)
            if acc:
                net = prop
                losses.append(float(-x1.data))
            else:
                losses.append(float(-x0.data))
            picks.append(int(acc))
            continue
        nets = [None] * (N + 1); nl = [None] * (N + 1)
        nets[0] = copy.deepcopy(net).to(device)
        if kind == "MP":
            p_s = [None] * (N + 1)
            p_s[0] = randn(s, 0) * 0.0005
            ranint = int(1 + float(stream_uniforms(seed, s, STREAM_PICK, 0, 1)[0]) * N)       # int(random.uniform(1, N + 1))
            sign = 1.0
            for i in range(N):
                if i >= ranint:
                    sign = -1.0
                j = i + 1
                nets[j] = copy.deepcopy(nets[i]).to(device)
                p_s[j] = copy.deepcopy(p_s[i])
                nl[i] = -loss_fn(nets[i](X), y)
                nl[i].backward()
                kick_drift(nets[j], nets[i], p_s[j], sign)
                nl[j] = -loss_fn(nets[j](X), y)
                nl[j].backward()
                p_s[j] += sign * step_size * _flat_grad(nets[j]).reshape(d) / 2
            ko = [float((p * p).sum() / 2) for p in p_s]
            B = hmc_weights(HMC_MP, [float(v.detach()) for v in nl], ko)
        else:
            depth = int(round(math.log2(N + 1)))
            ko = [0.0] * (N + 1); ki = [0.0] * (N + 1)
            for i in range(depth):
                j = 2 ** i
                for k in range(j):
                    p0 = randn(s, k + j) * 0.0005
                    nets[k + j] = copy.deepcopy(nets[k]).to(device)
                    p = copy.deepcopy(p0)
                    nl[k] = -loss_fn(nets[k](X), y)
                    nl[k].backward()
                    kick_drift(nets[k + j], nets[k], p, 1.0)
                    nl[k + j] = -loss_fn(nets[k + j](X), y)
                    nl[k + j].backward()
                    p += step_size * _flat_grad(nets[k + j]).reshape(d) / 2
                    ko[k + j] = float((p0 * p0).sum() / 2); ki[k + j] = float((p * p).sum() / 2)
            B = hmc_weights(HMC_TREE_CIFAR if kind == "PMP" else HMC_TREE_BNN, [float(v.detach()) for v in nl], ko, ki)
        I = int(draw_numpy(B.astype(np.float64), [u])[0])
        net = nets[I]
        losses.append(float(-nl[I].detach()))
        picks.append(I)
    return losses, picks, net


# ---- d-dimensional GLM heads (extension of the simple-net model; no reference counterpart: plain binary64) ----
def loglik_glm_f64(X, y, thetas, kind, scale=1.0):
    """kind 'logistic': sum_i log sigmoid(s_i x_i.theta), s_i = 2 y_i - 1.  kind 'gauss': theta = (coefficients, sigma),
    sum_i log N(y_i; x_i.coef, sigma^2) — lb.py:103-108 with a d-vector covariate.  Returns loglik / scale per row of thetas."""
    X = np.asarray(X, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    th = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
    out = np.empty(len(th))
    for p, t in enumerate(th):
        if kind == "logistic":
            u = (2.0 * (y > 0.5) - 1.0) * (X @ t)
            out[p] = -np.sum(np.maximum(-u, 0.0) + np.log1p(np.exp(-np.abs(u))))
        else:
            r = y - X @ t[:-1]
            sg = t[-1]
            out[p] = -0.5 * len(y) * np.log(2.0 * np.pi * sg * sg) - 0.5 * np.sum(r * r) / (sg * sg)
    return out / scale
