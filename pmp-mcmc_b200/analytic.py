"""The reference's analytic-target samplers on the B200 path.

Same names and signatures as simple_sampling/error/error.py and complex_nets/correlation/com_dim.py:

    normal(x, mu, sigma)            error.py:11-14      density of the 1-D target
    SP(hops, mu, sigma)             error.py:17-40      Barker single proposal, uniform(-0.25, 0.25) step
    MP(hops, mu, sigma, N)          error.py:43-77      N proposals, multi-proposal weights, N+1 draws per hop
    PSP(hops, mu, sigma, N)         error.py:78-134     binary prefetch tree (N+1 = 2^D nodes)
    PMP(hops, mu, sigma, N, deep)   error.py:137-190    (N+1)-ary tree of depth `deep`
    banana_distribution(x)          banana_data.ipynb cell 2
    PMP_dim(hops, mu, cov, N, dim)  com_dim.py:24-86    binary tree on N(0, I_dim), start 2.5*1, returns all resampled points

Return values follow the reference (1-D arrays after the 20 % burn-in cut; [hops*(N+1), dim] for PMP_dim).  Additions:
`seed=` (the reference is unseeded), `chains=` to run that many independent chains at once on the device (returns one more
leading axis) and `target=` ("normal" | "banana").  States are float32 on the device; weights are float64.
"""
import math

import numpy as np

from . import _lib as L
from . import dist as _dist


def normal(x, mu, sigma):
    return np.exp((-(x - mu) ** 2) / (2 * sigma ** 2)) / (sigma * np.sqrt(2 * np.pi))


def banana_distribution(x):
    x1, x2 = x[0], x[1]
    return np.exp(-(x1 ** 2) / 2) * np.exp(-((x2 - 2 * (x1 ** 2 - 5)) ** 2) / 2)


def _target_cfg(target, mu, sigma):
    if target == "normal":
        return dict(dim=1, target=L.TARGET_NORMAL1D, target_p0=float(mu), target_p1=float(sigma))
    if target == "banana":
        return dict(dim=2, target=L.TARGET_BANANA)
    raise ValueError("target must be 'normal' or 'banana'")


def _init_states(target, mu, sigma, seed, chains):
    """error.py:20,47,83,141: X0 ~ uniform(mu - sigma, mu + sigma), from the chain-initialisation stream."""
    dim = 1 if target == "normal" else 2
    out = np.empty((chains, dim), dtype=np.float32)
    for c in range(chains):
        u = L.stream_uniforms(seed, 0, 3, c << 32, dim)
        out[c] = (mu - sigma + 2.0 * sigma * u) if target == "normal" else (np.array([0.0, -10.0]) + (2.0 * u - 1.0))
    return out


def _run(cfg_kwargs, hops, seed, chains, init, ctx):
    ctx = ctx or _dist.default_context()
    ctx.configure(**cfg_kwargs)
    P, dim = ctx.P, cfg_kwargs["dim"]
    ctx.seed(seed, 0)
    if chains is None:
        ctx.set_state(init[0])
        ctx.trace_config(hops, L.TRACE_SAMPLES)
        ctx.run(hops)
        return ctx.read_trace()["samples"].reshape(hops * P, dim).astype(np.float64)
    ctx.chains_create(chains, init)
    ctx.chains_run(hops, record_samples=True)
    s = ctx.chains_read_samples()                     # [hops, P, dim, chains]
    return np.transpose(s, (3, 0, 1, 2)).reshape(chains, hops * P, dim).astype(np.float64)


def _cut(X, hops, P, chains):
    k = int(0.2 * hops * P)
    return X[k:, 0] if chains is None else X[:, k:, 0]


def SP(hops, mu, sigma, seed=0, chains=None, target="normal", ctx=None):
    cfg = dict(tree=L.TREE_FLAT, b=2, depth=1, algo=L.ALGO_BARKER, draw=L.DRAW_SINGLE, flags=L.FLAG_UNIFORM_PROPOSAL, alpha=0.25, **_target_cfg(target, mu, sigma))
    ctx = ctx or _dist.default_context()
    init = _init_states(target, mu, sigma, seed, chains or 1)
    ctx.configure(**cfg)
    ctx.seed(seed, 0)
    burn = int(hops * 0.2)
    if chains is None:                                # states[] holds the state BEFORE each move (error.py:25)
        ctx.set_state(init[0])
        ctx.trace_config(hops, L.TRACE_STATE)
        ctx.run(hops)
        st = np.concatenate([init[:1], ctx.read_trace()["state"][:-1]]).astype(np.float64)
        return st[burn:, 0] if target == "normal" else st[burn:]
    ctx.chains_create(chains, init)
    ctx.chains_run(hops, record_samples=True)
    s = np.transpose(ctx.chains_read_samples(), (3, 0, 1, 2))[:, :, 1, :]     # row 1 of a P=2 record = the new state of that hop
    st = np.concatenate([init[:, None, :], s[:, :-1, :]], axis=1).astype(np.float64)
    return st[:, burn:, 0] if target == "normal" else st[:, burn:]


def MP(hops, mu, sigma, N, seed=0, chains=None, target="normal", ctx=None):
    cfg = dict(tree=L.TREE_FLAT, b=N + 1, depth=1, algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=1.0, **_target_cfg(target, mu, sigma))
    X = _run(cfg, hops, seed, chains, _init_states(target, mu, sigma, seed, chains or 1), ctx)
    return _cut(X, hops, N + 1, chains) if target == "normal" else X


def PSP(hops, mu, sigma, N, seed=0, chains=None, target="normal", ctx=None):
    depth = int(math.log2(N + 1))
    if 2 ** depth != N + 1:
        raise ValueError("PSP needs N + 1 to be a power of two")
    cfg = dict(tree=L.TREE_BINARY, b=2, depth=depth, algo=L.ALGO_PSP, draw=L.DRAW_PYTHON, alpha=1.0, **_target_cfg(target, mu, sigma))
    X = _run(cfg, hops, seed, chains, _init_states(target, mu, sigma, seed, chains or 1), ctx)
    return _cut(X, hops, N + 1, chains) if target == "normal" else X


def PMP(hops, mu, sigma, N, deep, seed=0, chains=None, target="normal", ctx=None, quirk_level_mod=True):
    """quirk_level_mod=True keeps error.py:173's `% ((N+1)*(i+1))` (identical to the intended rule for deep <= 2)."""
    cfg = dict(tree=L.TREE_BARY, b=N + 1, depth=deep, algo=L.ALGO_PMP, draw=L.DRAW_PYTHON, alpha=1.0,
               flags=L.FLAG_QUIRK_LEVEL_MOD if quirk_level_mod else 0, **_target_cfg(target, mu, sigma))
    X = _run(cfg, hops, seed, chains, _init_states(target, mu, sigma, seed, chains or 1), ctx)
    return _cut(X, hops, (N + 1) ** deep, chains) if target == "normal" else X


def PMP_dim(hops, mu, cov, N, dim, sigma=0.5, seed=0, chains=None, ctx=None):
    """com_dim.py:24-86 (the module-global proposal std `sigma` is an argument here).  Only the target the script runs is
    supported: mu = 0, cov = I (com_dim.py:100-101)."""
    if np.any(np.asarray(mu) != 0) or not np.array_equal(np.asarray(cov), np.eye(dim)):
        raise ValueError("PMP_dim supports the reference's target N(0, I) only")
    depth = int(math.log2(N + 1))
    cfg = dict(tree=L.TREE_BINARY, b=2, depth=depth, dim=dim, target=L.TARGET_STDNORMAL, algo=L.ALGO_PSP, draw=L.DRAW_PYTHON,
               alpha=float(sigma), kernel_sigma=float(sigma))
    init = np.full((chains or 1, dim), 2.5, dtype=np.float32)            # com_dim.py:28
    return _run(cfg, hops, seed, chains, init, ctx)
