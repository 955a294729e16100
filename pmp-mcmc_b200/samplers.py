"""The reference's simple-net entry points (simple_net/lb.py) on the B200 hot path.

Same class names, constructor arguments, method names and return shapes as lb.py, so the experiment section of that
script (lb.py:377-423) runs unchanged against this module:

    BayesNet / BayesNet_o       lb.py:20-42, 91-108   (beta0, beta, sigma; loglik = sum log N(y; b0+b x, |sigma|) * 50 / n)
    log_trans_prob              lb.py:111-116
    MetropolisOptimizer         lb.py:47-85           fit → [num_steps, 3]
    GMOptimizer                 lb.py:122-186         fit → [num_steps*(N+1), 3]   (all resampled candidates per step)
    preMOptimizer               lb.py:189-279         fit → [num_steps, 3]          (binary prefetch tree, N+1 = 2^D nodes)
    GMpreOptimizerV2            lb.py:286-369         fit → [num_steps*(N+1)^deep, 3]

What changes underneath: `fit` runs the whole chain on the device (Philox proposals, the P x n sweep, the acceptance and
the trace all stay in HBM; one D2H copy at the end); `step(data, proposal_nets)` still takes externally supplied
proposals (the reference's stream-injection seam) and evaluates them with the same kernels.  The reference is unseeded;
here every optimizer takes `seed=` (default 0) and is reproducible.  Extra keyword arguments (`ctx`, `seed`,
`uniforms`) are additions; positional signatures are the reference's.
"""
import math

import numpy as np

from . import _lib as L
from . import dist as _dist

try:
    import torch
    import torch.nn as nn
except ImportError:  # torch is only needed for the nn.Module flavour of the nets
    torch = None
    nn = None


def _to_np(v):
    if torch is not None and isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


_Base = nn.Module if nn is not None else object


class BayesNet(_Base):
    """lb.py:91-108.  Parameters in named_parameters() order: beta0, beta, sigma."""

    def __init__(self, seed=42):
        super().__init__()
        if torch is not None:
            torch.random.manual_seed(seed)
            self.beta0 = nn.Parameter(torch.tensor([0.0]))
            self.beta = nn.Parameter(torch.tensor([0.0]))
            self.sigma = nn.Parameter(torch.tensor([1.0]))
        else:
            self.beta0, self.beta, self.sigma = np.zeros(1, np.float32), np.zeros(1, np.float32), np.ones(1, np.float32)

    def forward(self, data):
        return self.beta0 + self.beta * data["x"]

    def theta(self):
        return np.array([float(_to_np(self.beta0)[0]), float(_to_np(self.beta)[0]), float(_to_np(self.sigma)[0])], dtype=np.float32)

    def set_theta(self, th):
        th = np.asarray(th, dtype=np.float32)
        if torch is not None:
            with torch.no_grad():
                self.beta0.copy_(torch.tensor([th[0]])); self.beta.copy_(torch.tensor([th[1]])); self.sigma.copy_(torch.tensor([th[2]]))
        else:
            self.beta0[0], self.beta[0], self.sigma[0] = th
        return self

    def loglik(self, data, ctx=None):
        """One proposal-evaluation on the device (lb.py:103-108).  Returns a 0-d tensor like the reference."""
        v = loglik_batch(data, self.theta()[None, :], ctx=ctx)[0]
        return torch.tensor(v, dtype=torch.float32) if torch is not None else np.float32(v)


class BayesNet_o(BayesNet):
    """lb.py:20-42: the same model with logprior 0 and logpost = loglik."""

    def logprior(self):
        return torch.tensor(0.0) if torch is not None else 0.0

    def logpost(self, data, ctx=None):
        return self.loglik(data, ctx=ctx) + self.logprior()


def _net_from_theta(th, like=None):
    net = (type(like) if like is not None else BayesNet)()
    return net.set_theta(th)


def loglik_batch(data, thetas, ctx=None, scale=None):
    """[net.loglik(data) for net in nets] (the loop lb.py:150/214/312) as ONE P x n sweep.  thetas: [P,3]."""
    ctx = ctx or _dist.default_context()
    x, y = _to_np(data["x"]).reshape(-1), _to_np(data["y"]).reshape(-1)
    thetas = np.ascontiguousarray(thetas, dtype=np.float32)
    n = x.size
    ctx.configure(L.TREE_FLAT, b=len(thetas), dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA,
                  flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=(n / 50.0 if scale is None else scale))
    _dist.set_data_linear_sharded(ctx, x, y)
    ctx.write_proposals(thetas)
    return ctx.loglik()


def log_trans_prob(net, net_star):
    """lb.py:111-116: sum over parameters of log N(p; p*, 1).  Host arithmetic (three numbers); float64 like the reference."""
    a, b = net.theta().astype(np.float64), net_star.theta().astype(np.float64)
    v = float(np.sum(-0.5 * math.log(2 * math.pi) - 0.5 * (a - b) ** 2))
    return torch.tensor([v], dtype=torch.float64) if torch is not None else np.array([v])


class _DeviceSampler:
    """Shared plumbing: configuration of the C-ABI context for one of the reference's optimizers."""
    tree, algo, draw, flags = L.TREE_FLAT, L.ALGO_MP, L.DRAW_PYTHON, 0

    def __init__(self, net, alpha, ctx=None, seed=0):
        self.net = net
        self.alpha = alpha
        self.d = 3
        self.seed = seed
        self._ctx = ctx
        self._iteration = 0

    # subclasses: self._b, self._depth
    def _context(self):
        if self._ctx is None:
            self._ctx = _dist.default_context()
        return self._ctx

    def _configure(self, data):
        ctx = self._context()
        x, y = _to_np(data["x"]).reshape(-1), _to_np(data["y"]).reshape(-1)
        ctx.configure(self.tree, b=self._b, depth=self._depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=self.algo, draw=self.draw,
                      flags=self.flags, alpha=float(self.alpha), scale=x.size / 50.0)     # lb.py:108: loglik * 50 / n
        _dist.set_data_linear_sharded(ctx, x, y)
        return ctx

    def update(self, net):
        """lb.py:131-136: a copy of `net` with every parameter moved by N(0, alpha) — one node of the Philox stream."""
        z = L.stream_normals(self.seed, self._iteration, 3, 0, 3).astype(np.float32)   # chain-init stream: independent of fit's proposals
        self._iteration += 1
        th = net.theta()
        return _net_from_theta((th + (np.float32(self.alpha) * z).astype(np.float32)).astype(np.float32), like=net)

    def _run(self, data, num_steps, what):
        ctx = self._configure(data)
        ctx.set_state(self.net.theta())
        ctx.seed(self.seed, self._iteration)
        ctx.trace_config(num_steps, what)
        ctx.run(num_steps)
        tr = ctx.read_trace()
        self._iteration += num_steps
        self.net = _net_from_theta(ctx.get_state(), like=self.net)
        return ctx, tr

    def _step_external(self, data, proposal_nets, uniforms=None):
        """Evaluate caller-supplied proposals (dict/list index → net) and draw; returns (draw indices, next index)."""
        ctx = self._configure(data)
        P = ctx.P
        if len(proposal_nets) != P:
            raise ValueError("expected %d proposal nets, got %d" % (P, len(proposal_nets)))
        props = np.stack([proposal_nets[i].theta() for i in range(P)])
        ctx.set_state(props[0])
        ctx.seed(self.seed, self._iteration)
        ctx.write_proposals(props)
        ctx.loglik(read=False)
        idx, nxt = ctx.accept(uniforms)
        self._iteration += 1
        return idx, nxt


class MetropolisOptimizer(_DeviceSampler):
    """lb.py:47-85: one N(0, alpha) proposal per step, accept iff u < exp(logpost' - logpost)."""
    algo, draw = L.ALGO_MH, L.DRAW_SINGLE

    def __init__(self, net, alpha, ctx=None, seed=0):
        super().__init__(net, alpha, ctx, seed)
        self._b, self._depth = 2, 1

    def step(self, data=None, uniforms=None):
        ctx = self._configure(data)
        ctx.set_state(self.net.theta())
        ctx.seed(self.seed, self._iteration)
        ctx.propose()
        ctx.loglik(read=False)
        ctx.accept(uniforms)
        self._iteration += 1
        self.net = _net_from_theta(ctx.get_state(), like=self.net)
        return self.net

    def fit(self, data=None, num_steps=1000):
        _, tr = self._run(data, num_steps, L.TRACE_STATE)
        return tr["state"].astype(np.float64)


class GMOptimizer(_DeviceSampler):
    """lb.py:122-186: N proposals around the current state, weights lb.py:144-150, N+1 draws with replacement."""
    tree, algo, draw = L.TREE_FLAT, L.ALGO_MP, L.DRAW_PYTHON

    def __init__(self, net, alpha, N, ctx=None, seed=0):
        super().__init__(net, alpha, ctx, seed)
        self.N = N
        self._b, self._depth = N + 1, 1

    def step(self, data=None, proposal_nets=None, uniforms=None):
        idx, nxt = self._step_external(data, proposal_nets, uniforms)
        new_proposal_nets = {j: proposal_nets[int(i)] for j, i in enumerate(idx)}      # lb.py:158-160
        self.net = proposal_nets[int(nxt)]                                              # lb.py:162-163
        return new_proposal_nets

    def fit(self, data=None, num_steps=1000):
        ctx, tr = self._run(data, num_steps, L.TRACE_SAMPLES)
        return tr["samples"].reshape(num_steps * ctx.P, 3).astype(np.float64)           # lb.py:169,181-185


class preMOptimizer(_DeviceSampler):
    """lb.py:189-279: binary prefetch tree with N+1 = 2^D nodes, per-level Barker product (lb.py:216-240)."""
    tree, algo, draw = L.TREE_BINARY, L.ALGO_PSP, L.DRAW_PYTHON

    def __init__(self, net, alpha, N, ctx=None, seed=0):
        super().__init__(net, alpha, ctx, seed)
        self.N = N
        self._depth = int(math.log2(N + 1))
        if 2 ** self._depth != N + 1:
            raise ValueError("preMOptimizer needs N + 1 to be a power of two (lb.py:209)")
        self._b = 2

    def step(self, data=None, proposal_nets=None, uniforms=None):
        idx, nxt = self._step_external(data, proposal_nets, uniforms)
        new_proposal_nets = {j: proposal_nets[int(i)] for j, i in enumerate(idx)}
        self.net = proposal_nets[int(nxt)]
        return new_proposal_nets

    def fit(self, data=None, num_steps=1000):
        _, tr = self._run(data, num_steps, L.TRACE_STATE)                               # lb.py:263,275-278: the state per step
        return tr["state"].astype(np.float64)


class GMpreOptimizerV2(_DeviceSampler):
    """lb.py:286-369: (N+1)-ary tree of depth `deep`; per-level multi-proposal weights multiplied down the tree.
    `quirk_level_mod=True` reproduces the reference's `% ((N+1)*(i+1))` propagation (lb.py:330), identical for deep <= 2."""
    tree, algo, draw = L.TREE_BARY, L.ALGO_PMP, L.DRAW_PYTHON

    def __init__(self, net, alpha, N, deep, ctx=None, seed=0, quirk_level_mod=False):
        super().__init__(net, alpha, ctx, seed)
        self.N, self.deep = N, deep
        self._b, self._depth = N + 1, deep
        self.flags = L.FLAG_QUIRK_LEVEL_MOD if quirk_level_mod else 0

    def step(self, data=None, proposal_nets=None, uniforms=None):
        idx, nxt = self._step_external(data, proposal_nets, uniforms)
        new_proposal_nets = {j: proposal_nets[int(i)] for j, i in enumerate(idx)}
        self.net = proposal_nets[int(nxt)]
        return new_proposal_nets

    def fit(self, data=None, num_steps=1000):
        ctx, tr = self._run(data, num_steps, L.TRACE_SAMPLES)
        return tr["samples"].reshape(num_steps * ctx.P, 3).astype(np.float64)           # lb.py:350,364-368


MAX_CO_SCHEDULED = 32      # csrc/chain_persistent_multi.cuh: PERSIST_MAX_CHAINS


def fit_independent(trainers, data, num_steps=1000, max_group=MAX_CO_SCHEDULED):
    """Independent chains co-scheduled in one cooperative kernel (pmp_run_multi).

    The reference runs its experiments as independent repeats — error.py:191-213 (20 repeats per sampler), lb.py:377-423 (one
    chain per step size) — one after the other.  Here up to 32 `GMOptimizer` (or `preMOptimizer`) trainers with the same
    N run at once (more are run in groups of `max_group`; every group aliases the one device copy of the data that the first
    trainer's context owns): one chain alone leaves the sweep SMs idle while it is being accepted.  Every trainer keeps its own seed,
    step size and start state, gets exactly the trace `trainer.fit(data, num_steps)` would return (bit for bit), and ends in
    the same state.  Returns the list of traces in trainer order."""
    trainers = list(trainers)
    if not trainers:
        return []
    kind = type(trainers[0])
    if kind not in (GMOptimizer, preMOptimizer) or any(type(t) is not kind or t.N != trainers[0].N for t in trainers):
        raise ValueError("fit_independent takes GMOptimizer or preMOptimizer trainers of one kind and one N")
    what = L.TRACE_SAMPLES if kind is GMOptimizer else L.TRACE_STATE
    x, y = _to_np(data["x"]).reshape(-1), _to_np(data["y"]).reshape(-1)
    ctxs = []
    for k, t in enumerate(trainers):
        if t._ctx is None or any(t._ctx is c for c in ctxs):
            t._ctx = _dist.create_context()                    # every chain needs a context of its own
        ctx = t._ctx
        ctx.configure(t.tree, b=t._b, depth=t._depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=t.algo, draw=t.draw, flags=t.flags,
                      alpha=float(t.alpha), scale=x.size / 50.0)
        if k == 0:
            _dist.set_data_linear_sharded(ctx, x, y)
        else:
            ctx.share_data_from(ctxs[0])
        ctx.set_state(t.net.theta())
        ctx.seed(t.seed, t._iteration)
        ctx.trace_config(num_steps, what)
        ctxs.append(ctx)
    out = []
    max_group = max(1, min(int(max_group), MAX_CO_SCHEDULED))
    for lo in range(0, len(ctxs), max_group):
        group = ctxs[lo:lo + max_group]
        if len(group) == 1:
            group[0].run(num_steps)                          # a leftover chain runs alone: same bits (tests/test_gpu_multichain.py)
        else:
            L.run_multi(group, num_steps)
    for t, ctx in zip(trainers, ctxs):
        tr = ctx.read_trace()
        t._iteration += num_steps
        t.net = _net_from_theta(ctx.get_state(), like=t.net)
        out.append(tr["samples"].reshape(num_steps * ctx.P, 3).astype(np.float64) if kind is GMOptimizer else tr["state"].astype(np.float64))
    return out
