"""The reference's Bayesian FC samplers (complex_nets/Mnist/FC/{MH,MP,PMP}_FC.py) on the B200 path.

Same names as the scripts: `Model` (PMP_FC.py:21-36), `loss(net)` (PMP_FC.py:40-44: CrossEntropy(mean)/10 on the module-level
`X`, `y`), `MetropolisOptimizer` (MH_FC.py:73-134), `MPOptimizer` (MP_FC.py:77-158), `PMPOptimizer` (PMP_FC.py:79-186) with
`update`, `step(s, proposal_nets, proposal_nets_paras, para_num)` and `fit(num_steps) -> np.array(loss_list)`.
The scripts download MNIST at import; here the data are given with `set_data(X, y)` (BASELINE config 5 uses synthetic
MNIST-shaped data).  All P forward passes of an iteration run as one tcgen05 GEMM chain on the device (csrc/fc_sweep.cu).
`torch.multinomial` (an exponential race in current torch) is replaced by one inverse-CDF draw from the library's Philox
stream — same distribution, and reproducible; pass `uniforms=` to `step` to inject the draw.
"""
import copy
import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import dist as _dist

X = None          # [n, 28, 28] or [n, 784] float32 — module-level like the reference scripts
y = None          # [n] int64
_ctx = None
_data_version = 0
FC_DIM = 567434


def set_data(X_, y_, ctx=None):
    """Replaces the MNIST block PMP_FC.py:47-74: registers the training set and uploads this rank's shard."""
    global X, y, _ctx, _data_version
    X = torch.as_tensor(X_, dtype=torch.float32)
    y = torch.as_tensor(y_, dtype=torch.int64)
    _ctx = ctx or _dist.default_context()
    Xn = X.reshape(len(y), -1).numpy()
    lo, hi = _dist.shard_bounds(len(y), _ctx.world_size, _ctx.rank, align=128)
    _ctx.set_data_fc(Xn[lo:hi], y.numpy()[lo:hi], n_offset=lo, n_global=len(y))
    _data_version += 1


class Model(torch.nn.Module):
    """PMP_FC.py:21-36."""

    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(28 * 28, 512)
        self.fc2 = nn.Linear(512, 256)
        self.fc3 = nn.Linear(256, 128)
        self.fc4 = nn.Linear(128, 10)

    def forward(self, x):
        x = x.view(-1, 28 * 28)
        x = F.relu(self.fc1(x))
        x = F.relu(self.fc2(x))
        x = F.relu(self.fc3(x))
        return self.fc4(x)


def flatten(net):
    """torch.cat of the parameters in order (PMP_FC.py:173-174)."""
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy().astype(np.float32)


def unflatten(theta, like=None, model=None):
    net = (model or Model)() if like is None else copy.deepcopy(like)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            k = p.numel()
            p.copy_(torch.from_numpy(np.asarray(theta[off:off + k], dtype=np.float32)).view_as(p))
            off += k
    return net


def mean_ce_batch(thetas):
    """Mean cross-entropy of every row of `thetas` ([P, 567434]) over the registered data: one device sweep."""
    if _ctx is None:
        raise RuntimeError("fc.set_data(X, y) first")
    thetas = np.ascontiguousarray(thetas, dtype=np.float32)
    _ctx.configure(L.TREE_FLAT, b=len(thetas), dim=FC_DIM, target=L.TARGET_FC, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE,
                   flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=1.0)
    _ctx.write_proposals(thetas)
    return -_ctx.loglik()


@torch.no_grad()
def loss(net):
    """PMP_FC.py:40-44 (MP_FC.py:70-74): CrossEntropyLoss()(net(X), y) / 10 as a 0-d tensor."""
    return torch.tensor(mean_ce_batch(flatten(net)[None, :])[0] / 10.0, dtype=torch.float32)


class _FCBase:
    tree, algo, flags, draw, scale, temperature = L.TREE_FLAT, L.ALGO_MP, 0, L.DRAW_SINGLE, 10.0, 1.0
    _target, _dim = L.TARGET_FC, FC_DIM          # cnn.py reuses these samplers with the CNN target (the reference's optimizer classes are the same code)

    @staticmethod
    def _context():
        return _ctx

    def __init__(self, net, alpha, seed=0):
        self.net = net
        self.alpha = alpha
        self.first = 0
        self.loss = None
        self.loss_proposal = None
        self.lamb = 10000
        self.loss_fn = torch.nn.CrossEntropyLoss()
        self.N = 7
        self.d = sum(p.numel() for p in self.net.parameters())
        self.sigma = 1
        self.loss_list = []
        self.seed = seed
        self._iteration = 0

    @torch.no_grad()
    def update(self, net):
        """PMP_FC.py:96-102: every parameter moved by N(0, alpha); increments from the Philox chain-init stream."""
        th = flatten(net)
        z = L.stream_normals(self.seed, self._iteration, 3, 0, th.size).astype(np.float32)
        self._iteration += 1
        return unflatten(th + np.float32(self.alpha) * z, like=net)

    def _shape(self):
        raise NotImplementedError

    def _configure(self):
        tree, b, depth = self._shape()
        ctx = self._context()
        if ctx is None:
            raise RuntimeError("set_data(X, y) first")
        ctx.configure(tree, b=b, depth=depth, dim=self._dim, target=self._target, algo=self.algo, draw=self.draw, flags=self.flags,
                      alpha=float(self.alpha), scale=self.scale, kernel_sigma=float(self.sigma), mh_temperature=self.temperature)
        return ctx

    def _step_device(self, uniforms=None):
        """propose → sweep → accept on the device; returns (accepted index, mean CE of every node)."""
        ctx = self._configure()
        ctx.set_state(flatten(self.net))
        ctx.seed(self.seed, self._iteration)
        ctx.propose()
        lt = ctx.loglik()
        _, nxt = ctx.accept(uniforms)
        self._iteration += 1
        self.net = unflatten(ctx.get_state(), like=self.net)
        return nxt, -lt * self.scale

    def _step_external(self, proposal_nets, uniforms=None):
        ctx = self._configure()
        props = np.stack([flatten(n) for n in proposal_nets])
        ctx.set_state(props[0])
        ctx.seed(self.seed, self._iteration)
        ctx.write_proposals(props)
        lt = ctx.loglik()
        _, nxt = ctx.accept(uniforms)
        self._iteration += 1
        return nxt, -lt * self.scale


class MetropolisOptimizer(_FCBase):
    """MH_FC.py:73-134: accept iff u < exp(lamb * (loss - loss_proposal)), lamb = 10000, un-divided loss (MH_FC.py:67-71,99)."""
    algo, scale, temperature = L.ALGO_MH, 1.0, 10000.0

    def _shape(self):
        return L.TREE_FLAT, 2, 1

    def step(self, s, uniforms=None):
        nxt, ce = self._step_device(uniforms)
        self.loss_proposal = float(ce[1])
        self.loss = float(ce[nxt])
        self.loss_list.append(self.loss)
        return self.net

    def fit(self, num_steps=1000):
        for s in range(num_steps):
            self.step(s)
        return np.array(self.loss_list)


class MPOptimizer(_FCBase):
    """MP_FC.py:77-158: A_j = sum_k mean_dim logK(j,k) / (N+1) - loss_j, standardised, one multinomial draw."""
    algo, flags = L.ALGO_MP, L.FLAG_STANDARDIZE | L.FLAG_KERNEL_MEAN

    def _shape(self):
        return L.TREE_FLAT, self.N + 1, 1

    def step(self, s, proposal_nets=None, proposal_nets_paras=None, para_num=None, uniforms=None):
        if proposal_nets is None:
            nxt, ce = self._step_device(uniforms)
        else:
            nxt, ce = self._step_external(proposal_nets, uniforms)
            self.net = proposal_nets[nxt]
        self.loss = float(ce[nxt])                       # loss_fn(self.net(X), y).item(), MP_FC.py:125-127
        self.loss_list.append(self.loss)
        return self.net

    def fit(self, num_steps=1000):
        for s in range(num_steps):
            self.step(s)
        return np.array(self.loss_list)


class PMPOptimizer(_FCBase):
    """PMP_FC.py:79-186: binary prefetch tree (N+1 = 2^D), per-level Barker product on exp(-loss), standardised, one draw."""
    algo, flags = L.ALGO_PSP, L.FLAG_STANDARDIZE

    def _shape(self):
        depth = int(math.log2(self.N + 1))
        return L.TREE_BINARY, 2, depth

    step = MPOptimizer.step
    fit = MPOptimizer.fit
