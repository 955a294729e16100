"""The reference's network samplers for an ARBITRARY `loss(net)` callable (the CNN / LSTM scripts).

complex_nets/Mnist/{CNN,LSTM}/{MH,MP,PMP}_*.py repeat the FC optimizers byte for byte around a different `Model` and the
same `loss(net)` contract (a 0-d tensor, CrossEntropy / 10; un-divided in the MH scripts).  Here the P forward passes stay
with the caller's callable — any nn.Module, any device — and everything around them runs on the hot path:

    proposals     Philox increments for the flattened parameter vector on the device (pmp_propose), or the caller's own
                  `proposal_nets` (the reference's injection seam: step(s, proposal_nets, proposal_nets_paras, para_num))
    log-targets   -loss(net_p), handed over with pmp_write_logtarget (PMP_TARGET_EXTERNAL)
    weights+draw  the same acceptance kernels as the FC path: MP_FC.py:102-122 (mean-kernel term, standardised),
                  PMP_FC.py:105-143 (binary-tree Barker product, standardised), MH_FC.py:92-119 (exp(lamb * (loss - loss')))

A fused sweep for a given architecture (as csrc/fc_sweep.cu is for the 784-512-256-128-10 MLP) plugs in below this
interface without changing it.
"""
import copy
import math

import numpy as np
import torch

from . import _lib as L
from . import dist as _dist


def flatten(net):
    """torch.cat of the parameters in order (PMP_FC.py:173-174)."""
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy().astype(np.float32)


def unflatten(theta, like):
    net = copy.deepcopy(like)
    off = 0
    with torch.no_grad():
        for p in net.parameters():
            k = p.numel()
            p.copy_(torch.from_numpy(np.asarray(theta[off:off + k], dtype=np.float32)).view_as(p))
            off += k
    return net


class _Base:
    algo, flags, scale, temperature = L.ALGO_MP, 0, 1.0, 1.0

    def __init__(self, net, alpha, loss, N=7, seed=0, ctx=None):
        self.net, self.alpha, self.loss_callable, self.N, self.seed = net, alpha, loss, N, seed
        self.ctx = ctx or _dist.default_context()
        self.d = sum(p.numel() for p in net.parameters())
        self.sigma = 1
        self.loss = None
        self.loss_list = []
        self._iteration = 0

    def _shape(self):
        raise NotImplementedError

    @torch.no_grad()
    def update(self, net):
        """PMP_FC.py:96-102: every parameter moved by N(0, alpha); increments from the Philox chain-init stream."""
        th = flatten(net)
        z = L.stream_normals(self.seed, self._iteration, 3, 0, th.size).astype(np.float32)
        self._iteration += 1
        return unflatten(th + np.float32(self.alpha) * z, like=net)

    @torch.no_grad()
    def _step(self, proposal_nets=None, uniforms=None):
        tree, b, depth = self._shape()
        c = self.ctx
        c.configure(tree, b=b, depth=depth, dim=self.d, target=L.TARGET_EXTERNAL, algo=self.algo, draw=L.DRAW_SINGLE, flags=self.flags,
                    alpha=float(self.alpha), scale=self.scale, kernel_sigma=float(self.sigma), mh_temperature=self.temperature)
        c.seed(self.seed, self._iteration)
        if proposal_nets is None:
            c.set_state(flatten(self.net))
            c.propose()
            props = c.read_proposals()
            nets = [self.net] + [unflatten(props[p], like=self.net) for p in range(1, len(props))]
        else:
            nets = list(proposal_nets)
            props = np.stack([flatten(n) for n in nets])
            c.set_state(props[0])
            c.write_proposals(props)
        losses = np.array([float(self.loss_callable(n)) for n in nets], dtype=np.float64)      # the caller's P forward passes
        c.write_logtarget(-losses)
        _, nxt = c.accept(uniforms)
        self._iteration += 1
        self.net = nets[nxt]
        self.loss = float(losses[nxt])
        self.loss_list.append(self.loss)
        return nxt, losses

    def fit(self, num_steps=1000):
        for s in range(num_steps):
            self.step(s)
        return np.array(self.loss_list)


class MetropolisOptimizer(_Base):
    """MH_FC.py:73-134 / MH_CNN.py: accept iff u < exp(lamb * (loss - loss_proposal)), lamb = 10000."""
    algo, temperature = L.ALGO_MH, 10000.0

    def __init__(self, net, alpha, loss, seed=0, ctx=None):
        super().__init__(net, alpha, loss, N=1, seed=seed, ctx=ctx)
        self.lamb = 10000

    def _shape(self):
        return L.TREE_FLAT, 2, 1

    def step(self, s=0, uniforms=None):
        self.temperature = float(self.lamb)
        _, losses = self._step(None, uniforms)
        self.loss_proposal = float(losses[1])
        return self.net


class MPOptimizer(_Base):
    """MP_FC.py:77-158 / MP_CNN.py: A_j = sum_k mean_dim logK(j,k) / (N+1) - loss_j, standardised, one draw."""
    algo, flags = L.ALGO_MP, L.FLAG_STANDARDIZE | L.FLAG_KERNEL_MEAN

    def _shape(self):
        return L.TREE_FLAT, self.N + 1, 1

    def step(self, s=0, proposal_nets=None, proposal_nets_paras=None, para_num=None, uniforms=None):
        self._step(proposal_nets, uniforms)
        return self.net


class PMPOptimizer(_Base):
    """PMP_FC.py:79-186 / PMP_CNN.py: binary prefetch tree (N+1 = 2^D), per-level Barker product on exp(-loss), standardised."""
    algo, flags = L.ALGO_PSP, L.FLAG_STANDARDIZE

    def _shape(self):
        depth = int(math.log2(self.N + 1))
        if 2 ** depth != self.N + 1:
            raise ValueError("binary prefetch tree needs N+1 = 2^D nodes")
        return L.TREE_BINARY, 2, depth

    step = MPOptimizer.step
