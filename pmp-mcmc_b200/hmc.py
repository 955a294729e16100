"""The reference's gradient (HMC) samplers on the B200 path.

Same names as the scripts: `LeNet` (cifar_PMPhmc.py:25-56), `HMCOptimizer` (cifar_SPhmc.py:65-143), `MPHMCOptimizer` (cifar_MPhmc.py:67-153),
`PMPHMCOptimizer` (cifar_PMPhmc.py:65-172) and `bnnPMPHmc` ("Bayesian Network Training"/main.py:55-172), each with `step(...)` and `fit(num_steps=...)`
returning what the script returns.  The scripts read module-level `X, y, x_test, y_test`; here they are registered with `set_data`.

What runs where: the potential -CrossEntropy(net(X), y) and its gradient are the caller's network under torch autograd on the GPU (an arbitrary
module: LeNet with BatchNorm, torchbnn layers) — the same contract as nets.py for the likelihood callables.  Everything the scripts do on flat
parameter vectors is the device library's: the momentum draw (Philox stream instead of the unseeded torch.randn), the leapfrog half kicks and
drift, the kinetic energies, the acceptance weights of the tree / path and the categorical draw (csrc/hmc.cu through the C-ABI, device pointers of
torch tensors).  The module-copy semantics of the scripts are kept as they are (children are deep copies of their parents, `.grad` buffers are never
zeroed and therefore accumulate over repeated uses of a parent) because the trajectories depend on them.
"""
import copy
import math

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import dist as _dist

X = y = x_test = y_test = None
_ctx = None
device = "cuda:0"


def set_data(X_, y_, x_test_=None, y_test_=None, ctx=None):
    """Replaces the CIFAR / MNIST download blocks (cifar_PMPhmc.py:12-22, main.py:31-52)."""
    global X, y, x_test, y_test, _ctx, device
    _ctx = ctx or _dist.default_context()
    device = "cuda:%d" % _ctx.device
    X = torch.as_tensor(X_, dtype=torch.float32).to(device)
    y = torch.as_tensor(y_, dtype=torch.int64).to(device)
    x_test = X if x_test_ is None else torch.as_tensor(x_test_, dtype=torch.float32).to(device)
    y_test = y if y_test_ is None else torch.as_tensor(y_test_, dtype=torch.int64).to(device)


class Flatten(nn.Module):
    def forward(self, input):
        return input.view(input.size(0), -1)


class LeNet(nn.Module):
    """cifar_PMPhmc.py:33-56."""

    def __init__(self):
        super().__init__()
        self.model = nn.Sequential(
            nn.Conv2d(3, 6, kernel_size=5, stride=1, padding=0), nn.BatchNorm2d(6), nn.ReLU(inplace=True), nn.MaxPool2d(2, stride=2),
            nn.Conv2d(6, 16, kernel_size=5, stride=1, padding=0), nn.BatchNorm2d(16), nn.ReLU(inplace=True), nn.MaxPool2d(2, stride=2),
            Flatten(), nn.Linear(16 * 5 * 5, 120), nn.ReLU(inplace=True), nn.Linear(120, 84), nn.ReLU(inplace=True), nn.Linear(84, 10))

    def forward(self, x):
        return self.model(x)


def _flat(net):
    return torch.cat([p.data.reshape(-1) for p in net.parameters()]).contiguous()


def _flat_grad(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).contiguous()


def _assign(net, flat):
    off = 0
    for p in net.parameters():
        k = p.numel()
        p.data.copy_(flat[off:off + k].view_as(p.data))
        off += k


class _HMCBase:
    p_scale = 0.0005                                     # torch.randn(d) * 0.0005 in every script

    def __init__(self, net, alpha, N=1, seed=0):
        if _ctx is None:
            raise RuntimeError("hmc.set_data(X, y) first")
        self.net = net.to(device)
        self.alpha = alpha
        self.N = N
        self.d = sum(p.numel() for p in self.net.parameters())
        self.loss = torch.nn.CrossEntropyLoss().to(device)
        self.loss_list, self.train_acc, self.test_acc = [], [], []
        self.seed = seed
        self.picks = []
        self._theta = torch.empty(self.d, dtype=torch.float32, device=device)

    def _potential(self, net):
        """U(x) = -log p(x) with the scripts' sign: nets_loss = -CrossEntropy, and its gradient accumulated into net's .grad (cifar_PMPhmc.py:134-137)"""
        nl = -self.loss(net(X), y)
        nl.backward()
        return nl

    def _edge(self, parent, child, s, stream_index, step_size, sign=1.0, p=None):
        """one leapfrog step parent -> child (cifar_PMPhmc.py:128-162 / cifar_MPhmc.py:104-142): returns (nl_parent, nl_child, p, K(p0), K(p_final))"""
        nl_p = self._potential(parent)
        _ctx.seed(self.seed, s)
        fresh = p is None
        if fresh:
            p = torch.empty(self.d, dtype=torch.float32, device=device)
        torch.cuda.current_stream().synchronize()
        ke0 = _ctx.hmc_leapfrog_begin(_flat(parent), _flat_grad(parent), self._theta, p, step_size, sign, self.p_scale, stream_index, p_init=None if fresh else p)
        _assign(child, self._theta)
        nl_c = self._potential(child)
        torch.cuda.current_stream().synchronize()
        ke1 = _ctx.hmc_leapfrog_end(p, _flat_grad(child), step_size, sign)
        return nl_p, nl_c, p, ke0, ke1

    def _uniform(self, s, stream=1):
        return float(L.stream_uniforms(self.seed, s, stream, 0, 1)[0])

    def _accuracies(self, n_train, n_test):
        with torch.no_grad():
            self.train_acc.append((self.net(X).argmax(1) == y).type(torch.float).sum().item() / n_train)
            self.test_acc.append((self.net(x_test).argmax(1) == y_test).type(torch.float).sum().item() / n_test)


class HMCOptimizer(_HMCBase):
    """cifar_SPhmc.py:65-143: one leapfrog proposal, accept iff exp((-H_0 + H_1) * 1000) > rand."""

    def __init__(self, net, alpha, seed=0):
        super().__init__(net, alpha, 1, seed)

    def step(self, s, path_len=0.001, step_size=0.1):
        proposal_net = copy.deepcopy(self.net)
        x0, x1, p, k0, k1 = self._edge(self.net, proposal_net, s, 0, step_size)
        with torch.no_grad():
            x1 = -self.loss(proposal_net(X), y)                               # the script evaluates the proposal once more (cifar_SPhmc.py:119-120)
        _, acc = _ctx.hmc_accept(L.HMC_RULE_SP, [float(x0.detach()), float(x1.detach())], [k0, k1], u=self._uniform(s), temperature=1000.0)
        self.picks.append(acc)
        if acc:
            self.net = proposal_net
            self.loss_list.append(float(-x1.detach()))
        else:
            self.loss_list.append(float(-x0.detach()))

    def fit(self, data=None, num_steps=1000):
        for s in range(num_steps):
            self.step(s)
        return np.array(self.loss_list), np.array(self.train_acc), np.array(self.test_acc)


class MPHMCOptimizer(_HMCBase):
    """cifar_MPhmc.py:67-153: a leapfrog path of N nodes (the direction flips after a random node), weights relative to node 0."""

    def step(self, s, p_s, nets_loss):
        """p_s: per-node kinetic energies |p_s[j]|^2 / 2 (the script passes the momenta themselves and squares them here), nets_loss: -CE per node"""
        B, I = _ctx.hmc_accept(L.HMC_RULE_MP, [float(v.detach()) if hasattr(v, "detach") else float(v) for v in nets_loss], p_s, u=self._uniform(s))
        return I

    def fit(self, data=None, num_steps=1000, step_size=0.1):
        for s in range(num_steps):
            nets = [None] * (self.N + 1); nl = [None] * (self.N + 1); ke = [0.0] * (self.N + 1)
            nets[0] = copy.deepcopy(self.net).to(device)
            ranint = int(1 + self._uniform(s, stream=2) * self.N)              # int(random.uniform(1, N + 1)), cifar_MPhmc.py:100
            sign, p = 1.0, None
            for i in range(self.N):
                if i >= ranint:
                    sign = -1.0
                nets[i + 1] = copy.deepcopy(nets[i]).to(device)
                nl[i], nl[i + 1], p, k0, k1 = self._edge(nets[i], nets[i + 1], s, 0, step_size, sign, p)
                if i == 0:
                    ke[0] = k0
                ke[i + 1] = k1
            I = self.step(s, ke, nl)
            self.picks.append(I)
            self.net = nets[I]
            self._accuracies(50000, 10000)
        return np.array(self.loss_list), np.array(self.train_acc), np.array(self.test_acc)


class PMPHMCOptimizer(_HMCBase):
    """cifar_PMPhmc.py:65-172: binary prefetch tree whose edges are leapfrog steps; prod over levels of max(0, 1 - w_old/w_new) | min(1, w_new/w_old)."""
    rule = L.HMC_RULE_TREE_CIFAR
    record_loss = False                                  # the CIFAR script only prints the loss; main.py appends it

    def step(self, s, p_s, nets_loss, tree_deep):
        """p_s = (ke_out, ke_in): kinetic energies of the momentum drawn at the parent / arrived at the child, indexed by the child"""
        B, I = _ctx.hmc_accept(self.rule, [float(v.detach()) if hasattr(v, "detach") else float(v) for v in nets_loss], p_s[0], p_s[1], u=self._uniform(s))
        return I

    def fit(self, data=None, num_steps=1000, step_size=0.1):
        tree_deep = math.log2(self.N + 1)
        trajectory = []
        for s in range(num_steps):
            nets = [None] * (self.N + 1); nl = [None] * (self.N + 1)
            ko = [0.0] * (self.N + 1); ki = [0.0] * (self.N + 1)
            nets[0] = copy.deepcopy(self.net).to(device)
            for i in range(int(tree_deep)):
                j = int(math.pow(2, i))
                for k in range(j):
                    nets[k + j] = copy.deepcopy(nets[k]).to(device)
                    nl[k], nl[k + j], _, ko[k + j], ki[k + j] = self._edge(nets[k], nets[k + j], s, k + j, step_size)
            I = self.step(s, (ko, ki), nl, tree_deep)
            self.picks.append(I)
            self.net = nets[I]
            if self.record_loss:
                self.loss_list.append(-float(nl[I].detach()))
                trajectory.append(_flat(self.net)[:10].cpu().numpy().tolist())
            else:
                self._accuracies(50000, 10000)
        if self.record_loss:
            return np.array(self.loss_list), np.array(self.train_acc), np.array(self.test_acc), trajectory
        return np.array(self.loss_list), np.array(self.train_acc), np.array(self.test_acc)


class bnnPMPHmc(PMPHMCOptimizer):
    """"Bayesian Network Training"/main.py:55-172: the same tree with the normalised pair w_new / (w_new + w_old) per level; records the loss and the
    first ten parameters of every accepted state."""
    rule = L.HMC_RULE_TREE_BNN
    record_loss = True
