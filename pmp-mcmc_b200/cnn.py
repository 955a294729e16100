"""The reference's Bayesian CNN samplers (complex_nets/Mnist/CNN/{MH,MP,PMP}_CNN.py) on the B200 path.

Same names as the scripts: `Model` (PMP_CNN.py:22-44: conv 1->10 5x5, ReLU, maxpool 2, conv 10->20 3x3, ReLU, 2000-500-10, log_softmax),
`loss(net)` (PMP_CNN.py:48-52: CrossEntropy(mean)/10 on the module-level `X`, `y`), `MetropolisOptimizer` (MH_CNN.py:80-141), `MPOptimizer`
(MP_CNN.py:83-169), `PMPOptimizer` (PMP_CNN.py:87-194) — the optimizer classes are the FC scripts' classes verbatim, so they are shared with
fc.py.  All P forward passes of an iteration run on the device (csrc/cnn_sweep.cuh: float32 direct convolutions + one tcgen05 GEMM with the
500 -> 10 layer fused into its epilogue); nothing here falls back to torch for the sweep.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from . import dist as _dist
from . import fc as _fc

X = None          # [n, 1, 28, 28] float32 — module-level like the reference scripts
y = None          # [n] int64
_ctx = None
CNN_DIM = 1007590


def set_data(X_, y_, ctx=None):
    """Replaces the MNIST block PMP_CNN.py:55-84: registers the training set and uploads this rank's shard."""
    global X, y, _ctx
    X = torch.as_tensor(X_, dtype=torch.float32).reshape(-1, 1, 28, 28)
    y = torch.as_tensor(y_, dtype=torch.int64)
    _ctx = ctx or _dist.default_context()
    Xn = X.reshape(len(y), -1).numpy()
    lo, hi = _dist.shard_bounds(len(y), _ctx.world_size, _ctx.rank, align=128)
    _ctx.set_data_cnn(Xn[lo:hi], y.numpy()[lo:hi], n_offset=lo, n_global=len(y))


class Model(torch.nn.Module):
    """PMP_CNN.py:22-44."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(1, 10, 5)
        self.conv2 = nn.Conv2d(10, 20, 3)
        self.fc1 = nn.Linear(20 * 10 * 10, 500)
        self.fc2 = nn.Linear(500, 10)

    def forward(self, x):
        in_size = x.size(0)
        out = F.max_pool2d(F.relu(self.conv1(x)), 2, 2)
        out = F.relu(self.conv2(out)).view(in_size, -1)
        out = self.fc2(F.relu(self.fc1(out)))
        return F.log_softmax(out, dim=1)


flatten = _fc.flatten


def unflatten(theta, like=None):
    return _fc.unflatten(theta, like=like, model=Model)


def mean_ce_batch(thetas):
    """Mean cross-entropy of every row of `thetas` ([P, 1007590]) over the registered data: one device sweep."""
    if _ctx is None:
        raise RuntimeError("cnn.set_data(X, y) first")
    thetas = np.ascontiguousarray(thetas, dtype=np.float32)
    _ctx.configure(L.TREE_FLAT, b=len(thetas), dim=CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE,
                   flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=1.0)
    _ctx.write_proposals(thetas)
    return -_ctx.loglik()


@torch.no_grad()
def loss(net):
    """PMP_CNN.py:48-52 (MP_CNN.py:76-80): CrossEntropyLoss()(net(X), y) / 10 as a 0-d tensor."""
    return torch.tensor(mean_ce_batch(flatten(net)[None, :])[0] / 10.0, dtype=torch.float32)


class _CNNMixin:
    _target, _dim = L.TARGET_CNN, CNN_DIM

    @staticmethod
    def _context():
        return _ctx


class MetropolisOptimizer(_CNNMixin, _fc.MetropolisOptimizer):
    """MH_CNN.py:80-141 (un-divided loss, lamb = 10000)."""


class MPOptimizer(_CNNMixin, _fc.MPOptimizer):
    """MP_CNN.py:83-169."""


class PMPOptimizer(_CNNMixin, _fc.PMPOptimizer):
    """PMP_CNN.py:87-194."""
