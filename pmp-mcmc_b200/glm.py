"""d-dimensional linear / logistic regression on the GEMM sweep.

The reference's simple-net model is a 3-parameter linear-Gaussian regression with a scalar covariate (lb.py:100-108);
BASELINE.json names the same experiment "Bayesian logistic regression".  This module is the d-dimensional version of both:
the P x n sweep is one tcgen05 GEMM [n, d] x [d, P] with a fused softplus / square epilogue (csrc/fc_sweep.cu, GLM heads),
and the samplers around it are the same device-resident ones (`GMOptimizer` = MP, `preMOptimizer` = binary prefetch tree).

    loglik_batch(X, y, thetas, kind)                 log-likelihood of every row of thetas
    GLMSampler(X, y, kind, theta0, alpha, ...).fit(num_steps) -> [num_steps, dim] states
"""
import math

import numpy as np

from . import _lib as L
from . import dist as _dist

KINDS = {"logistic": L.TARGET_GLM_LOGISTIC, "gauss": L.TARGET_GLM_GAUSS}


def _set_data(ctx, X, y):
    X = np.ascontiguousarray(X, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1)
    lo, hi = _dist.shard_bounds(len(y), ctx.world_size, ctx.rank)      # multiples of 64 rows
    ctx.set_data_glm(X[lo:hi], y[lo:hi], n_offset=lo, n_global=len(y))


def loglik_batch(X, y, thetas, kind="logistic", scale=1.0, ctx=None):
    c = ctx or _dist.default_context()
    thetas = np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float32)
    c.configure(L.TREE_FLAT, b=len(thetas), dim=thetas.shape[1], target=KINDS[kind], algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=scale)
    _set_data(c, X, y)
    c.write_proposals(thetas)
    return c.loglik()


class GLMSampler:
    """algo 'MP' (N proposals about the current state, lb.py:139-164), 'PSP' (binary prefetch tree with N+1 = 2^D nodes, lb.py:206-258)
    or 'MH' (lb.py:55-73).  scale defaults to n / 50 like lb.py:108."""

    def __init__(self, X, y, kind, theta0, alpha, algo="MP", N=7, scale=None, seed=0, ctx=None):
        self.ctx = ctx or _dist.default_context()
        self.kind, self.alpha, self.N, self.seed = kind, alpha, N, seed
        theta0 = np.asarray(theta0, dtype=np.float32)
        n = len(y)
        scale = float(n) / 50.0 if scale is None else scale
        if algo == "MP":
            self.ctx.configure(L.TREE_FLAT, b=N + 1, dim=len(theta0), target=KINDS[kind], algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=alpha, scale=scale)
        elif algo == "PSP":
            depth = int(math.log2(N + 1))
            if 2 ** depth != N + 1:
                raise ValueError("PSP needs N+1 = 2^D nodes")
            self.ctx.configure(L.TREE_BINARY, depth=depth, dim=len(theta0), target=KINDS[kind], algo=L.ALGO_PSP, draw=L.DRAW_PYTHON, alpha=alpha, scale=scale)
        elif algo == "MH":
            self.ctx.configure(L.TREE_FLAT, b=2, dim=len(theta0), target=KINDS[kind], algo=L.ALGO_MH, draw=L.DRAW_SINGLE, alpha=alpha, scale=scale)
        else:
            raise ValueError("algo must be MP, PSP or MH")
        _set_data(self.ctx, X, y)
        self.ctx.set_state(theta0)
        self.ctx.seed(seed, 0)

    def fit(self, num_steps=1000):
        c = self.ctx
        c.trace_config(num_steps, L.TRACE_STATE | L.TRACE_NEXT)
        c.run(num_steps)
        tr = c.read_trace()
        self.accepted = tr["next"]
        return tr["state"]
