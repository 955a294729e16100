"""ncu target: CHAINS (8) co-scheduled chains x ITERS (50) iterations at the headline shape (one chain_persistent_multi_kernel launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
rng = np.random.default_rng(0)
n, P = 100000, 1024
ITERS = int(os.environ.get("ITERS", 50))
x = rng.uniform(-1, 1, n).astype(np.float32); y = (-1 + 2 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
ctxs = []
for i in range(int(os.environ.get('CHAINS', 8))):
    c = pm.Context(0)
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
    if i == 0: c.set_data_linear(x, y)
    else: c.share_data_from(ctxs[0])
    c.set_state([1, 1, 1]); c.seed(2024 + i, 0)
    ctxs.append(c)
L.run_multi(ctxs, ITERS)
print("ok", [c.iteration() for c in ctxs])
