"""Throughput of K co-scheduled chains at the headline shape (P=1024, n=100000) against one chain alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
rng = np.random.default_rng(0)
n, P, iters = int(os.environ.get("N", 100000)), 1024, 2000
x = rng.uniform(-1, 1, n).astype(np.float32); y = (-1 + 2 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
out = {}
for K in [int(k) for k in os.environ.get("KS", "1,2,4,6,8").split(",")]:
    ctxs = []
    for i in range(K):
        c = pm.Context(0)
        c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
        if i == 0: c.set_data_linear(x, y)
        else: c.share_data_from(ctxs[0])
        c.set_state([1, 1, 1]); c.seed(2024 + i, 0)
        ctxs.append(c)
    L.run_multi_timed(ctxs, 200)
    ms = min(L.run_multi_timed(ctxs, iters) for _ in range(3))
    out[K] = {"us_per_chain_iter": ms * 1e3 / (iters * K), "evals_per_s": P * iters * K / (ms * 1e-3)}
    for c in reversed(ctxs): c.close()
c = pm.Context(0)
c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(2024, 0)
c.run(200); ms = min(c.run_timed(iters)[0] for _ in range(3))
out["solo_pmp_run"] = {"us_per_chain_iter": ms * 1e3 / iters, "evals_per_s": P * iters / (ms * 1e-3)}
print(json.dumps(out, indent=1))
