"""Tuning sweep for the co-scheduled chain kernel on ONE GPU: chains K x acceptance CTAs x warp groups at several dataset sizes
(n = 100000/N emulates the per-GPU shard of an N-GPU run; the NVLink exchange adds ~1-2 us of latency per iteration on top)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
P, iters = 1024, 1000
out = []
for n in [int(v) for v in os.environ.get("NS", "100000,50000,25000,12500").split(",")]:
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, n).astype(np.float32); y = (-1 + 2 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
    for K in [int(k) for k in os.environ.get("KS", "8,16,24,32").split(",")]:
        ctxs = []
        for i in range(K):
            c = pm.Context(0)
            c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
            if i == 0: c.set_data_linear(x, y)
            else: c.share_data_from(ctxs[0])
            c.set_state([1, 1, 1]); c.seed(2024 + i, 0)
            ctxs.append(c)
        for acc in sorted({max(1, K // 4), max(1, K // 2), min(K, 24)}):
            for ng in (4, 8):
                os.environ["PMP_MULTI_ACCEPT"] = str(acc); os.environ["PMP_MULTI_GROUPS"] = str(ng)
                L.run_multi_timed(ctxs, 100)
                ms = min(L.run_multi_timed(ctxs, iters) for _ in range(2))
                r = {"n": n, "K": K, "accept_ctas": acc, "groups": ng, "us_per_chain_iter": round(ms * 1e3 / (iters * K), 3), "evals_per_s": P * iters * K / (ms * 1e-3)}
                out.append(r); print(json.dumps(r), flush=True)
        for c in reversed(ctxs): c.close()
