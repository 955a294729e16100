"""Single chain, P = 1024 flat MP: microseconds per iteration against n — is the sweep phase quantised to rounds of 32 warps x one 64-point chunk per sweep CTA?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
c = pm.Context(0)
for n in (500, 36864, 60000, 73728, 76000, 85000, 95000, 100000, 105000, 110592, 113000, 125000, 147456):
    x, y = synthetic_linear(n)
    c.configure(L.TREE_FLAT, b=1024, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=n / 100.0)
    c.set_data_linear(x, y); c.trace_config(0, 0); c.set_state([1, 1, 1]); c.seed(7, 0)
    c.run(200)
    best = min(c.run_timed(4000)[0] / 4000 * 1e3 for _ in range(3))
    print("n=%d chunks/CTA(18 CTAs per tile)=%.1f: %.2f us/iter" % (n, (n + 63) // 64 / 18.0, best), flush=True)
