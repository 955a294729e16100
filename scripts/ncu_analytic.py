"""Batched banana PMP chains (16 nodes, 2^20 chains, samples recorded) for ncu: one chains_kernel launch of ITERS iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
c = pm.Context(0)
c.configure(tree=L.TREE_BARY, b=4, depth=2, dim=2, target=L.TARGET_BANANA, algo=L.ALGO_PMP, draw=L.DRAW_PYTHON, alpha=1.0, flags=L.FLAG_QUIRK_LEVEL_MOD)
c.seed(0, 0); c.chains_create(1 << 20)
iters = int(os.environ.get("ITERS", 24))
ms = c.chains_run_timed(iters, True)
print("ms", ms, "algorithmic GB", (1 << 20) * iters * 16 * 2 * 4 / 1e9, "GB/s", (1 << 20) * iters * 16 * 2 * 4 / ms / 1e6)
