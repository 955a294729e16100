"""Short device-resident chain at the C3 shape (P=1024, n=100k) for ncu: ITERS iterations in one pmp_run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
n, P = int(os.environ.get("N", 100000)), int(os.environ.get("P", 1024))
x, y = synthetic_linear(n)
c = pm.Context(0)
c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(1, 0)
c.run(int(os.environ.get("ITERS", 50)))
print("state", c.get_state(), "launches", c.launch_count())
