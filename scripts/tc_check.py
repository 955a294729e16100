"""Tensor-core sweep vs FMA sweep vs binary64: accuracy and back-to-back launch time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from oracle import oracle as o
from conftest import synthetic_linear
c = pm.Context(0)
for n, P, scale in ((100000, 1024, 1000.0), (500, 1024, 10.0), (500, 4, 10.0), (100000, 4, 1000.0), (777, 130, 10.0), (50001, 2048, 100.0), (1000000, 1024, 1000.0), (64, 1, 1.0), (1, 3, 1.0)):
    x, y = synthetic_linear(n)
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=scale)
    c.set_data_linear(x, y)
    for state in ([1, 1, 1], [-1, 2, 0.5]):
        c.set_state(state); c.seed(1234, 0); c.propose()
        props = c.read_proposals()
        lt64 = o.loglik_linear_f64(x, y, props, scale)
        res = {}
        for fma in (0, 1):
            os.environ["PMP_SWEEP_TC"] = str(1 - fma)
            lt = c.loglik()
            lt2 = c.loglik()
            us = c.time_sweep(200) / 200 * 1e3
            res[fma] = (lt, us)
            print("n=%d P=%d state=%s %s: rel err vs f64 max %.2e  deterministic %s  sweep b2b %.2f us" % (n, P, state, "fma" if fma else "tc ", np.max(np.abs(lt - lt64) / np.abs(lt64)), np.array_equal(lt, lt2), us), flush=True)
        print("    tc vs fma max rel %.2e" % np.max(np.abs(res[0][0] - res[1][0]) / np.abs(res[1][0])))
os.environ["PMP_SWEEP_TC"] = "0"
