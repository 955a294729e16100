"""Times pmp_run at the C2/C3 shapes with the persistent cooperative kernel on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
c = pm.Context(0)
for n, P, scale in ((100000, 1024, 1000.0), (500, 1024, 10.0), (500, 4, 10.0), (100000, 4, 1000.0)):
    x, y = synthetic_linear(n)
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=scale)
    c.set_data_linear(x, y)
    for persistent, ptc in ((1, 1), (1, 0), (0, 0)):
        os.environ["PMP_PERSISTENT"] = str(persistent); os.environ["PMP_PERSISTENT_TC"] = str(ptc)
        c.trace_config(0, 0)
        c.set_state([1, 1, 1]); c.seed(1, 0)
        c.run(200)
        best = 1e9
        for rep in range(3):
            ms, _ = c.run_timed(4000)
            best = min(best, ms / 4000 * 1e3)
        print("n=%d P=%d persistent=%d tc=%d: %.2f us/iter  state %s" % (n, P, persistent, ptc, best, c.get_state()), flush=True)
