"""Reader-warp timeline of the persistent tensor-core chain kernel (build with EXTRA=-DPMP_TC_STAMPS)."""
import os, sys, ctypes
os.environ["PMP_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
n, P = 100000, 1024
x, y = synthetic_linear(n)
c = pm.Context(0)
c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(1, 0)
c.run(8)
buf = (ctypes.c_uint64 * (64 + 3072))()
c.L.pmp_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert c.L.pmp_debug_stamps(c.h, buf) == 0
v = np.array(list(buf), dtype=np.int64)
ep = v[896:896 + 128].reshape(2, 8, 8)[:, :, :6]; iss = v[832:848].reshape(2, 8)
base = ep[ep > 0].min()
for h in range(2):
    print("stage %d reader (q=0,j=0) per quad [top, full seen, ld1 done, ld2 done, elected/issued, tail done]:" % h)
    for i in range(8):
        if ep[h, i, 0] > 0: print("   quad %d: %s   refill issued at %s" % (h + 2 * i, [int(t - base) for t in ep[h, i]], int(iss[h, i] - base) if iss[h, i] > 0 else "-"))
