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
c.run(320)
buf = (ctypes.c_uint64 * (64 + 3072))()
c.L.pmp_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
assert c.L.pmp_debug_stamps(c.h, buf) == 0
v = np.array(list(buf), dtype=np.int64)
ct = v[64:64 + 3 * 444].reshape(-1, 3)
t0 = ct[:, 0].min()
st, en, sm = ct[:, 0] - t0, ct[:, 1] - t0, ct[:, 2]
print("CTA start ns: min %d p50 %d p90 %d max %d" % (st.min(), np.median(st), np.percentile(st, 90), st.max()))
print("CTA end   ns: min %d p50 %d p90 %d max %d" % (en.min(), np.median(en), np.percentile(en, 90), en.max()))
print("CTA dur   ns: min %d p50 %d p90 %d max %d" % ((en - st).min(), np.median(en - st), np.percentile(en - st, 90), (en - st).max()))
print("accept start - first CTA start ns:", v[48] - t0, " accept end:", v[54] - t0)
cnt = np.bincount(sm.astype(int), minlength=148)
print("CTAs per SM: min %d max %d ; SMs used %d" % (cnt.min(), cnt.max(), (cnt > 0).sum()))
order = np.argsort(st)
for i in list(order[:6]) + list(order[-6:]):
    print("  cta %3d sm %3d start %6d end %6d dur %6d" % (i, sm[i], st[i], en[i], en[i] - st[i]))
print("CTAs starting after 2us:", int((st > 2000).sum()))
