"""Phase stamps of the tensor-core sweep kernel (CTA 0) and the per-CTA start/end spread."""
import os, sys
os.environ.setdefault("PMP_SWEEP_TC", "1"), ctypes
os.environ["PMP_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
for n, P in ((100000, 1024), (100000, 4), (500, 1024)):
    x, y = synthetic_linear(n)
    c = pm.Context(0)
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
    c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(1, 0); c.propose()
    for rep in range(3):
        c.time_sweep(5)
        buf = (ctypes.c_uint64 * (64 + 3072))()
        c.L.pmp_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        assert c.L.pmp_debug_stamps(c.h, buf) == 0
        v = np.array(list(buf), dtype=np.int64)
        clk, ns = v[0:6], v[16:22]
        ctas = v[64:64 + 3 * 148].reshape(148, 3)
        ctas = ctas[ctas[:, 0] > 0]
        t0 = ctas[:, 0].min()
        print("n=%d P=%d: CTA0 cycles start→staged %d →operand built %d →pipeline done %d →flushed %d →end %d | ns total %d" % ((n, P) + tuple(np.diff(clk)) + (ns[5] - ns[0],)))
        print("   %d CTAs: start spread %d ns, end min/median/max %d/%d/%d ns after first start; durations min/med/max %d/%d/%d ns" % (
            len(ctas), ctas[:, 0].max() - t0, ctas[:, 1].min() - t0, np.median(ctas[:, 1]) - t0, ctas[:, 1].max() - t0,
            (ctas[:, 1] - ctas[:, 0]).min(), np.median(ctas[:, 1] - ctas[:, 0]), (ctas[:, 1] - ctas[:, 0]).max()))
    iss = v[832:848]; ep = v[896:896 + 128].reshape(16, 8)[:, :7]; base = clk[2]
    print("   issuer commit times (cycles after operand built):", [int(t - base) for t in iss if t > 0])
    print("   warp0 per unit [top, ld done, arrived, fma done, next full, next ld issued, tail done]:", [[int(t - base) for t in r] for r in ep if r[0] > 0])
    c.close()
