"""Tiny end-to-end workloads for compute-sanitizer (scripts/sanitize.sh): every kernel family once, sizes that finish under the tool's
10-100x slowdown.  WHICH = chain | multi | stepwise | fc | glm | analytic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
which = sys.argv[1] if len(sys.argv) > 1 else "chain"
rng = np.random.default_rng(0)
n = 3000
x = rng.uniform(-1, 1, n).astype(np.float32); y = (-1 + 2 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
c = pm.Context(0)
if which in ("chain", "stepwise"):
    for tree, b, depth, algo, draw in ((L.TREE_FLAT, 256, 1, L.ALGO_MP, L.DRAW_CUDA), (L.TREE_BINARY, 2, 6, L.ALGO_PSP, L.DRAW_PYTHON)):
        c.configure(tree, b=b, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=algo, draw=draw, alpha=0.02, scale=60.0)
        c.set_data_linear(x, y); c.set_state([0, 0, 1]); c.seed(3, 0)
        c.trace_config(6, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS)
        c.run(6)
        print(which, "next", c.read_trace()["next"])
elif which == "multi":
    K = 3
    ctxs = [c] + [pm.Context(0) for _ in range(K - 1)]
    for i, cc in enumerate(ctxs):
        cc.configure(L.TREE_FLAT, b=256, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=60.0)
        if i == 0: cc.set_data_linear(x, y)
        else: cc.share_data_from(c)
        cc.set_state([0, 0, 1]); cc.seed(3 + i, 0); cc.trace_config(6, L.TRACE_NEXT)
    L.run_multi(ctxs, 6)
    print("multi next", [cc.read_trace()["next"].tolist() for cc in ctxs])
    for cc in reversed(ctxs[1:]): cc.close()
elif which == "fc":
    nf = 300
    X = rng.standard_normal((nf, 784)).astype(np.float32); yl = rng.integers(0, 10, size=nf).astype(np.int64)
    th = (rng.uniform(-1, 1, 567434) * 0.04).astype(np.float32)
    for alpha in (1e-4, 1e-2):
        c.configure(L.TREE_BINARY, depth=2, dim=567434, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=alpha, scale=10.0)
        c.set_data_fc(X, yl); c.set_state(th); c.seed(1, 0); c.propose()
        print("fc lt", alpha, c.loglik())
elif which == "glm":
    ng, d = 1000, 20
    Xg = rng.standard_normal((ng, d)).astype(np.float32); yg = (rng.uniform(size=ng) < 0.5).astype(np.float32)
    c.configure(L.TREE_FLAT, b=50, dim=d, target=L.TARGET_GLM_LOGISTIC, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=100.0)
    c.set_data_glm(Xg, yg); c.write_proposals((0.3 * rng.standard_normal((50, d))).astype(np.float32))
    print("glm lt", c.loglik()[:4])
elif which == "analytic":
    c.configure(L.TREE_BARY, b=4, depth=2, dim=2, target=L.TARGET_BANANA, algo=L.ALGO_PMP, draw=L.DRAW_PYTHON, alpha=1.0, flags=L.FLAG_QUIRK_LEVEL_MOD)
    c.seed(0, 0); c.chains_create(4096); c.chains_run(4, True)
    print("analytic", c.chains_read_states()[:2])
c.close()
