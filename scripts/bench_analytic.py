"""Secondary benchmark: batched independent chains on the analytic targets (BASELINE configs 1 and 4): banana 2-D PMP (N=3, D=2 -> 16
nodes), 1-D normal MP, N(0, I_d) binary-tree PSP.  Bound: the HBM write of the recorded resampled points (4*dim bytes per node
evaluation); reports node evaluations/s and achieved GB/s against the measured HBM peak."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L

peaks = {}
try: peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except OSError: pass
hbm = peaks.get("hbm_gbs", 6650.0)
c = pm.Context(0)
out = []
CASES = [("banana PMP N=3 D=2 (16 nodes, dim 2)", dict(tree=L.TREE_BARY, b=4, depth=2, dim=2, target=L.TARGET_BANANA, algo=L.ALGO_PMP, draw=L.DRAW_PYTHON, alpha=1.0, flags=L.FLAG_QUIRK_LEVEL_MOD), 1 << 20, 24),
         ("normal 1-D MP N=3 (4 nodes, dim 1)", dict(tree=L.TREE_FLAT, b=4, depth=1, dim=1, target=L.TARGET_NORMAL1D, algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=1.0, target_p0=0.0, target_p1=1.0), 1 << 22, 48),
         ("N(0,I_40) PSP D=3 (8 nodes, dim 40)", dict(tree=L.TREE_BINARY, b=2, depth=3, dim=40, target=L.TARGET_STDNORMAL, algo=L.ALGO_PSP, draw=L.DRAW_PYTHON, alpha=0.5, kernel_sigma=0.5), 1 << 17, 24)]
for name, cfg, chains, iters in CASES:
    c.configure(**cfg); c.seed(0, 0)
    P, dim = c.P, cfg["dim"]
    c.chains_create(chains)
    for rec in (True, False):
        c.chains_run_timed(2, rec)
        ms = min(c.chains_run_timed(iters, rec) for _ in range(3))
        evals = float(chains) * iters * P
        wbytes = evals * dim * 4 if rec else 0.0
        out.append({"workload": name, "chains": chains, "iters": iters, "record_samples": rec, "ms": ms, "node_evals_per_s": evals / (ms * 1e-3),
                    "chain_iters_per_s": chains * iters / (ms * 1e-3),
                    "roofline": {"bound": "hbm", "achieved": wbytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": wbytes / (ms * 1e-3) / 1e9 / hbm,
                                 "what": "algorithmic bytes = 4*dim per node evaluation (the recorded resampled points); states stay in registers / L2"} if rec else None})
print(json.dumps(out, indent=1))
