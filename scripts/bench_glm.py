"""Secondary benchmark: the GLM log-target sweep (logistic / Gaussian head) at the headline shape P=1024, n=100000, d covariates.
One launch of fc_gemm2_kernel<256, EPI2_GLM> evaluates all P nodes over all rows."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L

n = int(os.environ.get("N", 100000)); P = int(os.environ.get("P", 1024)); d = int(os.environ.get("D", 64)); kind = os.environ.get("KIND", "logistic")
reps = int(os.environ.get("REPS", 20))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, d)).astype(np.float32); X[:, 0] = 1
true = rng.standard_normal(d).astype(np.float32) / np.sqrt(d)
y = (rng.uniform(size=n) < 1 / (1 + np.exp(-(X @ true)))).astype(np.float32) if kind == "logistic" else (X @ true + 0.5 * rng.standard_normal(n)).astype(np.float32)
c = pm.Context(0)
dim = d + (1 if kind == "gauss" else 0)
c.configure(L.TREE_FLAT, b=P, dim=dim, target=L.TARGET_GLM_LOGISTIC if kind == "logistic" else L.TARGET_GLM_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_PYTHON,
            flags=L.FLAG_NO_KERNEL_TERM, alpha=0.01, scale=n / 50.0)
c.set_data_glm(X, y)
th0 = np.concatenate([true, [1.0]]).astype(np.float32) if kind == "gauss" else true
c.set_state(th0); c.seed(1, 0); c.propose()
c.loglik(read=False); c.sync()
t0 = time.perf_counter()
for _ in range(reps):
    c.loglik(read=False)
c.sync()
dt = (time.perf_counter() - t0) / reps
iters = 200
c.run(20); t0 = time.perf_counter(); c.run(iters); t_chain = (time.perf_counter() - t0) / iters
pairs = float(n) * P
peaks = {}
try: peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except OSError: pass
print(json.dumps({"workload": "GLM %s head sweep, n=%d, P=%d, d=%d" % (kind, n, P, d), "seconds_per_sweep": dt, "proposal_evals_per_s": P / dt,
                  "chain_iteration_seconds": t_chain, "chain_evals_per_s": P / t_chain,
                  "algorithmic_tflops": 2.0 * d * pairs / dt / 1e12, "pairs_per_s": pairs / dt,
                  "roofline": {"bound": "sfu (XU pipe: ex2 + lg2 per (node, row) pair, float->fixed conversions)" if kind == "logistic" else "epilogue issue (TMEM read, square, shuffle transpose-reduce)",
                               "note": "ncu, logistic, d=64: XU pipe 85 % busy, tensor pipe 9 % (profiles/r1c_glm_logistic_ncu_full_summary.txt): the sweep is bound by the "
                                       "transcendentals of the fused epilogue, not by the contraction"}}))
