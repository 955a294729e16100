"""Secondary benchmark: the CNN log-target sweep (PMP_CNN.py shape: conv 1->10 5x5, pool, conv 10->20 3x3, 2000-500-10; n=60000 synthetic MNIST-shaped rows).
Reports proposal-evals/s for a batch of nodes and the algorithmic flop rate (2*1329000*n per node: 324 000 convolution MACs on the CUDA cores in float32,
1 005 000 dense MACs on tcgen05)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
CNN_DIM = 1007590
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(os.environ.get("N", 60000)); P = int(os.environ.get("P", 64)); reps = int(os.environ.get("REPS", 3))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, 784)).astype(np.float32); y = rng.integers(0, 10, size=n).astype(np.int64)
c = pm.Context(0)
c.configure(L.TREE_BINARY, depth=int(np.log2(P)), dim=CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
c.set_data_cnn(X, y)
c.set_state(np.load(os.path.join(ROOT, "tests", "golden", "cnn_theta0.npy"))); c.seed(1, 0)
c.propose(); c.sync()
c.loglik(read=False); c.sync()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); c.loglik(read=False); c.sync(); ts.append(time.perf_counter() - t0)
dt = min(ts)
conv_flop, dense_flop = 2.0 * 324000 * n * P, 2.0 * 1005000 * n * P
lt = c.loglik()
idx, nxt = c.accept()
print(json.dumps({"workload": "CNN (PMP_CNN.py:22-52) log-target sweep, n=%d, P=%d (binary tree)" % (n, P), "seconds_per_sweep": dt, "ms_per_node": 1e3 * dt / P,
                  "proposal_evals_per_s": P / dt, "algorithmic_tflops": (conv_flop + dense_flop) / dt / 1e12, "conv_fp32_tflop_per_sweep": conv_flop / 1e12,
                  "dense_tflop_per_sweep": dense_flop / 1e12, "lt_range": [float(lt.min()), float(lt.max())], "accepted": int(nxt)}))
c.close()
