"""Secondary benchmark: the FC log-target sweep (BASELINE config 5 shape: 784-512-256-128-10, n=60000 synthetic MNIST-shaped rows).
Reports proposal-evals/s for a batch of nodes and the tensor roofline fraction on ALGORITHMIC flops (2*566528*n per node)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
FC_DIM = 567434

n = int(os.environ.get("N", 60000)); P = int(os.environ.get("P", 64)); reps = int(os.environ.get("REPS", 3))
rng = np.random.default_rng(0)
X = rng.standard_normal((n, 784)).astype(np.float32); y = rng.integers(0, 10, size=n).astype(np.int64)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:                                     # torchrun: data rows sharded over the ranks, integer loss sums all-reduced with NCCL inside the library
    import torch, torch.distributed as td
    from pmp_mcmc_b200 import dist as pdist
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = pdist.create_context(local)
    lo, hi = pdist.shard_bounds(n, world, rank, align=128)
else:
    c = pm.Context(0)
    lo, hi = 0, n
c.configure(L.TREE_BINARY, depth=int(np.log2(P)), dim=FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
c.set_data_fc(X[lo:hi], y[lo:hi], n_offset=lo, n_global=n)
c.set_state(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'fc_theta0.npy'))); c.seed(1, 0)
import ctypes
t0 = time.perf_counter(); c.propose(); c.sync(); t_prop = time.perf_counter() - t0
c.loglik(read=False); c.sync()
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); c.loglik(read=False); c.sync(); ts.append(time.perf_counter() - t0)
dt = min(ts)
alg = 2.0 * 566528 * n * P
peaks = {}
try: peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except OSError: pass
peak = peaks.get("bf16_tflops_sustained", 1400.0)
lt = c.loglik()
idx, nxt = c.accept()
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); td.all_reduce(t, op=td.ReduceOp.MAX); dt = float(t.item())
    peak *= world
if rank == 0:
  print(json.dumps({"n_gpus": world, "workload": "FC 784-512-256-128-10 log-target sweep, n=%d, P=%d (binary tree)" % (n, P), "seconds_per_sweep": dt, "proposal_evals_per_s": P / dt,
                  "propose_seconds": t_prop, "algorithmic_tflops": alg / dt / 1e12, "hardware_tflops_bf16x3": 3 * alg / dt / 1e12 * (2496 * 512 + 1536 * 256 + 768 * 128 + 384 * 16) / (3 * 566528.0),
                  "mode": os.environ.get("PMP_FC_MODE", "auto"), "roofline": {"bound": "tensor", "achieved": alg / dt / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": alg / dt / 1e12 / peak, "note": "algorithmic flops; the bf16x3 split executes ~3.1x as many"},
                  "lt_range": [float(lt.min()), float(lt.max())], "accepted": int(nxt)}))
c.close()
if world > 1:
    td.destroy_process_group()
