"""First-contact script for the GPU box: parity spot checks + timing of sweep variants.  Not a test; prints a report."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from oracle import oracle as o
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_linear

c = pm.Context(0)
print("device", c.device_info())
print("fp32 peak TFLOP/s: FFMA", c.fp32_peak(False), "FFMA2", c.fp32_peak(True))

n, P = 100000, 1024
x, y = synthetic_linear(n)
c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
c.set_data_linear(x, y)
c.set_state([1, 1, 1]); c.seed(1234, 0)
c.propose()
props = c.read_proposals()
ref = o.propose(o.TREE_FLAT, P, 1, 3, 0.01, [1, 1, 1], 1234, 0)
print("proposals bit-exact:", np.array_equal(props.view(np.uint32), ref.view(np.uint32)), "max diff", np.abs(props - ref).max())
lt = c.loglik()
acc = o.sumsq_fixed_mirror(x, y, props, (n + 63) // 64 + 1)
lt_m = o.loglik_linear_from_fixed(acc, props, n, 1000.0)
lt64 = o.loglik_linear_f64(x, y, props, 1000.0)
print("loglik vs mirror rel", np.max(np.abs(lt - lt_m) / np.abs(lt_m)), " vs f64 rel", np.max(np.abs(lt - lt64) / np.abs(lt64)))
idx, nxt = c.accept()
print("accept draws[:8]", idx[:8], "next", nxt, "state", c.get_state())

def bench(label, iters=2000):
    c.set_state([1, 1, 1]); c.seed(1, 0)
    c.run(200)
    ms, _ = c.run_timed(iters)
    ms2, sw = c.run_timed(min(iters, 1000), sweep=True)
    print("%-40s %8.2f us/iter (graph)  | plain launches %8.2f us/iter, sweep kernel %7.2f us" % (label, ms / iters * 1e3, ms2 / min(iters, 1000) * 1e3, sw / min(iters, 1000) * 1e3), flush=True)

for tp in (16, 32, 64):
    for per_sm in (2, 3, 4):
        os.environ["PMP_SWEEP_TP"] = str(tp); os.environ["PMP_SWEEP_BLOCKS_PER_SM"] = str(per_sm)
        c.trace_config(0, 0)  # drops the cached graph
        c.propose(); sw = c.time_sweep(200) / 200 * 1e3
        bench("P=1024 n=100k TP=%d per_sm=%d [sweep b2b %.2f us]" % (tp, per_sm, sw))
os.environ["PMP_SWEEP_TP"] = "32"; os.environ["PMP_SWEEP_BLOCKS_PER_SM"] = "2"
os.environ["PMP_GRAPH_ITERS"] = "32"
# binary tree D=10 PSP python draw
c.configure(L.TREE_BINARY, depth=10, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_PSP, draw=L.DRAW_PYTHON, alpha=0.01, scale=2000.0)
bench("binary D=10 PSP")
x5, y5 = synthetic_linear(500)
c.set_data_linear(x5, y5)
for P_ in (4, 1024):
    c.configure(L.TREE_FLAT, b=P_, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=10.0)
    bench("n=500 P=%d MP" % P_)
print("launches", c.launch_count())
