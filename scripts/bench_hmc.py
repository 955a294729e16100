"""Secondary benchmark: the HMC leapfrog kernels (csrc/hmc.cu) on a flat parameter vector of D floats.  Bound: HBM — begin reads theta_parent, grad (and the
injected momentum) and writes theta_child, p: 16 (20) bytes per element; end reads p, grad and writes p: 12 bytes per element.  With the momentum drawn on the
device the binary64 quantile behind each normal bounds `begin` instead."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pmp_mcmc_b200 as pm
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = {}
try: peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except OSError: pass
hbm = peaks.get("hbm_gbs", 6541.5)
D = int(os.environ.get("D", 1 << 26))
c = pm.Context(0)
dev = "cuda:0"
theta = torch.randn(D, device=dev); grad = torch.randn(D, device=dev); child = torch.empty(D, device=dev); p = torch.randn(D, device=dev) * 0.0005
torch.cuda.synchronize()
out = []
def timed(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)     # the calls synchronise the stream themselves (they return the kinetic energy)
    return min(ts)
c.seed(1, 0)
t = timed(lambda: c.hmc_leapfrog_begin(theta, grad, child, p, 0.1, 1.0, 0.0005, 0, p_init=p))
out.append({"kernel": "hmc_begin_kernel (momentum injected)", "elements": D, "ms": t * 1e3, "gbs": 20.0 * D / t / 1e9, "frac_hbm": 20.0 * D / t / 1e9 / hbm})
t = timed(lambda: c.hmc_leapfrog_end(p, grad, 0.1, 1.0))
out.append({"kernel": "hmc_end_kernel", "elements": D, "ms": t * 1e3, "gbs": 12.0 * D / t / 1e9, "frac_hbm": 12.0 * D / t / 1e9 / hbm})
t = timed(lambda: c.hmc_leapfrog_begin(theta, grad, child, p, 0.1, 1.0, 0.0005, 0))
out.append({"kernel": "hmc_begin_kernel (Philox momentum, binary64 quantile)", "elements": D, "ms": t * 1e3, "gbs": 16.0 * D / t / 1e9, "frac_hbm": 16.0 * D / t / 1e9 / hbm, "normals_per_s": D / t})
print(json.dumps(out, indent=1))
c.close()
