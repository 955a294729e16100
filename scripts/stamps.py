import os, sys, ctypes
os.environ["PMP_DEBUG_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
for n, P in ((100000, 1024), (500, 4), (500, 1024)):
    x, y = synthetic_linear(n)
    c = pm.Context(0)
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
    c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(1, 0)
    c.run(320)
    buf = (ctypes.c_uint64 * (64 + 3072))()
    c.L.pmp_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    assert c.L.pmp_debug_stamps(c.h, buf) == 0
    v = np.array(list(buf)[:64], dtype=np.int64)
    sw_clk, sw_ns, ac_clk, ac_ns = v[0:6], v[16:22], v[32:39], v[48:55]
    print("n=%d P=%d" % (n, P))
    print("  sweep CTA0 phases (cycles): start→stage_issued %d →nodes_built %d →data_landed %d →compute_done %d →flushed %d" % tuple(np.diff(sw_clk)))
    print("  sweep CTA0 start→end (globaltimer ns): %d" % (sw_ns[5] - sw_ns[0]))
    print("  accept phases (cycles): start→lt %d →logw %d →max/exp %d →scan %d →draws %d →state/trace %d" % tuple(np.diff(ac_clk)))
    print("  accept post (cycles, stamp 6 -> 7): %d ; whole acceptance cycle pre -> end of post: %d" % (v[39] - v[38], v[39] - v[32]))
    print("  accept start→end ns: %d ; sweep CTA0 end → accept start ns: %d" % (ac_ns[6] - ac_ns[0], ac_ns[0] - sw_ns[5]))
    print("  sweep CTA0 start → accept end ns: %d" % (ac_ns[6] - sw_ns[0]))
    print("  [persistent] accept: wait_begin→(arrive seen)→body start %d cyc ; body end→released %d cyc ; sweep CTA0: wait %d, props %d, compute %d, flush %d, fence+arrive %d cyc" % (
        v[32] - v[40], v[41] - v[38], sw_clk[1] - sw_clk[0], sw_clk[2] - sw_clk[1], sw_clk[3] - sw_clk[2], sw_clk[4] - sw_clk[3], sw_clk[5] - sw_clk[4]))
    print("  [persistent] ns: accept wait_begin %d, body start %d, body end %d, released %d | sweep: iter start %d, go %d, flushed %d, arrived %d" % (
        v[56] - v[56], v[48] - v[56], v[54] - v[56], v[57] - v[56], sw_ns[0] - v[56], sw_ns[1] - v[56], sw_ns[4] - v[56], sw_ns[5] - v[56]))
    c.close()
