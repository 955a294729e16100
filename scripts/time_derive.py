"""Single chain (pmp_run; flat MP trees and the binary prefetch tree of 100000_PMP.cu): the acceptance publishing the accepted state only (PMP_DERIVE_NODES=1, default) against publishing every node (=0).
Prints microseconds per iteration and the SHA-256 of a 400-iteration trace in both modes (must be equal)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pmp_mcmc_b200 as pm
from pmp_mcmc_b200 import _lib as L
from conftest import synthetic_linear
c = pm.Context(0)
for n, P, scale in ((100000, 1024, 1000.0), (100000, -1024, 1000.0), (100000, 2048, 1000.0), (100000, 100, 1000.0), (500, 1024, 10.0), (500, 4, 10.0)):
    x, y = synthetic_linear(n)
    if P > 0: c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=scale)
    else:     # P < 0: the binary prefetch tree of 100000_PMP.cu (table rule as shipped), |P| = 2^depth nodes
        c.configure(L.TREE_BINARY, depth=(-P).bit_length() - 1, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=0.01, scale=scale, flags=L.FLAG_QUIRK_TABLE_CONST)
    c.set_data_linear(x, y)
    for derive in (0, 1):
        os.environ["PMP_DERIVE_NODES"] = str(derive)
        c.trace_config(1000, L.TRACE_STATE | L.TRACE_NEXT)
        c.set_state([1, 1, 1]); c.seed(7, 0)
        c.run(150); c.run(250)                   # two launches: the second starts from the nodes the first left behind
        tr = c.read_trace()
        h = hashlib.sha256(np.ascontiguousarray(tr["state"]).tobytes() + np.ascontiguousarray(tr["next"]).tobytes()).hexdigest()[:16]
        props = c.read_proposals()
        hp = hashlib.sha256(np.ascontiguousarray(props).tobytes()).hexdigest()[:16]
        c.trace_config(0, 0)
        best = 1e9
        for rep in range(3):
            ms, _ = c.run_timed(4000)
            best = min(best, ms / 4000 * 1e3)
        print("n=%d P=%d derive=%d: %.2f us/iter  trace %s  nodes %s  state %s" % (n, P, derive, best, h, hp, c.get_state()), flush=True)
