"""The reference's CUDA experiment programs (simple_net/**/*.cu `main()`) on the device-resident chain.

Each function keeps the program's parameters (N = proposals per step so P = N+1 nodes, num_steps, brin_in, alpha, start
state, SCALE literal) and writes the same files (sinks.py), but the loop the program runs on the host — mt19937 proposals,
3-4 blocking cudaMemcpy, one kernel launch, exp / std::discrete_distribution per iteration (500_PMP.cu:159-250) — is one
pmp_run here: proposals, sweep, weights, draws and the trace stay in HBM and come back in one copy.

    time_analysis("MP"|"PMP", ...)   500_MP.cu / 500_PMP.cu / 100000_MP.cu / 100000_PMP.cu / ess_per_s_{MP,PMP}.cu
    convergence("MH"|"MP"|"PMP", ...) conv_mh.cu / conv_mp.cu / conv_pmp.cu (general tree, N_step+1 children per node)
    convergence_with_cores(...)       convery_time_{MP,PMP}.cu (stops after set_time seconds)

The binary-tree PMP programs ship with the short, type-punned table upload: their transition term is a
constant.  `as_shipped=True` (default) reproduces that; False uses the intended table.
"""
import math
import time

import numpy as np

from . import _lib as L
from . import dist as _dist
from . import sinks


def _setup(kind, ctx, x, y, N, alpha, scale, theta0, seed, as_shipped, tree_deep=None, N_step=None):
    c = ctx or _dist.default_context()
    P = N + 1
    if kind == "MP":
        c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=alpha, scale=scale)
    elif kind == "PMP" and N_step is None:
        depth = int(math.log2(P))
        if 2 ** depth != P:
            raise ValueError("binary prefetch tree needs N+1 = 2^D nodes, got %d" % P)
        flags = L.FLAG_QUIRK_TABLE_CONST if as_shipped else 0
        c.configure(L.TREE_BINARY, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, flags=flags, alpha=alpha, scale=scale)
    elif kind == "PMP":
        if (N_step + 1) ** tree_deep != P:
            raise ValueError("general tree needs N+1 = (N_step+1)^tree_deep nodes (conv_pmp.cu:85-87), got %d" % P)
        c.configure(L.TREE_BARY, b=N_step + 1, depth=tree_deep, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=alpha, scale=scale)
    elif kind == "MH":
        c.configure(L.TREE_FLAT, b=2, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MH, draw=L.DRAW_SINGLE, alpha=alpha, scale=scale)
    else:
        raise ValueError("kind must be MH, MP or PMP")
    _dist.set_data_linear_sharded(c, np.ascontiguousarray(x, np.float32), np.ascontiguousarray(y, np.float32))
    c.set_state(np.asarray(theta0, np.float32))
    c.seed(seed, 0)
    return c


def time_analysis(kind, x=None, y=None, data_dir=None, N=3, num_steps=5000, brin_in=4000, alpha=0.01, scale=10.0, theta0=(1, 1, 1),
                  out_dir=None, seed=0, ctx=None, as_shipped=True):
    """500_MP.cu:79-264 / 500_PMP.cu:79-264 (n=500: scale 10), 100000_*.cu (scale 1000), ess_per_s_*.cu (scale 2000,
    num_steps 1e6, brin_in 3000).  `num_steps` counts the recorded iterations (the programs run num_steps + brin_in).
    Returns {"it_per_s", "seconds", "samples" [num_steps, P, 3], "weights" [num_steps, P], "state", "files"}."""
    if x is None:
        x, y = sinks.read_data_txt(data_dir)
    c = _setup(kind, ctx, x, y, N, alpha, scale, theta0, seed, as_shipped)
    P = N + 1
    if brin_in:
        c.trace_config(0, 0)
        c.run(brin_in)
    c.trace_config(num_steps, L.TRACE_SAMPLES | L.TRACE_LOGW)
    t0 = time.perf_counter()                                   # the programs start clock() at i == brin_in (500_MP.cu:166-168)
    c.run(num_steps)
    seconds = time.perf_counter() - t0
    tr = c.read_trace()
    weights = sinks.normalised_weights(tr["logw"])
    files = sinks.write_cuda_dump(out_dir, P, kind, tr["samples"], weights, seconds) if out_dir else {}
    c.trace_config(0, 0)
    return {"it_per_s": num_steps / seconds, "seconds": seconds, "samples": tr["samples"], "weights": weights, "state": c.get_state(), "files": files}


def convergence(kind, x=None, y=None, data_dir=None, N=7, num_steps=3000, alpha=0.02, scale=2000.0, theta0=(0, 0, 1), tree_deep=3, N_step=7,
                out_dir=None, seed=0, ctx=None):
    """conv_mh.cu (N ignored), conv_mp.cu (N=7, 3000 steps), conv_pmp.cu (N=511 = 8^3 - 1, general tree, 2000 steps): the current
    state after every iteration plus a per-iteration clock.  The chain runs device-resident, so the clock column is the
    run's total time spread evenly over the iterations (the programs call clock() on the host every iteration)."""
    if x is None:
        x, y = sinks.read_data_txt(data_dir)
    general = kind == "PMP"
    c = _setup(kind, ctx, x, y, 1 if kind == "MH" else N, alpha, scale, theta0, seed, False, tree_deep if general else None, N_step if general else None)
    P = 2 if kind == "MH" else N + 1
    c.trace_config(num_steps, L.TRACE_STATE | (L.TRACE_LOGW if kind != "MH" else 0))
    t0 = time.perf_counter()
    c.run(num_steps)
    seconds = time.perf_counter() - t0
    tr = c.read_trace()
    times = seconds * (np.arange(num_steps) + 1) / num_steps
    weights = sinks.normalised_weights(tr["logw"]) if kind != "MH" else None
    files = sinks.write_conv_trace(out_dir, kind, num_steps, tr["state"], times, weights, P) if out_dir else {}
    c.trace_config(0, 0)
    return {"it_per_s": num_steps / seconds, "seconds": seconds, "states": tr["state"], "times": times, "weights": weights, "files": files}


def convergence_with_cores(kind, x=None, y=None, data_dir=None, N=1023, num_steps=20000, set_time=180.0, alpha=0.01, scale=2000.0,
                           theta0=(1, 2.0, 0.5), block=1000, out_dir=None, seed=0, ctx=None, as_shipped=True):
    """convery_time_{MP,PMP}.cu:85-89,173: run until num_steps iterations or set_time seconds, whichever comes first (checked
    every `block` iterations here, every iteration there); writes the convergence files."""
    if x is None:
        x, y = sinks.read_data_txt(data_dir)
    c = _setup(kind, ctx, x, y, N, alpha, scale, theta0, seed, as_shipped)
    c.trace_config(num_steps, L.TRACE_STATE)
    done, t0, marks = 0, time.perf_counter(), []
    while done < num_steps and time.perf_counter() - t0 < set_time:
        k = min(block, num_steps - done)
        c.run(k)
        done += k
        marks.append((done, time.perf_counter() - t0))
    tr = c.read_trace()
    its, secs = np.array([m[0] for m in marks], float), np.array([m[1] for m in marks], float)
    times = np.interp(np.arange(done) + 1, np.concatenate([[0], its]), np.concatenate([[0], secs]))
    files = sinks.write_conv_trace(out_dir, kind, done, tr["state"][:done], times) if out_dir else {}
    c.trace_config(0, 0)
    return {"iterations": done, "seconds": float(secs[-1]) if len(secs) else 0.0, "states": tr["state"][:done], "times": times, "files": files}
