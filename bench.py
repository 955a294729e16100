#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path: multi-proposal MCMC on the 3-parameter linear-Gaussian model,
P = 1024 candidate states per iteration, n = 100 000 data points (BASELINE.json metric; the reference's
`100000_MP.cu` time-analysis shape: flat proposals, CUDA draw rule, SCALE 1000, alpha 0.01, theta0 = (1,1,1)).

A "step" is one block of ITERS_PER_STEP iterations of every chain, run device-resident.  Between steps L2 is flushed
(256 MB memset); inside a step the 0.8 MB dataset is re-read from L2 / shared memory by design — that is what a chain does.

Single GPU: the workload is CHAINS (default 8) INDEPENDENT chains of that shape co-scheduled in one cooperative kernel
(pmp_run_multi): one chain alone is a dependency loop that leaves the sweep SMs idle while it is being accepted, and
independent repeats are how the reference's experiments are run.  Each chain's trace is bit-identical to the chain run alone
(tests/test_gpu_multichain.py); `value` counts the proposal evaluations of all chains, `single_chain` in the same JSON line
is one chain alone (pmp_run), so both the throughput and the latency-bound figure are on record.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the dataset is sharded (strong scaling: n stays 100 000, as BASELINE
config 3 names it) and the per-node partial sums are all-reduced with NCCL inside the library.
`--impl reference` times the reference's CPU implementation of the same path (oracle port of lb.py's per-proposal
torch loop; /root/reference cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_NODES = 1024
N_DATA = 100000
ITERS_PER_STEP = 1000
CHAINS = 8
SCALE = 1000.0
ALPHA = 0.01
METRIC = "proposal-evals/sec"
UNIT = "proposal-evals/s"
# README.md:44 of the reference (V100): MP, n=100000, P=1024: 33473.53 us kernel + 1099.258 us host/copy per iteration
BASELINE_EVALS_PER_S = 1024 / ((33473.53 + 1099.258) * 1e-6)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full: the dataset is read from HBM once per launch and
# then lives in shared memory.  Single chain: profiles/r1b_chain_persistent_ncu_full_summary.txt (50-iteration launch);
# co-scheduled chains: profiles/r1c_chain_persistent_multi_ncu_full_summary.txt (8 chains x 1000 iterations: the launch the bench times).
NCU_DRAM_BYTES_PER_LAUNCH = 914944 + 2304
NCU_DRAM_BYTES_PER_LAUNCH_MULTI = 1131008 + 101376


def synthetic(n, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = (-1.0 + 2.0 * x + 0.5 * rng.standard_normal(n)).astype(np.float32)
    return x, y


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag.is_set():
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([c.strip() for c in line.split(",")])
        p.terminate()

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def cpu_port_evals_per_s(x, y, n_evals, threads=None):
    """The reference's CPU path for one sweep: a Python loop of BayesNet.loglik (lb.py:103-108, torch float32)."""
    from oracle import oracle
    import torch
    rng = np.random.default_rng(1)
    nets = (np.array([1, 1, 1], np.float32) + ALPHA * rng.standard_normal((n_evals, 3))).astype(np.float32)
    oracle.loglik_lb_torch(x, y, nets[:8], threads)      # warm-up
    t0 = time.perf_counter()
    oracle.loglik_lb_torch(x, y, nets, threads)
    dt = time.perf_counter() - t0
    return n_evals / dt, torch.get_num_threads(), dt


def reference_cuda_kernel(x, y, P):
    """The reference's own log_likelihood_kernel (100000_MP.cu:10-36, compiled from its source into oracle/_ref by oracle/Makefile)
    on this GPU: one launch evaluates P proposals.  Reported beside the CPU baseline; None when oracle/_ref was not built."""
    import ctypes
    path = os.path.join(ROOT, "oracle", "_ref", "libref_mp_100000.so")
    if not os.path.exists(path):
        return None
    try:
        lib = ctypes.CDLL(path)
        lib.ref_set_data.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        lib.ref_loglik.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
        rng = np.random.default_rng(1)
        nets = np.ascontiguousarray((np.array([1, 1, 1], np.float32) + ALPHA * rng.standard_normal((P, 3))).astype(np.float32))
        out = np.empty(P, np.float32)
        ms = ctypes.c_float()
        if lib.ref_set_data(x.ctypes.data, y.ctypes.data, len(x)) != 0 or lib.ref_loglik(nets.ctypes.data, P, out.ctypes.data, 5, ctypes.byref(ms)) != 0:
            return None
        return {"kernel_us": ms.value * 1e3, "value": P / (ms.value * 1e-3), "unit": UNIT,
                "what": "reference log_likelihood_kernel<<<ceil(P/256),256>>> recompiled for sm_100a, P=%d, n=%d, kernel time only (no host loop, no copies), mean of 5 launches" % (P, len(x))}
    except OSError:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    x, y = synthetic(N_DATA)
    evals = 256          # bounded sample of the 1024-proposal sweep per step
    vals = []
    for s in range(args.warmup + args.steps):
        v, cores, dt = cpu_port_evals_per_s(x, y, evals)
        if s >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals])) * 1e3 * (P_NODES / evals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": value / BASELINE_EVALS_PER_S, "dtype": "f32",
            "data": "synthetic", "iters_per_sec": value / P_NODES,
            "config": {"workload": "simple_net linear-Gaussian MP, P=1024, n=100000 (lb.py BayesNet.loglik per proposal on the host)", "P": P_NODES, "n": N_DATA},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d of the 1024 proposal-evaluations per step at n=100000, torch CPU float32 loop (oracle.loglik_lb_torch)" % evals},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--iters-per-step", type=int, default=ITERS_PER_STEP)
    ap.add_argument("--weak", action="store_true", help="n = 100000 per GPU instead of 100000 in total")
    ap.add_argument("--chains", type=int, default=CHAINS, help="independent chains co-scheduled on one GPU (1: a single chain with pmp_run)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L, dist as pdist
    if world > 1:
        import torch.distributed as td
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pdist.create_context(local)
    info = ctx.device_info()
    iters = args.iters_per_step
    n_global = N_DATA * world if args.weak else N_DATA
    x, y = synthetic(n_global)
    xp, yp = torch.from_numpy(x).pin_memory().numpy(), torch.from_numpy(y).pin_memory().numpy()     # pinned host buffers for the e2e arm

    # N = 1: the chains share one cooperative kernel.  N > 1: every chain's data are sharded over the ranks; the same kernel runs on
    # every GPU and exchanges the per-node integer sums through NVLink peer memory (pmp_peer_exchange_*, attached by
    # dist.create_context); PMP_PEER_XCHG=0 falls back to one stream + NCCL communicator per chain.
    chains = max(1, min(8, args.chains))

    def configure(c):
        c.configure(L.TREE_FLAT, b=P_NODES, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=ALPHA, scale=SCALE)

    configure(ctx)
    lo, hi = pdist.set_data_linear_sharded(ctx, xp, yp)
    ctxs = [ctx]
    for k in range(1, chains):
        c = pdist.create_context(local)
        configure(c)
        c.share_data_from(ctx)
        ctxs.append(c)

    def reset(seed0):
        for k, c in enumerate(ctxs):
            c.set_state([1, 1, 1]); c.seed(seed0 + k, 0)

    def run_step(timed):
        if chains == 1:
            if timed:
                return ctx.run_timed(iters)[0]
            ctx.run(iters)
            return None
        if timed:
            return L.run_multi_timed(ctxs, iters)
        L.run_multi(ctxs, iters)
        return None

    reset(2024)
    fused_multi = world > 1 and chains > 1 and all(c.peers_attached for c in ctxs) and os.environ.get("PMP_PEER_XCHG", "1") != "0"

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()
        for c in ctxs:
            c.sync()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm: inputs already in HBM, CUDA events on the ctx stream, max over ranks -------------------
    for _ in range(args.warmup):
        ctx.l2_flush(); run_step(False)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = sum(c.launch_count() for c in ctxs)
    step_ms = []
    for _ in range(args.steps):
        ctx.l2_flush()
        barrier()
        step_ms.append(max_over_ranks(run_step(True)))
    barrier()
    launches = sum(c.launch_count() for c in ctxs) - launches0
    total_s = sum(step_ms) * 1e-3
    clocks = sampler.summary()
    iters_per_s = args.steps * iters * chains / total_s            # chain iterations per second, all chains
    value = iters_per_s * P_NODES

    # ---- one chain alone (pmp_run): the latency-bound figure -------------------------------------------------------------
    single = None
    if chains > 1:
        ctx.set_state([1, 1, 1]); ctx.seed(2024, 0)
        ctx.l2_flush(); ctx.run(iters)
        ms1 = []
        for _ in range(3):
            ctx.l2_flush(); ctx.sync()
            ms1.append(ctx.run_timed(iters)[0])
        us1 = float(np.mean(ms1)) * 1e3 / iters
        single = {"value": P_NODES * 1e6 / us1, "unit": UNIT, "iters_per_sec": 1e6 / us1, "us_per_iter": us1,
                  "what": "one chain alone, chain_persistent_kernel via pmp_run, 3 steps of %d iterations" % iters}

    # ---- end-to-end arm: host buffers in, trace out, through the C-ABI the Python samplers call ----------------------
    e2e_s = []
    for c in ctxs:
        c.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS)
    outs = [c.trace_buffers(pinned=True) for c in ctxs]            # page-locked host buffers the traces are copied into
    for s in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        pdist.set_data_linear_sharded(ctx, xp, yp)                 # H2D: this rank's shard of x and y (the other chains alias it)
        for c in ctxs[1:]:
            c.share_data_from(ctx)
        reset(7)
        for c in ctxs:
            c.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS)
        run_step(False)
        trs = [c.read_trace(out=o) for c, o in zip(ctxs, outs)]                        # D2H per chain: states [iters,3] f32, accepted index [iters] i32, the P resampled indices [iters,P] i32
        dt = max_over_ranks(time.perf_counter() - t0)
        assert all(tr["n"] == iters for tr in trs)
        if s >= 2:
            e2e_s.append(dt)
    e2e_value = args.steps * iters * chains * P_NODES / sum(e2e_s)
    h2d = int((hi - lo) * 8 + 12 * chains)
    d2h = int(chains * iters * (16 + 4 * P_NODES))
    for c in ctxs:
        c.trace_config(0, 0)

    # ---- roofline of the dominant kernel: algorithmic flops / average launch duration ----------------------------------
    # Single GPU: the whole chain is ONE cooperative launch (chain_persistent_kernel: 147 sweep CTAs + 1 acceptance CTA), so
    # "the kernel's launch duration" is the timed region itself and the fraction below charges the sweep's roofline with the
    # acceptance and the two hand-offs per iteration as well.  Multi-GPU: the stepwise loop, dominant kernel = sweep_linear_kernel.
    flops_per_iter = 6.0 * (hi - lo) * P_NODES                   # 3 FP32 lane-ops (sub, fma, fma) per (node, point), DESIGN.md §4
    peak = max(ctx.fp32_peak(False), ctx.fp32_peak(True))         # measured FFMA/FFMA2 issue-rate microbenchmark (MEASURED_PEAKS.json has no FP32 figure)
    ctx.set_state([1, 1, 1]); ctx.seed(2024, 0); ctx.propose()
    reps = 200
    sweep_ms = ctx.time_sweep(reps) / reps
    sweep_achieved = flops_per_iter / (sweep_ms * 1e-3) / 1e12
    bytes_per_iter = 8.0 * (hi - lo) + 12.0 * P_NODES + 8.0 * P_NODES
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    persistent = world == 1 and os.environ.get("PMP_PERSISTENT", "1") != "0"
    iter_s = total_s / (args.steps * iters * chains)               # seconds per chain iteration
    traffic = None
    if chains > 1 and (world == 1 or fused_multi):
        kernel, kernel_us, launch_flops = "chain_persistent_multi_kernel<MP> (%d chains, one launch per step)" % chains, iter_s * 1e6 * iters * chains, flops_per_iter * iters * chains
        traffic = NCU_DRAM_BYTES_PER_LAUNCH_MULTI if world == 1 else None
    elif persistent:
        kernel, kernel_us, launch_flops = "chain_persistent_kernel<MP> (one launch per step)", iter_s * 1e6 * iters, flops_per_iter * iters
        traffic = NCU_DRAM_BYTES_PER_LAUNCH
    else:
        kernel, kernel_us, launch_flops = "sweep_linear_kernel<4,true>", sweep_ms * 1e3, flops_per_iter
    achieved = launch_flops / (kernel_us * 1e-6) / 1e12
    roofline = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "measured in this run by pmp_fp32_peak (FFMA/FFMA2 microbenchmark); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
                "note": "the sweep is bound by FP32 issue (3 lane-ops per node-point pair), not by HBM (0.8 MB, L2/shared-memory resident) nor by the tensor pipe; see DESIGN.md 4",
                "kernel": kernel, "kernel_us": kernel_us, "flops_per_launch": launch_flops,
                "sweep_kernel_alone": {"kernel": "sweep_linear_kernel<4,true>", "kernel_us": sweep_ms * 1e3, "achieved": sweep_achieved, "frac": sweep_achieved / peak,
                                       "what": "the stepwise loop's sweep kernel launched back to back (CUDA events on the ctx stream)"},
                "hbm": {"achieved_gbs": bytes_per_iter / iter_s / 1e9, "peak_gbs": hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback", "note": "0.8 MB of data per iteration: L2-resident, HBM is not the bound"}}

    line = None
    if rank == 0:
        cpu = None
        if world == 1:
            v, cores, dt = cpu_port_evals_per_s(x, y, 40960)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "40960 proposal-evaluations at n=100000 (40 sweeps of P=1024) with the lb.py per-proposal torch loop, %.1f s" % dt,
                   "reference_cuda_kernel_same_gpu": reference_cuda_kernel(x, y, P_NODES)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": float(np.mean(step_ms)), "higher_is_better": True, "scaling": "weak" if args.weak else "strong",
                "vs_baseline": value / BASELINE_EVALS_PER_S, "dtype": "f32", "data": "synthetic",
                "iters_per_sec": iters_per_s, "us_per_iter": 1e6 / iters_per_s, "single_chain": single,
                "config": {"workload": "simple_net linear-Gaussian multi-proposal MCMC (100000_MP.cu shape): P=1024 nodes, n=%d points, flat proposals alpha=0.01, SCALE=1000, CUDA draw rule; "
                                       "%d independent chain(s) per GPU%s" % (n_global, chains, (" co-scheduled in one cooperative kernel (pmp_run_multi)" if world == 1 else
                                                                   (" co-scheduled in one cooperative kernel per GPU, data rows sharded over the ranks, per-node sums exchanged through NVLink peer memory inside the kernel (pmp_run_multi)" if fused_multi
                                                                    else " on separate streams and NCCL communicators, data sharded over the ranks (pmp_run_multi)")) + "; iters_per_sec and value count all chains" if chains > 1 else ""),
                           "P": P_NODES, "n": n_global, "chains": chains, "iters_per_step": iters, "device": info["name"],
                           "l2": "flushed (256 MB memset) between steps; inside a step the dataset is re-read from L2 by design",
                           "baseline": "reference README.md:44, V100: (33473.53 + 1099.258) us per iteration at P=1024, n=100000"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": "set_data (pinned host x,y) + per chain: set_state, seed, %d iterations, read_trace (states, accepted indices and all P resampled indices per iteration); %d chain(s) per step, wall clock" % (iters, chains)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    for c in reversed(ctxs):
        c.close()
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
