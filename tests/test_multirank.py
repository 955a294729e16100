"""N > 1: host logic on CPU with gloo (world_size 2 and 3), and — on a box with >= 2 GPUs — the sharded chain with the
in-library NCCL all-reduce, compared bit for bit with the single-GPU chain."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, synthetic_linear


def _torchrun(nproc, args, port, timeout=500):
    """torchrun in its own process group; on a timeout the WHOLE group is killed (a hung rank must not keep spinning on the GPU)."""
    import signal
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_worker.py")] + args
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        err += "\n[killed after %d s]" % timeout
        return subprocess.CompletedProcess(cmd, -9, out, err)
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_and_rendezvous_gloo(tmp_path, world):
    out = tmp_path / "r.npz"
    r = _torchrun(world, ["--cpu", "--out", str(out), "--points", "10007"], 29610 + world)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert np.load(out)["world"] == world


def test_shard_bounds_properties():
    from pmp_mcmc_b200 import dist
    for n in (0, 1, 63, 64, 65, 500, 100000, 100003):
        for world in (1, 2, 3, 4, 8):
            b = [dist.shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:]):
                assert hi == lo2 and lo % 64 == 0
            assert all(hi >= lo for lo, hi in b)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_chain_equals_single_gpu_chain(tmp_path, world):
    """Sharded over `world` GPUs — a single chain, K = 8 chains co-scheduled in the fused peer-exchange kernel, 4 chains on streams +
    NCCL communicators (PMP_PEER_XCHG=0) — every trace must equal the same chain run alone on ONE GPU, bit for bit."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs >= %d GPUs (run under gpurun --gpus %d)" % (world, world))
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    n, iters, K = 30000, 40, 8
    out = tmp_path / "mg.npz"
    r = _torchrun(world, ["--out", str(out), "--points", str(n), "--iters", str(iters), "--chains", str(K)], 29633 + world)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    g = np.load(out)
    assert int(g["world"]) == world
    x, y = synthetic_linear(n, seed=21)
    c = pm.Context(0)
    try:
        for name, tree, b, depth, algo, draw, scale in (("mp", 0, 256, 1, L.ALGO_MP, L.DRAW_CUDA, 1000.0), ("psp", 1, 2, 6, L.ALGO_PSP, L.DRAW_PYTHON, 600.0)):
            c.configure(tree, b=b, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=algo, draw=draw, alpha=0.02, scale=scale)
            c.set_data_linear(x, y)
            c.set_state([-0.8, 1.7, 0.7]); c.seed(99, 0)
            c.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW)
            c.run(iters)
            tr = c.read_trace()
            for k in ("state", "next", "draws", "logw"):      # integer-exact partial sums → identical bits at any GPU count
                assert np.array_equal(tr[k], g[name + "_" + k]), (name, k)
        # K sharded chains in the fused kernel ("co<k>") and 4 on streams / communicators ("st<k>") equal their solo single-GPU runs
        tags = ["co%d" % k for k in range(K)] + ["st%d" % k for k in range(min(K, 4))]
        for tag in tags:
            seed = 99 + 24 * int(tag[2:])
            c.configure(0, b=256, depth=1, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=1000.0)
            c.set_data_linear(x, y)
            c.set_state([-0.8, 1.7, 0.7]); c.seed(seed, 0)
            c.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW)
            c.run(iters)
            tr = c.read_trace()
            for k in ("state", "next", "draws", "logw"):
                assert np.array_equal(tr[k], g[tag + "_" + k]), (tag, k)
        # FC and GLM sweeps: sharded rows + all-reduced integer sums give the single-GPU bits
        from oracle import oracle as o
        rng = np.random.default_rng(31)
        nf = 1500
        Xf = rng.standard_normal((nf, 784)).astype(np.float32); yf = rng.integers(0, 10, size=nf).astype(np.int64)
        c.configure(L.TREE_BINARY, depth=2, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
        c.set_data_fc(Xf, yf); c.set_state(o.fc_init_theta(2)); c.seed(21, 0); c.propose()
        assert np.array_equal(c.loglik(), g["fc_lt"])
        ng, dg = 5000, 20
        Xg = rng.standard_normal((ng, dg)).astype(np.float32); yg = (rng.uniform(size=ng) < 0.5).astype(np.float32)
        thg = (0.3 * rng.standard_normal((50, dg))).astype(np.float32)
        c.configure(L.TREE_FLAT, b=50, dim=dg, target=L.TARGET_GLM_LOGISTIC, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=100.0)
        c.set_data_glm(Xg, yg); c.write_proposals(thg)
        assert np.array_equal(c.loglik(), g["glm_lt"])
        ncn = 1100
        Xc = rng.standard_normal((ncn, 784)).astype(np.float32); yc = rng.integers(0, 10, size=ncn).astype(np.int64)
        c.configure(L.TREE_BINARY, depth=2, dim=o.CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
        c.set_data_cnn(Xc, yc); c.set_state(o.cnn_init_theta(2)); c.seed(21, 0); c.propose()
        assert np.array_equal(c.loglik(), g["cnn_lt"])
    finally:
        c.close()
