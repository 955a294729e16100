"""Worker for the multi-rank tests (launched by torch.distributed.run).  Backend nccl on GPUs, gloo on CPU (--cpu).
GPU mode: runs the sharded device-resident chain and writes rank 0's trace; the parent compares it with a 1-GPU run.
CPU mode: exercises the rendezvous / sharding host logic with a stand-in context (no kernels)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--out", required=True)
    ap.add_argument("--points", dest="n", type=int, default=30000)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--chains", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as td
    from conftest import synthetic_linear
    from pmp_mcmc_b200 import _lib as L, dist as pdist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    x, y = synthetic_linear(a.n, seed=21)
    if a.cpu:
        td.init_process_group("gloo")
        made = {}

        class FakeContext:                      # records what create_context hands to the library
            def __init__(self, device=0, world_size=1, rank=0, nccl_unique_id=None):
                made.update(device=device, world_size=world_size, rank=rank, uid=bytes(nccl_unique_id))
                self.world_size, self.rank = world_size, rank

            @staticmethod
            def nccl_unique_id():
                return bytes(range(128))

            def set_data_linear(self, xs, ys, n_offset=0, n_global=None):
                made.update(lo=n_offset, n_local=len(xs), n_global=n_global, sx=float(np.sum(xs, dtype=np.float64)))

            def peer_exchange_handle(self):          # a stand-in IPC handle that names its rank
                return bytes([self.rank]) * 64

            def peer_exchange_attach(self, handles):
                made.update(handles=bytes(handles))

        real = L.Context
        L.Context = FakeContext
        try:
            ctx = pdist.create_context(device=0)
            lo, hi = pdist.set_data_linear_sharded(ctx, x, y)
        finally:
            L.Context = real
        assert made["uid"] == bytes(range(128)) and made["world_size"] == world and made["rank"] == rank
        # dist.attach_peers: every rank receives all exchange-buffer handles, rank-major (what pmp_peer_exchange_attach expects)
        assert made["handles"] == b"".join(bytes([r]) * 64 for r in range(world))
        assert lo % 64 == 0 and made["n_global"] == a.n and made["n_local"] == hi - lo
        t = torch.tensor([float(hi - lo), made["sx"]], dtype=torch.float64)
        td.all_reduce(t)
        assert int(t[0].item()) == a.n                                   # shards tile the dataset exactly
        assert abs(t[1].item() - float(np.sum(x, dtype=np.float64))) < 1e-6
        if rank == 0:
            np.savez(a.out, ok=1, world=world)
        td.destroy_process_group()
        return
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = pdist.create_context(local)
    res = {}
    for name, tree, b, depth, algo, draw, scale in (("mp", 0, 256, 1, L.ALGO_MP, L.DRAW_CUDA, 1000.0), ("psp", 1, 2, 6, L.ALGO_PSP, L.DRAW_PYTHON, 600.0)):
        ctx.configure(tree, b=b, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=algo, draw=draw, alpha=0.02, scale=scale)
        pdist.set_data_linear_sharded(ctx, x, y)
        ctx.set_state([-0.8, 1.7, 0.7]); ctx.seed(99, 0)
        ctx.trace_config(a.iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW)
        ctx.run(a.iters)
        tr = ctx.read_trace()
        # every rank must hold the identical chain (replicated acceptance on identical reduced sums)
        t = torch.from_numpy(tr["state"].copy()).cuda()
        ref = t.clone(); td.broadcast(ref, src=0)
        assert torch.equal(t, ref)
        for k in ("state", "next", "draws", "logw"):
            res[name + "_" + k] = tr[k]
    # K sharded chains (pmp_run_multi): chain 0 repeats the "mp" run above, the others are independent chains with their own keys — each
    # must equal its solo run.  fused: one cooperative kernel per GPU, sums exchanged through NVLink peer memory inside it (all K chains);
    # streams: NCCL between kernels, one stream + communicator per chain (first 4 chains)
    K = max(2, a.chains)
    ctxs = [ctx] + [pdist.create_context(local) for _ in range(K - 1)]
    for c in ctxs:
        c.configure(0, b=256, depth=1, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=1000.0)
    pdist.set_data_linear_sharded(ctx, x, y)
    for c in ctxs[1:]:
        c.share_data_from(ctx)
    for mode, prefix, group in (("1", "co", ctxs), ("0", "st", ctxs[:min(K, 4)])):
        os.environ["PMP_PEER_XCHG"] = mode
        for k, c in enumerate(group):
            c.set_state([-0.8, 1.7, 0.7]); c.seed(99 + 24 * k, 0)
            c.trace_config(a.iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW)
        L.run_multi(group, a.iters)
        for k, c in enumerate(group):
            tr = c.read_trace()
            t = torch.from_numpy(tr["draws"].copy()).cuda()
            ref = t.clone(); td.broadcast(ref, src=0)
            assert torch.equal(t, ref)                       # replicated acceptance: identical on every rank
            for key in ("state", "next", "draws", "logw"):
                res["%s%d_%s" % (prefix, k, key)] = tr[key]
    os.environ["PMP_PEER_XCHG"] = "1"
    it0 = ctx.iteration()
    L.run_multi(ctxs, a.iters)                              # a second fused launch continues the exchange counters
    assert all(c.iteration() == it0 + a.iters for c in ctxs[:min(K, 4)])
    for c in reversed(ctxs[1:]):
        c.close()
    # FC and GLM sweeps on sharded rows: integer loss sums all-reduced inside the library → the same bits as one GPU
    from oracle import oracle as o
    rng = np.random.default_rng(31)
    nf = 1500
    Xf = rng.standard_normal((nf, 784)).astype(np.float32); yf = rng.integers(0, 10, size=nf).astype(np.int64)
    ctx.configure(L.TREE_BINARY, depth=2, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
    lo, hi = pdist.shard_bounds(nf, world, rank, align=128)
    ctx.set_data_fc(Xf[lo:hi], yf[lo:hi], n_offset=lo, n_global=nf)
    ctx.set_state(o.fc_init_theta(2)); ctx.seed(21, 0); ctx.propose()
    res["fc_lt"] = ctx.loglik()
    ng, dg = 5000, 20
    Xg = rng.standard_normal((ng, dg)).astype(np.float32); yg = (rng.uniform(size=ng) < 0.5).astype(np.float32)
    thg = (0.3 * rng.standard_normal((50, dg))).astype(np.float32)
    ctx.configure(L.TREE_FLAT, b=50, dim=dg, target=L.TARGET_GLM_LOGISTIC, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=100.0)
    lo, hi = pdist.shard_bounds(ng, world, rank)
    ctx.set_data_glm(Xg[lo:hi], yg[lo:hi], n_offset=lo, n_global=ng)
    ctx.write_proposals(thg)
    res["glm_lt"] = ctx.loglik()
    # CNN sweep on sharded rows (same integer loss sums)
    ncn = 1100
    Xc = rng.standard_normal((ncn, 784)).astype(np.float32); yc = rng.integers(0, 10, size=ncn).astype(np.int64)
    ctx.configure(L.TREE_BINARY, depth=2, dim=o.CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
    lo, hi = pdist.shard_bounds(ncn, world, rank, align=128)
    ctx.set_data_cnn(Xc[lo:hi], yc[lo:hi], n_offset=lo, n_global=ncn)
    ctx.set_state(o.cnn_init_theta(2)); ctx.seed(21, 0); ctx.propose()
    res["cnn_lt"] = ctx.loglik()
    if rank == 0:
        np.savez(a.out, world=world, **res)
    ctx.close()
    td.destroy_process_group()


if __name__ == "__main__":
    main()
