"""One process per GPU: context creation and data sharding.

The reference has no multi-GPU path.  Here the log-likelihood is a sum over independent data points, so
the data rows are split contiguously across ranks, every rank generates the same proposals from the same Philox
counters, and the only exchange per iteration is an NCCL all-reduce (sum, uint64 — exact, order-free) of the P partial
sums inside libpmp_b200.  torch.distributed is used for the rendezvous only (broadcast of the 128-byte NCCL id)."""
import os

import numpy as np

from . import _lib

CHUNK = 64   # csrc/common.cuh: shard boundaries are multiples of CHUNK so the integer sums are identical for any world size

_default_ctx = None


def world():
    """(rank, world_size, local_rank) from torch.distributed if initialised, else from the torchrun environment, else (0,1,0)."""
    try:
        import torch.distributed as td
        if td.is_available() and td.is_initialized():
            return td.get_rank(), td.get_world_size(), int(os.environ.get("LOCAL_RANK", td.get_rank()))
    except ImportError:
        pass
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard_bounds(n, world_size, rank, align=CHUNK):
    """Contiguous shard [lo, hi) of n rows for `rank`: equal numbers of CHUNK-sized blocks (±1), the ragged tail on the last rank."""
    blocks = (n + align - 1) // align
    lo_b = blocks * rank // world_size
    hi_b = blocks * (rank + 1) // world_size
    return min(n, lo_b * align), min(n, hi_b * align)


def create_context(device=None):
    """A Context for this rank.  world_size > 1 needs torch.distributed to be initialised (any backend) for the id broadcast."""
    rank, ws, local = world()
    dev = local if device is None else device
    if ws == 1:
        return _lib.Context(device=dev)
    import torch
    import torch.distributed as td
    if not td.is_initialized():
        raise RuntimeError("world_size > 1: initialise torch.distributed first (it carries the NCCL unique id)")
    uid = np.frombuffer(_lib.Context.nccl_unique_id(), dtype=np.uint8).copy() if rank == 0 else np.zeros(128, dtype=np.uint8)
    t = torch.from_numpy(uid)
    if td.get_backend() == "nccl":
        t = t.cuda(dev)
    td.broadcast(t, src=0)
    ctx = _lib.Context(device=dev, world_size=ws, rank=rank, nccl_unique_id=t.cpu().numpy().tobytes())
    if os.environ.get("PMP_PEER_XCHG", "1") != "0" and ws <= 8 and hasattr(ctx, "peer_exchange_handle"):
        attach_peers(ctx, dev)
    return ctx


def attach_peers(ctx, dev):
    """All-gather the CUDA IPC handles of the ranks' exchange buffers and map the peers' buffers (NVLink peer memory): the
    chain kernel then exchanges the per-node sums itself (pmp_run_multi) instead of calling NCCL between kernels."""
    import torch
    import torch.distributed as td
    mine = torch.from_numpy(np.frombuffer(ctx.peer_exchange_handle(), dtype=np.uint8).copy())
    if td.get_backend() == "nccl":
        mine = mine.cuda(dev)
    allh = [torch.empty_like(mine) for _ in range(ctx.world_size)]
    td.all_gather(allh, mine)
    ctx.peer_exchange_attach(b"".join(h.cpu().numpy().tobytes() for h in allh))


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = create_context()
    return _default_ctx


def set_data_linear_sharded(ctx, x, y):
    """Upload this rank's shard of (x, y); returns (lo, hi)."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1)
    lo, hi = shard_bounds(x.size, ctx.world_size, ctx.rank)
    ctx.set_data_linear(x[lo:hi], y[lo:hi], n_offset=lo, n_global=x.size)
    return lo, hi
