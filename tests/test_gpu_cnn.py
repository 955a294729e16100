"""GPU parity of the CNN sweep (csrc/cnn_sweep.cuh: float32 direct convolutions + tcgen05 fc1 with the fused 500 -> 10 head) against the binary64
forward pass, the reference's float32 torch evaluation of loss(net) (PMP_CNN.py:48-52) and the accepted indices of the reference's step()."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 784)).astype(np.float32), rng.integers(0, 10, size=n).astype(np.int64)


def _cfg(ctx, P, alpha, L, o, tree=None, depth=1, algo=None, flags=None):
    ctx.configure(L.TREE_FLAT if tree is None else tree, b=P, depth=depth, dim=o.CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_TABLE if algo is None else algo,
                  draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM if flags is None else flags, alpha=alpha, scale=10.0)


@pytest.mark.parametrize("n,P,alpha", [(1, 2, 1e-2), (7, 3, 1e-2), (300, 3, 1e-2), (1001, 5, 1e-4), (2500, 9, 1e-4)])
def test_cnn_logtarget_parity(ctx, n, P, alpha):
    """ragged row counts (n not a multiple of the 8-image groups, of 128 or of 256), P not a multiple of the node batch"""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    X, y = _data(n, seed=n)
    theta0 = o.cnn_init_theta(1)
    _cfg(ctx, P, alpha, L, o)
    ctx.set_data_cnn(X, y)
    ctx.set_state(theta0); ctx.seed(5, 0); ctx.propose()
    props = ctx.read_proposals()
    assert np.array_equal(props[0], theta0)
    lt = ctx.loglik()
    truth = np.array([-o.cnn_mean_ce_f64(X, y, props[p]) / 10.0 for p in range(P)])
    # float32 convolutions + bf16x3 contraction (~16 mantissa bits per product, fp32 accumulation): stated bound 2e-5 relative on the log-target
    np.testing.assert_allclose(lt, truth, rtol=2e-5)
    if n >= 300:
        d_dev, d_true = lt[1:] - lt[0], truth[1:] - truth[0]
        assert np.max(np.abs(d_dev - d_true)) <= 0.05 * np.max(np.abs(d_true)) + 2e-7
    ref32 = np.array([-o.cnn_loss_torch32(X, y, props[p]) for p in range(min(P, 3))])
    np.testing.assert_allclose(lt[: len(ref32)], ref32, rtol=2e-5)
    assert np.array_equal(lt, ctx.loglik())          # integer-exact NLL sums: bitwise repeatable


def test_cnn_shards_add_up_bit_exactly(ctx):
    """the per-node loss is an integer sum over rows: two shards evaluated separately add up to the full sweep's integer (what the all-reduce relies on)"""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    n, P = 640, 4
    X, y = _data(n, seed=3)
    theta0 = o.cnn_init_theta(2)
    _cfg(ctx, P, 1e-3, L, o)
    ctx.set_state(theta0); ctx.seed(2, 0)
    ctx.set_data_cnn(X, y); ctx.propose()
    props = ctx.read_proposals()
    full = ctx.loglik()
    parts = []
    for lo, hi in ((0, 384), (384, 640)):
        ctx.set_data_cnn(X[lo:hi], y[lo:hi], n_offset=lo, n_global=n)
        ctx.write_proposals(props)
        parts.append(ctx.loglik())
    # every shard reports -(its integer sum)/n_global/scale: exact in binary64 up to the final division
    np.testing.assert_allclose(parts[0] + parts[1], full, rtol=1e-15)


@pytest.mark.parametrize("tag", ["s", "full"])
@pytest.mark.parametrize("kind", ["PMP", "MP"])
def test_cnn_trained_model_accepted_index_is_the_references(ctx, kind, tag):
    """theta0 = the reference's CNN_model.pkl, alpha = 1e-4 (PMP_CNN.py:15,196-198), n = 256 and n = 60 000: the device's losses against the reference's
    float32 loss(net), its standardised weights against the reference's B, and — for every injected uniform farther from a boundary of the reference's
    cdf than the measured cdf discrepancy — the SAME accepted index as the reference's PMPOptimizer.step / MPOptimizer.step (tests/golden/cnn_step.npz)."""
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    G = np.load(os.path.join(ROOT, "tests", "golden", "cnn_step.npz"))
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "cnn_theta0.npy"))
    n = int(G[tag + "_n"])
    X = np.random.default_rng(int(G[tag + "_data_seed"])).standard_normal((n, 1, 28, 28)).astype(np.float32)
    y = G[tag + "_labels"].astype(np.int64)
    if kind == "PMP":
        _cfg(ctx, 2, float(G["alpha"]), L, o, tree=L.TREE_BINARY, depth=3, algo=L.ALGO_PSP, flags=L.FLAG_STANDARDIZE)
    else:
        _cfg(ctx, 8, float(G["alpha"]), L, o, algo=L.ALGO_MP, flags=L.FLAG_STANDARDIZE | L.FLAG_KERNEL_MEAN)
    ctx.set_data_cnn(X.reshape(n, -1), y)

    def fresh():
        ctx.set_state(theta0); ctx.seed(int(G["prop_seed"]), 0); ctx.propose()
        return ctx.loglik()
    lt = fresh()
    ref_loss, truth = G["%s_%s_loss" % (tag, kind)], G["%s_%s_truth" % (tag, kind)]
    np.testing.assert_allclose(-lt, ref_loss, rtol=2e-5)
    np.testing.assert_allclose(-lt, truth, rtol=2e-5)
    d_dev, d_true = (-lt)[1:] - (-lt)[0], truth[1:] - truth[0]
    assert np.max(np.abs(d_dev - d_true)) <= 0.05 * np.max(np.abs(d_true)), (d_dev, d_true)
    B = G["%s_%s_B" % (tag, kind)]
    cdf_ref = np.cumsum(B / B.sum())
    compared, cdf_gap = 0, None
    for u, i_ref in zip(G["u_grid"], G["%s_%s_I" % (tag, kind)]):
        fresh()
        idx, nxt = ctx.accept(np.array([float(u)]))
        if cdf_gap is None:
            w = o.weights_from_log(ctx.read_logweights())
            cdf_gap = float(np.max(np.abs(np.cumsum(w / w.sum()) - cdf_ref)))
            assert cdf_gap < 0.03, cdf_gap
        if np.min(np.abs(u - cdf_ref[:-1])) > cdf_gap:
            assert idx[0] == i_ref == nxt, (u, idx, i_ref)
            compared += 1
    assert compared >= 9, compared


def test_cnn_host_layer(ctx):
    """cnn.py: Model / loss / PMPOptimizer / MPOptimizer / MetropolisOptimizer with the reference's call pattern"""
    import torch
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import cnn
    G = np.load(os.path.join(ROOT, "tests", "golden", "cnn_step.npz"))
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "cnn_theta0.npy"))
    n = int(G["s_n"])
    X = np.random.default_rng(int(G["s_data_seed"])).standard_normal((n, 1, 28, 28)).astype(np.float32)
    y = G["s_labels"].astype(np.int64)
    cnn.set_data(X, y, ctx=ctx)
    net = cnn.unflatten(theta0)
    assert isinstance(net, cnn.Model) and np.array_equal(cnn.flatten(net), theta0)
    np.testing.assert_allclose(float(cnn.loss(net)), G["s_PMP_loss"][0], rtol=2e-5)
    with torch.no_grad():                                                      # the mirror Model is the reference's network
        ref = float(torch.nn.CrossEntropyLoss()(net(torch.from_numpy(X)), torch.from_numpy(y)) / 10)
    np.testing.assert_allclose(ref, G["s_PMP_loss"][0], rtol=1e-6)
    props = o.propose(o.TREE_BINARY, 2, 3, o.CNN_DIM, float(G["alpha"]), theta0, int(G["prop_seed"]), 0)
    nets = [cnn.unflatten(props[i]) for i in range(8)]
    opt = cnn.PMPOptimizer(cnn.unflatten(theta0), alpha=1e-4)
    opt.step(1, nets, [torch.from_numpy(p) for p in props], torch.tensor(o.CNN_DIM), uniforms=np.array([0.51]))
    assert opt.net is nets[int(G["s_PMP_I"][6])] and abs(opt.loss - 2.1876) < 2e-3
    for cls in (cnn.PMPOptimizer, cnn.MPOptimizer, cnn.MetropolisOptimizer):
        tr = cls(cnn.unflatten(theta0), alpha=1e-4, seed=4).fit(num_steps=3)
        assert tr.shape == (3,) and np.all(np.abs(tr - 2.1876) < 3e-3)


def test_cnn_device_resident_run_equals_stepwise(ctx):
    """pmp_run on the CNN target walks the same chain as the host-driven propose / loglik / accept sequence"""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    X, y = _data(700, seed=3)
    theta0 = o.cnn_init_theta(2)

    def setup():
        _cfg(ctx, 2, 1e-3, L, o, tree=L.TREE_BINARY, depth=2, algo=L.ALGO_PSP, flags=L.FLAG_STANDARDIZE)
        ctx.set_data_cnn(X, y); ctx.set_state(theta0); ctx.seed(21, 0)
    setup()
    nxts = []
    for _ in range(4):
        ctx.propose(); ctx.loglik(read=False)
        nxts.append(int(ctx.accept()[1]))
    ref = ctx.get_state()
    setup()
    ctx.trace_config(4, L.TRACE_NEXT)
    ctx.run(4)
    tr = ctx.read_trace()
    assert list(tr["next"]) == nxts and np.array_equal(ctx.get_state(), ref) and ctx.iteration() == 4
    ctx.trace_config(0, 0)


def test_cnn_mh_sampler_against_reference_loss(ctx):
    """MH_CNN.py:74-78,100-116: the un-divided loss of the current state and of one proposal, and the accept rule u < exp(10000 (loss - loss'))"""
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    G = np.load(os.path.join(ROOT, "tests", "golden", "cnn_step.npz"))
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "cnn_theta0.npy"))
    n = int(G["s_n"])
    X = np.random.default_rng(int(G["s_data_seed"])).standard_normal((n, 784)).astype(np.float32)
    y = G["s_labels"].astype(np.int64)
    props = o.propose(o.TREE_BINARY, 2, 3, o.CNN_DIM, float(G["alpha"]), theta0, int(G["prop_seed"]), 0)[:2]
    ctx.configure(L.TREE_FLAT, b=2, dim=o.CNN_DIM, target=L.TARGET_CNN, algo=L.ALGO_MH, draw=L.DRAW_SINGLE, alpha=float(G["alpha"]), scale=1.0, mh_temperature=10000.0)
    ctx.set_data_cnn(X, y); ctx.set_state(props[0]); ctx.seed(1, 0)
    ctx.write_proposals(props)
    lt = ctx.loglik()
    np.testing.assert_allclose(-lt, G["s_MH_loss"], rtol=2e-5)
    ratio = np.exp(10000.0 * (G["s_MH_loss"][0] - G["s_MH_loss"][1]))
    for u in (0.5 * min(ratio, 1.0), min(0.999999, 1.5 * ratio + 1e-3)):
        ctx.set_state(props[0]); ctx.write_proposals(props); ctx.loglik(read=False)
        idx, nxt = ctx.accept(np.array([u]))
        assert nxt == int(u < ratio), (u, ratio, nxt)
