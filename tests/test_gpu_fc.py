"""GPU parity of the FC sweep (tcgen05 bf16x3 GEMM chain with fused epilogues) against the binary64 forward pass and the
reference's float32 torch evaluation of loss(net) (PMP_FC.py:40-44)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 784)).astype(np.float32)        # MNIST is normalised to ~N(0,1) (PMP_FC.py:54)
    y = rng.integers(0, 10, size=n).astype(np.int64)
    return X, y


@pytest.mark.parametrize("mode", ["auto", "x3"])
@pytest.mark.parametrize("n,P,alpha", [(300, 3, 1e-2), (2000, 5, 1e-4), (4133, 9, 1e-4)])
def test_fc_logtarget_parity(ctx, n, P, alpha, mode, monkeypatch):
    """mode auto: the one-product delta formulation (fc_gemm3_kernel) where the steps are small (alpha sqrt(depth) <= 1e-3), the 3-product
    split otherwise; mode x3: the split everywhere.  Same bounds for both."""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    if mode != "auto":
        monkeypatch.setenv("PMP_FC_MODE", mode)
    X, y = _data(n, seed=n)
    theta0 = o.fc_init_theta(1)
    ctx.configure(L.TREE_FLAT, b=P, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP if False else L.ALGO_TABLE, draw=L.DRAW_SINGLE,
                  flags=L.FLAG_NO_KERNEL_TERM, alpha=alpha, scale=10.0)
    ctx.set_data_fc(X, y)
    ctx.set_state(theta0); ctx.seed(5, 0); ctx.propose()
    props = ctx.read_proposals()
    assert np.array_equal(props[0], theta0)
    lt = ctx.loglik()
    truth = np.array([-o.fc_mean_ce_f64(X, y, props[p]) / 10.0 for p in range(P)])
    # bf16x3 contraction: ~16 mantissa bits per product, fp32 accumulation → stated bound 2e-5 relative on the log-target
    np.testing.assert_allclose(lt, truth, rtol=2e-5)
    # what drives the acceptance is the DIFFERENCE between nodes: it must be resolved, not drowned in rounding noise
    d_dev, d_true = lt[1:] - lt[0], truth[1:] - truth[0]
    assert np.max(np.abs(d_dev - d_true)) <= 0.05 * np.max(np.abs(d_true)) + 2e-7
    # the reference's own float32 evaluation agrees with the truth no better than that
    ref32 = np.array([-o.fc_loss_torch32(X, y, props[p]) for p in range(min(P, 3))])
    np.testing.assert_allclose(lt[: len(ref32)], ref32, rtol=2e-5)
    assert np.array_equal(lt, ctx.loglik())          # integer-exact NLL sums: bitwise repeatable


def _golden():
    import os
    from conftest import ROOT
    from oracle import oracle as o
    G = np.load(os.path.join(ROOT, "tests", "golden", "fc_step.npz"))
    n = int(G["n"])
    rng = np.random.default_rng(int(G["data_seed"]))
    X = rng.standard_normal((n, 28, 28)).astype(np.float32)
    y = rng.integers(0, 10, size=n).astype(np.int64)
    return G, X, y, o.fc_init_theta(int(G["theta_seed"]))


@pytest.mark.parametrize("kind", ["PMP", "MP"])
def test_fc_step_against_reference_golden(ctx, kind):
    """Reference PMPOptimizer.step / MPOptimizer.step (PMP_FC.py:105-143, MP_FC.py:102-122) on synthetic MNIST-shaped data."""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    G, X, y, theta0 = _golden()
    if kind == "PMP":
        ctx.configure(L.TREE_BINARY, depth=3, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=float(G["alpha"]), scale=10.0)
    else:
        ctx.configure(L.TREE_FLAT, b=8, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_MP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE | L.FLAG_KERNEL_MEAN, alpha=float(G["alpha"]), scale=10.0)
    ctx.set_data_fc(X.reshape(len(y), -1), y)
    ctx.set_state(theta0); ctx.seed(int(G["prop_seed"]), 0); ctx.propose()
    props = ctx.read_proposals()
    lt = ctx.loglik()
    np.testing.assert_allclose(-lt, G[kind + "_loss"], rtol=2e-5)                               # loss(net) of the reference, float32
    truth = np.array([o.fc_mean_ce_f64(X, y, props[p]) / 10.0 for p in range(8)])
    idx, nxt = ctx.accept(np.array([float(G[kind + "_u"])]))
    A_dev = ctx.read_logweights()
    if kind == "PMP":
        A_true = o.standardize(o.psp_logweights(-truth, props[:, :4].astype(np.float64), 3, use_kernel=False))
    else:
        kt = o.mp_logweights(np.zeros(8), props.astype(np.float64) / np.sqrt(o.FC_DIM)) / 8.0     # sum_k mean_dim logK / P up to a constant
        A_true = o.standardize(-truth + (kt - kt.mean()))
    # standardisation blows differences of ~1e-6 up to O(1): the device (bf16x3) must track the binary64 weights; the reference's
    # own float32 B is the noisier of the two (corr with the truth is checked in tests/test_cpu_fc.py)
    assert np.max(np.abs(A_dev - A_true)) < 0.35, (A_dev, A_true)
    assert np.corrcoef(A_dev, A_true)[0, 1] > 0.98
    assert idx[0] == o.draw_blocked(o.weights_from_log(A_dev), [float(G[kind + "_u"])], "right")[0] == nxt
    assert np.array_equal(ctx.get_state(), props[nxt])


def test_fc_host_layer(ctx):
    """fc.py: Model / loss / PMPOptimizer / MPOptimizer / MetropolisOptimizer with the reference's call pattern."""
    import torch
    from oracle import oracle as o
    from pmp_mcmc_b200 import fc
    G, X, y, theta0 = _golden()
    fc.set_data(X, y, ctx=ctx)
    net = fc.unflatten(theta0)
    assert isinstance(net, fc.Model) and np.array_equal(fc.flatten(net), theta0)
    np.testing.assert_allclose(float(fc.loss(net)), G["PMP_loss"][0], rtol=2e-5)
    props = o.propose(o.TREE_BINARY, 2, 3, o.FC_DIM, float(G["alpha"]), theta0, int(G["prop_seed"]), 0)
    nets = [fc.unflatten(props[i]) for i in range(8)]
    opt = fc.PMPOptimizer(fc.unflatten(theta0), alpha=1e-4)
    opt.step(1, nets, [torch.from_numpy(p) for p in props], torch.tensor(o.FC_DIM), uniforms=np.array([float(G["PMP_u"])]))
    assert any(opt.net is n for n in nets) and abs(opt.loss - 2.31017) < 1e-3
    for cls in (fc.PMPOptimizer, fc.MPOptimizer, fc.MetropolisOptimizer):
        tr = cls(fc.unflatten(theta0), alpha=1e-4, seed=4).fit(num_steps=3)
        assert tr.shape == (3,) and np.all(np.abs(tr - 2.3102) < 2e-3)


def test_fc_device_resident_run_equals_stepwise(ctx):
    """pmp_run on the FC target (propose → GEMM chain → acceptance queued on the stream, no host round trip) walks the same
    chain as the host-driven propose / loglik / accept sequence."""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    X, y = _data(1500, seed=3)
    theta0 = o.fc_init_theta(2)

    def setup():
        ctx.configure(L.TREE_BINARY, depth=2, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-3, scale=10.0)
        ctx.set_data_fc(X, y); ctx.set_state(theta0); ctx.seed(21, 0)
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


@pytest.mark.parametrize("mode", ["delta", "x3"])
@pytest.mark.parametrize("tag", ["s", "full"])
@pytest.mark.parametrize("kind", ["PMP", "MP"])
def test_fc_trained_model_accepted_index_is_the_references(ctx, kind, tag, mode, monkeypatch):
    """theta0 = the reference's FC_model.pkl, alpha = 1e-4 (PMP_FC.py:15,188-189), n = 384 and the BASELINE size n = 60 000: the device's
    losses against the reference's float32 loss(net), its standardised weights against the reference's B, and — for every injected uniform
    whose distance to a boundary of the reference's cdf exceeds the measured cdf discrepancy — the SAME accepted index as the reference's
    PMPOptimizer.step / MPOptimizer.step (tests/golden/fc_step_trained.npz, generated from the reference's own code)."""
    import os
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    monkeypatch.setenv("PMP_FC_MODE", mode)
    G = np.load(os.path.join(ROOT, "tests", "golden", "fc_step_trained.npz"))
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "fc_theta0.npy"))
    n = int(G[tag + "_n"])
    X = np.random.default_rng(int(G[tag + "_data_seed"])).standard_normal((n, 28, 28)).astype(np.float32)
    y = G[tag + "_labels"].astype(np.int64)
    if kind == "PMP":
        ctx.configure(L.TREE_BINARY, depth=3, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=float(G["alpha"]), scale=10.0)
    else:
        ctx.configure(L.TREE_FLAT, b=8, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_MP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE | L.FLAG_KERNEL_MEAN, alpha=float(G["alpha"]), scale=10.0)
    ctx.set_data_fc(X.reshape(n, -1), y)

    def fresh():
        ctx.set_state(theta0); ctx.seed(int(G["prop_seed"]), 0); ctx.propose()
        return ctx.loglik()
    lt = fresh()
    ref_loss, truth = G["%s_%s_loss" % (tag, kind)], G["%s_%s_truth" % (tag, kind)]
    np.testing.assert_allclose(-lt, ref_loss, rtol=2e-5)
    # node differences (what the standardised weights are made of): ~1e-5 of the loss here; resolved to 5 % of the largest one
    d_dev, d_true = (-lt)[1:] - (-lt)[0], truth[1:] - truth[0]
    assert np.max(np.abs(d_dev - d_true)) <= 0.05 * np.max(np.abs(d_true)), (d_dev, d_true)
    B = G["%s_%s_B" % (tag, kind)]
    cdf_ref = np.cumsum(B / B.sum())
    u_grid, I_ref = G["u_grid"], G["%s_%s_I" % (tag, kind)]
    compared = 0
    cdf_gap = None
    for u, i_ref in zip(u_grid, I_ref):
        fresh()
        idx, nxt = ctx.accept(np.array([float(u)]))
        if cdf_gap is None:
            w = o.weights_from_log(ctx.read_logweights())
            cdf_gap = float(np.max(np.abs(np.cumsum(w / w.sum()) - cdf_ref)))
            assert cdf_gap < 0.03, cdf_gap            # stated bound: bf16-split contraction + float32 reference weights, after standardisation
        if np.min(np.abs(u - cdf_ref[:-1])) > cdf_gap:      # the reference's index is well defined at this uniform
            assert idx[0] == i_ref == nxt, (u, idx, i_ref)
            compared += 1
    assert compared >= 9, compared


def test_fc_delta_and_split_contractions_agree_on_a_deep_tree(ctx, monkeypatch):
    """P = 64 nodes of a depth-6 prefetch tree about the trained model: the one-product delta chain against the 3-product split and
    against binary64 — log-targets, and the node differences that the acceptance is made of."""
    import os
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "fc_theta0.npy"))
    n = 1000
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, 784)).astype(np.float32); y = rng.integers(0, 10, size=n).astype(np.int64)
    out = {}
    for mode in ("delta", "x3"):
        monkeypatch.setenv("PMP_FC_MODE", mode)
        ctx.configure(L.TREE_BINARY, depth=6, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
        ctx.set_data_fc(X, y); ctx.set_state(theta0); ctx.seed(9, 0); ctx.propose()
        out[mode] = ctx.loglik()
        assert np.array_equal(out[mode], ctx.loglik())            # integer loss sums: bitwise repeatable
    props = ctx.read_proposals()
    nodes = [0, 1, 2, 3, 17, 40, 63]
    truth = np.array([-o.fc_mean_ce_f64(X, y, props[p]) / 10.0 for p in nodes])
    for mode in ("delta", "x3"):
        np.testing.assert_allclose(out[mode][nodes], truth, rtol=2e-5)
        d_dev, d_true = out[mode][nodes][1:] - out[mode][0], truth[1:] - truth[0]
        assert np.max(np.abs(d_dev - d_true)) <= 0.05 * np.max(np.abs(d_true)), (mode, d_dev, d_true)
    dd, dx = out["delta"] - out["delta"][0], out["x3"] - out["x3"][0]
    assert np.max(np.abs(dd - dx)) <= 0.05 * np.max(np.abs(dx))


def test_fc_node_batch_size_does_not_change_a_bit(ctx, monkeypatch):
    """nodes per GEMM launch (8 on one GPU, up to 32 on a shard: the fused last layer is then loaded per tile instead of once per launch) is a scheduling
    choice: the integer loss sums of every node are the same bits"""
    import os
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "fc_theta0.npy"))
    X, y = _data(1300, seed=8)
    out = {}
    for nb in ("8", "20", "3"):
        monkeypatch.setenv("PMP_FC_BATCH", nb)
        ctx.configure(L.TREE_BINARY, depth=5, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
        ctx.set_data_fc(X, y); ctx.set_state(theta0); ctx.seed(9, 0); ctx.propose()
        out[nb] = ctx.loglik()
    assert np.array_equal(out["8"], out["20"]) and np.array_equal(out["8"], out["3"])
    truth = np.array([-o.fc_mean_ce_f64(X, y, ctx.read_proposals()[p]) / 10.0 for p in (0, 7, 31)])
    np.testing.assert_allclose(out["20"][[0, 7, 31]], truth, rtol=2e-5)


def test_fc_caller_supplied_small_increments_take_the_delta_chain(ctx):
    """the reference's step(s, proposal_nets, ...) hands over its own proposals (PMP_FC.py:105-143); when they are small increments about node 0 the library measures that
    and runs the same one-product delta chain as for its own proposals (identical bits); unrelated parameter vectors keep the 3-product split"""
    import os
    from conftest import ROOT
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    theta0 = np.load(os.path.join(ROOT, "tests", "golden", "fc_theta0.npy"))
    X, y = _data(900, seed=4)
    ctx.configure(L.TREE_BINARY, depth=3, dim=o.FC_DIM, target=L.TARGET_FC, algo=L.ALGO_PSP, draw=L.DRAW_SINGLE, flags=L.FLAG_STANDARDIZE, alpha=1e-4, scale=10.0)
    ctx.set_data_fc(X, y); ctx.set_state(theta0); ctx.seed(3, 0); ctx.propose()
    props = ctx.read_proposals()
    lt_dev = ctx.loglik()
    ctx.write_proposals(props)
    lt_ext = ctx.loglik()
    assert np.array_equal(lt_dev, lt_ext)
    far = props.copy(); far[5] = o.fc_init_theta(3)                      # one unrelated node: the whole sweep falls back to the split
    ctx.write_proposals(far)
    lt_far = ctx.loglik()
    truth = np.array([-o.fc_mean_ce_f64(X, y, far[p]) / 10.0 for p in (0, 3, 5)])
    np.testing.assert_allclose(lt_far[[0, 3, 5]], truth, rtol=2e-5)
    np.testing.assert_allclose(lt_far[[0, 3]], lt_dev[[0, 3]], rtol=2e-6)
