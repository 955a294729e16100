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


@pytest.mark.parametrize("n,P,alpha", [(300, 3, 1e-2), (2000, 5, 1e-4), (4133, 9, 1e-4)])
def test_fc_logtarget_parity(ctx, n, P, alpha):
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
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
