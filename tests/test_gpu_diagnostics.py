"""On-device trace diagnostics (pmp_trace_diagnostics) against a numpy restatement, and the ESS estimator on a process with a
known integrated autocorrelation time."""
import numpy as np
import pytest

from conftest import synthetic_linear



@pytest.mark.gpu
def test_trace_diagnostics_match_numpy(ctx):
    from pmp_mcmc_b200 import _lib as L, diagnostics as dg
    x, y = synthetic_linear(3000, seed=2)
    ctx.configure(L.TREE_FLAT, b=16, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=60.0)
    ctx.set_data_linear(x, y); ctx.set_state([-1, 2, 0.5]); ctx.seed(4, 0)
    iters = 3000
    ctx.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT)
    ctx.run(iters)
    tr = ctx.read_trace()
    ref = dg.reference_numpy(tr["state"], tr["next"], max_lag=64)
    dev = ctx.trace_diagnostics(max_lag=64)
    assert dev["n"] == iters
    np.testing.assert_allclose(dev["mean"], ref["mean"], rtol=1e-12)
    np.testing.assert_allclose(dev["var"], ref["var"], rtol=1e-9)
    np.testing.assert_allclose(dev["acov"], ref["acov"], rtol=1e-8, atol=1e-16)
    assert dev["msjd"] == pytest.approx(ref["msjd"], rel=1e-10) and dev["move_rate"] == pytest.approx(ref["move_rate"], rel=1e-12)
    s = dg.summarize(ctx, seconds=0.5, max_lag=64)
    assert s["ess"].shape == (3,) and np.all(s["ess"] > 1) and np.all(s["ess"] <= iters * 1.0001 + 1)
    np.testing.assert_allclose(s["ess_per_s"], s["ess"] / 0.5)
    ctx.trace_config(0, 0)


def test_ess_estimator_on_ar1():
    """AR(1) with coefficient phi has tau = (1 + phi) / (1 - phi): the Geyer estimate must land near n / tau."""
    from pmp_mcmc_b200 import diagnostics as dg
    rng = np.random.default_rng(0)
    n, phi = 200000, 0.8
    e = rng.standard_normal(n)
    x = np.empty(n); x[0] = e[0]
    for t in range(1, n):
        x[t] = phi * x[t - 1] + e[t]
    ref = dg.reference_numpy(x[:, None], max_lag=200)
    ess = dg.ess_from_acov(ref["acov"], n)[0]
    assert ess == pytest.approx(n * (1 - phi) / (1 + phi), rel=0.1)
    assert len(dg.skewness_of_batch_means(x, 2)) == 5
