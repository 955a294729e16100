"""GPU parity of the d-dimensional logistic / linear-Gaussian heads (one tcgen05 GEMM [n, d] x [d, P] with a fused softplus /
square epilogue and a warp-shuffle per-node reduction) against binary64, and the device-resident chain on them."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _data(n, d, kind, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)).astype(np.float32)
    X[:, 0] = 1.0                                       # intercept column
    true = rng.standard_normal(d).astype(np.float32) / np.sqrt(d)
    if kind == "logistic":
        y = (rng.uniform(size=n) < 1.0 / (1.0 + np.exp(-(X @ true)))).astype(np.float32)
    else:
        y = (X @ true + 0.5 * rng.standard_normal(n)).astype(np.float32)
    return X, y, true


@pytest.mark.parametrize("kind", ["logistic", "gauss"])
@pytest.mark.parametrize("n,d,P", [(1000, 5, 7), (20000, 64, 300), (4133, 130, 1024)])
def test_glm_logtarget_parity(ctx, kind, n, d, P):
    from oracle import oracle as o
    from pmp_mcmc_b200 import glm
    X, y, true = _data(n, d, kind, seed=n + d)
    rng = np.random.default_rng(1)
    th = (true + 0.3 * rng.standard_normal((P, d)) / np.sqrt(d)).astype(np.float32)
    if kind == "gauss":
        th = np.concatenate([th, rng.uniform(0.3, 1.5, (P, 1)).astype(np.float32)], axis=1)
    scale = n / 50.0
    lt = glm.loglik_batch(X, y, th, kind, scale, ctx=ctx)
    truth = o.loglik_glm_f64(X, y, th, kind, scale)
    # bf16x3 contraction (~16 mantissa bits per product), MUFU exp/log in the epilogue: stated bound 2e-5 relative
    np.testing.assert_allclose(lt, truth, rtol=2e-5)
    assert np.array_equal(lt, ctx.loglik())                      # integer sums: bitwise repeatable
    # differences between nodes are what the acceptance sees
    assert np.max(np.abs((lt - lt[0]) - (truth - truth[0]))) <= 2e-4 * np.max(np.abs(truth - truth[0])) + 1e-6


def test_glm_shards_add_up_bit_exactly(ctx):
    """The per-node sums are integers over 32-row partials: two shards (cut at a multiple of 64 rows) evaluated separately give
    log-targets whose integer parts add up to the unsharded run's — checked through the Gaussian head's closed form."""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    n, d, P = 6000, 17, 40
    X, y, true = _data(n, d, "logistic", seed=5)
    th = (true + 0.2 * np.random.default_rng(2).standard_normal((P, d))).astype(np.float32)

    def run(lo, hi):
        ctx.configure(L.TREE_FLAT, b=P, dim=d, target=L.TARGET_GLM_LOGISTIC, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=0.0, scale=1.0)
        ctx.set_data_glm(X[lo:hi], y[lo:hi], n_offset=lo, n_global=n)
        ctx.write_proposals(th)
        return ctx.loglik()
    whole, a, b = run(0, n), run(0, 2560), run(2560, n)
    fx = float(1 << 24)
    assert np.array_equal(np.rint(whole * fx), np.rint(a * fx) + np.rint(b * fx))


@pytest.mark.parametrize("kind,algo", [("logistic", "MP"), ("gauss", "PSP"), ("logistic", "MH")])
def test_glm_device_resident_chain(ctx, kind, algo):
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L, glm
    n, d = 3000, 6
    X, y, true = _data(n, d, kind, seed=9)
    theta0 = np.zeros(d + (1 if kind == "gauss" else 0), np.float32)
    if kind == "gauss":
        theta0[-1] = 1.0
    s = glm.GLMSampler(X, y, kind, theta0, alpha=0.05, algo=algo, N=7, seed=3, ctx=ctx)
    states = s.fit(300)
    assert states.shape == (300, len(theta0)) and np.all(np.isfinite(states))
    ll0 = o.loglik_glm_f64(X, y, theta0, kind)[0]
    ll1 = o.loglik_glm_f64(X, y, states[-1], kind)[0]
    assert ll1 > ll0                                           # the chain climbs from the start state
    # the device loop equals the host-driven propose / loglik / accept sequence
    ctx.set_state(theta0); ctx.seed(3, 0)
    nx = []
    for _ in range(12):
        ctx.propose(); ctx.loglik(read=False)
        nx.append(int(ctx.accept()[1]))
    assert nx == list(s.accepted[:12]) and np.array_equal(ctx.get_state(), states[11])
