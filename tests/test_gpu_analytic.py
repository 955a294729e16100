"""GPU parity of the analytic-target path (error.py / com_dim.py / banana): single chain through pmp_run and batched chains
(csrc/chains.cu) against the oracle, which itself replays the reference's runs exactly (tests/test_cpu_analytic.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# name, tree, b, depth, dim, target, algo, draw, alpha, flags, p0, p1, ks
CASES = [
    ("MP normal N=3", 0, 4, 1, 1, 1, 2, 0, 1.0, 0, 0.3, 1.7, 1.0),
    ("PSP normal N=7", 1, 2, 3, 1, 1, 3, 0, 1.0, 0, 0.0, 1.0, 1.0),
    ("PMP normal N=3 D=2", 2, 4, 2, 1, 1, 4, 0, 1.0, 1, 0.0, 1.0, 1.0),
    ("PMP normal N=2 D=3 quirk", 2, 3, 3, 1, 1, 4, 0, 1.0, 1, 0.0, 1.0, 1.0),
    ("PMP banana N=3 D=2", 2, 4, 2, 2, 2, 4, 0, 1.0, 0, 0.0, 1.0, 1.0),
    ("MP banana N=3", 0, 4, 1, 2, 2, 2, 0, 1.0, 0, 0.0, 1.0, 1.0),
    ("SP normal", 0, 2, 1, 1, 1, 1, 2, 0.25, 32, 0.0, 1.0, 1.0),
    ("com_dim d=10 D=3", 1, 2, 3, 10, 3, 3, 0, 0.5, 0, 0.0, 1.0, 0.5),
    ("com_dim d=40 D=5", 1, 2, 5, 40, 3, 3, 0, 0.5, 0, 0.0, 1.0, 0.5),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_batched_chains_against_oracle(ctx, case):
    from oracle import oracle as o
    _, tree, b, depth, dim, target, algo, draw, alpha, flags, p0, p1, ks = case
    ctx.configure(tree, b=b, depth=depth, dim=dim, target=target, algo=algo, draw=draw, flags=flags, alpha=alpha, scale=1.0,
                  kernel_sigma=ks, target_p0=p0, target_p1=p1)
    P, n_chains, iters, seed = ctx.P, 70, 9, 31
    rng = np.random.default_rng(7)
    init = (rng.normal(size=(n_chains, dim)) + (np.array([0.0, -10.0]) if target == 2 else 0.0)).astype(np.float32)
    ctx.seed(seed, 0)
    ctx.chains_create(n_chains, init)
    ctx.chains_run(iters, record_samples=True)
    samples = ctx.chains_read_samples()          # [iters, P, dim, chains]
    states = ctx.chains_read_states()
    for c in (0, 1, 33, 69):
        ref = o.analytic_chain(tree, b, depth, dim, target, algo, draw, alpha, seed, iters, init[c], chain=c, p0=p0, p1=p1, ks=ks,
                               quirk_level_mod=bool(flags & 1), uniform=bool(flags & 32))
        assert np.array_equal(samples[:, :, :, c], ref["samples"]), "chain %d" % c        # bit-exact nodes and identical draws
        assert np.array_equal(states[c], ref["states"][-1])
    # a second launch continues the same chains (iteration counter carries on)
    ctx.chains_run(3, record_samples=False)
    ref = o.analytic_chain(tree, b, depth, dim, target, algo, draw, alpha, seed, 12, init[5], chain=5, p0=p0, p1=p1, ks=ks,
                           quirk_level_mod=bool(flags & 1), uniform=bool(flags & 32))
    assert np.array_equal(ctx.chains_read_states()[5], ref["states"][-1])


@pytest.mark.parametrize("case", CASES[:6] + CASES[7:8], ids=[c[0] for c in CASES[:6] + CASES[7:8]])
def test_single_chain_run_against_oracle(ctx, case):
    """pmp_run on an analytic target (propose + general acceptance kernel) = chain 0 of the oracle (same stream elements)."""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    _, tree, b, depth, dim, target, algo, draw, alpha, flags, p0, p1, ks = case
    ctx.configure(tree, b=b, depth=depth, dim=dim, target=target, algo=algo, draw=draw, flags=flags, alpha=alpha, scale=1.0,
                  kernel_sigma=ks, target_p0=p0, target_p1=p1)
    init = np.linspace(-0.5, 0.7, dim).astype(np.float32) + (np.array([0.0, -10.0], np.float32) if target == 2 else 0)
    ctx.set_state(init); ctx.seed(17, 0)
    ctx.trace_config(8, L.TRACE_SAMPLES | L.TRACE_STATE)
    ctx.run(8)
    tr = ctx.read_trace()
    ref = o.analytic_chain(tree, b, depth, dim, target, algo, draw, alpha, 17, 8, init, chain=0, p0=p0, p1=p1, ks=ks, quirk_level_mod=bool(flags & 1))
    assert np.array_equal(tr["samples"], ref["samples"])
    assert np.array_equal(tr["state"], ref["states"])


def test_reference_named_entry_points(ctx):
    """error.py / com_dim.py signatures; statistical known answers (error.py's own acceptance criterion is the sample mean)."""
    from pmp_mcmc_b200 import analytic as A
    x = A.MP(3000, 0.0, 1.0, 3, seed=1, ctx=ctx)
    assert x.shape == (3000 * 4 - int(0.2 * 3000 * 4),) and abs(x.mean()) < 0.15 and abs(x.std() - 1) < 0.15
    x = A.PMP(800, 0.0, 1.0, 3, 2, seed=2, ctx=ctx)
    assert x.shape == (800 * 16 - int(0.2 * 800 * 16),) and abs(x.mean()) < 0.2
    x = A.PSP(2000, 0.0, 1.0, 7, seed=3, ctx=ctx)
    assert abs(x.mean()) < 0.2 and abs(x.std() - 1) < 0.2
    x = A.SP(20000, 0.0, 1.0, seed=4, ctx=ctx)
    assert x.shape == (16000,) and abs(x.mean()) < 0.35
    xs = A.MP(400, 0.0, 1.0, 3, seed=5, chains=256, ctx=ctx)
    assert xs.shape == (256, 400 * 4 - int(0.2 * 400 * 4)) and abs(xs.mean()) < 0.05
    d = A.PMP_dim(500, np.zeros(10), np.eye(10), 7, 10, seed=6, ctx=ctx)
    assert d.shape == (500 * 8, 10) and abs(d[2000:].mean()) < 0.6          # decays from 2.5 towards 0 (dimension_Chins_Parl.csv trend)
    bn = A.PMP(1500, 0, 1, 3, 2, seed=7, target="banana", ctx=ctx)
    assert bn.shape == (1500 * 16, 2) and abs(bn[4000:, 0].mean()) < 0.5 and -13 < bn[4000:, 1].mean() < -6
