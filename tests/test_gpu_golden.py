"""The CUDA path against the golden vectors produced by the REFERENCE's own Python code (tests/golden/lb_step.npz):
reference dataset (n=500), reference BayesNet.loglik values, reference step() draws with the recorded uniforms."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
G = os.path.join(ROOT, "tests", "golden")
CASES = {"mp": (0, 4, 1, "MP"), "mp8": (0, 8, 1, "MP"), "psp": (1, 2, 3, "PSP"), "pmp": (2, 4, 2, "PMP")}


@pytest.mark.parametrize("name", list(CASES))
def test_step_against_reference_golden(ctx, name):
    from pmp_mcmc_b200 import _lib as L
    g = np.load(os.path.join(G, "lb_step.npz"))
    tree, b, depth, algo = CASES[name]
    x, y = g["x"], g["y"]
    ctx.configure(tree, b=b, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=getattr(L, "ALGO_" + algo), draw=L.DRAW_PYTHON, alpha=0.05, scale=len(x) / 50.0)
    ctx.set_data_linear(x, y)
    P = ctx.P
    for k in range(len(g[name + "_u"])):
        ctx.set_state(g["state"]); ctx.seed(1, 0); ctx.propose()
        props = ctx.read_proposals()
        assert np.array_equal(props, g[name + "_props"])                                   # same bits as the fixture's proposals
        lt = ctx.loglik()
        np.testing.assert_allclose(lt, g[name + "_loglik"], rtol=1e-5)                     # reference BayesNet.loglik, torch float32
        u = np.concatenate([g[name + "_u"][k], [g[name + "_upick"][k]]])
        idx, nxt = ctx.accept(u)
        assert np.array_equal(idx, g[name + "_draws"][k])                                  # reference step(): resampled indices
        assert nxt == g[name + "_next"][k]                                                 # reference step(): new self.net
        assert np.array_equal(ctx.get_state(), props[nxt])


def test_samplers_module_matches_reference_step():
    """The lb.py-named host layer (pmp_mcmc_b200.samplers) driven the way the reference drives its optimizers."""
    from pmp_mcmc_b200 import samplers as S
    g = np.load(os.path.join(G, "lb_step.npz"))
    data = {"x": g["x"], "y": g["y"]}
    for name, make in (("mp", lambda net: S.GMOptimizer(net, 0.05, N=3)), ("psp", lambda net: S.preMOptimizer(net, 0.05, N=7)),
                       ("pmp", lambda net: S.GMpreOptimizerV2(net, 0.05, N=3, deep=2))):
        props = g[name + "_props"]
        nets = {i: S.BayesNet().set_theta(props[i]) for i in range(len(props))}
        np.testing.assert_allclose([float(nets[i].loglik(data)) for i in range(3)], g[name + "_loglik"][:3], rtol=1e-5)
        for k in range(3):
            opt = make(S.BayesNet().set_theta(g["state"]))
            u = np.concatenate([g[name + "_u"][k], [g[name + "_upick"][k]]])
            new_nets = opt.step(data, nets, uniforms=u)
            got = [int(np.flatnonzero([new_nets[j] is nets[i] for i in range(len(props))])[0]) for j in range(len(props))]
            assert got == list(g[name + "_draws"][k])
            assert opt.net is nets[int(g[name + "_next"][k])]
    np.testing.assert_allclose(float(S.log_trans_prob(nets[0], nets[1])), g["pmp_logtrans_0j"][1], rtol=1e-6)


def test_fit_shapes_and_posterior(ctx):
    """fit() return shapes of lb.py:169,263,350 and a statistical known answer: the chain finds beta0≈-1, beta≈2, sigma≈0.5."""
    from pmp_mcmc_b200 import samplers as S
    from conftest import synthetic_linear
    x, y = synthetic_linear(20000, seed=1)
    data = {"x": x, "y": y}
    tr = S.GMOptimizer(S.BayesNet(), 0.05, N=7, ctx=ctx, seed=3).fit(data, num_steps=400)
    assert tr.shape == (400 * 8, 3) and tr.dtype == np.float64
    tr2 = S.GMpreOptimizerV2(S.BayesNet(), 0.05, N=7, deep=2, ctx=ctx, seed=3).fit(data, num_steps=300)
    assert tr2.shape == (300 * 64, 3)
    tr3 = S.preMOptimizer(S.BayesNet(), 0.05, N=7, ctx=ctx, seed=3).fit(data, num_steps=1500)
    assert tr3.shape == (1500, 3)
    tr4 = S.MetropolisOptimizer(S.BayesNet_o(), 0.05, ctx=ctx, seed=3).fit(data, num_steps=3000)
    assert tr4.shape == (3000, 3)
    for t in (tr[-800:], tr2[-6400:], tr3[-300:], tr4[-500:]):
        m = t.mean(axis=0)
        assert abs(m[0] + 1) < 0.15 and abs(m[1] - 2) < 0.25 and abs(abs(m[2]) - 0.5) < 0.2, m
