"""GPU parity of the HMC variants (csrc/hmc.cu, hmc.py): acceptance weights and accepted indices against the reference's step() outputs
(tests/golden/hmc_step.npz), the leapfrog kernels against the scripts' float32 torch arithmetic, and whole fit() trajectories against the oracle's
restatement of the scripts' loops (itself pinned to the reference's fit() in tests/test_cpu_hmc.py)."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "hmc_step.npz"))


def _energies(case):
    nl, ps = G["A%d_nl" % case], G["A%d_ps" % case]
    P = len(nl)
    ke = lambda a, b: float((ps[a, b].astype(np.float32) ** 2).sum(dtype=np.float32) / np.float32(2))
    ko, ki = np.zeros(P), np.zeros(P)
    for c in range(1, P):
        par = c - (1 << (c.bit_length() - 1))
        ko[c], ki[c] = ke(par, c), ke(c, par)
    return nl, ko, ki, np.array([ke(j, 0) for j in range(P)])


@pytest.mark.parametrize("case", range(5))
def test_hmc_accept_against_reference_step(ctx, case):
    """device weights = the B of PMPHMCOptimizer.step / bnnPMPHmc.step / MPHMCOptimizer.step; the accepted index = the inverse-CDF draw from the
    reference's B wherever the uniform is farther from a cdf boundary than the float32 discrepancy of the weights"""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    nl, ko, ki, kmp = _energies(case)
    for kind, rule, a, b in (("PMP", L.HMC_RULE_TREE_CIFAR, ko, ki), ("BNN", L.HMC_RULE_TREE_BNN, ko, ki), ("MP", L.HMC_RULE_MP, kmp, None)):
        B_ref = G["A%d_%s_B" % (case, kind)].astype(np.float64)
        compared = 0
        for u in G["u_grid"]:
            w, idx = ctx.hmc_accept(rule, nl, a, b, u=float(u))
            np.testing.assert_allclose(w, B_ref, rtol=2e-6, atol=3e-7)
            if B_ref.sum() > 0:
                cdf = np.cumsum(B_ref / B_ref.sum())
                if np.min(np.abs(u - cdf[:-1])) > 1e-5:
                    assert idx == int(o.draw_numpy(B_ref, [u])[0]), (kind, u)
                    compared += 1
        assert compared >= 11 or B_ref.sum() == 0


def test_hmc_single_proposal_rule(ctx):
    """cifar_SPhmc.py:118-126: accept iff exp((-H_0 + H_1) * 1000) > rand"""
    from pmp_mcmc_b200 import _lib as L
    for nl, ke, u in (([-2.30, -2.2995], [1e-5, 2e-5], 0.5), ([-2.30, -2.3008], [1e-5, 2e-5], 0.5), ([-2.30, -2.3008], [1e-5, 2e-5], 0.4), ([-2.3, -2.2], [0.0, 0.0], 0.99)):
        w, idx = ctx.hmc_accept(L.HMC_RULE_SP, nl, ke, u=u)
        f = np.float32
        with np.errstate(over="ignore"):
            acc = np.exp(f(1000) * (-(f(ke[0]) + f(nl[0])) + (f(nl[1]) + f(ke[1]))), dtype=f)
        assert idx == int(acc > f(u))
        np.testing.assert_allclose(w[1], acc, rtol=1e-5)


def test_hmc_leapfrog_matches_the_scripts_float32_arithmetic(ctx):
    import torch
    from oracle import oracle as o
    dev = "cuda:%d" % ctx.device
    d, step = 100003, 0.1
    g0 = torch.Generator().manual_seed(1)
    theta = torch.randn(d, generator=g0).to(dev); grad = torch.randn(d, generator=g0).to(dev); grad2 = torch.randn(d, generator=g0).to(dev)
    for sign, si in ((1.0, 3), (-1.0, 0)):
        ctx.seed(77, 5)
        child = torch.empty(d, device=dev); p = torch.empty(d, device=dev)
        torch.cuda.synchronize()
        k0 = ctx.hmc_leapfrog_begin(theta, grad, child, p, step, sign, 0.0005, si)
        p0 = torch.from_numpy(o.hmc_momentum(77, 5, si, d)).to(dev) * 0.0005          # torch.randn(d) * 0.0005 with the library's stream
        pr = p0.clone(); pr += sign * step * grad / 2                                  # cifar_PMPhmc.py:141
        th = theta.clone(); th += sign * step * pr                                     # cifar_PMPhmc.py:145
        assert torch.equal(p, pr) and torch.equal(child, th)
        np.testing.assert_allclose(k0, float((p0.double() ** 2).sum() / 2), rtol=1e-12)
        k1 = ctx.hmc_leapfrog_end(p, grad2, step, sign)
        pr += sign * step * grad2 / 2                                                  # cifar_PMPhmc.py:162
        assert torch.equal(p, pr)
        np.testing.assert_allclose(k1, float((pr.double() ** 2).sum() / 2), rtol=1e-12)
        # injected momentum (the MP path carries p from node to node, cifar_MPhmc.py:106)
        q = pr.clone()
        torch.cuda.synchronize()
        ctx.hmc_leapfrog_begin(theta, grad, child, q, step, sign, 0.0005, 0, p_init=q)
        pr += sign * step * grad / 2
        assert torch.equal(q, pr)


@pytest.mark.parametrize("kind,N", [("PMP", 3), ("BNN", 3), ("MP", 3), ("SP", 1), ("PMP", 7)])
def test_hmc_fit_equals_restated_scripts(ctx, kind, N):
    """hmc.py (device leapfrog + device acceptance around torch autograd) walks the same trajectory as the scripts' loops restated in pure torch on the same GPU"""
    import copy
    import torch
    from torch import nn
    from oracle import oracle as o
    from pmp_mcmc_b200 import hmc
    n, seed, steps = 24, 11, 3
    Xn = np.random.default_rng(3).standard_normal((n, 3, 32, 32)).astype(np.float32)
    yn = np.random.default_rng(4).integers(0, 10, size=n).astype(np.int64)
    hmc.set_data(Xn, yn, ctx=ctx)
    torch.manual_seed(5)
    net = nn.Sequential(nn.Flatten(), nn.Linear(3 * 32 * 32, 24), nn.ReLU(), nn.Linear(24, 10)) if kind == "BNN" else hmc.LeNet()
    losses, picks, ref_net = o.hmc_fit_restated(kind, copy.deepcopy(net).to(hmc.device), hmc.X, hmc.y, steps, seed, N=N, device=hmc.device)
    cls = {"PMP": hmc.PMPHMCOptimizer, "BNN": hmc.bnnPMPHmc, "MP": hmc.MPHMCOptimizer, "SP": hmc.HMCOptimizer}[kind]
    opt = cls(copy.deepcopy(net), 0.001, seed=seed) if kind == "SP" else cls(copy.deepcopy(net), 0.001, N, seed=seed)
    opt.fit(num_steps=steps)
    assert opt.picks == picks
    a = torch.cat([p.detach().reshape(-1) for p in opt.net.parameters()]); b = torch.cat([p.detach().reshape(-1) for p in ref_net.parameters()])
    # the leapfrog arithmetic is bit-exact (test above); two autograd passes through cuDNN convolutions are not bit-reproducible run to run
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=0, atol=2e-6)
    if kind in ("BNN", "SP"):
        np.testing.assert_allclose(opt.loss_list, losses, rtol=1e-6)
