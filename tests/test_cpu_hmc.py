"""CPU checks of the HMC oracle (oracle/oracle.py hmc_*), pinned to tests/golden/hmc_step.npz = outputs of the reference's own
complex_nets/Cifar-10/cifar_{SP,MP,PMP}hmc.py and "Bayesian Network Training"/main.py code (oracle/make_golden.py hmc)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as o

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
def test_hmc_weights_against_reference_step(case):
    """PMPHMCOptimizer.step (cifar_PMPhmc.py:77-108), bnnPMPHmc.step (main.py:67-103), MPHMCOptimizer.step (cifar_MPhmc.py:79-86): case 3 has energies
    that underflow (0/0 -> NaN -> 1, x/0 -> Inf -> 1)"""
    nl, ko, ki, kmp = _energies(case)
    for kind, rule in (("PMP", o.HMC_TREE_CIFAR), ("BNN", o.HMC_TREE_BNN)):
        np.testing.assert_allclose(o.hmc_weights(rule, nl, ko, ki), G["A%d_%s_B" % (case, kind)], rtol=2e-6, atol=3e-7)   # 1 - w_old/w_new cancels: one ulp of a float32 exp shows up as ~1e-7 absolute
    np.testing.assert_allclose(o.hmc_weights(o.HMC_MP, nl, kmp), G["A%d_MP_B" % case], rtol=2e-6, atol=1e-7)


def _net(kind):
    import torch
    from torch import nn
    torch.manual_seed(5)
    if kind == "BNN":
        return nn.Sequential(nn.Flatten(), nn.Linear(3 * 32 * 32, 24), nn.ReLU(), nn.Linear(24, 10))
    from pmp_mcmc_b200 import hmc
    return hmc.LeNet()


@pytest.mark.parametrize("kind,N", [("PMP", 3), ("BNN", 3), ("MP", 3), ("SP", 1)])
def test_hmc_fit_restatement_against_reference_fit(kind, N):
    """two steps of the reference's fit() (generators replaced by this repo's streams) = oracle.hmc_fit_restated on the same torch build: accepted indices,
    recorded losses, final parameters"""
    import torch
    n, seed, steps = int(G["B_n"]), int(G["B_seed"]), int(G["B_steps"])
    X = torch.from_numpy(np.random.default_rng(3).standard_normal((n, 3, 32, 32)).astype(np.float32))
    y = torch.from_numpy(np.random.default_rng(4).integers(0, 10, size=n).astype(np.int64))
    losses, picks, net = o.hmc_fit_restated(kind, _net(kind), X, y, steps, seed, N=N)
    final = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).numpy()
    if kind != "SP":
        assert picks == list(G["B_%s_picks" % kind])
    np.testing.assert_allclose(final, G["B_%s_final" % kind], rtol=0, atol=1e-7)
    if len(G["B_%s_losses" % kind]):
        np.testing.assert_allclose(losses, G["B_%s_losses" % kind], rtol=1e-6)
