"""FC path, CPU side: the oracle against the golden vectors produced by the reference's PMP_FC.py / MP_FC.py / MH_FC.py."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as o

G = np.load(os.path.join(ROOT, "tests", "golden", "fc_step.npz"))


def golden_inputs():
    n = int(G["n"])
    rng = np.random.default_rng(int(G["data_seed"]))
    X = rng.standard_normal((n, 28, 28)).astype(np.float32)
    y = rng.integers(0, 10, size=n).astype(np.int64)
    return X, y, o.fc_init_theta(int(G["theta_seed"]))


def test_fc_dim_and_layout():
    assert o.FC_DIM == 567434
    parts = o.fc_unpack(np.arange(o.FC_DIM, dtype=np.float32))
    assert [p.shape for p in parts] == o.FC_SHAPES and parts[2][0, 0] == 512 * 784 + 512


@pytest.mark.parametrize("kind,tree,depth", [("PMP", o.TREE_BINARY, 3), ("MP", o.TREE_FLAT, 1)])
def test_loss_and_step_against_reference(kind, tree, depth):
    X, y, theta0 = golden_inputs()
    P = 8
    props = o.propose(tree, P if tree == o.TREE_FLAT else 2, depth, o.FC_DIM, float(G["alpha"]), theta0, int(G["prop_seed"]), 0)
    ref_loss = G[kind + "_loss"]
    got32 = np.array([o.fc_loss_torch32(X, y, props[p]) for p in range(P)])
    np.testing.assert_allclose(got32, ref_loss, rtol=2e-6)                                   # same float32 arithmetic as loss(net), PMP_FC.py:40-44
    truth = np.array([o.fc_mean_ce_f64(X, y, props[p]) / 10.0 for p in range(P)])
    np.testing.assert_allclose(truth, ref_loss, rtol=2e-6)
    # the draw: the reference's weights B (captured at torch.multinomial) + the recorded uniform give its accepted index
    assert o.draw_numpy(G[kind + "_B"], [float(G[kind + "_u"])])[0] == int(G[kind + "_I"])
    np.testing.assert_allclose(ref_loss[int(G[kind + "_I"])] * 10.0, float(G[kind + "_step_loss"]), rtol=1e-6)
    # the weight rule restated in binary64 on binary64 losses; the reference evaluates it in float32 on losses that differ in
    # the 6th digit, so its B is only a noisy version of this (documented in DESIGN.md §5): same ranking of the heavy nodes
    if kind == "PMP":
        A = o.standardize(o.psp_logweights(-truth, props[:, :8].astype(np.float64), depth, use_kernel=False))
    else:
        A = -truth.copy()          # the kernel term of MP_FC.py:107-114 is the same for all nodes to 1e-9 at alpha = 1e-4
        A = o.standardize(A)
    assert np.argmax(A) == np.argmax(G[kind + "_B"])
    assert np.corrcoef(A, np.log(G[kind + "_B"]))[0, 1] > 0.9


def test_mh_loss():
    X, y, theta0 = golden_inputs()
    props = o.propose(o.TREE_FLAT, 8, 1, o.FC_DIM, float(G["alpha"]), theta0, int(G["prop_seed"]), 0)
    got = np.array([o.fc_loss_torch32(X, y, props[p], div=1.0) for p in range(2)])
    np.testing.assert_allclose(got, G["MH_loss"], rtol=2e-6)


GT = np.load(os.path.join(ROOT, "tests", "golden", "fc_step_trained.npz"))


def trained_inputs(tag):
    """Inputs of the trained-model fixtures (oracle/make_golden.py golden_fc_trained): theta0 = the reference's FC_model.pkl flattened."""
    n = int(GT[tag + "_n"])
    rng = np.random.default_rng(int(GT[tag + "_data_seed"]))
    X = rng.standard_normal((n, 28, 28)).astype(np.float32)
    return X, GT[tag + "_labels"].astype(np.int64), np.load(os.path.join(ROOT, "tests", "golden", "fc_theta0.npy"))


def trained_rule_f64(kind, truth, props):
    """The weight rule of PMP_FC.py:119-140 / MP_FC.py:107-119 restated in binary64 on binary64 losses."""
    if kind == "PMP":
        return o.standardize(o.psp_logweights(-truth, props[:, :8].astype(np.float64), 3, use_kernel=False))
    kt = o.mp_logweights(np.zeros(8), props.astype(np.float64) / np.sqrt(o.FC_DIM)) / 8.0
    return o.standardize(-truth + (kt - kt.mean()))


@pytest.mark.parametrize("kind,tree,depth", [("PMP", o.TREE_BINARY, 3), ("MP", o.TREE_FLAT, 1)])
def test_trained_model_step_against_reference(kind, tree, depth):
    """The regime the reference runs (theta0 = FC_model.pkl, alpha = 1e-4, small loss, node differences ~1e-5 of it): the oracle's
    float32 loss is the reference's, and the binary64 restatement of the weight rule reproduces the reference's accepted index for
    every injected uniform (its float32 weights sit within 0.01 of the exact cdf here, unlike the random-init fixture)."""
    X, y, theta0 = trained_inputs("s")
    assert theta0.shape == (o.FC_DIM,) and theta0.dtype == np.float32
    props = o.propose(tree, 8 if tree == o.TREE_FLAT else 2, depth, o.FC_DIM, float(GT["alpha"]), theta0, int(GT["prop_seed"]), 0)
    ref_loss = GT["s_%s_loss" % kind]
    got32 = np.array([o.fc_loss_torch32(X, y, props[p]) for p in range(8)])
    np.testing.assert_allclose(got32, ref_loss, rtol=2e-6)
    truth = np.array([o.fc_mean_ce_f64(X, y, props[p]) / 10.0 for p in range(8)])
    np.testing.assert_allclose(truth, GT["s_%s_truth" % kind], rtol=1e-12)
    B = GT["s_%s_B" % kind]
    w = o.weights_from_log(trained_rule_f64(kind, truth, props))
    assert np.max(np.abs(np.cumsum(w / w.sum()) - np.cumsum(B / B.sum()))) < 0.01
    assert np.array_equal(o.draw_numpy(w, GT["u_grid"]), GT["s_%s_I" % kind])
    assert np.array_equal(o.draw_numpy(B, GT["u_grid"]), GT["s_%s_I" % kind])
