"""The oracle against the golden vectors produced by the REFERENCE's own Python code (oracle/make_golden.py)."""
import math
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as o

G = os.path.join(ROOT, "tests", "golden")
CASES = {"mp": (o.TREE_FLAT, 4, 1), "mp8": (o.TREE_FLAT, 8, 1), "psp": (o.TREE_BINARY, 2, 3), "pmp": (o.TREE_BARY, 4, 2)}


def logweights(name, lt, props):
    tree, b, depth = CASES[name]
    if name.startswith("mp"):
        return o.mp_logweights(lt, props)
    if name == "psp":
        return o.psp_logweights(lt, props, depth)
    return o.pmp_logweights(lt, props, b, depth)


@pytest.fixture(scope="module")
def lb():
    return np.load(os.path.join(G, "lb_step.npz"))


@pytest.mark.parametrize("name", list(CASES))
def test_proposals_are_reproducible(lb, name):
    tree, b, depth = CASES[name]
    assert np.array_equal(o.propose(tree, b, depth, 3, 0.05, lb["state"], 1, 0), lb[name + "_props"])


@pytest.mark.parametrize("name", list(CASES))
def test_loglik_restatements_match_reference_BayesNet_loglik(lb, name):
    """BayesNet.loglik (lb.py:103-108) as run by the reference on its own 500-point dataset."""
    x, y, props = lb["x"], lb["y"], lb[name + "_props"]
    ref = lb[name + "_loglik"]
    np.testing.assert_allclose(o.loglik_lb_torch(x, y, props), ref, rtol=1e-6)
    np.testing.assert_allclose(o.loglik_linear_f64(x, y, props, len(x) / 50.0), ref, rtol=1e-5)
    np.testing.assert_allclose(o.loglik_linear_refcuda(x, y, props, len(x) / 50.0), ref, rtol=1e-5)


@pytest.mark.parametrize("name", list(CASES))
def test_log_trans_prob(lb, name):
    props = lb[name + "_props"].astype(np.float64)
    got = np.array([o.log_kernel(props[0], props[j]) for j in range(len(props))])
    np.testing.assert_allclose(got, lb[name + "_logtrans_0j"], rtol=1e-6, atol=1e-9)   # reference evaluates scipy on float32 scalars


@pytest.mark.parametrize("name", list(CASES))
def test_step_draws_match_reference_step(lb, name):
    """GMOptimizer/preMOptimizer/GMpreOptimizerV2.step (lb.py:139-164, 206-258, 304-345) with the recorded uniforms:
    the oracle's weights + draw_numpy reproduce the reference's resampled indices and next state exactly."""
    props, lt = lb[name + "_props"].astype(np.float64), lb[name + "_loglik"]
    A = logweights(name, lt, props)
    w = o.weights_from_log(A)
    P = len(lt)
    for k in range(len(lb[name + "_u"])):
        draws = o.draw_numpy(w, lb[name + "_u"][k])
        assert np.array_equal(draws, lb[name + "_draws"][k])
        assert draws[o.pick_index(lb[name + "_upick"][k], P)] == lb[name + "_next"][k]
        assert np.array_equal(o.draw_blocked(w, lb[name + "_u"][k], "right"), lb[name + "_draws"][k])


def test_mh_logpost(lb):
    got = o.loglik_linear_f64(lb["x"], lb["y"], np.stack([lb["state"], lb["mp_props"][1]]), len(lb["x"]) / 50.0)
    np.testing.assert_allclose(got, lb["mh_logpost"], rtol=1e-5)


def test_analytic_targets():
    g = np.load(os.path.join(G, "analytic_targets.npz"))
    np.testing.assert_allclose(np.exp([o.log_normal1d(v, 0.3, 1.7) for v in g["xs"]]), g["normal_pdf"], rtol=1e-12)
    np.testing.assert_allclose(np.exp([o.log_banana(p) for p in g["pts2"]]), g["banana_pdf"], rtol=1e-10, atol=1e-300)
    for d in (10, 40):
        np.testing.assert_allclose([o.log_stdnormal(p) for p in g["mvn_pts_%d" % d]], np.log(g["mvn_pdf_%d" % d]), rtol=1e-10)
        # transition_prob (com_dim.py:18-21, sigma = 0.5): Gaussian kernel in the squared distance times constants
        pts = g["mvn_pts_%d" % d]
        ref = g["trans_%d" % d]
        got = np.array([1.0 / (math.sqrt(2 * math.pi) * 0.5) * math.exp(-0.5 * np.sum((p - pts[0]) ** 2) / 0.25) * 10 ** (d / 10) for p in pts])
        np.testing.assert_allclose(got, ref, rtol=1e-12)
