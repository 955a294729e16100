"""error.py's MP / PSP / PMP: the oracle restatement replays the REFERENCE's own runs bit for bit (tests/golden/error_chains.npz)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as o

G = np.load(os.path.join(ROOT, "tests", "golden", "error_chains.npz"))


@pytest.mark.parametrize("kind", ["MP", "PSP", "PMP", "PMP3"])
def test_error_py_replay_is_exact(kind):
    hops, N, deep = [int(v) for v in G[kind + "_args"]]
    mu, sigma = G[kind + "_musigma"]
    X = o.error_py_replay(kind.rstrip("3"), hops, mu, sigma, N, deep, float(G[kind + "_x0"]), G[kind + "_normals"], G[kind + "_u"], G[kind + "_picks"])
    assert np.array_equal(X, G[kind + "_X"])


def test_analytic_chain_is_deterministic_and_chain_ids_differ():
    a = o.analytic_chain(o.TREE_BARY, 4, 2, 2, 2, 4, 0, 1.0, 5, 6, [0.0, -10.0], chain=0)
    b = o.analytic_chain(o.TREE_BARY, 4, 2, 2, 2, 4, 0, 1.0, 5, 6, [0.0, -10.0], chain=0)
    c = o.analytic_chain(o.TREE_BARY, 4, 2, 2, 2, 4, 0, 1.0, 5, 6, [0.0, -10.0], chain=3)
    assert np.array_equal(a["samples"], b["samples"]) and not np.array_equal(a["samples"], c["samples"])
