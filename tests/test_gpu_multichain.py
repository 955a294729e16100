"""Co-scheduled independent chains (pmp_run_multi, chain_persistent_multi.cuh): every chain's trace must be bit-identical to
the same chain run alone with pmp_run — the kernel changes WHEN a chain's sweep runs, never what it computes."""
import numpy as np
import pytest

from conftest import synthetic_linear

pytestmark = pytest.mark.gpu

CASES = [
    # tree, b, depth, algo, draw, flags, P-ish label
    ("mp_flat_1024", dict(tree="FLAT", b=1024, depth=1, algo="MP", draw="CUDA", flags=0, alpha=0.01, scale=100.0)),
    ("psp_binary_d8", dict(tree="BINARY", b=2, depth=8, algo="PSP", draw="PYTHON", flags=0, alpha=0.02, scale=100.0)),
    ("mp_flat_37", dict(tree="FLAT", b=37, depth=1, algo="MP", draw="PYTHON", flags=0, alpha=0.05, scale=10.0)),
]


def _configure(c, L, k):
    c.configure(getattr(L, "TREE_" + k["tree"]), b=k["b"], depth=k["depth"], dim=3, target=L.TARGET_LINEAR_GAUSS, algo=getattr(L, "ALGO_" + k["algo"]),
                draw=getattr(L, "DRAW_" + k["draw"]), flags=k["flags"], alpha=k["alpha"], scale=k["scale"])


@pytest.mark.parametrize("name,k", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("n_chains", [1, 2, 3, 4, 8, 13, 32])
def test_co_scheduled_chains_equal_solo_chains(name, k, n_chains):
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    n, iters = 5000, 40
    x, y = synthetic_linear(n, seed=3)
    what = L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW
    rs = np.random.default_rng(17)
    starts = [np.array([0.0, 0.0, 1.0], np.float32), np.array([-1.0, 2.0, 0.5], np.float32)] + \
             [np.array([rs.uniform(-1, 1), rs.uniform(-1, 2), rs.uniform(0.5, 2)], np.float32) for _ in range(30)]
    solo = []
    for i in range(n_chains):
        c = pm.Context(0)
        _configure(c, L, k)
        c.set_data_linear(x, y); c.set_state(starts[i]); c.seed(100 + i, 0)
        c.trace_config(iters, what)
        c.run(iters)
        solo.append((c.read_trace(), c.get_state(), c.iteration()))
        c.close()
    ctxs = []
    for i in range(n_chains):
        c = pm.Context(0)
        _configure(c, L, k)
        if i == 0:
            c.set_data_linear(x, y)
        else:
            c.share_data_from(ctxs[0])
        c.set_state(starts[i]); c.seed(100 + i, 0)
        c.trace_config(iters, what)
        ctxs.append(c)
    L.run_multi(ctxs, iters)
    for i, c in enumerate(ctxs):
        tr, (ref, st, it) = c.read_trace(), solo[i]
        assert tr["n"] == ref["n"] == iters
        for key in ("state", "next", "draws"):
            assert np.array_equal(tr[key], ref[key]), (name, i, key)
        assert np.array_equal(tr["logw"].view(np.uint64), ref["logw"].view(np.uint64)), (name, i, "logw")
        assert np.array_equal(c.get_state(), st) and c.iteration() == it
    # the chains are independent: different seeds give different trajectories
    if n_chains > 1:
        assert not np.array_equal(ctxs[0].read_trace()["next"], ctxs[1].read_trace()["next"])
    # a second joint run continues every chain where it stopped (counters and state live on the device)
    L.run_multi(ctxs, 10)
    assert all(c.iteration() == iters + 10 for c in ctxs)
    for c in reversed(ctxs):
        c.close()


def test_run_multi_rejects_unshared_data():
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    x, y = synthetic_linear(1000, seed=1)
    a, b = pm.Context(0), pm.Context(0)
    for c in (a, b):
        _configure(c, L, CASES[0][1]); c.set_data_linear(x, y); c.set_state([0, 0, 1]); c.seed(1, 0)
    with pytest.raises(L.PmpError, match="share"):
        L.run_multi([a, b], 10)
    b.close(); a.close()


def test_fit_independent_equals_fit():
    """samplers.fit_independent: lb.py trainers co-scheduled; each gets the trace its own fit() returns."""
    from pmp_mcmc_b200 import samplers as S
    x, y = synthetic_linear(4000, seed=8)
    data = {"x": x, "y": y}
    solo = []
    for k, alpha in enumerate((0.01, 0.03, 0.1)):
        t = S.GMOptimizer(S.BayesNet(), alpha, N=7, seed=40 + k)
        solo.append((t.fit(data, 30), t.net.theta()))
    trainers = [S.GMOptimizer(S.BayesNet(), alpha, N=7, seed=40 + k) for k, alpha in enumerate((0.01, 0.03, 0.1))]
    traces = S.fit_independent(trainers, data, 30)
    for (ref, th), tr, t in zip(solo, traces, trainers):
        assert tr.shape == (30 * 8, 3) and np.array_equal(tr, ref) and np.array_equal(t.net.theta(), th)


def test_fit_independent_in_groups():
    """More trainers than one launch takes: groups of `max_group` chains, every group aliasing the first trainer's data; a leftover single
    chain runs alone.  Each trainer still gets its own fit() trace."""
    from pmp_mcmc_b200 import samplers as S
    x, y = synthetic_linear(3000, seed=9)
    data = {"x": x, "y": y}
    alphas = [0.01 * (1 + k) for k in range(10)]
    solo = []
    for k, alpha in enumerate(alphas):
        t = S.preMOptimizer(S.BayesNet(), alpha, N=7, seed=70 + k)
        solo.append(t.fit(data, 12))
    trainers = [S.preMOptimizer(S.BayesNet(), alpha, N=7, seed=70 + k) for k, alpha in enumerate(alphas)]
    traces = S.fit_independent(trainers, data, 12, max_group=3)          # 3 + 3 + 3 + 1
    for ref, tr in zip(solo, traces):
        assert np.array_equal(tr, ref)


def test_headline_shape_co_scheduled_equals_solo():
    """The launch bench.py times: P = 1024 flat MP, n = 100 000, CUDA draw rule, 8 chains co-scheduled — chain 0 and chain 7 against solo runs."""
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    n, iters, K = 100000, 60, 8
    x, y = synthetic_linear(n, seed=0)
    what = L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS

    def make(seed, owner=None):
        c = pm.Context(0)
        c.configure(L.TREE_FLAT, b=1024, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=1000.0)
        if owner is None:
            c.set_data_linear(x, y)
        else:
            c.share_data_from(owner)
        c.set_state([1, 1, 1]); c.seed(seed, 0); c.trace_config(iters, what)
        return c
    ctxs = [make(2024)]
    ctxs += [make(2024 + k, ctxs[0]) for k in range(1, K)]
    L.run_multi(ctxs, iters)
    got = [c.read_trace() for c in ctxs]
    for k in (0, K - 1):
        s = make(2024 + k)
        s.run(iters)
        ref = s.read_trace()
        for key in ("state", "next", "draws"):
            assert np.array_equal(got[k][key], ref[key]), (k, key)
        s.close()
    for c in reversed(ctxs):
        c.close()


@pytest.mark.parametrize("tree,P,n", [("FLAT", 1024, 20000), ("FLAT", 37, 5000), ("FLAT", 4, 500), ("FLAT", 2048, 3000), ("FLAT", 130, 64),
                                      ("PSP", 1024, 20000), ("PSP", 2048, 3000), ("PSP", 8, 500), ("PSP", 256, 64), ("TABLE", 1024, 5000)])
def test_state_only_handoff_equals_node_handoff(tree, P, n, monkeypatch):
    """Flat and binary trees in the persistent kernels: the acceptance publishes the accepted state and every reader derives its nodes (Handoff::state) —
    the traces, the draws, the log-weights and the nodes left behind must equal those of the hand-off that publishes every node (PMP_DERIVE_NODES=0),
    across launch boundaries and for a co-scheduled launch of one chain (chain_persistent_multi_kernel takes the same path for K = 1)."""
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    x, y = synthetic_linear(n, seed=5)
    what = L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW
    res = {}
    for mode in ("0", "1", "multi"):
        monkeypatch.setenv("PMP_DERIVE_NODES", "0" if mode == "0" else "1")
        c = pm.Context(0)
        if tree == "FLAT":
            c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=0.02, scale=max(n / 100.0, 1.0))
        elif tree == "PSP":          # binary prefetch tree: a node is the state plus the increments of its ancestors
            c.configure(L.TREE_BINARY, depth=P.bit_length() - 1, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_PSP, draw=L.DRAW_PYTHON, alpha=0.02, scale=max(n / 100.0, 1.0))
        else:                        # the 100000_PMP.cu shape (bench.py pmp_tree block)
            c.configure(L.TREE_BINARY, depth=P.bit_length() - 1, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=0.02, scale=max(n / 100.0, 1.0),
                        flags=L.FLAG_QUIRK_TABLE_CONST)
        c.set_data_linear(x, y); c.set_state([0.5, 1.0, 1.5]); c.seed(11, 0)
        c.trace_config(200, what)
        if mode == "multi":
            for iters in (3, 60, 2, 35):
                L.run_multi([c], iters)
        else:
            for iters in (3, 60, 2, 35):          # 2 iterations: the shortest launch the persistent kernel takes
                c.run(iters)
        tr = c.read_trace()
        res[mode] = (tr, c.read_proposals().copy(), c.get_state().copy())
        c.close()
    for mode in ("1", "multi"):
        for key in ("state", "next", "draws", "logw"):
            assert np.array_equal(res["0"][0][key], res[mode][0][key]), (mode, key)
        assert np.array_equal(res["0"][1], res[mode][1]), mode
        assert np.array_equal(res["0"][2], res[mode][2]), mode
