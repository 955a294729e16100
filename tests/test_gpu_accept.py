"""GPU parity tests of the acceptance step (log-weights, categorical draws, next state) through the C-ABI.

Bit-exact for the index work: given the same weights and uniforms the device's draws equal the oracle's mirror of the
device scan (oracle.draw_blocked) always, and the reference's own draw semantics (numpy searchsorted 'right' /
libstdc++ lower_bound) on every case tested here."""
import numpy as np
import pytest

from conftest import synthetic_linear

pytestmark = pytest.mark.gpu


def _o():
    from oracle import oracle
    return oracle


def _L():
    from pmp_mcmc_b200 import _lib
    return _lib


def _oracle_logweights(o, L, algo, lt, props, b, depth, flags, ks=1.0):
    use_k = not (flags & L.FLAG_NO_KERNEL_TERM)
    if algo == L.ALGO_MP:
        return o.mp_logweights(lt, props, ks, use_k)
    if algo == L.ALGO_PSP:
        return o.psp_logweights(lt, props, depth, ks, use_k)
    if algo == L.ALGO_PMP:
        return o.pmp_logweights(lt, props, b, depth, ks, use_k, bool(flags & L.FLAG_QUIRK_LEVEL_MOD))
    if algo == L.ALGO_TABLE:
        if not use_k and not (flags & L.FLAG_QUIRK_TABLE_CONST):
            return np.array(lt, dtype=np.float64)
        return o.table_logweights(lt, props, b, depth, ks, bool(flags & L.FLAG_QUIRK_TABLE_CONST))
    raise AssertionError


CASES = [
    # tree, b, depth, dim, algo, draw, flags
    (0, 4, 1, 3, "MP", "PYTHON", 0), (0, 8, 1, 3, "MP", "CUDA", 0), (0, 64, 1, 1, "MP", "PYTHON", 0), (0, 1024, 1, 3, "MP", "CUDA", 0),
    (1, 2, 3, 3, "PSP", "PYTHON", 0), (1, 2, 10, 3, "PSP", "PYTHON", 0), (1, 2, 5, 40, "PSP", "PYTHON", 0), (1, 2, 3, 2, "PSP", "SINGLE", "STD"),
    (2, 4, 2, 1, "PMP", "PYTHON", 0), (2, 8, 2, 3, "PMP", "PYTHON", 0), (2, 4, 3, 2, "PMP", "PYTHON", 0), (2, 3, 3, 1, "PMP", "PYTHON", "LEVELMOD"),
    (1, 2, 4, 3, "TABLE", "CUDA", 0), (1, 2, 10, 3, "TABLE", "CUDA", "CONST"), (2, 8, 3, 3, "TABLE", "CUDA", 0), (0, 1500, 1, 3, "MP", "CUDA", 0),
]


@pytest.mark.parametrize("tree,b,depth,dim,algo,draw,flag", CASES)
def test_accept_against_oracle(ctx, tree, b, depth, dim, algo, draw, flag):
    L, o = _L(), _o()
    algo_id = getattr(L, "ALGO_" + algo)
    draw_id = getattr(L, "DRAW_" + draw)
    flags = {0: 0, "STD": L.FLAG_STANDARDIZE, "LEVELMOD": L.FLAG_QUIRK_LEVEL_MOD, "CONST": L.FLAG_QUIRK_TABLE_CONST}[flag]
    ctx.configure(tree, b=b, depth=depth, dim=dim, target=L.TARGET_EXTERNAL, algo=algo_id, draw=draw_id, flags=flags, alpha=0.3, scale=1.0)
    P = ctx.P
    rng = np.random.default_rng(P * 7 + dim)
    state = rng.normal(size=dim).astype(np.float32)
    bb = 2 if tree == 1 else b
    for trial in range(3):
        ctx.set_state(state)
        ctx.seed(1000 + trial, trial)
        ctx.propose()
        props = ctx.read_proposals()
        lt = rng.normal(scale=3.0, size=P) - 50.0
        if trial == 2 and not (flags & L.FLAG_STANDARDIZE):   # standardising -inf is NaN in the reference too (PMP_FC.py:138-140)
            lt[rng.integers(0, P, size=max(1, P // 8))] = -np.inf     # dead nodes
            lt[0] = -40.0
        ctx.write_logtarget(lt)
        n_u = P + 1 if draw == "PYTHON" else (1 if draw == "SINGLE" else P)
        u = rng.random(n_u)
        idx, nxt = ctx.accept(u)
        A_dev = ctx.read_logweights()
        A_ref = _oracle_logweights(o, L, algo_id, lt, props.astype(np.float64), bb, depth, flags)
        if flags & L.FLAG_STANDARDIZE:
            A_ref = o.standardize(A_ref)
        fin = np.isfinite(A_ref)
        assert np.array_equal(fin, np.isfinite(A_dev))
        np.testing.assert_allclose(A_dev[fin], A_ref[fin], rtol=1e-11, atol=1e-9)
        # draws: exact against the mirror of the device scan on the device's own weights ...
        w = o.weights_from_log(A_dev)
        side = "left" if draw == "CUDA" else "right"
        nd = 1 if draw == "SINGLE" else P
        assert np.array_equal(idx, o.draw_blocked(w, u[:nd], side))
        # ... and equal to the reference's own draw semantics on the oracle's weights
        w_ref = o.weights_from_log(A_ref)
        ref_idx = o.draw_libstdcxx(w_ref, u[:nd]) if draw == "CUDA" else o.draw_numpy(w_ref, u[:nd])
        assert np.array_equal(idx, ref_idx)
        expect_next = idx[o.pick_index(u[P], P)] if draw == "PYTHON" else idx[0]
        assert nxt == expect_next
        assert np.array_equal(ctx.get_state(), props[nxt])
        state = props[nxt]


@pytest.mark.parametrize("algo,u,lt,expect", [("MH", 0.3, (-10.0, -10.5), 1), ("MH", 0.7, (-10.0, -10.5), 0), ("MH", 0.999, (-10.0, -9.0), 1),
                                              ("BARKER", 0.3, (-10.0, -10.5), 1), ("BARKER", 0.4, (-10.0, -10.5), 0), ("BARKER", 0.7, (-10.0, -9.0), 1)])
def test_single_proposal_rules(ctx, algo, u, lt, expect):
    """MH: u < exp(lt1-lt0) (lb.py:65-69).  Barker: u < w1/(w0+w1) (error.py:29-35)."""
    L = _L()
    ctx.configure(L.TREE_FLAT, b=2, dim=2, target=L.TARGET_EXTERNAL, algo=getattr(L, "ALGO_" + algo), draw=L.DRAW_SINGLE, alpha=0.1)
    ctx.set_state([0.5, -0.5]); ctx.seed(3, 0); ctx.propose()
    props = ctx.read_proposals()
    ctx.write_logtarget(np.array(lt))
    idx, nxt = ctx.accept(np.array([u]))
    assert nxt == expect and idx[0] == expect
    assert np.array_equal(ctx.get_state(), props[expect])


@pytest.mark.parametrize("tree,b,depth,algo,draw,flags,scale,n", [(0, 1024, 1, "MP", "CUDA", 0, 1000.0, 20000), (0, 4, 1, "MP", "PYTHON", 0, 10.0, 20000), (1, 2, 10, "PSP", "PYTHON", 0, 2000.0, 20000),
                                                                  (1, 2, 3, "PSP", "PYTHON", 0, 2000.0, 20000), (1, 2, 10, "TABLE", "CUDA", "CONST", 1000.0, 20000), (2, 8, 2, "PMP", "PYTHON", 0, 2000.0, 20000),
                                                                  (0, 2000, 1, "MP", "CUDA", 0, 2000.0, 20000), (0, 1024, 1, "MP", "CUDA", 0, 10.0, 500), (0, 4, 1, "MP", "CUDA", 0, 10.0, 500),
                                                                  (1, 2, 10, "PSP", "PYTHON", 0, 10.0, 500)])
@pytest.mark.parametrize("generic,persistent,tc", [(0, 1, 1), (0, 1, 0), (0, 0, 0), (0, 0, 1), (1, 0, 0)])
def test_device_resident_chain_replays_in_oracle(tree, b, depth, algo, draw, flags, scale, n, generic, persistent, tc, monkeypatch):
    """pmp_run (device-resident loop: fused sweep, acceptance kernel that also publishes the next nodes, CUDA graph) against a
    step-by-step oracle replay: same proposals (bit-exact), log-weights within 1e-6 relative, identical draw and accepted
    index sequences, identical states.  `generic` switches the acceptance to the general kernel, `persistent` the loop structure."""
    import pmp_mcmc_b200 as pm
    L, o = _L(), _o()
    monkeypatch.setenv("PMP_ACCEPT_GENERIC", str(generic))
    monkeypatch.setenv("PMP_SWEEP_TC", str(tc))               # 1: tensor-core sweep in the stepwise loop ...
    monkeypatch.setenv("PMP_PERSISTENT_TC", str(tc))          # ... and in the persistent kernel
    monkeypatch.setenv("PMP_PERSISTENT", str(persistent))     # 1: one cooperative kernel for the whole chain; 0: CUDA-graph replay of sweep + acceptance
    monkeypatch.setenv("PMP_GRAPH_ITERS", "8")
    iters, seed = 13, 77
    x, y = synthetic_linear(n, seed=5)
    fl = L.FLAG_QUIRK_TABLE_CONST if flags == "CONST" else 0
    c = pm.Context(0)
    try:
        c.configure(tree, b=b, depth=depth, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=getattr(L, "ALGO_" + algo), draw=getattr(L, "DRAW_" + draw),
                    flags=fl, alpha=0.02, scale=scale)
        P = c.P
        c.set_data_linear(x, y)
        state = np.array([-0.8, 1.7, 0.7], dtype=np.float32)
        c.set_state(state); c.seed(seed, 0)
        c.trace_config(iters, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_DRAWS | L.TRACE_LOGW | L.TRACE_SAMPLES)
        c.run(iters)
        tr = c.read_trace()
        assert tr["n"] == iters and c.iteration() == iters
        bb = 2 if tree == 1 else b
        near_ties = 0
        for it in range(iters):
            props = o.propose(tree, b, depth, 3, 0.02, state, seed, it)
            lt = o.loglik_linear_f64(x, y, props, scale)
            A = _oracle_logweights(o, L, getattr(L, "ALGO_" + algo), lt, props.astype(np.float64), bb, depth, fl)
            # log-weights are sums of differences of log-targets, each good to ~1e-7 relative: the absolute term scales with |lt|
            np.testing.assert_allclose(tr["logw"][it], A, rtol=1e-6, atol=1e-7 + 4e-7 * float(np.max(np.abs(lt[np.isfinite(lt)]))))
            u = o.stream_uniforms(seed, it, o.STREAM_DRAW, 0, P)
            w_dev = o.weights_from_log(tr["logw"][it])
            side = "left" if draw == "CUDA" else "right"
            assert np.array_equal(tr["draws"][it], o.draw_blocked(w_dev, u, side))
            ref_idx = o.draw_libstdcxx(o.weights_from_log(A), u) if draw == "CUDA" else o.draw_numpy(o.weights_from_log(A), u)
            # identical to the draws made from the binary64 log-targets, except where a uniform sits on a cdf boundary closer than
            # the float32 log-target noise (1e-5 relative, the north-star tolerance): such near-ties are verified, then followed
            differ = np.nonzero(tr["draws"][it] != ref_idx)[0]
            if differ.size:
                cdf = np.cumsum(o.weights_from_log(A)); cdf /= cdf[-1]
                for t in differ:
                    lo, hi = sorted((int(tr["draws"][it][t]), int(ref_idx[t])))
                    assert hi - lo == 1 and abs(u[t] - cdf[lo]) < 1e-5, "iteration %d draw %d: %d vs %d is not a near-tie (margin %.3e)" % (it, t, tr["draws"][it][t], ref_idx[t], abs(u[t] - cdf[lo]))
                near_ties += differ.size
                ref_idx = tr["draws"][it].copy()
            nxt = ref_idx[o.pick_index(o.stream_uniforms(seed, it, o.STREAM_PICK, 0, 1)[0], P)] if draw == "PYTHON" else ref_idx[0]
            assert tr["next"][it] == nxt
            assert np.array_equal(tr["samples"][it], props[ref_idx])
            state = props[nxt]
            assert np.array_equal(tr["state"][it], state)
        assert np.array_equal(c.get_state(), state)
        assert near_ties <= 2, "%d near-tie draws in %d" % (near_ties, iters * P)
    finally:
        c.close()


@pytest.mark.parametrize("tc", [1, 0])
@pytest.mark.parametrize("n,P", [(500, 4), (500, 1024), (64, 16), (100000, 4), (5000, 3000), (100000, 1024), (20000, 300)])
def test_loop_structures_agree(n, P, tc, monkeypatch):
    """The persistent cooperative kernel and the CUDA-graph replay of sweep + acceptance are two schedules of the same
    arithmetic: identical states, accepted indices and log-weight bits after the same number of iterations — also when
    there are fewer (tile, chunk) units than sweep CTAs (n = 500: 8 chunks) or more nodes than one CTA's tile."""
    import pmp_mcmc_b200 as pm
    L = _L()
    x, y = synthetic_linear(n, seed=9)
    out = []
    monkeypatch.setenv("PMP_SWEEP_TC", str(tc)); monkeypatch.setenv("PMP_PERSISTENT_TC", str(tc))      # same sweep arithmetic in both loops
    for persistent in (1, 0):
        monkeypatch.setenv("PMP_PERSISTENT", str(persistent))
        c = pm.Context(0)
        try:
            c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=max(n / 50.0, 1.0))
            c.set_data_linear(x, y); c.set_state([1, 1, 1]); c.seed(21, 0)
            c.trace_config(70, L.TRACE_STATE | L.TRACE_NEXT | L.TRACE_LOGW)
            c.run(70)
            tr = c.read_trace()
            out.append((c.get_state().copy(), tr["next"].copy(), tr["logw"].copy(), tr["state"].copy()))
        finally:
            c.close()
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2].view(np.uint64), out[1][2].view(np.uint64))
    assert np.array_equal(out[0][3].view(np.uint32), out[1][3].view(np.uint32))
    assert np.array_equal(out[0][0].view(np.uint32), out[1][0].view(np.uint32))
