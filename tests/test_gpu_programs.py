"""cuda_programs.py: the reference's .cu experiment programs on the device-resident chain, writing the reference's files."""
import os

import numpy as np
import pytest

from conftest import synthetic_linear

pytestmark = pytest.mark.gpu


def _solo_chain(ctx, L, cfg, x, y, theta0, seed, burn, steps, what):
    ctx.configure(**cfg)
    ctx.set_data_linear(x, y); ctx.set_state(np.asarray(theta0, np.float32)); ctx.seed(seed, 0)
    if burn:
        ctx.trace_config(0, 0); ctx.run(burn)
    ctx.trace_config(steps, what)
    ctx.run(steps)
    return ctx.read_trace()


@pytest.mark.parametrize("kind,N", [("MP", 3), ("MP", 255), ("PMP", 15)])
def test_time_analysis_program(ctx, tmp_path, kind, N):
    from pmp_mcmc_b200 import _lib as L, cuda_programs as cp, sinks
    x, y = synthetic_linear(500, seed=5)
    r = cp.time_analysis(kind, x=x, y=y, N=N, num_steps=60, brin_in=25, alpha=0.01, scale=10.0, theta0=(1, 1, 1), out_dir=str(tmp_path), seed=9, ctx=ctx)
    P = N + 1
    assert r["samples"].shape == (60, P, 3) and r["weights"].shape == (60, P)
    np.testing.assert_allclose(r["weights"].sum(1), 1.0, rtol=1e-12)
    # the same chain through the plain C-ABI calls: same bits
    if kind == "MP":
        cfg = dict(tree=L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=10.0)
    else:
        cfg = dict(tree=L.TREE_BINARY, depth=4, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, flags=L.FLAG_QUIRK_TABLE_CONST, alpha=0.01, scale=10.0)
    tr = _solo_chain(ctx, L, cfg, x, y, (1, 1, 1), 9, 25, 60, L.TRACE_SAMPLES | L.TRACE_NEXT)
    assert np.array_equal(tr["samples"], r["samples"])
    # 500_MP.cu:240-243: the next state is the first resampled candidate
    assert np.array_equal(r["state"], r["samples"][-1, 0])
    # files: the reference's names, six significant digits, iteration-major rows
    tag = "_MP" if kind == "MP" else "_"
    for col, j in (("beta0", 0), ("beta_true", 1), ("sigma_true", 2)):
        got = sinks.read_column(os.path.join(str(tmp_path), "%d%s%s.txt" % (P, tag, col)))
        np.testing.assert_allclose(got, r["samples"][:, :, j].reshape(-1), rtol=1e-5, atol=1e-12)
    np.testing.assert_allclose(sinks.read_column(os.path.join(str(tmp_path), "%d%sA.txt" % (P, tag))), r["weights"].reshape(-1).astype(np.float32), rtol=1e-5, atol=1e-30)
    assert float(open(os.path.join(str(tmp_path), "%d%stime.txt" % (P, tag))).readline()) > 0


@pytest.mark.parametrize("kind,kw", [("MH", {}), ("MP", dict(N=7)), ("PMP", dict(N=63, tree_deep=2, N_step=7))])
def test_convergence_program(ctx, tmp_path, kind, kw):
    from pmp_mcmc_b200 import cuda_programs as cp, sinks
    x, y = synthetic_linear(2000, seed=6)
    r = cp.convergence(kind, x=x, y=y, num_steps=80, alpha=0.02, scale=40.0, theta0=(0, 0, 1), out_dir=str(tmp_path), seed=3, ctx=ctx, **kw)
    assert r["states"].shape == (80, 3)
    pars, t_ms = sinks.load_conv_trace(str(tmp_path), kind, 80)            # the notebook's reader
    np.testing.assert_allclose(pars.T, r["states"], rtol=1e-5, atol=1e-12)
    assert np.all(np.diff(t_ms) > 0)
    # the chain moves towards the data-generating parameters (-1, 2, 0.5) from (0, 0, 1)
    d0 = np.linalg.norm(np.array([0, 0, 1.0]) - np.array([-1, 2, 0.5]))
    assert np.linalg.norm(r["states"][-1] - np.array([-1, 2, 0.5])) < d0
    if kind != "MH":
        assert r["weights"].shape[0] == 80 and os.path.exists(os.path.join(str(tmp_path), "%d_%sA.txt" % (kw["N"] + 1, kind)))


def test_convergence_with_cores_program(ctx, tmp_path):
    from pmp_mcmc_b200 import cuda_programs as cp
    x, y = synthetic_linear(1000, seed=7)
    r = cp.convergence_with_cores("MP", x=x, y=y, N=63, num_steps=300, set_time=60.0, scale=20.0, block=100, out_dir=str(tmp_path), seed=1, ctx=ctx)
    assert r["iterations"] == 300 and r["states"].shape == (300, 3) and len(r["times"]) == 300
    r2 = cp.convergence_with_cores("MP", x=x, y=y, N=63, num_steps=10 ** 7, set_time=0.2, scale=20.0, block=200, seed=1, ctx=ctx)
    assert 200 <= r2["iterations"] < 10 ** 7 and r2["iterations"] % 200 == 0          # stopped by the clock
