"""CPU tests: the oracle against known answers and golden vectors, the host-side stream, and the ABI surface."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, synthetic_linear
from oracle import oracle as o


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    assert o.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert o.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert o.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_normal_quantile_against_scipy():
    from scipy.stats import norm
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.random(20000), 10.0 ** -np.arange(1, 16), 1 - 10.0 ** -np.arange(1, 16)])
    z = np.array([o.lib().oracle_norm_ppf(float(v)) for v in u])
    np.testing.assert_allclose(z, norm.ppf(u), rtol=5e-15, atol=5e-15)
    p = np.concatenate([rng.random(5000) * 0.1 + 1e-17, 2.0 ** -np.arange(1, 53)])
    lg = np.array([o.lib().oracle_det_log(float(v)) for v in p])
    np.testing.assert_allclose(lg, np.log(p), rtol=4e-16)


def test_product_host_stream_equals_oracle_stream():
    """csrc/philox.cuh (host path, through the C-ABI) and oracle/pmp_oracle.c are independent restatements: same bits."""
    from pmp_mcmc_b200 import _lib as L
    for seed, it, stream in [(0, 0, 0), (123456789123, 77, 1), (2 ** 63 + 5, 2 ** 40 + 1, 3)]:
        assert np.array_equal(L.stream_uniforms(seed, it, stream, 5, 999), o.stream_uniforms(seed, it, stream, 5, 999))
        assert np.array_equal(L.stream_normals(seed, it, stream, 0, 20001), o.stream_normals(seed, it, stream, 0, 20001))
    z = o.stream_normals(9, 0, 0, 0, 400000)
    assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3 and abs((z ** 4).mean() - 3) < 5e-2


@pytest.mark.parametrize("tree,b,depth", [(0, 8, 1), (1, 2, 4), (2, 4, 3)])
def test_tree_structure(tree, b, depth):
    """Every non-root node is its parent plus one increment, parents as in lb.py:268-272 / 356-360."""
    dim, alpha = 2, 0.5
    st = np.array([0.25, -1.5], np.float32)
    p = o.propose(tree, b, depth, dim, alpha, st, 3, 4)
    P = o.num_nodes(tree, b, depth)
    bb = 2 if tree == 1 else b
    for node in range(1, P):
        parent = 0
        if tree != 0:
            s = 1
            while node >= s * bb:
                s *= bb
            parent = node % s
        z = o.stream_normals(3, 4, 0, node * dim, dim).astype(np.float32)
        assert np.array_equal(p[node], (p[parent] + (np.float32(alpha) * z).astype(np.float32)).astype(np.float32))


def test_loglik_restatements_agree():
    n = 20000
    x, y = synthetic_linear(n, seed=4)
    rng = np.random.default_rng(1)
    nets = (np.array([-1, 2, 0.5], np.float32) + 0.05 * rng.standard_normal((12, 3))).astype(np.float32)
    f64 = o.loglik_linear_f64(x, y, nets, 1000.0)
    np.testing.assert_allclose(o.loglik_linear_suffstat(x, y, nets, 1000.0), f64, rtol=1e-7)          # closed form
    np.testing.assert_allclose(o.loglik_linear_refcuda(x, y, nets, 1000.0), f64, rtol=2e-5)           # reference kernel arithmetic
    acc = o.sumsq_fixed_mirror(x, y, nets, n // 64 + 2)
    np.testing.assert_allclose(o.loglik_linear_from_fixed(acc, nets, n, 1000.0), f64, rtol=1e-7)      # the device's evaluation order
    np.testing.assert_allclose(o.loglik_lb_torch(x, y, nets), o.loglik_linear_f64(x, y, nets, n / 50.0), rtol=5e-6)


def test_fixed_point_sums_are_shard_invariant():
    """Integer partial sums of CHUNK-aligned shards add up to the single-shard sum exactly (what the NCCL all-reduce relies on)."""
    from pmp_mcmc_b200 import dist
    n = 10007
    x, y = synthetic_linear(n, seed=8)
    nets = np.array([[-1, 2, 0.5], [0, 0, 1], [1, 1, 1]], np.float32)
    whole = o.sumsq_fixed_mirror(x, y, nets, n // 64 + 9)
    for world in (2, 3, 8):
        tot = np.zeros(3, dtype=np.uint64)
        for r in range(world):
            lo, hi = dist.shard_bounds(n, world, r)
            assert lo % 64 == 0
            tot += o.sumsq_fixed_mirror(x[lo:hi], y[lo:hi], nets, n // 64 + 9)
        assert np.array_equal(tot, whole)


def test_draw_semantics_against_numpy_choice_and_pandas():
    """draw_numpy == RandomState.choice(p=) == pandas DataFrame.sample(weights=) on the same global state (lb.py:154-156)."""
    import pandas as pd
    w = np.random.default_rng(2).random(37) ** 4
    for seed in range(5):
        np.random.seed(seed)
        u = np.random.random_sample(37)
        np.random.seed(seed)
        ref = np.random.choice(37, size=37, replace=True, p=w / w.sum())
        assert np.array_equal(o.draw_numpy(w, u), ref)
        np.random.seed(seed)
        index = pd.DataFrame(np.linspace(0, 36, 37).astype(np.int32))
        got = index.sample(37, replace=True, weights=pd.DataFrame(w)[0]).values.reshape(-1)
        assert np.array_equal(o.draw_numpy(w, u), got)
        # the device's association (blocked scan, unnormalised compare) gives the same indices
        assert np.array_equal(o.draw_blocked(w, u, "right"), ref)


def test_draw_semantics_against_libstdcxx(tmp_path):
    """draw_libstdcxx == std::discrete_distribution fed the same 53-bit uniforms (500_MP.cu:218-222)."""
    import subprocess
    src = tmp_path / "dd.cpp"
    src.write_text(r'''
#include <random>
#include <cstdio>
#include <cstdint>
#include <vector>
struct Feed { typedef uint32_t result_type; std::vector<uint32_t> v; size_t i = 0;
  static constexpr result_type min() { return 0; } static constexpr result_type max() { return 0xffffffffu; }
  result_type operator()() { return v[i++]; } };
int main() { int n; if (scanf("%d", &n) != 1) return 1; std::vector<double> w(n); for (auto& x : w) if (scanf("%lf", &x) != 1) return 1;
  int m; if (scanf("%d", &m) != 1) return 1; Feed f; for (int k = 0; k < 2 * m; ++k) { unsigned u; if (scanf("%u", &u) != 1) return 1; f.v.push_back(u); }
  std::discrete_distribution<> d(w.begin(), w.end()); for (int k = 0; k < m; ++k) printf("%d\n", d(f)); }
''')
    exe = tmp_path / "dd"
    subprocess.run(["g++", "-O1", "-o", str(exe), str(src)], check=True)
    rng = np.random.default_rng(5)
    w = rng.random(29) ** 3
    words = rng.integers(0, 2 ** 32, size=(200, 2), dtype=np.uint64)
    u = (words[:, 0].astype(np.float64) + words[:, 1].astype(np.float64) * 4294967296.0) / 18446744073709551616.0   # generate_canonical<double,53>
    u = np.minimum(u, np.nextafter(1.0, 0.0))
    inp = "%d\n%s\n%d\n%s\n" % (len(w), " ".join(repr(float(v)) for v in w), len(u), " ".join("%d %d" % (a, b) for a, b in words))
    out = subprocess.run([str(exe)], input=inp, capture_output=True, text=True, check=True).stdout.split()
    assert np.array_equal(np.array(out, dtype=np.int32), o.draw_libstdcxx(w, u))


def test_abi_header_and_library_agree():
    """Every function include/pmp_b200.h declares is exported by libpmp_b200.so and bound in _lib.EXPORTS (no compute calls)."""
    from pmp_mcmc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "pmp_b200.h")).read()
    declared = set(re.findall(r"\b(pmp_[a-z0-9_]+)\s*\(", hdr)) - {"pmp_ctx", "pmp_config"}
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.pmp_abi_version() == 1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pmp_mcmc_b200 as pm
    with pytest.raises(pm.PmpError, match="no CPU path"):
        pm.Context(0)


def test_null_arguments_are_errors_not_crashes():
    """Argument validation comes before any CUDA call: NULL contexts / pointers return a negative status and a message, with or without a GPU."""
    import ctypes
    from pmp_mcmc_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(_lib.TREE_FLAT, 4, 1, 3, _lib.TARGET_LINEAR_GAUSS, _lib.ALGO_MP, _lib.DRAW_PYTHON, 0, 0.01, 1.0, 1.0, 0.0, 1.0, 1.0)
    calls = [lambda: lib.pmp_create(None, 0, 1, 0, None),
             lambda: lib.pmp_create(ctypes.byref(ctypes.c_void_p()), 0, 0, 0, None),
             lambda: lib.pmp_create(ctypes.byref(ctypes.c_void_p()), 0, 2, 5, None),
             lambda: lib.pmp_configure(None, ctypes.byref(cfg)),
             lambda: lib.pmp_set_data_linear(None, None, None, 0, 0, 0),
             lambda: lib.pmp_set_state(None, None, 3),
             lambda: lib.pmp_seed(None, 1, 0),
             lambda: lib.pmp_propose(None),
             lambda: lib.pmp_loglik(None, None),
             lambda: lib.pmp_accept(None, None, 0, None, None),
             lambda: lib.pmp_run(None, 10, 1),
             lambda: lib.pmp_run_multi(None, 1, 10, 1),
             lambda: lib.pmp_trace_config(None, 10, 1),
             lambda: lib.pmp_read_trace(None, 10, None, None, None, None, None, None),
             lambda: lib.pmp_hmc_accept(None, 0, 2, None, None, None, 0.5, 1.0, None, None),
             lambda: lib.pmp_nccl_unique_id(None)]
    for i, call in enumerate(calls):
        rc = call()
        assert rc < 0, (i, rc)
        assert len(lib.pmp_last_error()) > 0, i
    assert lib.pmp_destroy(None) in (0, -1, -2, -3, -4, -5)          # destroying nothing is not a crash either


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "pmp-mcmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle mirror", "").replace("the oracle", "").replace("oracle_blocked_cdf", "").replace("oracle/", "").replace("oracle.", "") or f.endswith((".cu", ".cuh")), f
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_glm_oracle_against_scipy():
    """oracle.loglik_glm_f64 (checker of the d-dimensional heads): logistic vs scipy's expit / log, Gaussian vs scipy.stats.norm."""
    from scipy import stats
    from scipy.special import expit
    from oracle import oracle as o
    rng = np.random.default_rng(3)
    X = rng.standard_normal((400, 7)); yb = (rng.uniform(size=400) < 0.4).astype(float); yr = rng.standard_normal(400)
    th = rng.standard_normal((5, 7)) * 0.5
    ll = o.loglik_glm_f64(X, yb, th, "logistic", scale=8.0)
    ref = np.array([np.sum(np.where(yb > 0.5, np.log(expit(X @ t)), np.log(expit(-(X @ t))))) for t in th]) / 8.0
    np.testing.assert_allclose(ll, ref, rtol=1e-12)
    thg = np.concatenate([th, rng.uniform(0.5, 2, (5, 1))], axis=1)
    lg = o.loglik_glm_f64(X, yr, thg, "gauss", scale=8.0)
    refg = np.array([stats.norm(X @ t[:-1], t[-1]).logpdf(yr).sum() for t in thg]) / 8.0
    np.testing.assert_allclose(lg, refg, rtol=1e-12)


def test_binary_tree_ancestors_split_by_tile():
    """The index arithmetic behind the state-only hand-off for binary trees (chain_persistent.cuh): node 128 t + i is the state plus the increments
    of its ancestors in creation order (accept.cuh for_each_ancestor: anc = node & (2^(l+1) - 1) for every set bit l, ascending), and that list is
    [low-bit prefixes of i, in tile 0] + [element i of the tiles whose index is a low-bit prefix of t, t itself last] — so a sweep CTA needs the
    normals of 1 + popcount(t) tiles."""
    PT = 128
    for depth in (1, 3, 7, 8, 10, 11):
        P = 1 << depth
        for node in range(P):
            want = [node & ((2 << l) - 1) for l in range(depth) if (node >> l) & 1]
            t, i = divmod(node, PT)
            got = [i & ((2 << l) - 1) for l in range(7) if (i >> l) & 1]                      # tile 0, list entry 0
            tiles = [0] + [t & ((2 << l) - 1) for l in range(4) if (t >> l) & 1]              # the CTA's tile list (kernel: tiles[])
            got += [tiles[k] * PT + i for k in range(1, len(tiles))]                          # element i of list entry k
            assert got == want, (depth, node, got, want)
            assert len(tiles) == 1 + bin(t).count("1") <= 5
            assert tiles[-1] == t
