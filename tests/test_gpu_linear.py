"""GPU parity tests of the linear-Gaussian hot path, through the C-ABI (pmp_mcmc_b200._lib.Context)."""
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


@pytest.mark.parametrize("tree,b,depth,dim", [(0, 4, 1, 3), (0, 1024, 1, 3), (1, 2, 3, 3), (1, 2, 10, 3), (2, 4, 2, 2), (2, 8, 3, 3),
                                              (2, 6, 4, 1), (1, 2, 5, 160), (0, 7, 1, 37)])
def test_proposals_bit_exact(ctx, tree, b, depth, dim):
    """Philox stream → normals → tree: identical bits to the CPU restatement (integer/bit work: exact)."""
    L, o = _L(), _o()
    tgt = L.TARGET_LINEAR_GAUSS if dim == 3 else L.TARGET_EXTERNAL
    ctx.configure(tree, b=b, depth=depth, dim=dim, target=tgt, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=0.37, scale=1.0,
                  flags=L.FLAG_NO_KERNEL_TERM)
    state = np.linspace(-1.5, 2.0, dim).astype(np.float32)
    for seed, it in [(0, 0), (0xDEADBEEFCAFE, 5), (7, (1 << 40) + 3)]:
        ctx.set_state(state)
        ctx.seed(seed, it)
        ctx.propose()
        got = ctx.read_proposals()
        ref = o.propose(tree, b, depth, dim, 0.37, state, seed, it)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
        assert np.array_equal(got[0], state)


@pytest.mark.parametrize("n,P,scale", [(500, 4, 10.0), (500, 1024, 10.0), (100000, 4, 1000.0), (100000, 1024, 1000.0), (100000, 512, 2000.0),
                                       (1, 4, 1.0), (63, 16, 1.0), (65, 1000, 1.0), (4097, 64, 5.0), (100003, 1296, 2000.0)])
@pytest.mark.parametrize("impl", ["tc", "fma"])
def test_loglik_parity(ctx, n, P, scale, impl, monkeypatch):
    """The P x n sweep, both implementations: `fma` = CUDA-core FFMA2 sweep (the default, bit-mirrored by the oracle),
    `tc` = tcgen05 residual GEMM + square-accumulate epilogue (PMP_SWEEP_TC=1)."""
    L, o = _L(), _o()
    monkeypatch.setenv("PMP_SWEEP_TC", "1" if impl == "tc" else "0")
    x, y = synthetic_linear(n, seed=n)
    ctx.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=scale)
    ctx.set_data_linear(x, y)
    ctx.set_state([-0.9, 1.9, 0.6])
    ctx.seed(n * 131 + P, 0)
    ctx.propose()
    props = ctx.read_proposals()
    lt = ctx.loglik()
    # (1) the integer sums are exactly the chunked mirror's (order/grid independent); only the final log() can differ in the last ulp
    mirror = o.loglik_linear_from_fixed(o.sumsq_fixed_mirror(x, y, props, (n + 63) // 64 + 1), props, n, scale)
    if impl == "fma":
        np.testing.assert_allclose(lt, mirror, rtol=1e-13, atol=0)
    else:   # bf16x3 operands are exact; what differs is the tensor core's fp32 accumulation of the 15 products (measured <= 2e-7)
        np.testing.assert_allclose(lt, mirror, rtol=5e-7, atol=0)
    # (2) ground truth (same float32 per-point arithmetic, exact sum): well inside the 1e-5 contract
    np.testing.assert_allclose(lt, o.loglik_linear_f64(x, y, props, scale), rtol=1e-6)
    # (3) the reference CUDA kernel's own arithmetic (serial float32 running sum, 500_MP.cu:16-20): its rounding noise
    #     alone is ~1e-5 relative at n=1e5 (measured 1.3e-5 against the exact sum), so the bound is 5e-5 there
    ref = o.loglik_linear_refcuda(x, y, props[: min(P, 64)], scale)
    np.testing.assert_allclose(lt[: min(P, 64)], ref, rtol=5e-5 if n > 10000 else 1e-5)
    # calling it again gives the same bits (acc was zeroed by the finalise step)
    assert np.array_equal(lt, ctx.loglik())


def test_loglik_vs_lb_python(ctx):
    """BayesNet.loglik (lb.py:103-108): sum log N(y; b0+b x, |sigma|) * 50 / n, torch float32 → 1e-5 relative."""
    L, o = _L(), _o()
    n, P = 100000, 64
    x, y = synthetic_linear(n, seed=3)
    ctx.configure(L.TREE_BARY, b=8, depth=2, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_PMP, draw=L.DRAW_PYTHON, alpha=0.05, scale=n / 50.0)
    ctx.set_data_linear(x, y)
    ctx.set_state([0, 0, 1]); ctx.seed(42, 0); ctx.propose()
    props = ctx.read_proposals()
    lt = ctx.loglik()
    np.testing.assert_allclose(lt, o.loglik_lb_torch(x, y, props), rtol=1e-5)


def test_negative_sigma_and_hopeless_nodes(ctx):
    """sigma enters squared (CUDA) / through abs (lb.py:106); absurd proposals saturate to -inf instead of wrapping."""
    L, o = _L(), _o()
    x, y = synthetic_linear(5000, seed=9)
    ctx.configure(L.TREE_FLAT, b=4, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=100.0)
    ctx.set_data_linear(x, y)
    props = np.array([[-1, 2, 0.5], [-1, 2, -0.5], [1e6, 1e6, 1e-6], [-1, 2, 0.0]], dtype=np.float32)
    ctx.write_proposals(props)
    lt = ctx.loglik()
    assert lt[0] == lt[1] and np.isfinite(lt[0])
    assert lt[2] == -np.inf and lt[3] == -np.inf
    ctx.set_state(props[0])
    idx, nxt = ctx.accept()
    assert set(idx.tolist()) <= {0, 1}


def _ref_lib(name):
    import ctypes, os
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", name)
    if not os.path.exists(p):
        pytest.skip("oracle/_ref not built (reference tree was not mounted at build time)")
    lib = ctypes.CDLL(p)
    fp = ctypes.POINTER(ctypes.c_float)
    lib.ref_set_data.argtypes = [fp, fp, ctypes.c_int]
    lib.ref_loglik.argtypes = [fp, ctypes.c_int, fp, ctypes.c_int, fp]
    lib.ref_loglik_ex.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, fp]
    return lib


@pytest.mark.parametrize("libname,n,scale,P,algo_flags", [("libref_mp_500.so", 500, 10.0, 64, "mp"), ("libref_mp_100000.so", 100000, 1000.0, 256, "mp"),
                                                          ("libref_pmp_500.so", 500, 10.0, 64, "pmp"), ("libref_pmp_100000.so", 100000, 1000.0, 1024, "pmp")])
def test_against_reference_cuda_kernel(ctx, libname, n, scale, P, algo_flags):
    """The reference's own log_likelihood_kernel, compiled from its source (oracle/Makefile), on the same GPU and inputs:
    A[p] = loglik/SCALE + proposal-kernel terms (500_MP.cu:10-36 / 500_PMP.cu:10-33 as shipped, table bug included)."""
    import ctypes
    L, o = _L(), _o()
    ref = _ref_lib(libname)
    x, y = synthetic_linear(n, seed=11)
    fp = ctypes.POINTER(ctypes.c_float)
    assert ref.ref_set_data(x.ctypes.data_as(fp), y.ctypes.data_as(fp), n) == 0
    if algo_flags == "mp":
        ctx.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.01, scale=scale)
    else:
        D = int(np.log2(P))
        ctx.configure(L.TREE_BINARY, depth=D, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=0.01, scale=scale,
                      flags=L.FLAG_QUIRK_TABLE_CONST)
    ctx.set_data_linear(x, y)
    ctx.set_state([1, 1, 1]); ctx.seed(99, 0); ctx.propose()
    props = ctx.read_proposals()
    ctx.loglik(read=False)
    u = np.full(P, 0.5)
    ctx.accept(u)
    A = ctx.read_logweights()
    out = np.zeros(P, dtype=np.float32)
    ms = ctypes.c_float()
    assert ref.ref_loglik(props.ctypes.data_as(fp), P, out.ctypes.data_as(fp), 1, ctypes.byref(ms)) == 0
    # the reference accumulates in a float32 running sum (n + P*3 roundings): 5e-5 relative at n=1e5 (see test_loglik_parity)
    np.testing.assert_allclose(A, out.astype(np.float64), rtol=5e-5 if n > 10000 else 2e-5)


@pytest.mark.parametrize("libname,n,scale,b,D,fix", [
    ("libref_conv_pmp.so", 100000, 2000.0, 8, 3, 0),      # conv_pmp.cu as shipped: P = 512, N_step = 7, tree_deep = 3 (conv_pmp.cu:85-87), table upload bug included
    ("libref_conv_pmp.so", 100000, 2000.0, 8, 3, 1),      # the same kernel fed the whole table as floats: its transition arithmetic (conv_pmp.cu:22-33)
    ("libref_conv_pmp.so", 5000, 2000.0, 4, 2, 1),
    ("libref_conv_pmp.so", 5000, 2000.0, 3, 4, 1),
    ("libref_pmp_500.so", 500, 10.0, 2, 6, 1),            # binary table kernel (500_PMP.cu:23-30) with the table uploaded correctly
    ("libref_pmp_100000.so", 100000, 1000.0, 2, 10, 1),
])
def test_against_reference_tree_kernels(ctx, libname, n, scale, b, D, fix):
    """General-tree kernel of conv_pmp.cu:10-36 and the binary one of 500_PMP.cu:10-33, compiled from the reference's source, with the
    transition table rebuilt by oracle/ref_harness.cu (conv_pmp.cu:182-221) — as shipped (PMP_FLAG_QUIRK_TABLE_CONST) and with the table
    uploaded in full (the intended rule: PMP_ALGO_TABLE without the quirk flag)."""
    import ctypes
    L, o = _L(), _o()
    ref = _ref_lib(libname)
    P = b ** D
    x, y = synthetic_linear(n, seed=13)
    fp = ctypes.POINTER(ctypes.c_float)
    assert ref.ref_set_data(x.ctypes.data_as(fp), y.ctypes.data_as(fp), n) == 0
    tree = L.TREE_BINARY if b == 2 else L.TREE_BARY
    ctx.configure(tree, b=b, depth=D, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_TABLE, draw=L.DRAW_CUDA, alpha=0.02, scale=scale,
                  flags=0 if fix else L.FLAG_QUIRK_TABLE_CONST)
    ctx.set_data_linear(x, y)
    ctx.set_state([0, 0, 1] if b != 2 else [1, 1, 1]); ctx.seed(7, 0); ctx.propose()      # conv_pmp.cu:129 starts at (0, 0, 1)
    props = ctx.read_proposals()
    ctx.loglik(read=False)
    ctx.accept(np.full(P, 0.5))
    A = ctx.read_logweights()
    out = np.zeros(P, dtype=np.float32)
    ms = ctypes.c_float()
    assert ref.ref_loglik_ex(props.ctypes.data_as(fp), P, D, b - 1, fix, out.ctypes.data_as(fp), 1, ctypes.byref(ms)) == 0
    np.testing.assert_allclose(A, out.astype(np.float64), rtol=5e-5 if n > 10000 else 2e-5)
    # and the oracle's restatement of the same rule agrees with the reference kernel too
    lt = o.loglik_linear_f64(x, y, props, scale)
    np.testing.assert_allclose(o.table_logweights(lt, props.astype(np.float64), b, D, quirk_const=not fix), out.astype(np.float64), rtol=5e-5 if n > 10000 else 2e-5)


def test_against_reference_mh_and_conv_mp_kernels(ctx):
    """conv_mh.cu:10-26 (<<<1,1>>>, two hard-coded candidates, /2000) and conv_mp.cu's MP kernel (N = 7, /2000) compiled from source."""
    import ctypes
    L, o = _L(), _o()
    fp = ctypes.POINTER(ctypes.c_float)
    n = 100000
    x, y = synthetic_linear(n, seed=17)
    # MH: A[0], A[1] are the two log-likelihoods / 2000; the device MH rule accepts iff u < exp(A1 - A0) (conv_mh.cu:152)
    ref = _ref_lib("libref_conv_mh.so")
    assert ref.ref_set_data(x.ctypes.data_as(fp), y.ctypes.data_as(fp), n) == 0
    ctx.configure(L.TREE_FLAT, b=2, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MH, draw=L.DRAW_CUDA, alpha=0.02, scale=2000.0)
    ctx.set_data_linear(x, y)
    ctx.set_state([0, 0, 1]); ctx.seed(3, 0); ctx.propose()
    props = ctx.read_proposals()
    lt = ctx.loglik()
    out = np.zeros(2, dtype=np.float32)
    ms = ctypes.c_float()
    assert ref.ref_loglik_ex(props.ctypes.data_as(fp), 2, 1, 1, 0, out.ctypes.data_as(fp), 1, ctypes.byref(ms)) == 0
    np.testing.assert_allclose(lt, out.astype(np.float64), rtol=5e-5)
    ratio = np.exp(float(out[1]) - float(out[0]))
    for u in (0.999 * min(ratio, 1.0), min(1.0 - 1e-9, 1.001 * ratio)):
        ctx.set_state([0, 0, 1]); ctx.seed(3, 0); ctx.propose(); ctx.loglik(read=False)
        idx, nxt = ctx.accept(np.array([u]))
        assert nxt == (1 if u < np.exp(lt[1] - lt[0]) else 0)
    # MP, conv_mp.cu shape
    ref = _ref_lib("libref_conv_mp.so")
    assert ref.ref_set_data(x.ctypes.data_as(fp), y.ctypes.data_as(fp), n) == 0
    ctx.configure(L.TREE_FLAT, b=8, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_CUDA, alpha=0.02, scale=2000.0)
    ctx.set_data_linear(x, y)
    ctx.set_state([0, 0, 1]); ctx.seed(4, 0); ctx.propose()
    props = ctx.read_proposals()
    ctx.loglik(read=False); ctx.accept(np.full(8, 0.5))
    A = ctx.read_logweights()
    out = np.zeros(8, dtype=np.float32)
    assert ref.ref_loglik_ex(props.ctypes.data_as(fp), 8, 1, 1, 0, out.ctypes.data_as(fp), 1, ctypes.byref(ms)) == 0
    np.testing.assert_allclose(A, out.astype(np.float64), rtol=5e-5)


@pytest.mark.parametrize("dim,tree,b,depth", [(40000, 1, 2, 5), (40001, 1, 2, 5), (50001, 2, 3, 3), (40000, 1, 2, 6)])
def test_long_vector_tree_proposals_bit_exact(ctx, dim, tree, b, depth):
    """level-wise proposal kernel for long parameter vectors (FC / CNN sized): one shared Philox block per element pair, tail quantiles compacted per warp —
    the same bits as the oracle's per-element restatement, for even and odd dimensions (odd: the pairs straddle the node boundary)"""
    from oracle import oracle as o
    from pmp_mcmc_b200 import _lib as L
    rng = np.random.default_rng(dim)
    state = rng.standard_normal(dim).astype(np.float32)
    ctx.configure(tree, b=b, depth=depth, dim=dim, target=L.TARGET_EXTERNAL, algo=L.ALGO_TABLE, draw=L.DRAW_SINGLE, flags=L.FLAG_NO_KERNEL_TERM, alpha=3e-3, scale=1.0)
    ctx.set_state(state); ctx.seed(77, 5); ctx.propose()
    dev = ctx.read_proposals()
    ref = o.propose(tree, b, depth, dim, 3e-3, state, 77, 5)
    assert dev.shape == ref.shape and np.array_equal(dev.view(np.uint32), ref.view(np.uint32))
