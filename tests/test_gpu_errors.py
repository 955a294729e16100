"""Error behaviour of the C-ABI (include/pmp_b200.h): every misuse returns a negative status with a message (PmpError through the Python
binding) — nothing crashes, nothing falls back to another path — and the context stays usable afterwards.  The reference has no error handling
at all (raw cudaMalloc pointers, no checks); these are the contracts the drop-in adds."""
import ctypes

import numpy as np
import pytest

from conftest import synthetic_linear

pytestmark = pytest.mark.gpu


def _ctx():
    import pmp_mcmc_b200 as pm
    return pm.Context(0)


def _linear(c, L, P=8):
    c.configure(L.TREE_FLAT, b=P, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=0.01, scale=10.0)


BAD_CONFIGS = [
    (r"depth 14 out of range", dict(tree="BINARY", b=2, depth=14)),
    (r"depth 0 out of range", dict(tree="BINARY", b=2, depth=0)),
    ("must be >= 1", dict(tree="FLAT", b=0)),
    ("out of range", dict(tree="FLAT", b=8193)),
    ("out of range", dict(tree="BARY", b=8, depth=5)),
    ("dim must be", dict(tree="FLAT", b=4, dim=0)),
    ("unknown target", dict(tree="FLAT", b=4, target=99)),
    ("unknown algo", dict(tree="FLAT", b=4, algo=99)),
    ("unknown draw", dict(tree="FLAT", b=4, draw=99)),
    ("dim 3", dict(tree="FLAT", b=4, dim=2)),
    ("NORMAL1D has dim 1", dict(tree="FLAT", b=4, dim=3, target="NORMAL1D")),
    ("BANANA has dim 2", dict(tree="FLAT", b=4, dim=3, target="BANANA")),
    ("P == 2", dict(tree="FLAT", b=4, algo="MH")),
    ("BINARY tree", dict(tree="FLAT", b=4, algo="PSP")),
    ("BARY/BINARY", dict(tree="FLAT", b=4, algo="PMP")),
    ("scale must be non-zero", dict(tree="FLAT", b=4, scale=0.0)),
    ("kernel_sigma > 0", dict(tree="FLAT", b=4, kernel_sigma=0.0)),
]


@pytest.mark.parametrize("msg,kw", BAD_CONFIGS, ids=["%s-%d" % (m.replace(" ", "_"), i) for i, (m, _) in enumerate(BAD_CONFIGS)])
def test_configure_rejects(msg, kw):
    from pmp_mcmc_b200 import _lib as L
    c = _ctx()
    kw = dict(kw)
    tree = getattr(L, "TREE_" + kw.pop("tree"))
    for key, prefix in (("target", "TARGET_"), ("algo", "ALGO_")):
        if isinstance(kw.get(key), str):
            kw[key] = getattr(L, prefix + kw[key])
    kw.setdefault("dim", 3)
    with pytest.raises(L.PmpError, match=msg):
        c.configure(tree, **kw)
    _linear(c, L)                                  # the context is still good
    x, y = synthetic_linear(500)
    c.set_data_linear(x, y); c.set_state([0, 0, 1]); c.seed(1, 0)
    c.run(5)
    assert np.all(np.isfinite(c.get_state()))
    c.close()


def test_calls_in_the_wrong_order():
    from pmp_mcmc_b200 import _lib as L
    c = _ctx()
    with pytest.raises(L.PmpError, match="not configured"):
        c.run(10)
    with pytest.raises(L.PmpError, match="not configured"):
        c.set_state([0, 0, 1])
    with pytest.raises(L.PmpError, match="not configured"):
        c.propose()
    _linear(c, L)
    with pytest.raises(L.PmpError, match="data not set"):
        c.set_state([0, 0, 1]); c.propose(); c.loglik()
    x, y = synthetic_linear(300)
    c.set_data_linear(x, y)
    with pytest.raises(L.PmpError, match="pmp_accept before"):
        c.configure(L.TREE_FLAT, b=8, dim=3, target=L.TARGET_LINEAR_GAUSS, algo=L.ALGO_MP, draw=L.DRAW_PYTHON, alpha=0.01, scale=10.0)
        c.set_data_linear(x, y); c.set_state([0, 0, 1]); c.propose(); c.accept()
    c.loglik()
    with pytest.raises(L.PmpError, match="uniforms"):
        c.accept(uniforms=np.zeros(3))
    idx = c.accept()
    assert idx is not None
    with pytest.raises(L.PmpError, match="not traced"):
        c.trace_config(4, L.TRACE_STATE); c.run(4)
        st = np.zeros((4, 3), np.float32); dr = np.zeros((4, 8), np.int32); n = ctypes.c_int64()
        c._chk(c.L.pmp_read_trace(c.h, 4, st.ctypes.data_as(ctypes.c_void_p), None, dr.ctypes.data_as(ctypes.c_void_p), None, None, ctypes.byref(n)))
    with pytest.raises(L.PmpError, match="iters < 0"):
        c.run(-1)
    with pytest.raises(L.PmpError, match="max_iters < 0"):
        c.trace_config(-1, L.TRACE_STATE)
    c.close()


def test_shapes_and_counts_are_checked():
    from pmp_mcmc_b200 import _lib as L
    c = _ctx()
    _linear(c, L)
    with pytest.raises(L.PmpError, match="dim 2 != configured 3"):
        c.set_state([0, 1])
    with pytest.raises(L.PmpError, match=r"!= P\*dim"):
        c.write_proposals(np.zeros((7, 3), np.float32))
    with pytest.raises(L.PmpError, match="!= P"):
        c.write_logtarget(np.zeros(7))
    x, y = synthetic_linear(100)
    with pytest.raises(L.PmpError, match="bad shard"):
        c.set_data_linear(x, y, n_offset=50, n_global=120)
    with pytest.raises(L.PmpError, match="bad shard"):
        c.set_data_linear(x, y, n_offset=0, n_global=50)
    with pytest.raises(L.PmpError, match="x/y NULL"):
        c._chk(c.L.pmp_set_data_linear(c.h, None, None, 10, 0, 10))
    c.set_data_linear(x[:0], y[:0])                 # an empty dataset is a valid shard
    c.close()


def test_feature_specific_requirements():
    import torch
    from pmp_mcmc_b200 import _lib as L
    c = _ctx()
    _linear(c, L)
    with pytest.raises(L.PmpError, match="analytic targets only"):
        c.chains_create(128)
    x, y = synthetic_linear(300)
    c.set_data_linear(x, y); c.set_state([0, 0, 1]); c.seed(1, 0)
    with pytest.raises(L.PmpError, match="STATE trace"):
        c.trace_config(0, 0); c.run(8); c.trace_diagnostics()
    c.trace_config(8, L.TRACE_STATE)
    with pytest.raises(L.PmpError, match="at least two"):
        c.trace_diagnostics()
    with pytest.raises(L.PmpError):
        c.hmc_accept(L.HMC_RULE_TREE_BNN, np.zeros(2049), np.zeros(2049), np.zeros(2049))      # more nodes than the acceptance kernel holds
    with pytest.raises(L.PmpError):
        c.hmc_accept(99, np.zeros(4), np.zeros(4), np.zeros(4))
    t = torch.zeros(16, device="cuda:0")
    with pytest.raises(L.PmpError):
        c._chk(c.L.pmp_hmc_leapfrog_begin(c.h, None, t.data_ptr(), t.data_ptr(), t.data_ptr(), None, 16, 0.1, 1.0, 0.0005, 0, ctypes.byref(ctypes.c_double())))
    other = _ctx()
    _linear(other, L)
    with pytest.raises(L.PmpError, match="owns no linear-Gaussian data"):
        c.share_data_from(other)
    with pytest.raises(L.PmpError, match="given twice"):
        L.run_multi([c, c], 10)
    with pytest.raises(L.PmpError, match="iters out of range"):
        L.run_multi([c], 1)
    other.close(); c.close()
    c.close()                                       # closing twice is harmless


def test_device_out_of_range_and_bad_world():
    import pmp_mcmc_b200 as pm
    from pmp_mcmc_b200 import _lib as L
    with pytest.raises(L.PmpError, match="out of range"):
        pm.Context(4096)
    with pytest.raises(L.PmpError, match="bad world_size/rank"):
        pm.Context(0, world_size=2, rank=2)
    with pytest.raises(L.PmpError, match="NCCL unique id"):
        pm.Context(0, world_size=2, rank=0)
