"""ctypes binding of libpmp_b200.so (include/pmp_b200.h).  No fallback: if the library is missing it is built with
nvcc; if it cannot be built or loaded, importing raises."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpmp_b200.so")

# enums of include/pmp_b200.h
TREE_FLAT, TREE_BINARY, TREE_BARY = 0, 1, 2
TARGET_LINEAR_GAUSS, TARGET_NORMAL1D, TARGET_BANANA, TARGET_STDNORMAL, TARGET_FC, TARGET_EXTERNAL, TARGET_GLM_LOGISTIC, TARGET_GLM_GAUSS, TARGET_CNN = range(9)
ALGO_MH, ALGO_BARKER, ALGO_MP, ALGO_PSP, ALGO_PMP, ALGO_TABLE = range(6)
DRAW_PYTHON, DRAW_CUDA, DRAW_SINGLE = range(3)
FLAG_QUIRK_LEVEL_MOD, FLAG_QUIRK_TABLE_CONST, FLAG_STANDARDIZE, FLAG_KERNEL_MEAN, FLAG_NO_KERNEL_TERM, FLAG_UNIFORM_PROPOSAL = 1, 2, 4, 8, 16, 32
HMC_RULE_SP, HMC_RULE_MP, HMC_RULE_TREE_CIFAR, HMC_RULE_TREE_BNN = range(4)
TRACE_STATE, TRACE_NEXT, TRACE_DRAWS, TRACE_SAMPLES, TRACE_LOGW = 1, 2, 4, 8, 16

EXPORTS = [
    "pmp_last_error", "pmp_abi_version", "pmp_create", "pmp_destroy", "pmp_nccl_unique_id", "pmp_device_info",
    "pmp_configure", "pmp_num_nodes", "pmp_set_data_linear", "pmp_set_state", "pmp_get_state", "pmp_seed",
    "pmp_get_iteration", "pmp_propose", "pmp_read_proposals", "pmp_write_proposals", "pmp_loglik", "pmp_write_logtarget",
    "pmp_accept", "pmp_read_logweights", "pmp_trace_config", "pmp_run", "pmp_sync", "pmp_read_trace", "pmp_trace_reset",
    "pmp_run_timed", "pmp_launch_count", "pmp_fp32_peak", "pmp_l2_flush", "pmp_chains_create", "pmp_chains_run",
    "pmp_chains_read_states", "pmp_chains_read_samples", "pmp_chains_run_timed", "pmp_set_data_fc", "pmp_stream_uniforms",
    "pmp_stream_normals", "pmp_time_sweep", "pmp_share_data", "pmp_run_multi", "pmp_run_multi_timed",
    "pmp_peer_exchange_handle", "pmp_peer_exchange_attach", "pmp_trace_diagnostics", "pmp_set_data_glm", "pmp_set_data_cnn",
    "pmp_hmc_leapfrog_begin", "pmp_hmc_leapfrog_end", "pmp_hmc_accept",
]


class Config(ctypes.Structure):
    _fields_ = [
        ("tree", ctypes.c_int32), ("b", ctypes.c_int32), ("depth", ctypes.c_int32), ("dim", ctypes.c_int32),
        ("target", ctypes.c_int32), ("algo", ctypes.c_int32), ("draw", ctypes.c_int32), ("flags", ctypes.c_uint32),
        ("alpha", ctypes.c_float), ("scale", ctypes.c_float), ("kernel_sigma", ctypes.c_float),
        ("target_p0", ctypes.c_float), ("target_p1", ctypes.c_float), ("mh_temperature", ctypes.c_float),
    ]


class PmpError(RuntimeError):
    pass


_lib = None


def build(force=False):
    """Compile libpmp_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", CSRC, "--no-print-directory", "clean"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", CSRC, "--no-print-directory"], check=True, stdout=subprocess.DEVNULL)


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "pmp_b200.h"))
    return any(os.path.getmtime(s) > t for s in srcs)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        build()
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    L.pmp_last_error.restype = ctypes.c_char_p
    L.pmp_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, vp]
    L.pmp_destroy.argtypes = [vp]
    L.pmp_nccl_unique_id.argtypes = [vp]
    L.pmp_device_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.c_char_p, i32]
    L.pmp_configure.argtypes = [vp, ctypes.POINTER(Config)]
    L.pmp_num_nodes.argtypes = [vp]
    L.pmp_set_data_linear.argtypes = [vp, vp, vp, i64, i64, i64]
    L.pmp_set_state.argtypes = [vp, vp, i32]
    L.pmp_get_state.argtypes = [vp, vp, i32]
    L.pmp_seed.argtypes = [vp, u64, u64]
    L.pmp_get_iteration.argtypes = [vp, ctypes.POINTER(u64)]
    L.pmp_propose.argtypes = [vp]
    L.pmp_read_proposals.argtypes = [vp, vp, i64]
    L.pmp_write_proposals.argtypes = [vp, vp, i64]
    L.pmp_loglik.argtypes = [vp, vp]
    L.pmp_write_logtarget.argtypes = [vp, vp, i64]
    L.pmp_accept.argtypes = [vp, vp, i64, vp, vp]
    L.pmp_read_logweights.argtypes = [vp, vp, i64]
    L.pmp_trace_config.argtypes = [vp, i64, u32]
    L.pmp_run.argtypes = [vp, i64, i32]
    L.pmp_sync.argtypes = [vp]
    L.pmp_read_trace.argtypes = [vp, i64, vp, vp, vp, vp, vp, ctypes.POINTER(i64)]
    L.pmp_trace_reset.argtypes = [vp]
    L.pmp_run_timed.argtypes = [vp, i64, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    L.pmp_launch_count.argtypes = [vp, ctypes.POINTER(i64)]
    L.pmp_time_sweep.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_float)]
    L.pmp_fp32_peak.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    L.pmp_l2_flush.argtypes = [vp]
    L.pmp_chains_create.argtypes = [vp, i64, vp]
    L.pmp_chains_run.argtypes = [vp, i64, i32]
    L.pmp_chains_read_states.argtypes = [vp, vp]
    L.pmp_chains_read_samples.argtypes = [vp, vp, i64]
    L.pmp_chains_run_timed.argtypes = [vp, i64, i32, ctypes.POINTER(ctypes.c_float)]
    L.pmp_set_data_fc.argtypes = [vp, vp, vp, i64, i64, i64]
    L.pmp_set_data_cnn.argtypes = [vp, vp, vp, i64, i64, i64]
    L.pmp_hmc_leapfrog_begin.argtypes = [vp, vp, vp, vp, vp, vp, i64, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint64, vp]
    L.pmp_hmc_leapfrog_end.argtypes = [vp, vp, vp, i64, ctypes.c_float, ctypes.c_float, vp]
    L.pmp_hmc_accept.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_double, ctypes.c_double, vp, vp]
    L.pmp_stream_uniforms.argtypes = [u64, u64, u32, u64, i64, vp]
    L.pmp_stream_normals.argtypes = [u64, u64, u32, u64, i64, vp]
    L.pmp_share_data.argtypes = [vp, vp]
    L.pmp_set_data_glm.argtypes = [vp, vp, vp, i64, i64, i64, i32]
    L.pmp_trace_diagnostics.argtypes = [vp, i32, vp, vp, vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    L.pmp_peer_exchange_handle.argtypes = [vp, vp]
    L.pmp_peer_exchange_attach.argtypes = [vp, vp, i32]
    L.pmp_run_multi.argtypes = [ctypes.POINTER(vp), i32, i64, i32]
    L.pmp_run_multi_timed.argtypes = [ctypes.POINTER(vp), i32, i64, ctypes.POINTER(ctypes.c_float)]
    if L.pmp_abi_version() != 1:
        raise PmpError("libpmp_b200.so ABI version mismatch")
    _lib = L
    return L


def _handles(ctxs):
    arr = (ctypes.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    return arr


def run_multi(ctxs, iters, sync=True):
    """Co-scheduled independent chains (pmp_run_multi): one cooperative kernel sweeps chain B while chain A is accepted."""
    ctxs[0]._chk(ctxs[0].L.pmp_run_multi(_handles(ctxs), len(ctxs), iters, 1 if sync else 0))


def run_multi_timed(ctxs, iters):
    ms = ctypes.c_float()
    ctxs[0]._chk(ctxs[0].L.pmp_run_multi_timed(_handles(ctxs), len(ctxs), iters, ctypes.byref(ms)))
    return ms.value


def stream_uniforms(seed, iteration, stream, idx0, count):
    """Host evaluation of the library's Philox stream (no GPU needed): `count` uniforms in [0,1)."""
    out = np.empty(count, dtype=np.float64)
    rc = load().pmp_stream_uniforms(seed, iteration, stream, idx0, count, ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise PmpError(load().pmp_last_error().decode())
    return out


def stream_normals(seed, iteration, stream, idx0, count):
    out = np.empty(count, dtype=np.float64)
    rc = load().pmp_stream_normals(seed, iteration, stream, idx0, count, ctypes.c_void_p(out.ctypes.data))
    if rc != 0:
        raise PmpError(load().pmp_last_error().decode())
    return out


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class Context:
    """One pmp_ctx: a device, a stream, the chain state and (world_size > 1) an NCCL communicator."""

    def __init__(self, device=0, world_size=1, rank=0, nccl_unique_id=None):
        self.L = load()
        self.h = ctypes.c_void_p()
        uid = ctypes.create_string_buffer(bytes(nccl_unique_id), 128) if nccl_unique_id is not None else None
        self._chk(self.L.pmp_create(ctypes.byref(self.h), device, world_size, rank, uid))
        self.world_size, self.rank, self.device = world_size, rank, device
        self.peers_attached = False
        self.cfg = None
        self.P = 0

    def _chk(self, rc):
        if rc != 0:
            raise PmpError("libpmp_b200: %s (status %d)" % (self.L.pmp_last_error().decode(), rc))

    @staticmethod
    def nccl_unique_id():
        L = load()
        buf = ctypes.create_string_buffer(128)
        rc = L.pmp_nccl_unique_id(buf)
        if rc != 0:
            raise PmpError("libpmp_b200: %s (status %d)" % (L.pmp_last_error().decode(), rc))
        return buf.raw

    def close(self):
        if self.h:
            self.L.pmp_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device_info(self):
        sm, ma, mi = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        name = ctypes.create_string_buffer(256)
        self._chk(self.L.pmp_device_info(self.h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), name, 256))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "name": name.value.decode()}

    def configure(self, tree, b=2, depth=1, dim=3, target=TARGET_LINEAR_GAUSS, algo=ALGO_MP, draw=DRAW_PYTHON, flags=0,
                  alpha=0.01, scale=1.0, kernel_sigma=1.0, target_p0=0.0, target_p1=1.0, mh_temperature=1.0):
        cfg = Config(tree, b, depth, dim, target, algo, draw, flags, alpha, scale, kernel_sigma, target_p0, target_p1,
                     mh_temperature)
        self._chk(self.L.pmp_configure(self.h, ctypes.byref(cfg)))
        self.cfg = cfg
        self.P = self.L.pmp_num_nodes(self.h)
        return self

    def set_data_linear(self, x, y, n_offset=0, n_global=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        assert x.shape == y.shape and x.ndim == 1
        self._chk(self.L.pmp_set_data_linear(self.h, _ptr(x), _ptr(y), x.size, n_offset, x.size if n_global is None else n_global))

    def set_state(self, theta):
        th = np.ascontiguousarray(theta, dtype=np.float32).reshape(-1)
        self._chk(self.L.pmp_set_state(self.h, _ptr(th), th.size))

    def get_state(self):
        out = np.empty(self.cfg.dim, dtype=np.float32)
        self._chk(self.L.pmp_get_state(self.h, _ptr(out), out.size))
        return out

    def seed(self, seed, iteration=0):
        self._chk(self.L.pmp_seed(self.h, seed, iteration))

    def iteration(self):
        it = ctypes.c_uint64()
        self._chk(self.L.pmp_get_iteration(self.h, ctypes.byref(it)))
        return it.value

    def propose(self):
        self._chk(self.L.pmp_propose(self.h))

    def read_proposals(self):
        out = np.empty((self.P, self.cfg.dim), dtype=np.float32)
        self._chk(self.L.pmp_read_proposals(self.h, _ptr(out), out.size))
        return out

    def write_proposals(self, props):
        p = np.ascontiguousarray(props, dtype=np.float32)
        self._chk(self.L.pmp_write_proposals(self.h, _ptr(p), p.size))

    def loglik(self, read=True):
        out = np.empty(self.P, dtype=np.float64) if read else None
        self._chk(self.L.pmp_loglik(self.h, _ptr(out)))
        return out

    def write_logtarget(self, lt):
        a = np.ascontiguousarray(lt, dtype=np.float64)
        self._chk(self.L.pmp_write_logtarget(self.h, _ptr(a), a.size))

    def n_draws(self):
        return 1 if self.cfg.algo in (ALGO_MH, ALGO_BARKER) or self.cfg.draw == DRAW_SINGLE else self.P

    def accept(self, uniforms=None, read=True):
        u = np.ascontiguousarray(uniforms, dtype=np.float64) if uniforms is not None else None
        idx = np.empty(self.n_draws(), dtype=np.int32) if read else None
        nxt = ctypes.c_int32(-1)
        self._chk(self.L.pmp_accept(self.h, _ptr(u), u.size if u is not None else 0, _ptr(idx),
                                    ctypes.byref(nxt) if read else None))
        return idx, nxt.value

    def read_logweights(self):
        out = np.empty(self.P, dtype=np.float64)
        self._chk(self.L.pmp_read_logweights(self.h, _ptr(out), out.size))
        return out

    def trace_config(self, max_iters, what):
        self._chk(self.L.pmp_trace_config(self.h, max_iters, what))
        self._trace_what, self._trace_cap = what, max_iters

    def trace_reset(self):
        self._chk(self.L.pmp_trace_reset(self.h))

    def run(self, iters, sync=True):
        self._chk(self.L.pmp_run(self.h, iters, 1 if sync else 0))

    def sync(self):
        self._chk(self.L.pmp_sync(self.h))

    def peer_exchange_handle(self):
        """64-byte CUDA IPC handle of this rank's exchange buffer (world_size > 1)."""
        buf = ctypes.create_string_buffer(64)
        self._chk(self.L.pmp_peer_exchange_handle(self.h, buf))
        return buf.raw

    def peer_exchange_attach(self, handles):
        """handles: the 64-byte handles of all ranks, rank-major (bytes of length 64 * world_size)."""
        self._chk(self.L.pmp_peer_exchange_attach(self.h, ctypes.c_char_p(bytes(handles)), len(handles) // 64))
        self.peers_attached = True

    def share_data_from(self, owner):
        """Alias `owner`'s device copy of the linear-Gaussian data (no copy); `owner` must stay alive while this ctx uses it."""
        self._chk(self.L.pmp_share_data(self.h, owner.h))
        self._data_owner = owner

    def trace_buffers(self, max_iters=None, pinned=False):
        """Host arrays shaped for read_trace(out=...); pinned=True allocates them page-locked through torch."""
        n = self._trace_cap if max_iters is None else min(max_iters, self._trace_cap)
        w, P, d = self._trace_what, self.P, self.cfg.dim

        def mk(shape, dt):
            if pinned:
                import torch
                return torch.empty(shape, dtype={np.float32: torch.float32, np.int32: torch.int32, np.float64: torch.float64}[dt]).pin_memory().numpy()
            return np.empty(shape, dt)
        return {"state": mk((n, d), np.float32) if w & TRACE_STATE else None, "next": mk((n,), np.int32) if w & TRACE_NEXT else None,
                "draws": mk((n, P), np.int32) if w & TRACE_DRAWS else None, "samples": mk((n, P, d), np.float32) if w & TRACE_SAMPLES else None,
                "logw": mk((n, P), np.float64) if w & TRACE_LOGW else None}

    def read_trace(self, max_iters=None, out=None):
        n = self._trace_cap if max_iters is None else min(max_iters, self._trace_cap)
        b = out if out is not None else self.trace_buffers(max_iters)
        state, nxt, draws, samples, logw = b["state"], b["next"], b["draws"], b["samples"], b["logw"]
        rec = ctypes.c_int64()
        self._chk(self.L.pmp_read_trace(self.h, n, _ptr(state), _ptr(nxt), _ptr(draws), _ptr(samples), _ptr(logw), ctypes.byref(rec)))
        r = rec.value
        cut = lambda a: a[:r] if a is not None else None
        return {"n": r, "state": cut(state), "next": cut(nxt), "draws": cut(draws), "samples": cut(samples), "logw": cut(logw)}

    def trace_diagnostics(self, max_lag=256):
        """Device reduction of the STATE trace: mean, variance, autocovariances up to max_lag, MSJD, move rate."""
        d = self.cfg.dim
        mean, var = np.empty(d), np.empty(d)
        acov = np.zeros((max_lag + 1, d))
        msjd, mv, rows = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        self._chk(self.L.pmp_trace_diagnostics(self.h, max_lag, _ptr(mean), _ptr(var), _ptr(acov), ctypes.byref(msjd), ctypes.byref(mv), ctypes.byref(rows)))
        return {"n": rows.value, "mean": mean, "var": var, "acov": acov[: min(max_lag, rows.value - 1) + 1], "msjd": msjd.value, "move_rate": mv.value}

    def run_timed(self, iters, sweep=False):
        total, sw = ctypes.c_float(), ctypes.c_float()
        self._chk(self.L.pmp_run_timed(self.h, iters, ctypes.byref(total), ctypes.byref(sw) if sweep else None))
        return total.value, (sw.value if sweep else None)

    def time_sweep(self, reps):
        ms = ctypes.c_float()
        self._chk(self.L.pmp_time_sweep(self.h, reps, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        n = ctypes.c_int64()
        self._chk(self.L.pmp_launch_count(self.h, ctypes.byref(n)))
        return n.value

    def fp32_peak(self, packed=True):
        t = ctypes.c_double()
        self._chk(self.L.pmp_fp32_peak(self.h, 1 if packed else 0, ctypes.byref(t)))
        return t.value

    def l2_flush(self):
        self._chk(self.L.pmp_l2_flush(self.h))

    # batched analytic chains
    def chains_create(self, n_chains, init_states=None):
        a = np.ascontiguousarray(init_states, dtype=np.float32) if init_states is not None else None
        self._chk(self.L.pmp_chains_create(self.h, n_chains, _ptr(a)))
        self.n_chains = n_chains

    def chains_run(self, iters, record_samples=False):
        self._chk(self.L.pmp_chains_run(self.h, iters, 1 if record_samples else 0))
        self._chain_iters = iters if record_samples else 0

    def chains_run_timed(self, iters, record_samples=False):
        ms = ctypes.c_float()
        self._chk(self.L.pmp_chains_run_timed(self.h, iters, 1 if record_samples else 0, ctypes.byref(ms)))
        self._chain_iters = iters if record_samples else 0
        return ms.value

    def chains_read_states(self):
        out = np.empty((self.n_chains, self.cfg.dim), dtype=np.float32)
        self._chk(self.L.pmp_chains_read_states(self.h, _ptr(out)))
        return out

    def chains_read_samples(self):
        out = np.empty((self._chain_iters, self.P, self.cfg.dim, self.n_chains), dtype=np.float32)
        self._chk(self.L.pmp_chains_read_samples(self.h, _ptr(out), out.size))
        return out

    def set_data_glm(self, X, y, n_offset=0, n_global=None):
        """X [n, d] float32, y [n] float32 ({0,1} labels for the logistic head, responses for the Gaussian head)."""
        X = np.ascontiguousarray(X, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32).reshape(-1)
        if X.ndim != 2 or len(X) != len(y):
            raise ValueError("X must be [n, d] and y [n]")
        self._chk(self.L.pmp_set_data_glm(self.h, _ptr(X), _ptr(y), len(y), n_offset, len(y) if n_global is None else n_global, X.shape[1]))

    # ---- HMC variants (hmc.cu): every tensor argument is a torch CUDA float32 tensor (device pointers cross the C-ABI)
    def hmc_leapfrog_begin(self, theta_parent, grad_parent, theta_child, p_child, step, sign=1.0, p_scale=0.0005, stream_index=0, p_init=None):
        ke = ctypes.c_double(0.0)
        self._chk(self.L.pmp_hmc_leapfrog_begin(self.h, theta_parent.data_ptr(), grad_parent.data_ptr(), theta_child.data_ptr(), p_child.data_ptr(),
                                                 None if p_init is None else p_init.data_ptr(), theta_parent.numel(), step, sign, p_scale, stream_index, ctypes.byref(ke)))
        return ke.value

    def hmc_leapfrog_end(self, p_child, grad_child, step, sign=1.0):
        ke = ctypes.c_double(0.0)
        self._chk(self.L.pmp_hmc_leapfrog_end(self.h, p_child.data_ptr(), grad_child.data_ptr(), p_child.numel(), step, sign, ctypes.byref(ke)))
        return ke.value

    def hmc_accept(self, rule, nets_loss, ke_out, ke_in=None, u=0.5, temperature=1000.0):
        nl = np.ascontiguousarray(nets_loss, dtype=np.float64)
        ko = np.ascontiguousarray(ke_out, dtype=np.float64)
        ki = None if ke_in is None else np.ascontiguousarray(ke_in, dtype=np.float64)
        w = np.zeros(len(nl), dtype=np.float32)
        idx = ctypes.c_int32(-1)
        self._chk(self.L.pmp_hmc_accept(self.h, rule, len(nl), _ptr(nl), _ptr(ko), None if ki is None else _ptr(ki), float(u), float(temperature), _ptr(w), ctypes.byref(idx)))
        return w, int(idx.value)

    def set_data_cnn(self, X, labels, n_offset=0, n_global=None):
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(len(labels), -1)
        if X.shape[1] != 784:
            raise ValueError("the CNN target takes 28x28 images")
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        self._chk(self.L.pmp_set_data_cnn(self.h, _ptr(X), _ptr(lab), len(lab), n_offset, len(lab) if n_global is None else n_global))

    def set_data_fc(self, X, labels, n_offset=0, n_global=None):
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(len(labels), -1)
        lab = np.ascontiguousarray(labels, dtype=np.int64)
        self._chk(self.L.pmp_set_data_fc(self.h, _ptr(X), _ptr(lab), len(lab), n_offset, len(lab) if n_global is None else n_global))
