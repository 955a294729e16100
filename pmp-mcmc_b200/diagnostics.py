"""Chain diagnostics: ESS, ESS/s, MSJD/s, move rate, skewness.

The reference reports ESS per second and MSJD per second for MP / PMP (README.md:56, `ess_compare.pdf`, `msjd_compare.pdf` in
simple_net/MP_and_PMP_ESS_per_sec_and_MSJD_per_sec/) from million-line text dumps processed offline, and the skewness of
batch means in skewness/skewness.ipynb cell 1.  Here the reductions over the trace (means, variances, autocovariances, jump
distances) run on the device over the trace ring (`pmp_trace_diagnostics`) and only a few hundred numbers come back.
"""
import numpy as np


def ess_from_acov(acov, n):
    """Effective sample size per coordinate from autocovariances acov[k, j] (k = 0..L): Geyer's initial positive sequence —
    sum consecutive pairs rho_{2m} + rho_{2m+1} while they stay positive; ESS = n / (-1 + 2 * sum)."""
    acov = np.asarray(acov, dtype=np.float64)
    L, d = acov.shape
    out = np.empty(d)
    for j in range(d):
        if acov[0, j] <= 0:
            out[j] = float(n)
            continue
        rho = acov[:, j] / acov[0, j]
        s, m = 0.0, 0
        while 2 * m + 1 < L:
            pair = rho[2 * m] + rho[2 * m + 1]
            if pair <= 0:
                break
            s += pair
            m += 1
        tau = max(2.0 * s - 1.0, 1.0 / n)
        out[j] = n / tau
    return out


def summarize(ctx, seconds=None, max_lag=256):
    """Diagnostics of the STATE (+ NEXT) trace recorded on `ctx`: {"n", "mean", "var", "ess", "msjd", "move_rate"} and, when the
    run's wall-clock `seconds` is given, "ess_per_s" and "msjd_per_s" (the reference's per-second metrics)."""
    d = ctx.trace_diagnostics(max_lag)
    out = {"n": d["n"], "mean": d["mean"], "var": d["var"], "ess": ess_from_acov(d["acov"], d["n"]), "msjd": d["msjd"], "move_rate": d["move_rate"]}
    if seconds:
        out["ess_per_s"] = out["ess"] / seconds
        out["msjd_per_s"] = out["msjd"] * d["n"] / seconds          # squared distance travelled per second
    return out


def reference_numpy(states, nexts=None, max_lag=256):
    """The same quantities with numpy (what the device kernels are checked against)."""
    x = np.asarray(states, dtype=np.float64)
    n, dim = x.shape
    m = x.mean(0)
    L = min(max_lag, n - 1)
    acov = np.stack([((x[: n - k] - m) * (x[k:] - m)).sum(0) / n for k in range(L + 1)])
    msjd = float((np.diff(x, axis=0) ** 2).sum(1).mean()) if n > 1 else 0.0
    return {"n": n, "mean": m, "var": acov[0], "acov": acov, "msjd": msjd, "move_rate": float((np.asarray(nexts) != 0).mean()) if nexts is not None else -1.0}


def skewness_of_batch_means(samples, num_chains):
    """skewness.ipynb cell 1 (`skewness_fun`): split the sample column into num_chains * 10^i batches (i = 0..4), standardise the
    batch means with the overall mean / unbiased std and average their cubes."""
    s = np.asarray(samples, dtype=np.float64).reshape(-1)
    std, mean = np.std(s, ddof=1), np.mean(s)
    out = []
    for i in range(5):
        nb = num_chains * 10 ** i
        ln = int(s.shape[0] / nb)
        if ln < 1:
            break
        bm = s[: nb * ln].reshape(nb, ln).mean(1)
        out.append(float(np.mean(((bm - mean) / std) ** 3)))
    return out
