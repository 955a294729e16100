"""Trace sinks in the reference's on-disk formats, so its analysis notebooks run unchanged on traces
produced by the device-resident chains.

The reference's CUDA programs write plain text, one number per line, with C++ `ostream << float` formatting (printf "%g",
six significant digits):

  convergence runs   conv_mh.cu:98-112,157-164 / conv_mp.cu:164-176,254-259,273-282 / conv_pmp.cu:153-162,284-289
      <ALGO>_beta0_<steps>.txt  <ALGO>_beta1_<steps>.txt  <ALGO>_sigma_<steps>.txt   current state after every iteration
      <ALGO>_time<steps>.txt                                                       seconds since the start, after every iteration
      <P>_MPA.txt (MP) / <P>_PMPA.txt (PMP)                                        normalised weights, iteration-major
  ESS / time-analysis dumps   500_MP.cu:155-176,226-264, ess_per_s_MP.cu:164-176,263-271, ess_per_s_PMP.cu:154-166,267-275
      <P>_MPbeta0.txt  <P>_MPbeta_true.txt  <P>_MPsigma_true.txt  <P>_MPA.txt  <P>_MPtime.txt      (MP)
      <P>_beta0.txt    <P>_beta_true.txt    <P>_sigma_true.txt    <P>_A.txt    <P>_time.txt        (PMP)
      rows: the P resampled candidates of every recorded iteration (beta0 / beta1 / sigma columns), the P normalised
      weights of every recorded iteration, and ONE number in the time file (seconds of the whole run)
  data_trans.py:10-16   text column → <name>.npy
  par_conv_analy.ipynb cell 1   reads the convergence files back into pars[3, steps] and times (ms)

Reference quirk kept out: `std::accumulate(A_hat.begin(), A_hat.end(), 0)` (500_MP.cu:215) sums the weights in `int`, so the
"normalised" weights it logs are divided by a truncated sum; here they are divided by the true sum.
"""
import os

import numpy as np


def format_float(v):
    """`ostream << float`: the float32 value widened to double and printed with %g."""
    return "%g" % float(np.float32(v))


def write_column(path, values):
    """One number per line, C++ default float formatting."""
    a = np.asarray(values, dtype=np.float32).reshape(-1).astype(np.float64)
    with open(path, "w") as f:
        np.savetxt(f, a, fmt="%g")
    return path


def read_column(path):
    """How the notebooks read a column back: float(line.strip()) per line."""
    with open(path) as f:
        return np.array([float(line.strip()) for line in f if line.strip()], dtype=np.float64)


def normalised_weights(logw):
    """exp(A) / sum exp(A) per iteration from the log-weights trace [iters, P] (the .cu programs shift by a hand-tuned
    adjust_A instead of the maximum, 500_MP.cu:88-98,207-213)."""
    A = np.asarray(logw, dtype=np.float64)
    w = np.exp(A - A.max(axis=-1, keepdims=True))
    return w / w.sum(axis=-1, keepdims=True)


def write_conv_trace(out_dir, algo, num_steps, states, times_s, weights=None, P=None):
    """conv_{mh,mp,pmp}.cu outputs.  states [steps, 3] (beta0, beta1, sigma); times_s [steps] cumulative seconds."""
    os.makedirs(out_dir, exist_ok=True)
    states = np.asarray(states, dtype=np.float32).reshape(-1, 3)
    stem = os.path.join(out_dir, algo + "_")
    paths = {
        "beta0": write_column(stem + "beta0_%d.txt" % num_steps, states[:, 0]),
        "beta1": write_column(stem + "beta1_%d.txt" % num_steps, states[:, 1]),
        "sigma": write_column(stem + "sigma_%d.txt" % num_steps, states[:, 2]),
        "time": write_column(stem + "time%d.txt" % num_steps, times_s),
    }
    if weights is not None:
        paths["A"] = write_column(os.path.join(out_dir, "%d_%sA.txt" % (P, algo)), np.asarray(weights).reshape(-1))
    return paths


def load_conv_trace(folder, algo, num_steps):
    """par_conv_analy.ipynb cell 1: pars = vstack(beta0s, beta1s, sigmas) [3, steps], times in milliseconds."""
    cols = [read_column(os.path.join(folder, "%s_%s_%d.txt" % (algo, k, num_steps))) for k in ("beta0", "beta1", "sigma")]
    times = read_column(os.path.join(folder, "%s_time%d.txt" % (algo, num_steps))) * 1000
    return np.vstack(cols), times


def write_cuda_dump(out_dir, P, kind, samples, weights, elapsed_s):
    """The ESS / time-analysis dumps.  samples [iters, P, 3]: parameters of the P resampled candidates of every recorded
    iteration (data_log); weights [iters, P] (A_log); elapsed_s: seconds of the whole run (the only line of the time file)."""
    os.makedirs(out_dir, exist_ok=True)
    s = np.asarray(samples, dtype=np.float32).reshape(-1, 3)
    tag = "_MP" if kind == "MP" else "_"
    stem = os.path.join(out_dir, "%d%s" % (P, tag))
    paths = {
        "beta0": write_column(stem + "beta0.txt", s[:, 0]),
        "beta_true": write_column(stem + "beta_true.txt", s[:, 1]),
        "sigma_true": write_column(stem + "sigma_true.txt", s[:, 2]),
        "A": write_column(stem + "A.txt", np.asarray(weights).reshape(-1)),
    }
    with open(stem + "time.txt", "w") as f:
        f.write("%g" % float(elapsed_s))
    paths["time"] = stem + "time.txt"
    return paths


def txt_to_npy(path, out=None):
    """data_trans.py:10-16."""
    a = read_column(path)
    out = out or os.path.splitext(path)[0] + ".npy"
    np.save(out, a)
    return out


def read_data_txt(folder):
    """get_data() of the .cu programs (500_MP.cu:63-76): whitespace-separated floats in data_x.txt / data_y.txt."""
    x = np.loadtxt(os.path.join(folder, "data_x.txt"), dtype=np.float32).reshape(-1)
    y = np.loadtxt(os.path.join(folder, "data_y.txt"), dtype=np.float32).reshape(-1)
    if len(x) != len(y):
        raise ValueError("data_x.txt and data_y.txt differ in length (%d, %d)" % (len(x), len(y)))
    return x, y
