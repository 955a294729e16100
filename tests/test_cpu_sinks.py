"""Trace sinks: the files must look exactly like what the reference's CUDA programs write (C++ `ostream << float`, one number
per line) and must read back through the code the reference's notebooks / data_trans.py use."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from pmp_mcmc_b200 import sinks


def test_float_formatting_matches_cpp_ostream(tmp_path):
    vals = np.array([1.0, -0.99873, 2.0000123, 0.5, 1e-5, 123456.789, 3.4e-12, 1234567.0, 0.1, 100.0, -7.25e20, 0.000123456789], np.float32)
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++ to produce the C++ formatting")
    src = tmp_path / "p.cpp"
    src.write_text("#include <iostream>\n#include <cstdio>\nint main(){float v; while (fread(&v, 4, 1, stdin) == 1) std::cout << v << \"\\n\"; return 0;}\n")
    exe = tmp_path / "p"
    subprocess.run([gxx, "-O1", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], input=vals.tobytes(), stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")[:-1]
    assert out == [sinks.format_float(v) for v in vals]
    p = sinks.write_column(str(tmp_path / "c.txt"), vals)
    assert open(p).read().split("\n")[:-1] == out


def test_conv_trace_files_read_back_like_the_notebook(tmp_path):
    rng = np.random.default_rng(0)
    steps = 50
    states = rng.standard_normal((steps, 3)).astype(np.float32)
    times = np.cumsum(rng.uniform(1e-4, 2e-4, steps))
    w = rng.uniform(size=(steps, 8)); w /= w.sum(1, keepdims=True)
    files = sinks.write_conv_trace(str(tmp_path), "MP", steps, states, times, w, 8)
    assert sorted(os.path.basename(f) for f in files.values()) == sorted(["MP_beta0_50.txt", "MP_beta1_50.txt", "MP_sigma_50.txt", "MP_time50.txt", "8_MPA.txt"])
    pars, t_ms = sinks.load_conv_trace(str(tmp_path), "MP", steps)          # par_conv_analy.ipynb cell 1
    assert pars.shape == (3, steps)
    np.testing.assert_allclose(pars.T, states, rtol=1e-5, atol=1e-12)       # six significant digits
    np.testing.assert_allclose(t_ms, times * 1000, rtol=1e-5)
    assert len(sinks.read_column(files["A"])) == steps * 8


@pytest.mark.parametrize("kind,names", [("MP", ["16_MPbeta0.txt", "16_MPbeta_true.txt", "16_MPsigma_true.txt", "16_MPA.txt", "16_MPtime.txt"]),
                                         ("PMP", ["16_beta0.txt", "16_beta_true.txt", "16_sigma_true.txt", "16_A.txt", "16_time.txt"])])
def test_cuda_dump_files(tmp_path, kind, names):
    rng = np.random.default_rng(1)
    samples = rng.standard_normal((20, 16, 3)).astype(np.float32)
    w = sinks.normalised_weights(rng.standard_normal((20, 16)) * 50)
    np.testing.assert_allclose(w.sum(1), 1.0, rtol=1e-12)
    files = sinks.write_cuda_dump(str(tmp_path), 16, kind, samples, w, 1.2345678)
    assert sorted(os.path.basename(f) for f in files.values()) == sorted(names)
    np.testing.assert_allclose(sinks.read_column(files["sigma_true"]), samples[:, :, 2].reshape(-1), rtol=1e-5, atol=1e-12)
    assert float(open(files["time"]).readline()) == pytest.approx(1.23457)    # skewness.ipynb cell 0: float(file.readline())
    npy = sinks.txt_to_npy(files["sigma_true"])                                 # data_trans.py
    assert np.load(npy).shape == (20 * 16,)


def test_read_data_txt(tmp_path):
    x = np.linspace(-1, 1, 37).astype(np.float32)
    (tmp_path / "data_x.txt").write_text(" ".join("%g" % v for v in x))
    (tmp_path / "data_y.txt").write_text("\n".join("%g" % (2 * v) for v in x))
    xr, yr = sinks.read_data_txt(str(tmp_path))
    np.testing.assert_allclose(xr, x, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(yr, 2 * x, rtol=1e-5, atol=1e-7)
