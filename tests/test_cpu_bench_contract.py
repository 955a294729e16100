"""The one JSON line `bench.py` prints is a contract with the driver (task statement (4)): this checks the committed line of the final round-2 run
(profiles/r2_bench_line.json, written by `python bench.py` on a B200) for every key the contract names, and bench.py's own flags."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_committed_bench_line_has_every_contract_key():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "proposal-evals/sec" and d["higher_is_better"] is True and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    # value = P * iterations / time of the timed region
    iters = d["config"]["iters_per_step"] * d["steps"]
    assert abs(d["value"] - d["config"]["P"] * iters / (d["ms_per_step"] * 1e-3 * d["steps"])) / d["value"] < 1e-6
    # the other rows of the reference's table and the parity hash travel in the same line
    for k in ("co_scheduled", "n500", "pmp_binary_d10", "analytic", "fc", "cnn", "parity"):
        assert k in d, k
    assert d["parity"]["co_scheduled_chain0_equals_solo"] is True and len(d["parity"]["trace_sha256"]) == 64


def test_bench_flags():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout, flag
