"""ref_loader.py — import the REFERENCE's own Python definitions from /root/reference.  TEST INFRASTRUCTURE ONLY.

The full tree exists only in the build container; `make -C oracle` stages the few Python scripts this loader reads, unmodified, into
oracle/_ref/pysrc (git-ignored), which travels to the GPU box — only bench.py's reference arm / cpu_baseline leg uses them there.  The reference
scripts run experiments at import time, so only their definition part is exec'd: lb.py lines 1-376,
error.py 1-190, com_dim.py 1-86; matplotlib (absent here) is stubbed.  Used by oracle/make_golden.py to pin
oracle/oracle.py and pmp_oracle.c against the real code, and by the CPU test-suite when the tree is present."""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pysrc")   # oracle/Makefile `stage`: unmodified copies of the scripts
REF = os.environ.get("PMP_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(REF) and os.path.isdir(_STAGED):
    REF = _STAGED          # the GPU box: only the staged Python files travel (enough for load_lb / load_error / load_com_dim / load_fc)


def available():
    """The full reference tree (data files, notebooks, checkpoints) — needed by make_golden.py."""
    return os.path.isdir(REF) and REF != _STAGED


def python_available():
    """The reference's Python definitions (the mounted tree or the staged copies) — enough for the bench's reference arm."""
    return os.path.isfile(os.path.join(REF, "simple_net", "lb.py"))


def _stub_matplotlib():
    if "matplotlib" not in sys.modules:
        m = types.ModuleType("matplotlib")
        p = types.ModuleType("matplotlib.pyplot")
        m.pyplot = p
        sys.modules["matplotlib"] = m
        sys.modules["matplotlib.pyplot"] = p


def _exec_head(relpath, nlines, extra_globals=None):
    _stub_matplotlib()
    path = os.path.join(REF, relpath)
    with open(path, encoding="utf-8-sig") as f:
        src = "".join(f.readlines()[:nlines])
    ns = {"__name__": "pmp_reference_" + os.path.basename(relpath).replace(".", "_")}
    if extra_globals:
        ns.update(extra_globals)
    exec(compile(src, path, "exec"), ns)
    return ns


def load_lb():
    """simple_net/lb.py: BayesNet, BayesNet_o, log_trans_prob, MetropolisOptimizer, GMOptimizer, preMOptimizer, GMpreOptimizerV2."""
    return _exec_head("simple_net/lb.py", 376)


def load_error():
    """simple_sampling/error/error.py: normal, SP, MP, PSP, PMP."""
    return _exec_head("simple_sampling/error/error.py", 190)


def load_com_dim(sigma=0.5):
    """complex_nets/correlation/com_dim.py: normal, transition_prob, PMP (reads the module-global `sigma`, set at L102)."""
    return _exec_head("complex_nets/correlation/com_dim.py", 86, {"sigma": sigma})


def load_banana():
    """banana_distribution from banana_data.ipynb cell 2 lines 1-5."""
    import json
    import numpy as np
    with open(os.path.join(REF, "simple_sampling/error/banana/banana_data.ipynb")) as f:
        nb = json.load(f)
    cell = [c for c in nb["cells"] if c["cell_type"] == "code"][2]
    src = "".join(cell["source"][:5])
    ns = {"np": np}
    exec(compile(src, "banana_data.ipynb#cell2", "exec"), ns)
    return ns["banana_distribution"]


def load_fc(kind, X, y, device="cpu"):
    """complex_nets/Mnist/FC/{MH,MP,PMP}_FC.py: Model, loss and the optimizer class, with the MNIST download replaced by
    injected globals X [n,28,28] float32, y [n] int64 (BASELINE config 5: synthetic MNIST-shaped data)."""
    import copy
    import math
    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch import nn
    path = os.path.join(REF, "complex_nets/Mnist/FC/%s_FC.py" % kind)
    with open(path, encoding="utf-8-sig") as f:
        lines = f.readlines()
    ns = {"torch": torch, "F": F, "nn": nn, "copy": copy, "math": math, "np": np, "tqdm": lambda it: it, "device": device,
          "X": X, "y": y, "x_test": X[:16], "y_test": y[:16], "batch_size": int(X.shape[0]), "N": 7, "alpha": 1e-4}
    ranges = {"PMP": [(20, 44), (77, 186)], "MP": [(20, 36), (69, 74), (76, 164)], "MH": [(17, 34), (66, 71), (72, 135)]}[kind]
    src = ""
    for a, b in ranges:
        src += "".join(lines[a:b]) + "\n"
    exec(compile(src, path, "exec"), ns)
    return ns


def load_cnn(kind, X, y, device="cpu"):
    """complex_nets/Mnist/CNN/{MH,MP,PMP}_CNN.py: Model (conv 1->10 5x5, pool, conv 10->20 3x3, 2000-500-10, log_softmax), loss and the
    optimizer class, with the MNIST download replaced by injected globals X [n,1,28,28] float32, y [n] int64."""
    import copy
    import math
    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch import nn
    path = os.path.join(REF, "complex_nets/Mnist/CNN/%s_CNN.py" % kind)
    with open(path, encoding="utf-8-sig") as f:
        lines = f.readlines()
    ns = {"torch": torch, "F": F, "nn": nn, "copy": copy, "math": math, "np": np, "tqdm": lambda it: it, "device": device,
          "X": X, "y": y, "x_test": X[:16], "y_test": y[:16], "batch_size": int(X.shape[0]), "N": 7, "alpha": 1e-4}
    ranges = {"PMP": [(21, 52), (86, 194)], "MP": [(19, 44), (75, 80), (82, 169)], "MH": [(17, 42), (73, 78), (79, 141)]}[kind]
    src = ""
    for a, b in ranges:
        src += "".join(lines[a:b]) + "\n"
    exec(compile(src, path, "exec"), ns)
    return ns


def load_hmc(kind, X, y, device="cpu"):
    """The HMC variants: complex_nets/Cifar-10/cifar_{SP,MP,PMP}hmc.py (LeNet + the optimizer class) and
    "Bayesian Network Training"/main.py (class bnnPMPHmc), with the dataset download replaced by injected globals X, y."""
    import copy
    import math
    import random
    import numpy as np
    import torch
    import torch.nn.functional as F
    from torch import nn
    rel, ranges = {"PMP": ("complex_nets/Cifar-10/cifar_PMPhmc.py", [(24, 56), (64, 172)]),
                   "MP": ("complex_nets/Cifar-10/cifar_MPhmc.py", [(26, 58), (66, 153)]),
                   "SP": ("complex_nets/Cifar-10/cifar_SPhmc.py", None),
                   "BNN": ("complex_nets/Bayesian Network Training/main.py", [(54, 172)])}[kind]
    path = os.path.join(REF, rel)
    with open(path, encoding="utf-8-sig") as f:
        lines = f.readlines()
    if ranges is None:                       # SP: from `class Flatten` to the end of HMCOptimizer.fit
        a = next(i for i, l in enumerate(lines) if l.startswith("class Flatten"))
        b = next(i for i, l in enumerate(lines) if l.startswith("network = ") or l.startswith("network=") or l.startswith("init_network"))
        ranges = [(a, b)]
    ns = {"torch": torch, "F": F, "nn": nn, "copy": copy, "math": math, "np": np, "random": random, "tqdm": lambda it: it, "device": device,
          "X": X, "y": y, "x_test": X[:8], "y_test": y[:8], "print": lambda *a, **k: None}
    src = ""
    for a, b in ranges:
        src += "".join(lines[a:b]) + "\n"
    exec(compile(src, path, "exec"), ns)
    return ns
