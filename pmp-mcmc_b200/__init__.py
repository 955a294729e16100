"""pmp_mcmc_b200 — B200-native hot path of PMP-MCMC (proposal generation, proposals x data likelihood sweep,
prefetch-tree expansion + multi-proposal acceptance) behind the reference's own sampler entry points.

    _lib        ctypes binding of csrc/libpmp_b200.so (C-ABI in include/pmp_b200.h); no CPU fallback
    samplers    lb.py names: BayesNet, MetropolisOptimizer, GMOptimizer, preMOptimizer, GMpreOptimizerV2
    analytic    error.py / com_dim.py names: SP, MP, PSP, PMP, normal, banana_distribution
    fc          PMP_FC.py / MP_FC.py / MH_FC.py names: Model, loss, MetropolisOptimizer, MPOptimizer, PMPOptimizer
    cnn         PMP_CNN.py / MP_CNN.py / MH_CNN.py names (Model, loss, the three optimizers) on the device CNN sweep
    nets        the same three network samplers around an arbitrary loss(net) callable (CNN / LSTM scripts)
    cuda_programs  the .cu experiment programs' main() (time analysis, convergence, ESS dumps) on the device-resident chain
    sinks       trace files in the reference's text formats + the readers its notebooks use
    dist        one-process-per-GPU data sharding helpers (torch.distributed for the rendezvous, NCCL inside the library)
"""
from . import _lib  # noqa: F401
from ._lib import Context, PmpError  # noqa: F401

__all__ = ["Context", "PmpError"]
