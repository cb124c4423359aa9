"""scs_b200 — B200-native hot path of SelfConcordantSmoothOptimization.jl's proximal SCORE solvers.

Python host mirror of the reference interface over the C ABI in include/scs_b200.h (the Julia shim in
../julia/SCSB200.jl binds the same symbols).  All arithmetic runs in ../libscs_b200.so on an sm_100a GPU.
"""
from ._capi import LIB_PATH, EXPORTS, ScsError, UnsupportedError  # noqa: F401
from .api import (Context, default_context, Problem, ProblemGeneric, get_P, Solution, iterate,  # noqa: F401
                  batch_plan, batch_shard,
                  LogisticLoss, LeastSquaresLoss, QuadFormLoss,
                  PHuberSmootherL1L2, PHuberSmootherIndBox, PHuberSmootherGL, ExponentialSmootherIndBox,
                  LogExpSmootherIndBox, OsBaSmootherL1L2, OsBaSmootherGL,
                  ProxNSCORE, ProxGGNSCORE, ProxLQNSCORE)
from .dist import shard_rows, context_from_env  # noqa: F401
