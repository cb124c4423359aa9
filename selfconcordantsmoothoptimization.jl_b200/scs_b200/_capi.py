"""ctypes binding of include/scs_b200.h — the same symbols julia/SCSB200.jl `ccall`s.

There is deliberately no fallback: if the shared library is missing or no B200 is visible, every
call raises.  Nothing in this package computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCS_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libscs_b200.so")

SCS_OK, SCS_INVALID_ARG, SCS_UNSUPPORTED, SCS_NOT_SPD, SCS_CUDA_ERROR, SCS_NCCL_ERROR, SCS_OOM, SCS_STATE_ERROR = range(8)

LOSS_LOGISTIC, LOSS_LEASTSQUARES, LOSS_QUADFORM = 0, 1, 2
LABELS_LITERAL, LABELS_CONSISTENT = 0, 1
REG_KINDS = {"l1": 0, "l2": 1, "indbox": 2, "gl": 3}
SMOOTH_PHUBER_L1L2, SMOOTH_PHUBER_INDBOX, SMOOTH_PHUBER_GL, SMOOTH_EXP_INDBOX, SMOOTH_LOGEXP_INDBOX, SMOOTH_OSBA_L1L2, SMOOTH_OSBA_GL = range(7)
METHOD_N, METHOD_GGN, METHOD_LQN = 0, 1, 2
WEIGHTS_NEWTON, WEIGHTS_GGN = 0, 1
STAGES = ("forward", "adjoint", "gram", "solve", "vector", "allreduce", "fused", "gram_finalize", "residues", "reserved")

# every symbol include/scs_b200.h declares (tests check the library exports all of them)
EXPORTS = (
    "scs_version", "scs_last_error", "scs_comm_unique_id", "scs_ctx_create", "scs_ctx_destroy", "scs_ctx_sync",
    "scs_ctx_stream", "scs_problem_create", "scs_problem_create_csc", "scs_problem_is_sparse", "scs_problem_sparsify", "scs_problem_create_synthetic", "scs_problem_destroy",
    "scs_problem_read_rows", "scs_set_regularizer", "scs_set_smoother", "scs_set_method", "scs_set_L",
    "scs_set_gram_mode", "scs_get_gram_path", "scs_set_gram_bits", "scs_get_gram_info", "scs_get_gram_signed", "scs_set_stream_mode", "scs_get_stream_path", "scs_set_active_rows", "scs_set_batches", "scs_set_test_problem", "scs_get_test_history", "scs_method_init", "scs_objective", "scs_step", "scs_solve", "scs_loss_eval", "scs_gram", "scs_linear_solve",
    "scs_smoother_eval", "scs_prox", "scs_reg_value", "scs_get_counters", "scs_set_profiling", "scs_get_stage_ms", "scs_measure_i8_peak", "scs_i8_pipe_probe",
)


class ScsError(RuntimeError):
    """Base.error-class failure reported by the library (status code + scs_last_error())."""

    def __init__(self, code, msg):
        super().__init__(f"scs_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class UnsupportedError(ScsError):
    """The request has no GPU implementation; it is rejected instead of being run on the CPU."""


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C selfconcordantsmoothoptimization.jl_b200/csrc`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "scs_version": ([], i32),
        "scs_last_error": ([], C.c_char_p),
        "scs_comm_unique_id": ([vp], i32),
        "scs_ctx_create": ([i32, i32, i32, vp, C.POINTER(vp)], i32),
        "scs_ctx_destroy": ([vp], i32),
        "scs_ctx_sync": ([vp], i32),
        "scs_ctx_stream": ([vp, C.POINTER(u64)], i32),
        "scs_problem_create": ([vp, _dp, i64, i64, i64, _dp, i32, dbl, i32, C.POINTER(vp)], i32),
        "scs_problem_create_csc": ([vp, _ip, _ip, _dp, i64, i64, i64, _dp, i32, dbl, i32, i32, C.POINTER(vp)], i32),
        "scs_problem_is_sparse": ([vp, C.POINTER(i32), _ip], i32),
        "scs_problem_sparsify": ([vp], i32),
        "scs_problem_create_synthetic": ([vp, i64, i64, i64, i64, i32, dbl, i32, u64, dbl, C.POINTER(vp)], i32),
        "scs_problem_destroy": ([vp], i32),
        "scs_problem_read_rows": ([vp, i64, i64, _dp, _dp], i32),
        "scs_set_regularizer": ([vp, i32, dbl, dbl, _ip, i64, _ip, _dp, i64, _dp, i64], i32),
        "scs_set_smoother": ([vp, i32, dbl, _dp, i64, _dp, i64], i32),
        "scs_set_method": ([vp, i32, i32, i32, i32], i32),
        "scs_set_L": ([vp, i32, dbl], i32),
        "scs_method_init": ([vp], i32),
        "scs_set_gram_mode": ([vp, i32], i32),
        "scs_get_gram_path": ([vp, C.POINTER(i32)], i32),
        "scs_set_gram_bits": ([vp, i32], i32),
        "scs_get_gram_info": ([vp, C.POINTER(i32), C.POINTER(i32)], i32),
        "scs_get_gram_signed": ([vp, _ip, C.POINTER(i32)], i32),
        "scs_set_test_problem": ([vp, vp], i32),
        "scs_get_test_history": ([vp, _dp, i64, _ip], i32),
        "scs_set_active_rows": ([vp, i64, i64], i32),
        "scs_set_batches": ([vp, i64, _ip], i32),
        "scs_set_stream_mode": ([vp, i32], i32),
        "scs_get_stream_path": ([vp, C.POINTER(i32)], i32),
        "scs_objective": ([vp, _dp, _dp, _dp], i32),
        "scs_step": ([vp, _dp, _dp, i64, _dp, _dp, _dp], i32),
        "scs_solve": ([vp, _dp, _dp, i64, dbl, dbl, _dp, _dp, _dp, _dp, _dp, _dp, _ip, _ip], i32),
        "scs_loss_eval": ([vp, _dp, i32, _dp, _dp, _dp, _dp, _dp], i32),
        "scs_gram": ([vp, _dp, i32, _dp], i32),
        "scs_linear_solve": ([vp, _dp, _dp, i64, _dp, C.POINTER(i32)], i32),
        "scs_smoother_eval": ([vp, _dp, _dp, _dp], i32),
        "scs_prox": ([vp, _dp, _dp, dbl, _dp], i32),
        "scs_reg_value": ([vp, _dp, _dp], i32),
        "scs_get_counters": ([vp, _ip, i32], i32),
        "scs_set_profiling": ([vp, i32], i32),
        "scs_get_stage_ms": ([vp, _dp, _ip, i32], i32),
        "scs_measure_i8_peak": ([vp, dbl, _dp, _dp], i32),
        "scs_i8_pipe_probe": ([vp, i32, _dp], i32),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def check(code):
    if code == SCS_OK:
        return
    msg = lib().scs_last_error().decode("utf-8", "replace")
    if code == SCS_UNSUPPORTED:
        raise UnsupportedError(code, msg)
    raise ScsError(code, msg)


def dptr(a):
    """Pointer to a C-contiguous / F-contiguous float64 numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(_dp)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int64
    return a.ctypes.data_as(_ip)


def vec(x, m=None):
    v = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
    if m is not None and v.shape[0] != m:
        raise ValueError(f"expected a vector of length {m}, got {v.shape[0]}")
    return v
