"""The C-ABI library must load on a CPU-only box and export every symbol include/scs_b200.h declares.
No compute call is made here (there is no GPU); the error path of scs_ctx_create is exercised instead."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scs_[a-z_A-Z0-9]+)\s*\(", src)))


def test_header_and_binding_agree(scs):
    assert sorted(scs.EXPORTS) == header_symbols()


def test_library_exports_every_declared_symbol(scs):
    lib = C.CDLL(scs.LIB_PATH)
    for name in header_symbols():
        assert hasattr(lib, name), f"libscs_b200.so does not export {name}"
    assert lib.scs_version() == 100


def test_no_cpu_fallback_without_gpu(scs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is covered on CPU boxes")
    with pytest.raises(scs.ScsError) as e:
        scs.Context(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_arbitrary_f_is_rejected_explicitly(scs):
    import numpy as np
    with pytest.raises(scs.UnsupportedError):
        scs.Problem(np.zeros((3, 2)), np.zeros(3), np.zeros(2), lambda A, y, x: 0.0, 0.1, ctx=object())
    with pytest.raises(scs.UnsupportedError):
        scs.ProblemGeneric(np.zeros(2), lambda x: 0.0, 0.1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".jl")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, fn
