"""The three descriptions of the C ABI must agree symbol by symbol and argument by argument: the header
(include/scs_b200.h), the ctypes table the tests drive (scs_b200/_capi.py) and the `ccall`s of the Julia shim
(julia/SCSB200.jl) — the shim cannot be executed in this image (no Julia), so this is what keeps it
from drifting.  Arguments are compared by kind: pointer / 32-bit int / 64-bit int / double."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200")


def _header_prototypes():
    txt = open(os.path.join(ROOT, "include", "scs_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    protos = {}
    for m in re.finditer(r"\b(?:const\s+char\s*\*|int)\s+(scs_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        name, args = m.group(1), " ".join(m.group(2).split())
        kinds = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    kinds.append("ptr")
                elif re.match(r"(u?int64_t)\b", a):
                    kinds.append("i64")
                elif a.startswith("double"):
                    kinds.append("f64")
                elif re.match(r"(unsigned\s+)?int\b", a):
                    kinds.append("i32")
                else:
                    raise AssertionError(f"{name}: unrecognised C type in '{a}'")
        protos[name] = kinds
    return protos


def _ctypes_kind(t):
    if t in (C.c_int, C.c_uint):
        return "i32"
    if t in (C.c_int64, C.c_uint64):
        return "i64"
    if t is C.c_double:
        return "f64"
    return "ptr"  # c_void_p, c_char_p, POINTER(...)


def _julia_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cint": "i32", "Int64": "i64", "UInt64": "i64", "Float64": "f64", "Cdouble": "f64"}[t]


def _julia_ccalls():
    txt = open(os.path.join(PKG, "julia", "SCSB200.jl")).read()
    txt = "\n".join(l.split("#")[0] if not l.lstrip().startswith("#") else "" for l in txt.splitlines())
    calls = []
    for m in re.finditer(r"ccall\(\(:(scs_\w+),\s*LIB\),\s*(\w+),\s*\(", txt):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        inner = txt[m.end():i - 1]
        parts, cur, d = [], "", 0
        for ch in inner:
            if ch == "," and d == 0:
                parts.append(cur)
                cur = ""
            else:
                d += {"{": 1, "}": -1}.get(ch, 0)
                cur += ch
        if cur.strip():
            parts.append(cur)
        calls.append((m.group(1), m.group(2), [_julia_kind(p) for p in parts if p.strip()]))
    return calls


def test_header_matches_ctypes_table():
    import sys
    sys.path.insert(0, PKG)
    from scs_b200 import _capi as K
    L = K.lib()
    protos = _header_prototypes()
    assert set(protos) == set(K.EXPORTS), set(protos) ^ set(K.EXPORTS)
    for name, kinds in protos.items():
        fn = getattr(L, name)
        assert [_ctypes_kind(t) for t in fn.argtypes] == kinds, (name, kinds, fn.argtypes)


def test_julia_shim_matches_header():
    protos = _header_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 12
    for name, ret, kinds in calls:
        assert name in protos, f"shim calls {name}, which the header does not declare"
        assert ret == ("Cstring" if name == "scs_last_error" else "Cint"), (name, ret)
        assert kinds == protos[name], (name, kinds, protos[name])
    # the entry points INTEGRATION.md maps the reference's call sites to must all be bound
    bound = {c[0] for c in calls}
    for need in ("scs_ctx_create", "scs_problem_create", "scs_problem_create_csc", "scs_set_regularizer",
                 "scs_set_smoother", "scs_set_method", "scs_set_L", "scs_method_init", "scs_objective", "scs_step",
                 "scs_set_active_rows"):
        assert need in bound, need
