"""Regenerates tests/golden/*.json from the oracle (run from the repo root: python tests/golden/make_golden.py).

What these fixtures are — and are not: the reference cannot run in this image (no julia), so these are NOT
outputs of the reference.  They freeze the oracle restatement's iterates so that (a) the oracle cannot drift
silently and (b) the GPU parity tests compare against committed numbers, not only against a live recomputation.
The oracle itself is pinned to the reference by tests/test_oracle_reference_fixtures.py.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import scs_oracle as O  # noqa: E402


def main():
    out = os.path.dirname(os.path.abspath(__file__))
    for name in cases.CASES:
        method, model, reg, hmu, kw = cases.build(name, O)
        sol = O.iterate(method, model, reg, hmu, **kw)
        rec = {
            "case": name,
            "n": int(model.A.shape[0]),
            "m": int(model.A.shape[1]),
            "A_checksum": float(np.sum(model.A * np.cos(np.arange(model.A.size).reshape(model.A.shape, order="F") % 97))),
            "x": [float.hex(float(v)) for v in sol.x],
            "obj": [float.hex(float(v)) for v in sol.obj],
            "fval": [float.hex(float(v)) for v in sol.fval],
            "pri_res_norm": [None if v is None else float.hex(float(v)) for v in sol.pri_res_norm],
            "rel": [float.hex(float(v)) for v in sol.rel],
            "objrel": [float.hex(float(v)) for v in sol.objrel],
            "epochs": int(sol.epochs),
            "support": [int(i) for i in np.nonzero(sol.x)[0]],
            "x_after_step1": [float.hex(float(v)) for v in sol.iterates[0]],
        }
        with open(os.path.join(out, name + ".json"), "w") as f:
            json.dump(rec, f, indent=0)
        print(name, "epochs", sol.epochs, "hist", len(sol.obj), "nnz", len(rec["support"]))


if __name__ == "__main__":
    main()
