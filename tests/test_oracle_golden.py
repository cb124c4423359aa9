"""The oracle must reproduce the committed golden iterates bit for bit (guards against silent oracle drift)."""
import json
import os

import numpy as np
import pytest

import cases
from oracle import scs_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLD, name + ".json")) as f:
        g = json.load(f)
    for k in ("x", "obj", "fval", "rel", "objrel", "x_after_step1"):
        g[k] = np.array([float.fromhex(v) for v in g[k]])
    g["pri_res_norm"] = [None if v is None else float.fromhex(v) for v in g["pri_res_norm"]]
    return g


@pytest.mark.parametrize("name", cases.CASES)
def test_oracle_matches_golden(name):
    g = load_golden(name)
    method, model, reg, hmu, kw = cases.build(name, O)
    sol = O.iterate(method, model, reg, hmu, **kw)
    assert sol.epochs == g["epochs"]
    assert len(sol.obj) == len(g["obj"])
    # same machine + same BLAS => identical bits; allow 1e-13 so a different OpenBLAS build does not false-alarm
    np.testing.assert_allclose(sol.x, g["x"], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(np.array(sol.obj), g["obj"], rtol=1e-13)
    assert [int(i) for i in np.nonzero(sol.x)[0]] == g["support"]
