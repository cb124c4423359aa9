"""Reference-run pin: when tools/julia_crosscheck.jl has been run (it needs a Julia binary with the reference package and
its dependencies — absent from the build image), its dump of the REAL package's Solution for the tests/cases.py problems
is diffed here against the committed oracle goldens.  Until then the test skips and parity stays "unpinned at 1e-10"
(DESIGN.md §2).  If `julia` is on PATH and a checkout of the reference is reachable (SCS_REFERENCE_PKG, default /root/reference)
the dump is produced on the fly."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
JDIR = os.path.join(GOLD, "julia")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def _unhex(v):
    return [np.nan if x is None else float.fromhex(x) for x in v]


def _maybe_run_julia(tmp):
    jl, pkg = shutil.which("julia"), os.environ.get("SCS_REFERENCE_PKG", "/root/reference")
    if not jl or not os.path.isdir(os.path.join(pkg, "src")):
        return False
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dump_case_inputs.py"), tmp], check=True)
    r = subprocess.run([jl, f"--project={pkg}", os.path.join(ROOT, "tools", "julia_crosscheck.jl"), tmp, JDIR],
                       capture_output=True, text=True)
    return r.returncode == 0


@pytest.mark.parametrize("name", cases.CASES)
def test_reference_run_matches_oracle_golden(name, tmp_path_factory):
    path = os.path.join(JDIR, name + ".json")
    if not os.path.exists(path) and not _maybe_run_julia(str(tmp_path_factory.getbasetemp() / "scs_cases")):
        pytest.skip("no Julia run of the reference available (tools/julia_crosscheck.jl has not been executed): "
                    "parity is pinned to the reference only through its own 1e-6 / 1e-3 test assertions")
    if not os.path.exists(path):
        pytest.skip("julia ran but produced no dump for this case")
    ref = json.load(open(path))
    gold = json.load(open(os.path.join(GOLD, name + ".json")))
    xr, xg = np.array(_unhex(ref["x"])), np.array(_unhex(gold["x"]))
    assert ref["epochs"] == gold["epochs"]
    assert len(ref["obj"]) == len(gold["obj"])
    assert np.linalg.norm(xr - xg) <= 1e-10 * max(np.linalg.norm(xr), 1e-300)
    orf, og = np.array(_unhex(ref["obj"])), np.array(_unhex(gold["obj"]))
    fin = np.isfinite(orf)
    assert np.array_equal(fin, np.isfinite(og))
    assert np.max(np.abs(orf[fin] - og[fin]) / np.abs(orf[fin])) <= 1e-10
    assert ref["support"] == gold["support"]
