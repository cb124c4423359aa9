#!/usr/bin/env python
"""Writes the inputs of the tests/cases.py problems as raw little-endian float64 files for tools/julia_crosscheck.jl:

    python tools/dump_case_inputs.py /tmp/scs_cases        # <case>_A.bin (column-major), <case>_y.bin, <case>_x0.bin, cases.tsv

The synthetic data comes from the seeded Philox generator in oracle/synth.py, which Julia does not have; dumping the
bits is what guarantees both sides see the same inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/scs_cases"
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "cases.tsv"), "w") as idx:
        for name in cases.CASES:
            A, y, x0 = cases.data(name)
            np.asfortranarray(A).ravel(order="F").astype("<f8").tofile(os.path.join(out, name + "_A.bin"))
            np.asarray(y, dtype="<f8").tofile(os.path.join(out, name + "_y.bin"))
            np.asarray(x0, dtype="<f8").tofile(os.path.join(out, name + "_x0.bin"))
            idx.write(f"{name}\t{A.shape[0]}\t{A.shape[1]}\n")
            print(name, A.shape)


if __name__ == "__main__":
    main()
