import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "selfconcordantsmoothoptimization.jl_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def scs():
    """The host mirror package; GPU tests fail loudly (ImportError) if libscs_b200.so was not built."""
    import scs_b200
    scs_b200._capi.lib()
    return scs_b200
