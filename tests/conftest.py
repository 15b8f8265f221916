import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pv_oracle
    pv_oracle.lib()
    return pv_oracle


@pytest.fixture(scope="session")
def pvlib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    import subprocess
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "audiomod_b200", "csrc")])
    from audiomod_b200 import _lib
    return _lib.lib()
