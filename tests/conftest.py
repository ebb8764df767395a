import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """ctypes handle on the CPU oracle (oracle/_build/liborb_oracle.so); built on demand with make."""
    from tests import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def orbx():
    """The product package (ctypes binding of the C-ABI library)."""
    import wut_cuda_orb_slam3_b200 as pkg
    return pkg
