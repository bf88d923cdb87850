import os
import sys

# Ranks of a sharded chain run as threads on ONE device in tests/test_sharded.py: their persistent kernels wait for each other, so
# every stream involved needs its own hardware queue (the default of 8 connections lets two of them share one).  Must be set
# before CUDA initialises; irrelevant for one rank per GPU.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def po():
    """the CPU oracle (test infrastructure), built on demand"""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def brr():
    """the product package; builds the C-ABI library in-tree when it is missing"""
    import bayesrrcpp_b200 as b
    if not os.path.exists(b.LIB_PATH):
        from bayesrrcpp_b200 import build
        build.build()
    return b


HYP = dict(sigma0=0.01, v0E=1e-4, s02E=1e-3, v0G=1e-4, s02G=1e-3)
CVA = [1e-4, 1e-3, 1e-2]
