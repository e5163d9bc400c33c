import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine_lib():
    """The in-tree CUDA engine library (built on demand; nvcc cross-compiles without a GPU)."""
    from soapnuke_b200 import abi, build
    build.build_engine()
    return abi.load_engine()


@pytest.fixture(scope="session")
def coretest_lib():
    from helpers import load_coretest
    return load_coretest()
