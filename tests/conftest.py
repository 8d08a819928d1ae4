import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """C oracle (test infrastructure): built on demand from oracle/zip_oracle.c"""
    from oracle import cbind

    cbind.build()
    cbind.lib()
    return cbind


@pytest.fixture(scope="session")
def ctx():
    from zinc_b200 import default_context

    return default_context()
