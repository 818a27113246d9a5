import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("halo2-svd041_b200")


@pytest.fixture(scope="session")
def handle(pkg):
    # GPU tests must exercise the CUDA library: no skip-on-missing, no fallback.
    h = pkg.Handle()
    yield h
    h.close()
