import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("go-curdleproofs_b200")


@pytest.fixture(scope="session")
def ctx(pkg):
    """A library context on cuda:0.  GPU tests must run the CUDA path: if the
    library or the device is missing this raises (no silent fallback)."""
    c = pkg.Context(0)
    yield c
    c.close()
