import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cozk():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("co-zkvms_b200")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle - the checker, never the thing under test in the gpu tier."""
    from oracle import orc as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ctx(cozk):
    """One engine context on cuda:0 for the whole gpu tier.  Fails loudly if the CUDA library is missing."""
    c = cozk.Context()
    yield c
    c.close()
