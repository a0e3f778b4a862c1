import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def bp():
    """The product package (directory name has hyphens, so import it through importlib)."""
    import importlib
    return importlib.import_module("dnn-for-speech-enhancement_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    return oracle_py
