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


@pytest.fixture
def parity_log(record_property, request):
    """parity_log(name, achieved, bound): records the ACHIEVED error of a parity comparison (junit property, stdout,
    and one JSON line in gpurun_out/parity_errors.jsonl when that directory exists — copied to profiles/ per round) and
    asserts it against its bound.  Bounds are kept at <= 2x what the B200 runs of this round measured."""
    import json

    def log(name, achieved, bound):
        achieved, bound = float(achieved), float(bound)
        record_property(name, f"{achieved:.3e} (bound {bound:.1e})")
        print(f"[parity] {request.node.name} :: {name}: achieved {achieved:.3e}  bound {bound:.1e}")
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            try:
                with open(os.path.join(out_dir, "parity_errors.jsonl"), "a") as f:
                    f.write(json.dumps({"test": request.node.name, "what": name, "achieved": achieved,
                                        "bound": bound}) + "\n")
            except OSError:
                pass
        assert achieved <= bound, f"{name}: achieved {achieved:.3e} > bound {bound:.1e}"
    return log
