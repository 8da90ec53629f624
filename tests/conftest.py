import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


@pytest.fixture(scope="session")
def built():
    """libbftq.so + oracle lib present (built in-tree; never JIT-cached elsewhere)."""
    from bftkv_b200 import build as b
    b.build()
    b.build_oracle()
    return True


@pytest.fixture(scope="session")
def engine(built):
    from bftkv_b200 import Engine
    e = Engine(0)          # raises loudly if the CUDA extension or the GPU is missing
    yield e
    e.close()
