import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(f[:-len(".trace.gz")] for f in os.listdir(GOLDEN_DIR) if f.endswith(".trace.gz"))


@pytest.fixture(scope="session")
def golden():
    from tracefile import read_trace
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = read_trace(os.path.join(GOLDEN_DIR, name + ".trace.gz"))
        return cache[name]

    return get
