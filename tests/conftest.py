import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def read_case(name):
    d = os.path.join(GOLDEN, name)
    meta = {}
    for line in open(os.path.join(d, "case.txt")):
        k, v = line.rstrip("\n").split(" ", 1)
        meta[k] = v
    return d, meta


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
