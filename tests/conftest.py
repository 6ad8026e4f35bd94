import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "reference_stored.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_large():
    """Oracle pins at the BASELINE sizes (C oracle run once on the CPU; tests/golden/make_golden_large.py)."""
    with open(os.path.join(ROOT, "tests", "golden", "large_sizes.json")) as f:
        return json.load(f)
