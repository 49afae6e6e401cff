import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name), "r", encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_train():
    return load_golden("golden_train.json")


@pytest.fixture(scope="session")
def golden_cases():
    return load_golden("golden_cases.json")


@pytest.fixture(scope="session")
def golden_primitives():
    return load_golden("golden_primitives.json")


@pytest.fixture(scope="session")
def taylorswift():
    with open(os.path.join(GOLDEN, "taylorswift.txt"), "r", encoding="utf-8") as f:
        return f.read()
