import glob
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    # a hung GPU test must not eat the box's time budget: every test gets a timeout (pytest-timeout is in the image)
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(420))


def golden_names():
    return sorted(os.path.basename(p)[: -len(".json.gz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.json.gz")))


def load_golden(name):
    from tensororder_b200.plan_format import PortablePlan

    return PortablePlan.load(os.path.join(GOLDEN_DIR, name + ".json.gz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden
