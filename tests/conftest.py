"""pytest wiring: `-m "not gpu"` runs on the CPU build box, `-m gpu` on a real B200."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run with -m gpu on a B200")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as ge

    ge.build()
    return ge


@pytest.fixture(scope="session")
def pkg(built):
    return built.load_package()


@pytest.fixture(scope="session")
def orc(built):
    return built.load_oracle()


@pytest.fixture(scope="session")
def golden():
    import json

    with open(os.path.join(REPO, "tests", "golden", "reference_vectors.json")) as f:
        return json.load(f)
