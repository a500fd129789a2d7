import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full BASELINE-size cases (minutes of CPU-oracle time); part of -m gpu, deselect with -m 'gpu and not slow'")


@pytest.fixture(scope="session")
def hydrob200():
    import hydrob200 as m
    return m


@pytest.fixture(scope="session")
def oracle():
    import oracle as m
    m.build()
    return m
