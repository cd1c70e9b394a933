import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0 for the whole GPU session (the ctx is not re-entrant; tests run serially)."""
    import slideo_b200
    c = slideo_b200.Context(slideo_b200.default_config(keep_matches=1))
    yield c
    c.close()
