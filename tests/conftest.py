import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build librayhs_b200.so / liborc.so if a source is newer (a no-op on the GPU box: the .so files travel)."""
    from rayhs_b200 import build

    build.build_all()
    yield
