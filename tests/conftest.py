import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


@pytest.fixture(scope="session")
def native_libs():
    """Builds (if stale) the helper libraries that run on the CPU."""
    from veloci_b200 import build

    return {"index": build.build_index_lib(), "oracle": build.build_oracle()}
