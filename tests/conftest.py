import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (always) and make sure the CUDA library exists (nvcc cross-compiles on CPU)."""
    from ndt_slam_b200 import build
    build.build_oracle()
    if not build.LIB_CUDA.exists():
        build.build_cuda()
    yield
