import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the checker libraries (and, where nvcc exists, the product .so) are built."""
    from slak_b200 import build
    if not os.path.exists(build.LIB):
        build.build_lib()
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        build.build_oracle()
    yield


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
