import os
import sys

import pytest

# several in-process ranks share one GPU in the loopback tests and wait on each other's flags in-stream: give every
# stream its own hardware queue (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def lib_built():
    from dftfe_b200 import build

    return build.build()
