import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pytorch-a2c-ppo-acktr-gail_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

# tests/test_p2p_gpu.py runs several emulated ranks of the peer-memory exchange on the streams of ONE device: each needs a
# hardware queue of its own (read by the driver when the context is created)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
