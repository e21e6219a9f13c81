import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# The library compiles per-step specialised kernels (NVRTC, seconds per program) when a program meets a large batch.
# In this process they stay off so that the suite runs in seconds on the templated kernels; tests/test_gpu_jit.py
# forces them for every program in a subprocess and compares bit for bit, and bench.py checks the timed path (which
# uses them) against the CPU port at full size.  HB_JIT set by the caller wins.
os.environ.setdefault("HB_JIT", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
