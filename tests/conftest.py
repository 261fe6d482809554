import os
import sys

# before torch/MKL run anything: the live-reference tests execute the reference's
# BLAS calls, and MKL's per-CPU kernels differ in FMA use (see oracle/ref_shim.py)
os.environ.setdefault("MKL_CBWR", "COMPATIBLE")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU test")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
