import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def og():
    from oracle import oracle
    oracle.build()
    oracle.set_num_threads(min(8, os.cpu_count() or 1))
    return oracle


@pytest.fixture(scope="session")
def vsb():
    import vsb200
    return vsb200.binding


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    torch.cuda.set_device(0)
    return torch
