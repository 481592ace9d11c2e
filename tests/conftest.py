import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True, scope="session")
def _bounded_cpu_threads():
    # the oracle runs thousands of tiny torch CPU ops; on a many-core GPU host the default
    # (one thread per core) makes each of them slower, not faster
    import torch
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    yield
