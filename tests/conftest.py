import os
import sys

import pytest

# Several emulated slabs in one process wait for each other ON THE DEVICE (tests/test_peer_slab_gpu.py):
# give every stream its own hardware queue so that no waiting kernel sits in front of the kernel it waits for.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def oracle_mt():
    from oracle.pyoracle import Oracle
    return Oracle(threads=True)


GOLDEN = os.path.join(ROOT, "tests", "golden")
