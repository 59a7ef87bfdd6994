import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """bsg context on cuda:0.  GPU tests FAIL (not skip) when the extension is missing."""
    import bloomsearch_b200 as bs
    if not _has_gpu():
        pytest.skip("no CUDA device in this process")
    c = bs.Context(0)
    yield c
    c.close()
