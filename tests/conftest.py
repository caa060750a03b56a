import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle as orc
    orc.build()


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    """PfemContext factory; a missing library or device is an ERROR on a GPU run, never a skip."""
    from pfem_b200.capi import PfemContext

    def make(dim):
        return PfemContext(dim, 0)
    return make
