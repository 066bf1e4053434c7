import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (test infrastructure).  Built on demand with `make -C oracle`."""
    from oracle import loader
    if not os.path.exists(loader.ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return loader.load_oracle()


@pytest.fixture(scope="session")
def product_lib():
    from esvo_b200 import capi
    return capi.load_product()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
