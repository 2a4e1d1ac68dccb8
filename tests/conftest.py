import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from compression_algorithms_b200.device import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ob():
    """oracle bindings (test infrastructure only)"""
    from oracle import bindings
    if not bindings.have_port():
        pytest.skip("oracle port not built: make -C oracle port")
    return bindings
