import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dkd():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def ops(dkd):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dkd_b200 import ops as _ops, _lib
    _lib.load()  # raises if the library is missing: no fallback
    return _ops
