import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the unmodified reference compiled in place)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session")
def po():
    from oracle import pyoracle
    pyoracle.port()
    return pyoracle


@pytest.fixture(scope="session")
def synth(pkg):
    return pkg.synth


@pytest.fixture(scope="session")
def cuda(pkg):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test collected on a machine without CUDA; run with -m 'not gpu'")
    pkg.init(0)
    return torch
