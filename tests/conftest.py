import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled by oracle/Makefile (checker only)."""
    from oracle import ref
    if not ref.available() and not ref.build_ref():
        pytest.skip("oracle/_ref/libsolver2d_ref.so not available (reference sources not mounted)")
    return ref.load()


@pytest.fixture(scope="session")
def dev():
    """The product library's device ABI. Fails loudly if the CUDA library is missing."""
    from solver2d_b200 import device
    return device.Device()
