import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One library context for the GPU tests.  Fails (does not skip) when the CUDA extension or
    the device is missing: -m gpu on a box without either must be red."""
    import cuauv_vision_pipeline_b200 as bv
    c = bv.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ref_available():
    from oracle import ref_balance
    return ref_balance.available()
