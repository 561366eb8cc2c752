import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native pieces once (no-op when up to date / prebuilt on the GPU box)."""
    import __graft_entry__ as g

    g.build()


def has_gpu() -> bool:
    from pis_b200 import capi

    return capi.load().pisb_device_count() > 0
