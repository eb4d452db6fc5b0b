import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


def _have_gpu():
    try:
        from cfs_spmv_b200 import capi
        return capi.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    """binds the process to cuda:0 through the C ABI; fails loudly (no skip, no
    fallback) when the library or the device is missing under -m gpu."""
    from cfs_spmv_b200 import capi
    capi.init(0)
    return capi
