import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the CUDA library and the CPU oracle are built (no-op when up to date)."""
    import __graft_entry__ as g

    g.build(only_if_missing=True)
    return True


@pytest.fixture(scope="session")
def gpu(built):
    from vpp_b200 import capi
    import ctypes as C

    n = C.c_int(0)
    rc = capi.lib.vppb_device_count(C.byref(n))
    if rc != 0 or n.value == 0:
        pytest.fail("a test marked gpu ran without a CUDA device: " + capi.lib.vppb_last_error().decode())
    capi.check(capi.lib.vppb_init(0))
    return True
