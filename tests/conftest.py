import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cfg():
    from automatedvaletparking_b200.hostcfg import make_avp_config
    return make_avp_config(max_pops=20000)


@pytest.fixture(scope="session")
def native_built():
    """Build libavp_b200.so (nvcc cross-compiles without a GPU) and the oracle."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def device_planner(native_built):
    from automatedvaletparking_b200.batch import DevicePlanner
    os.environ.setdefault("AVP_HOST_TIMEOUT_S", "300")
    dp = DevicePlanner(max_pops=20000)
    yield dp
    dp.close()
