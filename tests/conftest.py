import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unverified_on_gpu(fn):
    """GPU tests of code written after the round's last GPU run (the gpurun budget was spent): they have passed with the oracle standing
    in for the device (test logic) and the code they exercise compiles for sm_100a, but they have not run on a B200 yet.  Opt in with
    PG_TEST_UNVERIFIED=1; the default `pytest -m gpu` run holds only tests that have been green on the hardware."""
    return pytest.mark.skipif(os.environ.get("PG_TEST_UNVERIFIED") != "1", reason="written after the last GPU run; set PG_TEST_UNVERIFIED=1")(fn)


@pytest.fixture(scope="session")
def oracle_backend():
    import oracle
    return oracle.backend()


@pytest.fixture(scope="session")
def gpu_backend():
    from presto_b200._lib import load
    return load()
