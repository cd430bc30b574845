import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle import cpu_oracle
    return cpu_oracle.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import cpu_oracle
    if not cpu_oracle.ref_available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return cpu_oracle.Ref()
