import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle.restatement()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref/libpomref.so not built (needs /root/reference)")
    return oracle.reference()
