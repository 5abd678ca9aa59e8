import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    from multi_task_breast_cancer_b200 import build, _lib
    build.build()
    return _lib.load()
