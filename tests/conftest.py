import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on the B200 box)")


@pytest.fixture(scope="session")
def mavi():
    return pkg


@pytest.fixture(scope="session")
def oracle():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly (no fallback) when it is missing."""
    return pkg.load_library()
