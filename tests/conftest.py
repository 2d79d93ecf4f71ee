import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """The CPU checkers (oracle restatement; compiled reference when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)


@pytest.fixture(scope="session")
def pkg():
    import _pkg
    return _pkg.load()


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    """The product package on a machine with a GPU; -m gpu tests only."""
    lib = pkg.load_library()
    if lib.b2n_device_count() < 1:
        pytest.fail("gpu test selected but libb2nav sees no CUDA device")
    return pkg
