import os
import subprocess
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def hostsim_lib():
    """CPU test double of the device runtime (tests/hostsim); never part of the product."""
    so = os.path.join(ROOT, "tests", "hostsim", "libngb200_hostsim.so")
    r = subprocess.run(["make", "-s", "hostsim"], cwd=ROOT, capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(so):
        pytest.fail("could not build tests/hostsim: " + r.stdout + r.stderr)
    from parity_util import pkg
    return pkg.Library(so)


@pytest.fixture(scope="session")
def cuda_lib():
    from parity_util import pkg
    lib = pkg.library()
    assert lib.backend == "cuda-sm_100a"
    return lib
