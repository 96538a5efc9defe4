import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (test infrastructure only), built on demand with gcc."""
    import ctypes
    import __graft_entry__ as ge
    from opesci_fd_b200 import abi
    return abi.bind(ctypes.CDLL(ge.build_oracle()))


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Building needs nvcc only; running needs a GPU."""
    import __graft_entry__ as ge
    from opesci_fd_b200 import abi
    ge.build_cuda()
    return abi.load_library()
