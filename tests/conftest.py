"""pytest configuration: the `gpu` marker (tests that need a B200) and shared fixtures.

CPU suite (`-m "not gpu"`): the oracle against the golden vectors generated from the reference, host logic, the
C-ABI library's exported symbols, and a world_size-2 gloo run of the data-parallel host logic.
GPU suite (`-m gpu`): parity of the CUDA path (through the C ABI) with the oracle.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ub():
    """The ctypes mirror of include/unet_b200.h; builds the library if the .so is missing."""
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if not os.path.exists(pkg.LIB_PATH):
        ge.build()
    pkg.lib()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    import unet_oracle
    return unet_oracle


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
