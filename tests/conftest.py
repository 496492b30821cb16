import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The C-ABI library is built in-tree and git-ignored: if a fresh checkout runs the tests before
    `__graft_entry__.build()`, build it here (nvcc cross-compiles without a GPU; ~35 s once)."""
    lib = os.path.join(ROOT, "lanedetection_end2end_b200", "liblanefit_b200.so")
    if os.path.exists(lib):
        return
    try:
        from lanedetection_end2end_b200.csrc import build as b
        if os.path.exists(b.NVCC):
            b.build()
    except Exception as e:          # the tests that need the library then say so themselves
        sys.stderr.write("could not build liblanefit_b200.so: %s\n" % e)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _restore_conv_mode():
    """Tests switch ops_net.CONV_MODE; every test starts from (and leaves) the library default."""
    try:
        from lanedetection_end2end_b200 import ops_net
    except Exception:
        yield
        return
    default = ops_net.CONV_MODE
    yield
    ops_net.CONV_MODE = default
