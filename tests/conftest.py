import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The suites load confignet_b200/lib/libconfignet_b200.so; build it when a fresh checkout has none
    (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "confignet_b200", "lib", "libconfignet_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
