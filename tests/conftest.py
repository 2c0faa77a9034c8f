import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _build_once():
    lib = os.path.join(ROOT, "lc2is_b200", "csrc", "liblc2is_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


_build_once()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
