import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree shared library (built by __graft_entry__.build(); the .so travels to the GPU box)."""
    from divergen_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    return _lib.load()
