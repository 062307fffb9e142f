import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree (git-ignored); build it when a fresh checkout runs the tests before build()."""
    import importlib

    b = importlib.import_module("disentangled-subject-to-vid_b200._build")
    if not os.path.exists(b.LIB_PATH):
        b.build()
    return b.LIB_PATH
