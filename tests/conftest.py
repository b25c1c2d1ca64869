import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """The in-tree .so is git-ignored: a fresh checkout has none until `__graft_entry__.build()` ran. Building it
    here (only when it is missing altogether) is test infrastructure — the product path itself still refuses to
    run without the library (versband_b200._lib.lib)."""
    import shutil

    from versband_b200 import build

    if not build.LIB.exists() and (os.path.exists(build.NVCC) or shutil.which("nvcc")):
        build.build()  # hosts without nvcc: the tests that need the library fail on their own, loudly
    yield
