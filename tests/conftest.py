import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """GPU tests selected on a box without a device (or without the built CUDA library) must FAIL, not skip: a silent
    skip would hide a missing native path.  Every `gpu` test gets a guard fixture that raises before the test body."""
    for item in items:
        if item.get_closest_marker("gpu") is not None:
            item.fixturenames.insert(0, "_require_gpu_and_native_library")


@pytest.fixture()
def _require_gpu_and_native_library():
    import torch
    assert torch.cuda.is_available(), "a test marked `gpu` was selected but no CUDA device is visible (run with -m 'not gpu' on CPU boxes)"
    import mil_b200
    assert mil_b200.lib().milb200_version() > 0, "libmilb200.so did not load: build it with `python __graft_entry__.py build`"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
