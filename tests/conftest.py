import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    from tests import util
    return util.Golden()


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device: skipped (not failed) where there is none."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    # a library that does not load must fail the tests loudly, so only the device count decides
    from usearch12_b200 import capi
    have = capi.lib().usb_device_count() > 0
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device (usb_device_count() == 0)")
        for it in gpu_items:
            it.add_marker(skip)
