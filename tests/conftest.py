import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference")


def _have_gpu():
    try:
        from firecode_b200 import _lib

        return _lib.load().fc_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    ref_ok = os.path.isdir("/root/reference/firecode")
    for item in items:
        if "reference" in item.keywords and not ref_ok:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present on this box"))


@pytest.fixture(scope="session")
def gpu():
    if not _have_gpu():
        pytest.fail("GPU test selected but no CUDA device / library: the CUDA path must run (no fallback)")
    return True
