import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs the reference tree (/root/reference, or its install under oracle/_ref)")


def pytest_collection_modifyitems(config, items):
    have_ref = (os.path.isfile("/root/reference/splicedice/SPLICEDICE.py")
                or os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "splicedice", "SPLICEDICE.py")))
    skip_ref = pytest.mark.skip(reason="reference tree not present on this machine")
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
