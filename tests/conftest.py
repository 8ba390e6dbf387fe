import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:          # tests/parity_helpers.py (a package named `tests` exists elsewhere on some boxes)
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/add_gym")
    skip_ref = pytest.mark.skip(reason="/root/reference is not present on this machine")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)
