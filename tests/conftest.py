import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs a copy of the reference (baseline/_ref or /root/reference)")


def pytest_collection_modifyitems(config, items):
    from _phiml import reference_path
    have_ref = reference_path() is not None
    skip_ref = pytest.mark.skip(reason="no copy of the reference (baseline/_ref or /root/reference)")
    for item in items:
        if 'reference' in item.keywords and not have_ref:
            item.add_marker(skip_ref)
