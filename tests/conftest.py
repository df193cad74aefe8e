import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a host without a CUDA device.  On a GPU box nothing is
    skipped: a missing libb200nuts.so must fail loudly there (there is no CPU fallback to hide behind)."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as exc:                        # pragma: no cover
        reason = "torch unavailable: %s" % exc
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
