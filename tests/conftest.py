import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests skip on a machine without a CUDA device.  With a device present nothing is
    skipped: a missing product library must fail loudly there (no CPU fallback)."""
    reason = None
    try:
        import torch

        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as e:  # pragma: no cover
        reason = f"torch unavailable: {e}"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import pyoracle

    pyoracle.build()
