import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_usable() -> bool:
    """True when the CUDA library is built and fc_create finds a device (the product has no CPU path)."""
    try:
        from freecappuccino_b200 import lib
        ctx = lib.Context(0)
        ctx.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the `gpu` tests instead of failing them; with an explicit
    `-m gpu` they still run (and fail loudly) so that a missing device or library cannot pass silently."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no CUDA device / libfcapp_cuda.so not built (the product has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
