import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter):
    """Report the symbol-flip count of the session (target 0; a flip is only tolerated on a rounding boundary)."""
    try:
        import common
    except ImportError:
        return
    if common.FLIP_LOG:
        terminalreporter.write_line(common.flip_summary())
        for label, tot, n, bad in common.FLIP_LOG:
            if n:
                terminalreporter.write_line(f"  flips: {label or '?'}: {n} of {tot} ({bad} unexplained)")
