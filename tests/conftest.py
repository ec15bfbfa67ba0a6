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


@pytest.fixture(autouse=True)
def _reset_lift_mode(request):
    """GPU tests may switch the process-wide arithmetic mode of the lifting kernels; restore the default."""
    yield
    if request.node.get_closest_marker("gpu") is not None:
        from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import ops
        ops.set_lift_mode("tc")
