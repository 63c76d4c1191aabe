import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "www24-rat_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(params=["fp32", "tf32", "fp16"])
def precision(request):
    """run a GPU test in both arithmetic modes of the RAT-block projections."""
    import rat_native
    from rat_native.engine import set_precision
    from tests import gpu_util
    set_precision(request.param)
    gpu_util.PREC["mode"] = request.param
    yield request.param
    set_precision("fp16")     # back to the library default
    gpu_util.PREC["mode"] = "fp16"
