import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(pytest.mark.timeout(180, method="thread"))  # a wedged kernel must fail the test, not hang the box
            if not has_gpu:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
