import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a GPU test on a box without a GPU is an error of the invocation, not a skip: the driver runs
    # `-m "not gpu"` here and `-m gpu` on the B200.  Only guard against accidental plain `pytest`.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (run with gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
