import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


# Test files that import the reference's whole Python stack (scene / gaussian_renderer / train-loop glue staged under
# baseline/_ref/py) and run it under both bindings: collected LAST, so the kernel-level parity files -- which depend on
# nothing but this repo, the oracle and the reference extension -- always run first under `-x`.
_GLUE_FILES = ("test_gpu_dropin_render.py", "test_gpu_fast_glue.py")


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda it: os.path.basename(str(it.fspath)) in _GLUE_FILES)      # stable: order kept within each group
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

TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)
