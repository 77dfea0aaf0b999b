import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must not silently pass: skip loudly instead.
    if _has_cuda():
        return
    if os.environ.get("HB_EMU") == "1":
        # the CPU model of the library (tests/emu) is loaded instead of libhuffb200.so: the tests that go through the C ABI
        # with host buffers run as they are; the ones that need torch CUDA tensors get the model's engine (see `eng` fixtures)
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
