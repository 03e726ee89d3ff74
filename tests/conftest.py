import os
import sys

import pytest

# The data-parallel tests emulate two ranks on ONE GPU whose kernels wait for each other on the device.  With the default
# 8 hardware work queues, two engines' streams can alias onto one queue and serialise behind a waiting kernel (a deadlock
# that cannot happen across real GPUs); must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# Same emulation: the gradient exchange kernel runs one persistent CTA per SM and waits for its peers' kernels; up to four
# emulated ranks must fit on the one GPU together.
os.environ.setdefault("SCNB_DP_GRID", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
