"""pytest configuration: registers the ``gpu`` marker and shared fixtures."""

import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The user's forward/backward pass is torch's own (cuDNN); its convolutions default to TF32 on
# sm_100, which is below the fp32 parity tolerance of this path.  The product never touches this flag.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
