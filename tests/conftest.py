import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel_err(a, b):
    """max|a-b| / max|b| -- the per-tensor relative error BASELINE.json's tolerances refer to."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    if b.numel() == 0:
        return 0.0
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


@pytest.fixture
def scan_impl():
    """Force a selective-scan kernel family for the test (``mtts_set_scan_impl``), back to automatic afterwards."""
    from mamba_tts_project_b200 import _lib
    yield _lib.set_scan_impl
    _lib.set_scan_impl(None)
