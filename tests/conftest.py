import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_small():
    import torch
    return torch.load(os.path.join(GOLDEN, "unet_small.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_full():
    import torch
    return torch.load(os.path.join(GOLDEN, "unet_full.pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_processing():
    import torch
    return torch.load(os.path.join(GOLDEN, "processing.pt"), weights_only=False)
