import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "birdnet-go_b200"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(REPO, "tests", "golden", "birdnet_v24_golden.npz"))


@pytest.fixture(scope="session")
def audio():
    import birdnet_oracle as bo
    x, _ = bo.read_wav(os.path.join(bo.ASSETS, "soundscape.wav"))
    y, _ = bo.read_wav(os.path.join(bo.ASSETS, "tawnyowl.wav"))
    return {"soundscape": x, "tawnyowl": y}


@pytest.fixture(scope="session")
def lib_path():
    from birdnet_b200 import build
    return build.build()
