import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)

    return load


@pytest.fixture(autouse=True)
def _settle_gpu_between_tests(request):
    """GPU tests build models, CUDA graphs and pinned staging buffers; what one test leaves behind (reference cycles
    holding graphs, streams, pinned blocks) is collected and the device drained BEFORE the next test starts, not at
    whatever point of its graph capture the garbage collector happens to run."""
    yield
    if "gpu" in request.keywords:
        import gc

        import torch

        if torch.cuda.is_available():
            gc.collect()
            torch.cuda.synchronize()
