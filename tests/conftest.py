import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_sessionstart(session):
    """The in-tree CUDA library normally travels with the snapshot; if it is missing (fresh checkout), build it
    before any test imports the package (nvcc cross-compiles sm_100a without a GPU)."""
    if not os.path.exists(os.path.join(ROOT, "neural_graph_mapping_b200", "libngm_b200.so")):
        import __graft_entry__

        __graft_entry__.build()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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
