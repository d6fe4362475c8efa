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


# The no-grad fp16 render has two pipelines: the three stage kernels (default) and the opt-in single fused kernel
# (NGM_RENDER_FUSED=1).  The modules that exercise that render run every test under both.
_BOTH_RENDER_PIPELINES = {"test_gpu_tc", "test_gpu_exchange", "test_gpu_hardening", "test_gpu_fullsize",
                          "test_gpu_tc_multiround", "test_gpu_packed_weights", "test_gpu_edge_cases", "test_gpu_multi"}


@pytest.fixture(autouse=True)
def render_pipeline(request, monkeypatch):
    mode = getattr(request, "param", None)
    if mode is not None:
        monkeypatch.setenv("NGM_RENDER_FUSED", "1" if mode == "fused" else "0")
    yield mode


def pytest_generate_tests(metafunc):
    if metafunc.module.__name__.split(".")[-1] in _BOTH_RENDER_PIPELINES and "render_pipeline" in metafunc.fixturenames:
        metafunc.parametrize("render_pipeline", ["stages", "fused"], indirect=True)
